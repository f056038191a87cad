import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.debug_sm100 import run
import ast
case = ast.literal_eval(sys.argv[1])
run(*case, bwd="bwd" in sys.argv)
