"""Build/packaging — same make-driven pattern as the reference's setup.py (setup.py:26-71: a custom
build_ext shells out to `make`), but sm_100a only and with no TensorFlow requirement: the C-ABI
library libfa_b200.so is always built; the TensorFlow op shim only where TensorFlow is importable."""
import os
import subprocess

from setuptools import Extension, setup
from setuptools.command.build_ext import build_ext

ROOT = os.path.dirname(os.path.abspath(__file__))


class MakeBuild(build_ext):
    def run(self):
        csrc = os.path.join(ROOT, "tf_flash_attention_b200", "csrc")
        subprocess.check_call(["make", "-j", str(min(8, os.cpu_count() or 1))], cwd=csrc)
        subprocess.check_call(["make", "tf_shim"], cwd=csrc)


setup(
    name="tf_flash_attention_b200",
    version="0.1.0",
    description="B200-native (sm_100a, tcgen05/TMEM/TMA) drop-in engine for tf_flash_attention",
    packages=["tf_flash_attention_b200", "tf_flash_attention_b200.tests"],
    package_data={"tf_flash_attention_b200": ["libfa_b200.so", "kernel/flash_attention.so"]},
    ext_modules=[Extension("tf_flash_attention_b200._native", sources=[])],
    cmdclass={"build_ext": MakeBuild},
)
