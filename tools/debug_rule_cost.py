"""Developer tool: per-kernel times of the same shape under different rules (is a rule's bookkeeping showing up?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tf_flash_attention_b200 import _capi, flash_attention as fa
dt = {"f16": torch.float16, "f32": torch.float32, "f64": torch.float64}[sys.argv[1] if len(sys.argv) > 1 else "f64"]
dims = int(sys.argv[2]) if len(sys.argv) > 2 else 1
shape = (1, 8, 32, 1024) if dims == 1 else (1, 8, 32, 32, 32)
g = torch.Generator(device="cuda").manual_seed(0)
Q, K, V = ((torch.rand(shape, generator=g, device="cuda", dtype=torch.float32) * 4 - 2).to(dt).requires_grad_(True) for _ in range(3))
dO = (torch.rand(shape, generator=g, device="cuda", dtype=torch.float32) * 4 - 2).to(dt)
big = max(shape[3:])
calls = {"full": lambda: (fa.full_1d if dims == 1 else fa.full_2d)(Q, K, V, "none_front"),
         "causal": lambda: (fa.causal_1d if dims == 1 else fa.causal_2d)(Q, K, V, "none_front"),
         "local(w=max)": lambda: (fa.local_1d if dims == 1 else fa.local_2d)(Q, K, V, big, 0, False, "none_front"),
         "local_causal(w=max)": lambda: (fa.local_1d if dims == 1 else fa.local_2d)(Q, K, V, big, 0, True, "none_front")}
for name, fn in calls.items():
    for _ in range(3):
        O = fn(); torch.autograd.grad(O, (Q, K, V), dO)
    torch.cuda.synchronize()
    _capi.lib.fa_kernel_timing(1)
    for _ in range(5):
        O = fn(); torch.autograd.grad(O, (Q, K, V), dO)
    torch.cuda.synchronize()
    _capi.lib.fa_kernel_timing(0)
    per = {}
    for k, ms in _capi.kernel_timings():
        per.setdefault(k, []).append(ms)
    print(name, {k: round(float(np.mean(v)) * 1000, 1) for k, v in per.items()}, "us", flush=True)
