"""One channel-first and one channel-last C2 forward (+ backward with --backward) for an ncu capture:
ncu --set full --clock-control none -k regex:fwd_kernel -o gpurun_out/cl python tools/ncu_channel_last.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tf_flash_attention_b200 import flash_attention as fa  # noqa: E402

bwd = "--backward" in sys.argv
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda: (torch.rand((4, 16, 128, 8192), generator=g, device="cuda") * 4 - 2).half()  # noqa: E731
Q, K, V, dO = mk(), mk(), mk(), mk()
cl = lambda x: x.permute(0, 3, 1, 2).contiguous()  # noqa: E731
for layout, (q, k, v, d_o) in (("channel_first", (Q, K, V, dO)), ("channel_last", (cl(Q), cl(K), cl(V), cl(dO)))):
    if bwd:
        q.requires_grad_(True), k.requires_grad_(True), v.requires_grad_(True)
        o = fa.causal_1d(q, k, v, "none_front", layout=layout)
        torch.autograd.grad(o, (q, k, v), d_o)
    else:
        with torch.no_grad():
            fa.causal_1d(q, k, v, "none_front", layout=layout)
torch.cuda.synchronize()
print("done")
