"""Developer probe: fp32 forward (3xTF32) + backward (bf16x3 tensor-core kernels) against the dense oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dense_attention as da
from tf_flash_attention_b200 import _capi, flash_attention as fa

def run(dims, rule, mode, w, s, c, batch, d, vd, qs, ks, seed=0):
    rng = np.random.default_rng(seed)
    Q, K, V, dO = da.random_inputs(rng, np.float32, batch, d, vd, qs, ks)
    ref = da.attention(Q, K, V, dims, rule, mode, w, s, c, dO=dO)
    tq, tk, tv = (torch.from_numpy(x).cuda().requires_grad_(True) for x in (Q, K, V))
    if rule == "full": O = (fa.full_1d if dims == 1 else fa.full_2d)(tq, tk, tv, mode)
    elif rule == "causal": O = (fa.causal_1d if dims == 1 else fa.causal_2d)(tq, tk, tv, mode)
    else: O = (fa.local_1d if dims == 1 else fa.local_2d)(tq, tk, tv, w, s, c, mode)
    fp = _capi.lib.fa_last_path()
    dQ, dK, dV = torch.autograd.grad(O, (tq, tk, tv), torch.from_numpy(dO).cuda())
    torch.cuda.synchronize()
    bp = _capi.lib.fa_last_path()
    msg = f"{dims}d {rule:6s} {mode:11s} w{w} s{s} c{int(c)} b{batch} d{d} q{qs} k{ks} fwd={fp} bwd={bp} O={np.abs(O.detach().cpu().numpy()-ref['O']).max():.2e}"
    for n, g in (("dQ", dQ), ("dK", dK), ("dV", dV)):
        g = g.cpu().numpy().astype(np.float64)
        e = np.abs(g - ref[n]) / np.maximum(1, np.abs(ref[n]))
        idx = np.unravel_index(np.argmax(e), e.shape)
        msg += f" {n}={e.max():.2e}@{idx[-1]}(ref {ref[n][idx]:.1f})"
    print(msg, flush=True)

if len(sys.argv) > 1:
    _capi.lib.fa_set_path_override(int(sys.argv[1]))
for c in [
    (1, "full", "none_front", 1, 0, 0, (1,), 64, 64, (128,), (64,)),
    (1, "full", "none_front", 1, 0, 0, (2,), 64, 64, (256,), (320,)),
    (1, "causal", "none_front", 1, 0, 0, (2,), 64, 64, (512,), (512,)),
    (1, "causal", "scale_front", 1, 0, 0, (2,), 64, 64, (200,), (328,)),
    (1, "local", "none_front", 5, 1, 1, (2,), 64, 64, (520,), (520,)),
    (2, "local", "none_front", 4, 0, 1, (2,), 64, 64, (24, 32), (24, 32)),
    (1, "full", "scale_end", 1, 0, 0, (2,), 64, 64, (1024,), (8192,)),
]:
    run(*c)
