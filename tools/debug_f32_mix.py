"""Developer probe: which side limits the fp32 gradient accuracy? forward {generic, 3xTF32} x backward {generic, bf16x3}."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dense_attention as da
from tf_flash_attention_b200 import _capi, flash_attention as fa
for case in [(1, "causal", "scale_end", (2,), (1000,), (88,)), (1, "causal", "none_front", (2, 2), (1024,), (1024,))]:
    dims, rule, mode, batch, qs, ks = case
    rng = np.random.default_rng(3)
    Q, K, V, dO = da.random_inputs(rng, np.float32, batch, 64, 64, qs, ks)
    ref = da.attention(Q, K, V, dims, rule, mode, 1, 0, 0, dO=dO)
    tq, tk, tv, tdo = (torch.from_numpy(x).cuda() for x in (Q, K, V, dO))
    for fo in (1, 0):
        _capi.lib.fa_set_path_override(fo)
        O, l, m = fa.causal_1d(tq, tk, tv, mode, True)
        torch.cuda.synchronize()
        for bo in (1, 0):
            _capi.lib.fa_set_path_override(bo)
            g = fa.attention_backward(1, "causal", tq, tk, tv, O, l, m, tdo, mode)
            torch.cuda.synchronize()
            errs = [float((np.abs(x.cpu().numpy().astype(np.float64) - ref[n]) / np.maximum(1, np.abs(ref[n]))).max()) for x, n in zip(g, ("dQ", "dK", "dV"))]
            print(case[3:], "fwd", "generic" if fo else "3xTF32", "bwd", "generic" if bo else "bf16x3", "path", _capi.lib.fa_last_path(), " ".join(f"{e:.2e}" for e in errs), flush=True)
_capi.lib.fa_set_path_override(0)
