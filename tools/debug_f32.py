import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dense_attention as da
from tf_flash_attention_b200 import _capi, flash_attention as fa
def run(dims, rule, mode, w, s, c, batch, d, vd, qs, ks, seed=0):
    rng = np.random.default_rng(seed)
    Q, K, V, dO = da.random_inputs(rng, np.float32, batch, d, vd, qs, ks)
    ref = da.attention(Q, K, V, dims, rule, mode, w, s, c)
    tq, tk, tv = (torch.from_numpy(x).cuda() for x in (Q, K, V))
    if rule == "full": O, l, m = (fa.full_1d if dims == 1 else fa.full_2d)(tq, tk, tv, mode, True)
    elif rule == "causal": O, l, m = (fa.causal_1d if dims == 1 else fa.causal_2d)(tq, tk, tv, mode, True)
    else: O, l, m = (fa.local_1d if dims == 1 else fa.local_2d)(tq, tk, tv, w, s, c, mode, True)
    torch.cuda.synchronize()
    path = _capi.lib.fa_last_path()
    On = O.cpu().numpy().astype(np.float64)
    err = np.abs(On - ref["O"])
    live = np.isfinite(ref["m"])
    ln, mn = l.cpu().numpy().astype(np.float64), m.cpu().numpy().astype(np.float64)
    lse = mn[live] + np.log(np.maximum(ln[live], 1e-300))
    print(f"{dims}d {rule:6s} {mode:11s} w{w} s{s} c{int(c)} b{batch} d{d} vd{vd} q{qs} k{ks} path={path} O err max={err.max():.3e} nbad(1e-5)={(err>1e-5).sum()} nan={np.isnan(On).sum()} lse_err={np.abs(lse-(ref['m'][live]+np.log(ref['l'][live]))).max():.2e} m_err={np.abs(mn[live]-ref['m'][live]).max():.2e}", flush=True)
for c in [
    (1, "full", "none_front", 1, 0, 0, (1,), 64, 64, (128,), (64,)),
    (1, "full", "none_front", 1, 0, 0, (2,), 64, 64, (256,), (320,)),
    (1, "causal", "none_front", 1, 0, 0, (2,), 64, 64, (512,), (512,)),
    (1, "full", "scale_end", 1, 0, 0, (2,), 64, 64, (1024,), (8192,)),
    (1, "local", "scale_front", 32, 0, 0, (8,), 32, 16, (1024,), (2048,)),
    (1, "causal", "scale_front", 1, 0, 0, (2,), 32, 32, (100,), (204,)),
    (2, "local", "none_front", 4, 0, 1, (2,), 64, 64, (24, 32), (24, 32)),
    (1, "local", "none_front", 2, 0, 0, (2,), 32, 32, (640,), (128,)),
]:
    run(*c)
