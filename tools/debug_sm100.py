"""Developer probe (not a test): runs the tcgen05 paths on a few shapes and prints error stats."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import dense_attention as da
from tf_flash_attention_b200 import _capi, flash_attention as fa

def run(dims, rule, mode, w, s, c, batch, d, vd, qs, ks, seed=0, bwd=False):
    rng = np.random.default_rng(seed)
    Q, K, V, dO = da.random_inputs(rng, np.float16, batch, d, vd, qs, ks)
    ref = da.attention(Q, K, V, dims, rule, mode, w, s, c, dO=dO if bwd else None)
    tq, tk, tv = (torch.from_numpy(x).cuda().requires_grad_(bwd) for x in (Q, K, V))
    if rule == "full":
        O, l, m = (fa.full_1d if dims == 1 else fa.full_2d)(tq, tk, tv, mode, True)
    elif rule == "causal":
        O, l, m = (fa.causal_1d if dims == 1 else fa.causal_2d)(tq, tk, tv, mode, True)
    else:
        O, l, m = (fa.local_1d if dims == 1 else fa.local_2d)(tq, tk, tv, w, s, c, mode, True)
    torch.cuda.synchronize()
    path = _capi.lib.fa_last_path()
    On = O.detach().cpu().numpy().astype(np.float64)
    err = np.abs(On - ref["O"])
    bad = np.argwhere(err > 2e-3)
    msg = f"{dims}d {rule:6s} {mode:11s} w{w} s{s} c{int(c)} b{batch} d{d} vd{vd} q{qs} k{ks} path={path} O err max={err.max():.3e} nbad={len(bad)} nan={np.isnan(On).sum()}"
    if len(bad):
        msg += f" first_bad={bad[0].tolist()} got={On[tuple(bad[0])]:.4f} ref={ref['O'][tuple(bad[0])]:.4f}"
    live = np.isfinite(ref["m"])
    ln, mn = l.cpu().numpy().astype(np.float64), m.cpu().numpy().astype(np.float64)
    if live.any():
        lse = mn[live] + np.log(np.maximum(ln[live], 1e-300))
        msg += f" lse_err={np.abs(lse - (ref['m'][live] + np.log(ref['l'][live]))).max():.2e}"
    if bwd:
        dQ, dK, dV = torch.autograd.grad(O, (tq, tk, tv), torch.from_numpy(dO).cuda())
        torch.cuda.synchronize()
        msg += f" bwd_path={_capi.lib.fa_last_path()}"
        for n, g in (("dQ", dQ), ("dK", dK), ("dV", dV)):
            g = g.cpu().numpy().astype(np.float64)
            e = np.abs(g - ref[n]) / np.maximum(1, np.abs(ref[n]))
            msg += f" {n}={e.max():.2e}(nan={np.isnan(g).sum()})"
    print(msg, flush=True)

if __name__ == "__main__":
    bwd = "bwd" in sys.argv
    cases = [
        (1, "full", "none_front", 1, 0, 0, (1,), 128, 128, (256,), (128,)),
        (1, "full", "none_front", 1, 0, 0, (1,), 128, 128, (256,), (256,)),
        (1, "full", "none_front", 1, 0, 0, (2,), 128, 128, (512,), (640,)),
        (1, "causal", "none_front", 1, 0, 0, (2,), 128, 128, (1024,), (1024,)),
        (1, "causal", "none_front", 1, 0, 0, (3,), 64, 64, (768,), (768,)),
        (1, "full", "none_front", 1, 0, 0, (2,), 64, 128, (200,), (328,)),
        (1, "causal", "scale_end", 1, 0, 0, (2,), 128, 64, (128,), (1024,)),
        (1, "local", "scale_front", 32, 0, 0, (2,), 64, 64, (512,), (1024,)),
        (1, "local", "none_front", 5, 2, 1, (1,), 128, 128, (520,), (520,)),
        (2, "local", "none_front", 4, 0, 1, (2,), 64, 64, (24, 32), (24, 32)),
        (2, "causal", "scale_front", 1, 0, 0, (1,), 128, 128, (16, 24), (32, 24)),
        (1, "causal", "none_front", 1, 0, 0, (4,), 128, 128, (4096,), (4096,)),
    ]
    for c in cases:
        run(*c, bwd=bwd)
