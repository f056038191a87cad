"""Developer tool: dump the per-event clock stamps of CTA 0 of the fwd kernel (FA_DBG_TIMELINE build).
Optional argument: a fa_set_path_override value (10 / 11 / 12 = the forward hand-off variants)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tf_flash_attention_b200 import _capi, flash_attention as fa
buf = torch.zeros(8 * 128 * 4, dtype=torch.int64, device="cuda")
_capi.lib.fa_debug_set_buffer.argtypes = [C.c_void_p]
_capi.lib.fa_debug_set_buffer(buf.data_ptr())
B, d, S = 8, 128, 8192
g = torch.Generator(device="cuda").manual_seed(0)
Q, K, V = ((torch.rand((B, d, S), generator=g, device="cuda") * 4 - 2).half() for _ in range(3))
if len(sys.argv) > 1:
    _capi.lib.fa_set_path_override(int(sys.argv[1]))
for _ in range(3):
    fa.causal_1d(Q, K, V, "none_front")
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(8, 128, 4)
t0 = t[t > 0].min()
names = ["WG0 S seen", "WG0 ld done", "WG0 exps done", "WG0 arrived", "WG1 S seen", "WG1 ld done", "WG1 exps done", "WG1 arrived", "MMA p0 seen", "MMA qk0 iss", "MMA p1 seen", "MMA qk1 iss"]
print("j  " + " ".join(f"{n:>13s}" for n in names))
for j in range(16, 40):
    row = list(t[0, j]) + list(t[1, j]) + [t[2, j, 0], t[2, j, 1], t[3, j, 0], t[3, j, 1]]
    print(f"{j:2d} " + " ".join(f"{(int(v) - int(t0)) if v else -1:13d}" for v in row))
d = t[0, 2:60]
print("WG0 avg: S->ld", (d[:, 1] - d[:, 0]).mean(), " ld->exps", (d[:, 2] - d[:, 1]).mean(), " exps->arrive", (d[:, 3] - d[:, 2]).mean(), " period", np.diff(d[:, 0]).mean())
d = t[1, 2:60]
print("WG1 avg: S->ld", (d[:, 1] - d[:, 0]).mean(), " ld->exps", (d[:, 2] - d[:, 1]).mean(), " exps->arrive", (d[:, 3] - d[:, 2]).mean())
m0, m1 = t[2, 2:60], t[3, 2:60]
print("MMA avg (WG0): p_half seen -> pv_lo issued", (m0[:, 3] - m0[:, 2]).mean(), " pv_lo issued -> p_ready seen", (m0[:, 0] - m0[:, 3]).mean(),
      " p_ready seen -> qk issued", (m0[:, 1] - m0[:, 0]).mean(), " qk0 issued -> p_half(WG1) seen", (m1[:, 2] - m0[:, 1]).mean(),
      " period", np.diff(m0[:, 0]).mean())
print("WG0: exps done(arrive issue) rel S seen:", (t[0, 2:60, 2] - t[0, 2:60, 0]).mean(), " WG1 S seen rel WG0 S seen", (t[1, 2:60, 0] - t[0, 2:60, 0]).mean())
print("MMA avg: p0 seen -> qk0 issued", (m0[:, 1] - m0[:, 0]).mean(), " p1 seen -> qk1 issued", (m1[:, 1] - m1[:, 0]).mean(), " WG0 arrive -> MMA sees", (m0[:, 0] - t[0, 2:60, 3]).mean(), " qk0 issued -> WG0 sees next S", (t[0, 3:61, 0] - m0[:, 1]).mean())

for w in (0, 1):
    a, e = t[w, 2:60], t[4 + w, 2:60]
    print(f"WG{w} fine: ld done -> max done", (e[:, 0] - a[:, 1]).mean(), " -> 2 chunks exp'd + stored", (e[:, 1] - e[:, 0]).mean(),
          " -> chunk 3 exp'd", (e[:, 2] - e[:, 1]).mean(), " -> wait::st + p_half arrive", (e[:, 3] - e[:, 2]).mean(),
          " -> chunk 4 exp'd, sums (exps done)", (a[:, 2] - e[:, 3]).mean())
