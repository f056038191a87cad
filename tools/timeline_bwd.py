"""Developer tool: per-event clock stamps of CTA 0 (key tile 0 of head 0: all 128 query sub-tiles live) of the
fused backward kernel, from a -DFA_DBG_TIMELINE build (tools/build_variant.sh tl fa_bwd_f16_sm100.cu -DFA_DBG_TIMELINE;
FA_B200_LIB=build_ab/libfa_tl.so python tools/timeline_bwd.py)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tf_flash_attention_b200 import _capi, flash_attention as fa
buf = torch.zeros(4 * 128 * 4, dtype=torch.int64, device="cuda")
_capi.lib.fa_debug_set_buffer_bwd.argtypes = [C.c_void_p]
_capi.lib.fa_debug_set_buffer_bwd(buf.data_ptr())
B, d, S = 8, 128, 8192
g = torch.Generator(device="cuda").manual_seed(0)
Q, K, V, dO = ((torch.rand((B, d, S), generator=g, device="cuda") * 4 - 2).half() for _ in range(4))
Q.requires_grad_(True); K.requires_grad_(True); V.requires_grad_(True)
for _ in range(3):
    O = fa.causal_1d(Q, K, V, "none_front")
    torch.autograd.grad(O, (Q, K, V), dO)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(4, 128, 4)
t0 = t[t > 0].min()
names = ["SM s_full", "SM ld done", "SM math done", "SM arrived", "MMA p_ready", "MMA S issued", "MMA dq_free", "MMA dP issued", "RED dq_full", "RED freed", "RED done"]
print("t   " + " ".join(f"{n:>13s}" for n in names))
for j in range(40, 60):
    row = list(t[0, j]) + list(t[1, j]) + list(t[2, j, :3])
    print(f"{j:3d} " + " ".join(f"{(int(v) - int(t0)) if v else -1:13d}" for v in row))
sl = slice(10, 120)
sm, mm, rd = t[0, sl].astype(np.float64), t[1, sl].astype(np.float64), t[2, sl].astype(np.float64)
print("period per sub-tile (MMA p_ready to p_ready):", np.diff(mm[:, 0]).mean())
print("softmax: s_full->ld", (sm[:, 1] - sm[:, 0]).mean(), " ld->math", (sm[:, 2] - sm[:, 1]).mean(), " math->arrived", (sm[:, 3] - sm[:, 2]).mean(), " total", (sm[:, 3] - sm[:, 0]).mean())
print("softmax arrive -> MMA sees p_ready:", (mm[:, 0] - sm[:, 3]).mean())
print("MMA: p_ready -> S issued", (mm[:, 1] - mm[:, 0]).mean(), " wait dq_free", (mm[:, 2] - mm[:, 1]).mean(), " -> dP issued", (mm[:, 3] - mm[:, 2]).mean())
print("MMA dP issued(t) -> softmax sees s_full(t+2):", (t[0, 12:122, 0] - t[1, 10:120, 3]).astype(np.float64).mean())
print("reduce: MMA p_ready -> dq_full seen", (rd[:, 0] - mm[:, 0]).mean(), " dq_full -> freed", (rd[:, 1] - rd[:, 0]).mean(), " freed -> done", (rd[:, 2] - rd[:, 1]).mean())
x3 = t[3, sl].astype(np.float64)
print("MMA detail: p_ready -> dq issued", (x3[:, 0] - mm[:, 0]).mean(), " -> dv/dk issued", (x3[:, 1] - x3[:, 0]).mean(), " -> stage t+2 full", (x3[:, 2] - x3[:, 1]).mean(), " -> S issued", (mm[:, 1] - x3[:, 2]).mean())
print("producer: loads of sub-tile t+2 issued -> MMA sees stage full:", (t[3, 10:120, 2] - t[3, 12:122, 3]).astype(np.float64).mean(), "  MMA p_ready(t) -> producer issues loads(t+2):", (t[3, 12:122, 3] - t[1, 10:120, 0]).astype(np.float64).mean())
