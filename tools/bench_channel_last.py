"""Channel-last (f3) against channel-first on the same problem: device time of the forward and of forward + backward,
(a) channel-first tensors (the reference's layout), (b) channel-last tensors read directly through the 4-D tensor
maps, (c) channel-last tensors through the adapter kernel either side of the channel-first op (what round 1 shipped).
Usage: python tools/bench_channel_last.py [--batch 16 --heads 16 --d 128 --seq 8192 --rule causal] [--steps 10]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tf_flash_attention_b200 import _capi, flash_attention as fa  # noqa: E402


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--heads", type=int, default=16)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--seq", type=int, default=8192)
    ap.add_argument("--rule", default="causal")
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    g = torch.Generator(device="cuda").manual_seed(1)
    mk = lambda: (torch.rand((a.batch, a.heads, a.d, a.seq), generator=g, device="cuda") * 4 - 2).half()  # noqa: E731
    Q, K, V, dO = mk(), mk(), mk(), mk()
    cl = lambda x: x.permute(0, 3, 1, 2).contiguous()  # noqa: E731
    Ql, Kl, Vl, dOl = cl(Q), cl(K), cl(V), cl(dO)
    op = fa.causal_1d if a.rule == "causal" else (lambda q, k, v, mode, **kw: fa.full_1d(q, k, v, mode, **kw))

    def run(q, k, v, d_o, backward, **kw):
        def step():
            if backward:
                q.requires_grad_(True), k.requires_grad_(True), v.requires_grad_(True)
                o = op(q, k, v, "none_front", **kw)
                torch.autograd.grad(o, (q, k, v), d_o)
            else:
                with torch.no_grad():
                    op(q, k, v, "none_front", **kw)
        return step

    def adapter(backward):
        def step():
            q, k, v = (fa.from_channel_last(x) for x in (Ql, Kl, Vl))
            if backward:
                q.requires_grad_(True), k.requires_grad_(True), v.requires_grad_(True)
                o = op(q, k, v, "none_front")
                fa.to_channel_last(o)
                for gr in torch.autograd.grad(o, (q, k, v), fa.from_channel_last(dOl)):
                    fa.to_channel_last(gr)
            else:
                with torch.no_grad():
                    fa.to_channel_last(op(q, k, v, "none_front"))
        return step

    nnz = a.seq * (a.seq + 1) // 2 if a.rule == "causal" else a.seq * a.seq
    flops_f = 4.0 * nnz * a.d * a.batch * a.heads
    out = {"shape": f"[{a.batch},{a.heads},{a.d},{a.seq}] {a.rule} fp16"}
    for name, bwd, fl in (("fwd", False, flops_f), ("fwd+bwd", True, 3.5 * flops_f)):
        # the kernels run under the board's power cap: interleave the three arms twice and keep the better time of each,
        # so that the order (a cool GPU for whichever arm runs first) does not decide the comparison
        t_cf = t_cl = t_ad = float("inf")
        for _ in range(2):
            _capi.lib.fa_launch_count(1)
            t_cl = min(t_cl, timed(run(Ql, Kl, Vl, dOl, bwd, layout="channel_last"), a.steps))
            launches = _capi.lib.fa_launch_count(1) / (a.steps + 3)
            t_cf = min(t_cf, timed(run(Q, K, V, dO, bwd), a.steps))
            t_ad = min(t_ad, timed(adapter(bwd), a.steps))
        out[name] = {"channel_first_ms": round(t_cf, 3), "channel_last_direct_ms": round(t_cl, 3),
                     "channel_last_adapter_ms": round(t_ad, 3), "direct_launches_per_step": launches,
                     "direct_tflops": round(fl / t_cl * 1e-9, 1), "channel_first_tflops": round(fl / t_cf * 1e-9, 1),
                     "adapter_tflops": round(fl / t_ad * 1e-9, 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
