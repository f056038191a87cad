"""Generates tests/golden/refkernel_golden.npz ON A B200: runs the UNMODIFIED reference CUDA kernel
(oracle/_ref/libref_fa.so, built from /root/reference by oracle/ref_build/Makefile; driver
oracle/refkernel.py) on a fixed list of seeded cases and stores its outputs (O, l, m, dQ, dK, dV).
tests/test_oracle_vs_refkernel.py then checks the NumPy oracle against these on the CPU, which pins
the oracle's numerics against the real reference implementation.

  gpurun -- 'python tests/golden/make_refkernel_golden.py'   ->  gpurun_out/refkernel_golden.npz
  cp gpurun_out/refkernel_golden.npz tests/golden/
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, ROOT)

CASES = [
    # name, dtype, dims, rule, sync, w, s, c, batch, d, v_d, q_shape, k_shape
    ("full1d_f64", "float64", 1, "full", "none_front", 1, 0, 0, (2,), 16, 12, (40,), (56,)),
    ("causal1d_f64_scale_end", "float64", 1, "causal", "scale_end", 1, 0, 0, (2,), 16, 16, (24,), (96,)),
    ("local1d_f64_stride", "float64", 1, "local", "scale_front", 3, 1, 1, (2,), 8, 8, (64,), (48,)),
    ("local2d_f64_causal", "float64", 2, "local", "none_front", 3, 0, 1, (2,), 8, 8, (6, 9), (6, 9)),
    ("causal2d_f64_scale_front", "float64", 2, "causal", "scale_front", 1, 0, 0, (1,), 8, 10, (4, 6), (8, 6)),
    ("readme_like_f32", "float32", 1, "local", "scale_front", 8, 0, 0, (2,), 32, 16, (64,), (128,)),
    ("causal1d_f32", "float32", 1, "causal", "none_front", 1, 0, 0, (2, 2), 24, 24, (100,), (100,)),
    ("full2d_f32_scale_end", "float32", 2, "full", "scale_end", 1, 0, 0, (2,), 16, 16, (5, 7), (10, 7)),
    ("causal1d_f16", "float16", 1, "causal", "none_front", 1, 0, 0, (2,), 32, 32, (96,), (96,)),
    ("local1d_f16_masked_rows", "float16", 1, "local", "none_front", 2, 0, 0, (2,), 16, 16, (80,), (24,)),
]


def inputs(case):
    name, dtype, dims, rule, sync, w, s, c, batch, d, v_d, qs, ks = case
    from oracle import dense_attention as da
    seed = sum(ord(ch) for ch in name)
    rng = np.random.default_rng(seed)
    return da.random_inputs(rng, np.dtype(dtype), batch, d, v_d, qs, ks)


if __name__ == "__main__":
    import torch
    from oracle import refkernel
    out = {}
    for case in CASES:
        name, dtype, dims, rule, sync, w, s, c, batch, d, v_d, qs, ks = case
        Q, K, V, dO = inputs(case)
        tq, tk, tv, tdo = (torch.from_numpy(x).cuda() for x in (Q, K, V, dO))
        O, l, m = refkernel.forward(tq, tk, tv, dims, rule, sync, w, s, bool(c))
        dQ, dK, dV = refkernel.backward(tq, tk, tv, O, l, m, tdo, dims, rule, sync, w, s, bool(c))
        torch.cuda.synchronize()
        for key, t in (("O", O), ("l", l), ("m", m), ("dQ", dQ), ("dK", dK), ("dV", dV)):
            out[f"{name}/{key}"] = t.cpu().numpy()
        print(name, "ok", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "refkernel_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
