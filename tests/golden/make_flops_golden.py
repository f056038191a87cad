"""Generates tests/golden/flops_golden.json from the REFERENCE's own EstimateForwardFlops
(flash_attention.cu:2069-2144) called on the host through oracle/_ref/libref_fa.so
(built unmodified from /root/reference by oracle/ref_build/Makefile). Run in the build container."""
import json
import os
import random
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import refkernel  # noqa: E402

if __name__ == "__main__":
    rng = random.Random(4242)
    out = []
    fixed = [
        (1, 1, "local", "scale_front", 32, 0, 0, 8, 32, 16, (1024,), (2048,)),       # C1
        (0, 1, "causal", "none_front", 1, 0, 0, 256, 128, 128, (8192,), (8192,)),    # C2
        (0, 2, "local", "none_front", 8, 0, 1, 256, 64, 64, (64, 64), (64, 64)),     # C3
        (0, 1, "full", "scale_end", 1, 0, 0, 256, 64, 64, (1024,), (8192,)),         # C4
        (2, 1, "full", "scale_end", 1, 0, 0, 16, 64, 64, (1024,), (8192,)),
    ]
    cases = list(fixed)
    for _ in range(60):
        dims = rng.choice([1, 2])
        if dims == 1:
            qs, ks = (rng.randint(1, 5000),), (rng.randint(1, 5000),)
        else:
            qs = (rng.randint(1, 70), rng.randint(1, 70))
            ks = (rng.randint(1, 70), rng.randint(1, 70))
        cases.append((rng.choice([0, 1, 2]), dims, rng.choice(["full", "causal", "local"]),
                      rng.choice(["none_front", "scale_front", "scale_end"]), rng.randint(1, 40), rng.randint(0, 3),
                      rng.choice([0, 1]), 3, rng.choice([8, 32, 64, 128]), rng.choice([8, 32, 64, 128]), qs, ks))
    for (dt, dims, rule, mode, w, s, c, b, d, vd, qs, ks) in cases:
        f = refkernel.estimate_flops(dt, dims, rule, mode, w, s, c, b, d, vd, qs, ks, 232448)
        out.append({"dtype": dt, "dims": dims, "rule": rule, "sync_mode": mode, "window_size": w,
                    "log2_stride_size": s, "is_causal": c, "batch": b, "d": d, "v_d": vd, "q_shape": list(qs),
                    "k_shape": list(ks), "smem": 232448, "flops": f})
    path = os.path.join(ROOT, "tests", "golden", "flops_golden.json")
    json.dump(out, open(path, "w"), separators=(",", ":"))
    print("wrote", len(out), "cases to", path)
