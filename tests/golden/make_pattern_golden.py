"""Generates tests/golden/pattern_golden.json by running the REFERENCE's own host code
(sync_methods.cc + the policies of flash_attention.h, compiled unmodified by
oracle/ref_build/Makefile into oracle/_ref/ref_pattern) on a fixed list of cases.

Run in the build container (needs /root/reference):   python tests/golden/make_pattern_golden.py
The masks are stored row-wise as hex strings of the packed bits.
"""
import json
import os
import random
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_pattern")


def run_case(dims, rule, mode, w, s, c, qs, ks):
    args = [BIN, str(dims), rule, mode, str(w), str(s), str(int(c))] + [str(v) for v in qs] + [str(v) for v in ks]
    out = subprocess.run(args, check=True, capture_output=True, text=True).stdout.splitlines()
    ref = [int(v) for v in out[0].split()[1:]]
    qo = [int(v) for v in out[1].split()]
    ko = [int(v) for v in out[2].split()]
    rows = out[3:]
    mask = np.array([[ch == "1" for ch in row] for row in rows], dtype=bool).reshape(len(qo), len(ko))
    return {"dims": dims, "rule": rule, "sync_mode": mode, "window_size": w, "log2_stride_size": s,
            "is_causal": int(c), "q_shape": list(qs), "k_shape": list(ks), "ref_shape": ref,
            "q_order": qo, "k_order": ko, "nnz": int(mask.sum()),
            "mask_hex": np.packbits(mask, axis=1).tobytes().hex()}


def cases():
    rng = random.Random(20261018)
    out = []
    # the docstring alignment tables (flash_attention/flash_attention.py:30-69)
    for mode in ("none_front", "scale_front", "scale_end"):
        out.append((1, "causal", mode, 1, 0, 0, (6,), (3,)))
        out.append((1, "causal", mode, 1, 0, 0, (3,), (6,)))
        out.append((2, "causal", mode, 1, 0, 0, (4, 4), (2, 2)))
        out.append((2, "causal", mode, 1, 0, 0, (2, 2), (4, 4)))
    # the README example (README.md:66-71), shrunk 8x in length, same ratios/window rule
    out.append((1, "local", "scale_front", 4, 0, 0, (128,), (256,)))
    # every rule x sync mode, 1-D and 2-D, ragged / non-dividing sizes, strides, causal flags
    for dims in (1, 2):
        for rule in ("full", "causal", "local"):
            for mode in ("none_front", "scale_front", "scale_end"):
                for _ in range(6 if rule == "local" else 2):
                    if dims == 1:
                        qs, ks = (rng.randint(1, 97),), (rng.randint(1, 97),)
                    else:
                        qs = (rng.randint(1, 11), rng.randint(1, 11))
                        ks = (rng.randint(1, 11), rng.randint(1, 11))
                    w, s, c = rng.randint(1, 7), rng.randint(0, 3), rng.randint(0, 1)
                    out.append((dims, rule, mode, w, s, c, qs, ks))
    # a C3-like 2-D causal window and a C4-like cross attention, small
    out.append((2, "local", "none_front", 3, 0, 1, (12, 12), (12, 12)))
    out.append((1, "full", "scale_end", 1, 0, 0, (16,), (128,)))
    out.append((1, "causal", "scale_end", 1, 0, 0, (16,), (128,)))
    return out


if __name__ == "__main__":
    golden = [run_case(*c) for c in cases()]
    path = os.path.join(HERE, "pattern_golden.json")
    with open(path, "w") as f:
        json.dump(golden, f, separators=(",", ":"))
    print(f"wrote {len(golden)} cases to {path} ({os.path.getsize(path)} bytes)")
