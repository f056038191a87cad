"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

# BASELINE.json north_star tolerances: max-abs on O and on the gradients
TOL = {np.dtype(np.float16): 2e-3, np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


def load_pattern_golden():
    with open(os.path.join(GOLDEN, "pattern_golden.json")) as f:
        cases = json.load(f)
    for c in cases:
        q = int(np.prod(c["q_shape"]))
        k = int(np.prod(c["k_shape"]))
        packed = np.frombuffer(bytes.fromhex(c["mask_hex"]), dtype=np.uint8).reshape(q, -1)
        c["mask"] = np.unpackbits(packed, axis=1)[:, :k].astype(bool)
    return cases


def case_id(c):
    return (f"{c['dims']}d-{c['rule']}-{c['sync_mode']}-w{c['window_size']}s{c['log2_stride_size']}"
            f"c{c['is_causal']}-q{'x'.join(map(str, c['q_shape']))}-k{'x'.join(map(str, c['k_shape']))}")


def scaled_err(got, ref):
    """max |got-ref| / max(1,|ref|): the absolute tolerance of BASELINE.json applied relative to
    magnitude once |ref| exceeds 1 (an fp16 output cannot carry 2e-3 absolute beyond |x|>=4)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref)))) if ref.size else 0.0


def max_abs_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref))) if ref.size else 0.0
