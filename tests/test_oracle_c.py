"""The plain-C restatement (oracle/dense_attention.c, built by oracle/Makefile) against the pattern golden
produced by the reference's own host code and against the NumPy oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import dense_attention as da
from oracle import pattern
from tests.helpers import case_id, load_pattern_golden

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SO = os.path.join(ROOT, "oracle", "_build", "liboracle_c.so")


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return C.CDLL(SO)


def _shape(s):
    return (C.c_int32 * 2)(*(list(s) + [1])[:2])


def test_c_pattern_matches_reference_golden(lib):
    rules = {"full": 0, "causal": 1, "local": 2}
    syncs = {"none_front": 0, "scale_front": 1, "scale_end": 2}
    for c in load_pattern_golden():
        q, k = int(np.prod(c["q_shape"])), int(np.prod(c["k_shape"]))
        mask = np.zeros((q, k), dtype=np.uint8)
        rc = lib.oc_pattern(c["dims"], rules[c["rule"]], syncs[c["sync_mode"]], c["window_size"],
                            c["log2_stride_size"], c["is_causal"], _shape(c["q_shape"]), _shape(c["k_shape"]),
                            mask.ctypes.data_as(C.c_void_p))
        assert rc == 0
        assert np.array_equal(mask.astype(bool), c["mask"]), case_id(c)


@pytest.mark.parametrize("rule,sync,w,s,c,qs,ks", [("causal", "scale_end", 1, 0, 0, (9,), (23,)),
                                                    ("local", "none_front", 3, 1, 1, (17,), (12,)),
                                                    ("local", "scale_front", 2, 0, 0, (3, 5), (6, 5)),
                                                    ("full", "none_front", 1, 0, 0, (4, 3), (2, 7))])
def test_c_dense_matches_numpy_oracle(lib, rule, sync, w, s, c, qs, ks):
    rng = np.random.default_rng(0)
    B, d, vd = 2, 5, 4
    Q, K, V, dO = da.random_inputs(rng, np.float64, (B,), d, vd, qs, ks)
    q, k = int(np.prod(qs)), int(np.prod(ks))
    Qf, Kf, Vf, dOf = (np.ascontiguousarray(x.reshape(B, -1, n)) for x, n in ((Q, q), (K, k), (V, k), (dO, q)))
    mask = pattern.tests_mask(qs, ks, sync, rule, w, s, bool(c))
    m8 = np.ascontiguousarray(mask.astype(np.uint8))
    O, l, m = np.zeros((B, vd, q)), np.zeros((B, q)), np.zeros((B, q))
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib.oc_forward.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64] + [C.c_void_p] * 7
    lib.oc_backward.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64] + [C.c_void_p] * 8
    assert lib.oc_forward(B, d, vd, q, k, p(Qf), p(Kf), p(Vf), p(m8), p(O), p(l), p(m)) == 0
    rO, rl, rm = da.forward(Qf, Kf, Vf, mask)
    assert np.max(np.abs(O - rO)) < 1e-12 and np.max(np.abs(l - rl)) < 1e-12
    assert np.array_equal(np.isfinite(m), np.isfinite(rm)) and np.max(np.abs(m[np.isfinite(rm)] - rm[np.isfinite(rm)])) < 1e-12
    dQ, dK, dV = np.zeros_like(Qf), np.zeros_like(Kf), np.zeros_like(Vf)
    assert lib.oc_backward(B, d, vd, q, k, p(Qf), p(Kf), p(Vf), p(dOf), p(m8), p(dQ), p(dK), p(dV)) == 0
    rdQ, rdK, rdV = da.backward(Qf, Kf, Vf, mask, dOf)
    for got, ref in ((dQ, rdQ), (dK, rdK), (dV, rdV)):
        assert np.max(np.abs(got - ref)) < 1e-11
