"""Pins the oracle's NUMERICS against the real reference: tests/golden/refkernel_golden.npz holds the
outputs of the unmodified reference CUDA kernel (run on a B200 by
tests/golden/make_refkernel_golden.py through oracle/_ref/libref_fa.so). The NumPy oracle must agree
with them — tightly in fp64/fp32, and within the reference tests' own tolerance in fp16 (the
reference accumulates in half precision: flash_attention.cu:284-286,335)."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import dense_attention as da

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_refkernel_golden", os.path.join(HERE, "golden", "make_refkernel_golden.py"))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)
GOLD = np.load(os.path.join(HERE, "golden", "refkernel_golden.npz"))


@pytest.mark.parametrize("case", gen.CASES, ids=lambda c: c[0])
def test_oracle_matches_reference_kernel(case):
    name, dtype, dims, rule, sync, w, s, c, batch, d, v_d, qs, ks = case
    Q, K, V, dO = gen.inputs(case)
    ref = da.attention(Q, K, V, dims, rule, sync, w, s, bool(c), dO=dO)
    nq, nk = int(np.prod(qs)), int(np.prod(ks))
    # fp64 / fp32: far tighter than the reference's own 1e-6*N (test_base.py:203-226);
    # fp16: the reference's own bound 1e-3*N
    tol = {"float64": (1e-11, 1e-11, 1e-11), "float32": (2e-5, 2e-4, 2e-4),
           "float16": (1e-3 * nk, 1e-3 * nq, 1e-3 * nq)}[dtype]
    got_O = GOLD[f"{name}/O"].astype(np.float64)
    assert np.max(np.abs(got_O - ref["O"])) <= tol[0]
    for key, t in (("dQ", tol[0] if dtype != "float32" else tol[1]), ("dK", tol[1]), ("dV", tol[2])):
        assert np.max(np.abs(GOLD[f"{name}/{key}"].astype(np.float64) - ref[key])) <= t, key
    # l, m: the pair reproduces the log-sum-exp; fully masked rows keep l = 0 and the 0xFA sentinel
    l, m = GOLD[f"{name}/l"].astype(np.float64), GOLD[f"{name}/m"]
    empty = ~np.isfinite(ref["m"])
    assert np.all(l[empty] == 0)
    assert np.all(m[empty].view(np.uint8) == 0xFA)
    assert np.all(got_O[np.broadcast_to(np.expand_dims(empty, -dims - 1), got_O.shape)] == 0)
    live = ~empty
    lse = m.astype(np.float64)[live] + np.log(l[live])
    lse_ref = ref["m"][live] + np.log(ref["l"][live])
    assert np.max(np.abs(lse - lse_ref)) <= {"float64": 1e-11, "float32": 1e-4, "float16": 0.1}[dtype]
