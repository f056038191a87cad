"""Self-consistency of the dense oracle (oracle/dense_attention.py): the closed-form backward
(internal_test.cu:413-511) against finite differences of the forward (test_1d.py:69-76), and the
fully-masked-row convention."""
import numpy as np

from oracle import dense_attention as da
from oracle import pattern


def test_backward_matches_finite_differences():
    rng = np.random.default_rng(3)
    Q, K, V, dO = da.random_inputs(rng, np.float64, (2,), 5, 4, (7,), (9,))
    mask = pattern.tests_mask((7,), (9,), "scale_end", "local", 3, 0, True)
    dQ, dK, dV = da.backward(Q, K, V, mask, dO)

    def loss(q, k, v):
        return float((da.forward(q, k, v, mask)[0] * dO).sum())

    eps = 1e-6
    for name, X, dX in (("Q", Q, dQ), ("K", K, dK), ("V", V, dV)):
        for _ in range(12):
            idx = tuple(int(rng.integers(0, n)) for n in X.shape)
            Xp, Xm = X.copy(), X.copy()
            Xp[idx] += eps
            Xm[idx] -= eps
            args_p = {"Q": (Xp, K, V), "K": (Q, Xp, V), "V": (Q, K, Xp)}[name]
            args_m = {"Q": (Xm, K, V), "K": (Q, Xm, V), "V": (Q, K, Xm)}[name]
            fd = (loss(*args_p) - loss(*args_m)) / (2 * eps)
            assert abs(fd - dX[idx]) < 1e-6, (name, idx, fd, dX[idx])


def test_fully_masked_rows_are_zero():
    rng = np.random.default_rng(4)
    # causal, scale_front with more Q than K: K orders 0,4,8.. ; every Q row sees key 0 ... use
    # local causal with stride so that some rows have no key at all
    qs, ks = (10,), (3,)
    mask = pattern.tests_mask(qs, ks, "none_front", "local", 1, 0, False)  # only |dq-dk|<1
    assert not mask[5:].any()
    Q, K, V, dO = da.random_inputs(rng, np.float32, (1,), 4, 4, qs, ks)
    O, l, m = da.forward(Q, K, V, mask)
    assert np.all(O[:, :, 3:] == 0) and np.all(l[:, 3:] == 0) and np.all(np.isneginf(m[:, 3:]))
    dQ, dK, dV = da.backward(Q, K, V, mask, dO)
    assert np.all(dQ[:, :, 3:] == 0)


def test_sentinel_values():
    assert da.sentinel(np.float16) == np.float16(-57152.0)
    assert abs(float(da.sentinel(np.float32)) - (-6.5161e35)) / 6.5161e35 < 1e-4
    assert abs(float(da.sentinel(np.float64)) - (-2.5075e284)) / 2.5075e284 < 1e-3
