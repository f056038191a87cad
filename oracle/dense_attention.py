"""TEST INFRASTRUCTURE — dense masked-softmax attention, the parity oracle.

Restates, in NumPy float64 on the dtype-rounded inputs:
  forward   flash_attention/tests/test_1d.py:69-76 (1-D), test_2d.py:97-109 (2-D):
              logit = einsum(Q,K)/sqrt(d); where(mask, logit, dtype.min);
              softmax; where(mask, p, 0); einsum(p, V)
            (rows with no attended key therefore give O = 0)
  backward  flash_attention/kernel/internal_test.cu:413-511 closed form:
              dV = P^T dO ; D = rowsum(dO * O) ; dS = P * (dO V^T - D) / sqrt(d) ;
              dQ = dS K ; dK = dS^T Q
  l, m      flash_attention/kernel/flash_attention.cu:915-1035: m = row max of the
            scaled, masked logits; l = sum exp(logit - m); fully masked rows keep
            l = 0 and m = the 0xFA.. byte-pattern sentinel
            (flash_attention_forward.cc:360-365, type_util.h:43-45).

Tensors are channel-first: Q [B, d, q], K [B, d, k], V [B, v_d, k] with the
sequence axes already flattened row-major and every batch axis flattened into B.
"""
import numpy as np

from . import pattern

SENTINEL_BITS = {np.dtype(np.float16): np.uint16(0xFAFA),
                 np.dtype(np.float32): np.uint32(0xFAFAFAFA),
                 np.dtype(np.float64): np.uint64(0xFAFAFAFAFAFAFAFA)}


def sentinel(dtype):
    """type_util.h:43-45 GetNegInfApprox(): every byte 0xFA."""
    dtype = np.dtype(dtype)
    return np.array([SENTINEL_BITS[dtype]]).view(dtype)[0]


def l_dtype(dtype):
    """flash_attention.h:181-185: l is float for half, T otherwise."""
    return np.dtype(np.float32) if np.dtype(dtype) == np.float16 else np.dtype(dtype)


def flatten_inputs(Q, K, V, seq_dims):
    """[batch..., c, seq...] -> [B, c, n] (flash_attention_forward.cc:97-140)."""
    def f(x):
        x = np.asarray(x)
        seq = x.shape[x.ndim - seq_dims:]
        c = x.shape[x.ndim - seq_dims - 1]
        batch = x.shape[: x.ndim - seq_dims - 1]
        return x.reshape((int(np.prod(batch, dtype=np.int64)), c, int(np.prod(seq)))), batch, seq
    Qf, qb, qs = f(Q)
    Kf, kb, ks = f(K)
    Vf, vb, vs = f(V)
    return Qf, Kf, Vf, qb, qs, ks


def forward(Q, K, V, mask, return_p=False):
    """Q [B,d,q], K [B,d,k], V [B,vd,k], mask bool [q,k] -> O [B,vd,q] float64,
    l [B,q], m [B,q] (float64; m = -inf and l = 0 on fully masked rows)."""
    Q64, K64, V64 = (np.asarray(x, dtype=np.float64) for x in (Q, K, V))
    d = Q64.shape[1]
    logit = np.einsum("bcq,bck->bqk", Q64, K64) / np.sqrt(np.float64(d))
    neg = np.where(mask[None], logit, -np.inf)
    m = neg.max(axis=-1)
    m_safe = np.where(np.isfinite(m), m, 0.0)
    e = np.where(mask[None], np.exp(logit - m_safe[..., None]), 0.0)
    l = e.sum(axis=-1)
    p = e / np.where(l > 0, l, 1.0)[..., None]
    O = np.einsum("bqk,bck->bcq", p, V64)
    if return_p:
        return O, l, m, p
    return O, l, m


def backward(Q, K, V, mask, dO):
    """Returns dQ [B,d,q], dK [B,d,k], dV [B,vd,k] in float64."""
    Q64, K64, V64, dO64 = (np.asarray(x, dtype=np.float64) for x in (Q, K, V, dO))
    d = Q64.shape[1]
    O, l, m, p = forward(Q64, K64, V64, mask, return_p=True)
    dV = np.einsum("bqk,bcq->bck", p, dO64)
    dP = np.einsum("bcq,bck->bqk", dO64, V64)
    D = np.einsum("bcq,bcq->bq", dO64, O)
    dS = p * (dP - D[..., None]) / np.sqrt(np.float64(d))
    dQ = np.einsum("bqk,bck->bcq", dS, K64)
    dK = np.einsum("bqk,bcq->bck", dS, Q64)
    return dQ, dK, dV


def attention(Q, K, V, seq_dims, rule, sync_mode, window_size=1, log2_stride_size=0, is_causal=False, dO=None):
    """Convenience wrapper on un-flattened channel-first tensors; returns arrays
    shaped like the reference op outputs (float64)."""
    Qf, Kf, Vf, batch, qs, ks = flatten_inputs(Q, K, V, seq_dims)
    mask = pattern.tests_mask(qs, ks, sync_mode, rule, window_size, log2_stride_size, is_causal)
    O, l, m = forward(Qf, Kf, Vf, mask)
    out = {"O": O.reshape(batch + (Vf.shape[1],) + tuple(qs)),
           "l": l.reshape(batch + tuple(qs)), "m": m.reshape(batch + tuple(qs)), "mask": mask}
    if dO is not None:
        dOf = np.asarray(dO).reshape(Qf.shape[0], Vf.shape[1], Qf.shape[2])
        dQ, dK, dV = backward(Qf, Kf, Vf, mask, dOf)
        out.update(dQ=dQ.reshape(np.shape(Q)), dK=dK.reshape(np.shape(K)), dV=dV.reshape(np.shape(V)))
    return out


def random_inputs(rng, dtype, batch, d, v_d, q_shape, k_shape):
    """U(-2,2) like the reference tests (test_base.py:170-173)."""
    def u(shape):
        return rng.uniform(-2.0, 2.0, size=shape).astype(dtype)
    q_shape, k_shape = tuple(q_shape), tuple(k_shape)
    batch = tuple(batch)
    return (u(batch + (d,) + q_shape), u(batch + (d,) + k_shape),
            u(batch + (v_d,) + k_shape), u(batch + (v_d,) + q_shape))


def forward_backward_chunked(Q, K, V, dO, q_shape, k_shape, sync_mode, rule, window_size=1,
                             log2_stride_size=0, is_causal=False, rows=512):
    """Same math as forward()/backward() for ONE batch element (Q [d,q], K [d,k], V [vd,k],
    dO [vd,q] or None) but walking the query rows in chunks so that the BASELINE.json sizes
    (8192 x 8192) fit in memory. Returns dict of float64 arrays."""
    Q64, K64, V64 = (np.asarray(x, dtype=np.float64) for x in (Q, K, V))
    d, nq = Q64.shape
    nk = K64.shape[1]
    (qc, kc), (ql, kl) = pattern.tests_locations(q_shape, k_shape, sync_mode)
    stride = 1 << int(log2_stride_size)
    O = np.zeros((V64.shape[0], nq))
    l = np.zeros(nq)
    m = np.full(nq, -np.inf)
    out = {"O": O, "l": l, "m": m}
    if dO is not None:
        dO64 = np.asarray(dO, dtype=np.float64)
        dQ, dK, dV = np.zeros_like(Q64), np.zeros_like(K64), np.zeros_like(V64)
        out.update(dQ=dQ, dK=dK, dV=dV)
    for s in range(0, nq, rows):
        e = min(nq, s + rows)
        idx_diff = ql[s:e, None] - kl[None, :]
        if rule == "full":
            mask = np.ones(idx_diff.shape, dtype=bool)
        elif rule == "causal":
            mask = idx_diff >= 0
        else:
            mask = np.ones(idx_diff.shape, dtype=bool)
            for dd in range(qc.shape[1]):
                diff = np.abs(qc[s:e, None, dd] - kc[None, :, dd])
                mask &= (diff % stride == 0) & (diff // stride < int(window_size))
            if is_causal:
                mask &= idx_diff >= 0
        logit = (Q64[:, s:e].T @ K64) / np.sqrt(np.float64(d))
        neg = np.where(mask, logit, -np.inf)
        mm = neg.max(axis=-1)
        ms = np.where(np.isfinite(mm), mm, 0.0)
        ex = np.where(mask, np.exp(logit - ms[:, None]), 0.0)
        ll = ex.sum(axis=-1)
        p = ex / np.where(ll > 0, ll, 1.0)[:, None]
        O[:, s:e] = (p @ V64.T).T
        l[s:e], m[s:e] = ll, mm
        if dO is not None:
            dOc = dO64[:, s:e]                      # [vd, r]
            dV += (p.T @ dOc.T).T                   # [vd, k]
            dP = dOc.T @ V64                        # [r, k]
            D = (dOc * O[:, s:e]).sum(axis=0)       # [r]
            dS = p * (dP - D[:, None]) / np.sqrt(np.float64(d))
            dQ[:, s:e] = (dS @ K64.T).T
            dK += (dS.T @ Q64[:, s:e].T).T
    return out
