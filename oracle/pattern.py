"""TEST INFRASTRUCTURE — the attended-index pattern, restated twice on the CPU.

(1) ``tests_formulation`` follows the reference's Python tests:
      locations per sync mode   flash_attention/tests/test_1d.py:9-50, test_2d.py:11-78
      masks                     flash_attention/tests/test_base.py:18-67
(2) ``kernel_formulation`` follows the reference's kernel-side host code:
      sync descriptors          flash_attention/kernel/sync_methods.cc:8-111
      order map                 flash_attention/kernel/sync_methods.h:56-85
      policies (Check)          flash_attention/kernel/flash_attention.h:13-140
      padding rule              flash_attention/kernel/flash_attention.cu:927

Both are pure NumPy integer arithmetic (int64 here; the reference is int32 and
every tested size fits). ``tests/test_oracle_pattern.py`` pins (1) == (2) ==
the golden masks produced by the reference's real code (oracle/_ref/ref_pattern).

Unlike the reference tests, ``window_size`` and ``log2_stride_size`` are
parameters here (the reference derives them from the tensor shape,
test_base.py:54-58, which makes its un-strided local cases effectively "full").
"""
import numpy as np

SYNC_MODES = ("none_front", "scale_front", "scale_end")
RULES = ("full", "causal", "local")


def _as_shape(s):
    s = tuple(int(v) for v in (s if hasattr(s, "__len__") else (s,)))
    assert len(s) in (1, 2) and all(v >= 1 for v in s)
    return s


# --------------------------------------------------------------------------- #
# (1) the tests' formulation
# --------------------------------------------------------------------------- #
def tests_locations(q_shape, k_shape, sync_mode):
    """Returns ((Q_coords[q, D], K_coords[k, D]), (Q_l[q], K_l[k])) with the
    sequence flattened row-major. test_1d.py:9-50 / test_2d.py:11-78."""
    q_shape, k_shape = _as_shape(q_shape), _as_shape(k_shape)
    assert len(q_shape) == len(k_shape)
    D = len(q_shape)
    maxs = [max(a, b) for a, b in zip(q_shape, k_shape)]

    def one(shape):
        grids = np.meshgrid(*[np.arange(n, dtype=np.int64) for n in shape], indexing="ij")
        coords = []
        for g, n, mx in zip(grids, shape, maxs):
            if sync_mode == "none_front":
                c = g
            elif sync_mode == "scale_front":
                c = g * (mx // n)
            elif sync_mode == "scale_end":
                c = (g + 1) * (mx // n) - 1
            else:
                raise ValueError(f"Unsupported sync_mode: {sync_mode}")
            coords.append(c.reshape(-1))
        coords = np.stack(coords, axis=-1)  # [n, D] in TF axis order (y, x)
        if D == 1:
            l = coords[:, 0]
        else:
            # l = y * max_width + x     (test_2d.py:22,26,46,56,75)
            l = coords[:, 0] * maxs[1] + coords[:, 1]
        return coords, l

    qc, ql = one(q_shape)
    kc, kl = one(k_shape)
    return (qc, kc), (ql, kl)


def tests_mask(q_shape, k_shape, sync_mode, rule, window_size=1, log2_stride_size=0, is_causal=False):
    """bool [q, k]. test_base.py:33-67."""
    (qc, kc), (ql, kl) = tests_locations(q_shape, k_shape, sync_mode)
    idx_diff = ql[:, None] - kl[None, :]
    if rule == "full":
        return np.ones_like(idx_diff, dtype=bool)
    if rule == "causal":
        return idx_diff >= 0
    if rule == "local":
        diff = np.abs(qc[:, None, :] - kc[None, :, :])
        stride = 1 << int(log2_stride_size)
        ok = np.all((diff % stride == 0) & (diff // stride < int(window_size)), axis=-1)
        if is_causal:
            ok &= idx_diff >= 0
        return ok
    raise ValueError(rule)


# --------------------------------------------------------------------------- #
# (2) the kernel-side formulation
# --------------------------------------------------------------------------- #
def _log2i(n):  # cute_ext/algorithms.h:8-16 (host branch): floor(log2(n))
    return 0 if n < 2 else 1 + _log2i(n >> 1)


def sync_descriptors(q_shape, k_shape, sync_mode):
    """sync_methods.cc:8-111. Returns dict with lists indexed innermost-first
    (index 0 = TF's last axis), like the reference's SequenceDescriptorPack."""
    q_shape, k_shape = _as_shape(q_shape), _as_shape(k_shape)
    ref, qd, kd = [], {"shape": [], "stride": [], "offset": []}, {"shape": [], "stride": [], "offset": []}
    for dim in range(len(q_shape) - 1, -1, -1):
        Q_dim, K_dim = q_shape[dim], k_shape[dim]
        max_dim = max(Q_dim, K_dim)
        ref_dim = 1 << _log2i(max_dim)
        if ref_dim < max_dim:
            ref_dim <<= 1
        ref.append(ref_dim)
        for desc, n in ((qd, Q_dim), (kd, K_dim)):
            desc["shape"].append(n)
            if sync_mode == "none_front":
                st, off = 1, 0
            elif sync_mode == "scale_front":
                st, off = max_dim // n, 0
            elif sync_mode == "scale_end":
                st = max_dim // n
                off = st - 1
            else:
                raise ValueError(f"Unsupported sync_mode: {sync_mode}")
            desc["stride"].append(st)
            desc["offset"].append(off)
    return {"reference_shape": ref, "Q": qd, "K": kd}


def order_map(ref, desc, n_total=None):
    """sync_methods.h:56-85: linear (row-major, TF) index -> order in the pow-2
    reference grid. order = sum_i (off_i + c_i*stride_i) * prod_{e<i} ref_e."""
    shape = desc["shape"]
    n = int(np.prod(shape)) if n_total is None else n_total
    idx = np.arange(n, dtype=np.int64)
    order = np.zeros(n, dtype=np.int64)
    mult = 1
    for i in range(len(shape)):
        c = idx % shape[i] if i < len(shape) - 1 else idx
        idx = idx // shape[i]
        order += (desc["offset"][i] + c * desc["stride"][i]) * mult
        mult *= ref[i]
    return order


def map_to_coords(order, ref):
    """flash_attention.h:13-25 (shift / mask on the pow-2 grid); innermost first."""
    out, shift = [], 0
    for s in ref:
        out.append((order >> shift) & (s - 1))
        shift += _log2i(s)
    return out


def kernel_mask(q_shape, k_shape, sync_mode, rule, window_size=1, log2_stride_size=0, is_causal=False):
    """bool [q, k] from the kernel-side formulas (flash_attention.h:45-140)."""
    pack = sync_descriptors(q_shape, k_shape, sync_mode)
    ref = pack["reference_shape"]
    qo = order_map(ref, pack["Q"])
    ko = order_map(ref, pack["K"])
    Qo, Ko = qo[:, None], ko[None, :]
    if rule == "full":
        return np.ones((qo.size, ko.size), dtype=bool)
    if rule == "causal":
        return Qo >= Ko
    if rule == "local":
        ok = np.ones((qo.size, ko.size), dtype=bool)
        if is_causal:
            ok &= ~(Qo < Ko)
        rem = (1 << int(log2_stride_size)) - 1
        for qc, kc in zip(map_to_coords(qo, ref), map_to_coords(ko, ref)):
            diff = np.abs(qc[:, None] - kc[None, :])
            ok &= ((diff & rem) == 0) & ((diff >> int(log2_stride_size)) < int(window_size))
        return ok
    raise ValueError(rule)


def kernel_orders(q_shape, k_shape, sync_mode):
    pack = sync_descriptors(q_shape, k_shape, sync_mode)
    ref = pack["reference_shape"]
    return ref, order_map(ref, pack["Q"]), order_map(ref, pack["K"])


# --------------------------------------------------------------------------- #
# nnz (attended pairs) — the unit of the unmasked-FLOP metric
# --------------------------------------------------------------------------- #
def nnz(q_shape, k_shape, sync_mode, rule, window_size=1, log2_stride_size=0, is_causal=False, chunk=2048):
    """Exact number of attended (q,k) pairs, computed in row chunks so that it
    also works at the BASELINE.json sizes (8192^2)."""
    q_shape, k_shape = _as_shape(q_shape), _as_shape(k_shape)
    (qc, kc), (ql, kl) = tests_locations(q_shape, k_shape, sync_mode)
    total = 0
    stride = 1 << int(log2_stride_size)
    for s in range(0, ql.size, chunk):
        e = min(ql.size, s + chunk)
        idx_diff = ql[s:e, None] - kl[None, :]
        if rule == "full":
            total += (e - s) * kl.size
            continue
        if rule == "causal":
            total += int(np.count_nonzero(idx_diff >= 0))
            continue
        ok = np.ones(idx_diff.shape, dtype=bool)
        for d in range(qc.shape[1]):
            diff = np.abs(qc[s:e, None, d] - kc[None, :, d])
            ok &= (diff % stride == 0) & (diff // stride < int(window_size))
        if is_causal:
            ok &= idx_diff >= 0
        total += int(np.count_nonzero(ok))
    return total
