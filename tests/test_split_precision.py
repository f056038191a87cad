"""Host-side check of the split-precision arithmetic the fp32 tensor-core kernels rely on (no GPU needed):
* forward (fa_fwd_f32_sm100.cu): x = hi + lo with hi = x truncated to TF32 (10 explicit mantissa bits);
  a.b ~ hi.hi + hi.lo + lo.hi                      -> relative error of a dot product ~2^-21;
* backward (fa_bwd_f32_sm100.cu): x = x0 + x1 + x2 with bf16 pieces (8 bits each, round to nearest);
  a.b ~ a0b0 + (a0b1 + a1b0) + (a0b2 + a1b1 + a2b0) -> relative error ~2^-24 (fp32 level).
The emulation rounds exactly like the device code (bit masks / round-to-nearest-even on the fp32 bit pattern) and
accumulates in float64, so what is measured is the representation error of the split alone."""
import numpy as np


def to_bf16(x):
    """float32 -> nearest bf16 (round to nearest even), returned as float32."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


def split_bf16x3(x):
    x = np.asarray(x, dtype=np.float32)
    p0 = to_bf16(x)
    r1 = (x - p0).astype(np.float32)
    p1 = to_bf16(r1)
    p2 = to_bf16((r1 - p1).astype(np.float32))
    return p0, p1, p2


def split_tf32(x):
    x = np.asarray(x, dtype=np.float32)
    hi = (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = (x - hi).astype(np.float32)
    lo_tf32 = (lo.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)   # the MMA reads lo as TF32 too
    return hi, lo_tf32


def test_bf16x3_pieces_reconstruct_fp32_exactly_enough():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(100000) * 10 ** rng.uniform(-6, 6, 100000)).astype(np.float32)
    p0, p1, p2 = split_bf16x3(x)
    rec = p0.astype(np.float64) + p1.astype(np.float64) + p2.astype(np.float64)
    rel = np.abs(rec - x.astype(np.float64)) / np.abs(x.astype(np.float64))
    assert rel.max() <= 2.0 ** -24        # three 8-bit pieces carry the 24-bit significand
    assert np.all(np.abs(p1) <= np.abs(p0) * 2.0 ** -7) and np.all(np.abs(p2) <= np.abs(p0) * 2.0 ** -15)


def test_six_piece_products_reach_fp32_accuracy_and_three_tf32_products_do_not():
    rng = np.random.default_rng(1)
    n, k = 2000, 64
    a = rng.uniform(-2, 2, (n, k)).astype(np.float32)
    b = rng.uniform(-2, 2, (n, k)).astype(np.float32)
    exact = np.einsum("nk,nk->n", a.astype(np.float64), b.astype(np.float64))
    norm = np.einsum("nk,nk->n", np.abs(a).astype(np.float64), np.abs(b).astype(np.float64))
    A, B = split_bf16x3(a), split_bf16x3(b)
    pairs = [(0, 2), (1, 1), (2, 0), (0, 1), (1, 0), (0, 0)]          # the order the kernels issue them in
    six = sum(np.einsum("nk,nk->n", A[i].astype(np.float64), B[j].astype(np.float64)) for i, j in pairs)
    ah, al = split_tf32(a)
    bh, bl = split_tf32(b)
    three = sum(np.einsum("nk,nk->n", x.astype(np.float64), y.astype(np.float64)) for x, y in ((ah, bh), (ah, bl), (al, bh)))
    err6 = np.abs(six - exact) / norm
    err3 = np.abs(three - exact) / norm
    assert err6.max() <= 2.0 ** -22                      # dropped terms a1b2, a2b1, a2b2: <= 3 * 2^-24 per product
    assert err3.max() <= 2.0 ** -19 and err3.max() > err6.max()       # 3xTF32: lo.lo dropped, lo truncated to TF32
    print(f"bf16x3 six products: {err6.max():.2e}   3xTF32: {err3.max():.2e}   (relative to sum |a||b|)")
