// fa_rules.h — the reference's sync modes and masking rules as one small POD that
// host code and every kernel evaluate with the SAME inline functions.
//
// Behaviour restated from the reference (nothing copied; CuTe is not used):
//   sync modes  -> per-dim (stride, offset) + power-of-two reference grid
//                  flash_attention/kernel/sync_methods.cc:8-111
//   order map   -> order = sum_i (off_i + c_i*stride_i) * prod_{e<i} ref_e
//                  flash_attention/kernel/sync_methods.h:56-85
//   rules       -> Full / Causal / Local Check()   flash_attention/kernel/flash_attention.h:45-140
//   padding     -> q < size && k < size            flash_attention/kernel/flash_attention.cu:927
// The tile classifier (skip / partial / full) is new: it is conservative by
// construction (interval arithmetic on coordinates), unlike the reference's
// LocalAttentionPolicy::IsSkipped bounding box (flash_attention.h:98-115), which can
// under-cover when a Q tile wraps a 2-D row.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FA_HD __host__ __device__ __forceinline__
#else
#define FA_HD inline
#endif

enum { FA_TILE_SKIP = 0, FA_TILE_PARTIAL = 1, FA_TILE_FULL = 2 };

struct FaSeqMap {
  int32_t n0;       // innermost extent (TF last axis); 1-D: local length of this call
  int32_t n1;       // outer extent (1 for 1-D)
  int32_t stride0, stride1;
  int32_t off0, off1;
  int32_t base0;    // ring: global index of local element 0 (1-D only)
  int32_t total;    // number of local elements = n0*n1
};

struct FaRule {
  int32_t dims;         // 1 or 2
  int32_t rule;         // 0 full, 1 causal, 2 local
  int32_t window;       // local
  int32_t log2_stride;  // local
  int32_t causal;       // causal rule, or local with is_causal
  int32_t ref_log2_0;   // log2 of the innermost reference extent
  int32_t ref0, ref1;   // reference grid extents (powers of two)
  FaSeqMap q, k;
};

struct FaPos {
  int32_t c0, c1;   // coordinates in the reference grid (c1 = 0 for 1-D)
  int32_t order;    // (c1 << ref_log2_0) + c0
};

FA_HD int32_t fa_ceil_log2(int32_t n) {
  int32_t l = 0;
  while ((int64_t(1) << l) < n) ++l;
  return l;
}

// Builds the rule POD. Returns 0, or a negative FA_EINVAL_* (values from fa_b200.h).
// q_shape/k_shape are in TF order (outer, inner).
inline int fa_make_rule(int32_t seq_dims, int32_t rule, int32_t window, int32_t log2_stride,
                        int32_t is_causal, int32_t sync_mode, const int32_t* q_shape,
                        const int32_t* k_shape, int32_t q_base, int32_t k_base,
                        int32_t q_full, int32_t k_full, FaRule* out) {
  if (seq_dims != 1 && seq_dims != 2) return -3;
  if (rule < 0 || rule > 2) return -4;
  if (sync_mode < 0 || sync_mode > 2) return -5;
  if (rule == 2) {
    if (window < 1) return -6;
    if (log2_stride < 0 || log2_stride >= 31) return -7;
    int64_t sw = int64_t(window) << log2_stride;
    if (sw > 0x7fffffffLL) return -7;
  }
  FaRule r;
  r.dims = seq_dims;
  r.rule = rule;
  r.window = rule == 2 ? window : 1;
  r.log2_stride = rule == 2 ? log2_stride : 0;
  r.causal = (rule == 1) || (rule == 2 && is_causal);
  // innermost-first, like the reference's SequenceDescriptorPack
  int32_t qn[2] = {1, 1}, kn[2] = {1, 1}, qfull[2] = {1, 1}, kfull[2] = {1, 1};
  for (int i = 0; i < seq_dims; ++i) {
    qn[i] = q_shape[seq_dims - 1 - i];
    kn[i] = k_shape[seq_dims - 1 - i];
    if (qn[i] < 1 || kn[i] < 1) return -8;
    qfull[i] = qn[i];
    kfull[i] = kn[i];
  }
  if (seq_dims == 1) {
    if (q_full > 0) qfull[0] = q_full;
    if (k_full > 0) kfull[0] = k_full;
    if (q_base < 0 || k_base < 0 || int64_t(q_base) + qn[0] > qfull[0] || int64_t(k_base) + kn[0] > kfull[0])
      return -8;
  } else if (q_base || k_base || q_full || k_full) {
    return -8;
  }
  int32_t ref[2] = {1, 1}, qs[2] = {1, 1}, ks[2] = {1, 1}, qo[2] = {0, 0}, ko[2] = {0, 0};
  for (int i = 0; i < seq_dims; ++i) {
    int32_t mx = qfull[i] > kfull[i] ? qfull[i] : kfull[i];
    int32_t lg = fa_ceil_log2(mx);
    if (lg > 30) return -8;
    ref[i] = int32_t(1) << lg;
    if (sync_mode != 0) {
      qs[i] = mx / qfull[i];
      ks[i] = mx / kfull[i];
      if (sync_mode == 2) {
        qo[i] = qs[i] - 1;
        ko[i] = ks[i] - 1;
      }
    }
  }
  if (int64_t(ref[0]) * ref[1] > 0x7fffffffLL) return -8;  // orders must fit int32
  if (int64_t(qn[0]) * qn[1] > 0x7fffffffLL || int64_t(kn[0]) * kn[1] > 0x7fffffffLL) return -8;
  r.ref0 = ref[0];
  r.ref1 = ref[1];
  r.ref_log2_0 = fa_ceil_log2(ref[0]);
  r.q = FaSeqMap{qn[0], qn[1], qs[0], qs[1], qo[0], qo[1], q_base, qn[0] * qn[1]};
  r.k = FaSeqMap{kn[0], kn[1], ks[0], ks[1], ko[0], ko[1], k_base, kn[0] * kn[1]};
  *out = r;
  return 0;
}

// Position of local element `idx` (row-major over the local sequence).
FA_HD FaPos fa_pos(const FaRule& r, const FaSeqMap& s, int32_t idx) {
  FaPos p;
  if (r.dims == 1) {
    p.c0 = s.off0 + (idx + s.base0) * s.stride0;
    p.c1 = 0;
    p.order = p.c0;
  } else {
    int32_t y = idx / s.n0;
    int32_t x = idx - y * s.n0;
    p.c0 = s.off0 + x * s.stride0;
    p.c1 = s.off1 + y * s.stride1;
    p.order = (p.c1 << r.ref_log2_0) + p.c0;
  }
  return p;
}

FA_HD int32_t fa_iabs(int32_t v) { return v < 0 ? -v : v; }

// The element rule (reference Check()); padding is the caller's business.
FA_HD bool fa_attend(const FaRule& r, const FaPos& q, const FaPos& k) {
  if (r.rule == 0) return true;
  if (r.causal && q.order < k.order) return false;
  if (r.rule == 1) return true;
  const int32_t rem = (int32_t(1) << r.log2_stride) - 1;
  int32_t d0 = fa_iabs(q.c0 - k.c0);
  if ((d0 & rem) != 0 || (d0 >> r.log2_stride) >= r.window) return false;
  if (r.dims == 2) {
    int32_t d1 = fa_iabs(q.c1 - k.c1);
    if ((d1 & rem) != 0 || (d1 >> r.log2_stride) >= r.window) return false;
  }
  return true;
}

// Closed interval of coordinates/orders covered by local elements [lo, hi] (hi clamped
// by the caller to total-1). Conservative for 2-D ranges that span several rows.
struct FaBox {
  int32_t c0_lo, c0_hi, c1_lo, c1_hi, ord_lo, ord_hi;
};

FA_HD FaBox fa_box(const FaRule& r, const FaSeqMap& s, int32_t lo, int32_t hi) {
  FaBox b;
  FaPos a = fa_pos(r, s, lo), z = fa_pos(r, s, hi);
  b.ord_lo = a.order;
  b.ord_hi = z.order;
  b.c1_lo = a.c1;
  b.c1_hi = z.c1;
  if (r.dims == 1 || a.c1 == z.c1) {
    b.c0_lo = a.c0;
    b.c0_hi = z.c0;
  } else {
    b.c0_lo = s.off0;
    b.c0_hi = s.off0 + (s.n0 - 1) * s.stride0;
  }
  return b;
}

// interval helpers: min and max of |a-b| for a in [alo,ahi], b in [blo,bhi]
FA_HD int32_t fa_min_absdiff(int32_t alo, int32_t ahi, int32_t blo, int32_t bhi) {
  if (ahi < blo) return blo - ahi;
  if (bhi < alo) return alo - bhi;
  return 0;
}
FA_HD int32_t fa_max_absdiff(int32_t alo, int32_t ahi, int32_t blo, int32_t bhi) {
  int32_t a = ahi - blo, b = bhi - alo;
  return a > b ? a : b;
}

FA_HD int64_t fa_i64max(int64_t a, int64_t b) { return a > b ? a : b; }
FA_HD int64_t fa_i64min(int64_t a, int64_t b) { return a < b ? a : b; }
// smallest multiple of S (> 0) that is >= v
FA_HD int64_t fa_ceil_mult(int64_t v, int64_t S) {
  const int64_t q = v >= 0 ? (v + S - 1) / S : -((-v) / S);
  return q * S;
}

// Classifies the block of local q rows [q_lo, q_hi] x local k columns [k_lo, k_hi]
// (both already clamped to valid elements). SKIP => no pair attends. FULL => every pair
// attends. Anything else is PARTIAL and must evaluate fa_attend per element.
FA_HD int fa_classify(const FaRule& r, int32_t q_lo, int32_t q_hi, int32_t k_lo, int32_t k_hi) {
  if (r.rule == 0) return FA_TILE_FULL;
  FaBox q = fa_box(r, r.q, q_lo, q_hi), k = fa_box(r, r.k, k_lo, k_hi);
  bool full = true;
  if (r.causal) {
    if (q.ord_hi < k.ord_lo) return FA_TILE_SKIP;
    if (q.ord_lo < k.ord_hi) full = false;
  }
  if (r.rule == 2) {
    const int64_t sw = int64_t(r.window) << r.log2_stride;  // |delta| must be < sw
    if (int64_t(fa_min_absdiff(q.c0_lo, q.c0_hi, k.c0_lo, k.c0_hi)) >= sw) return FA_TILE_SKIP;
    if (r.dims == 2 && int64_t(fa_min_absdiff(q.c1_lo, q.c1_hi, k.c1_lo, k.c1_hi)) >= sw) return FA_TILE_SKIP;
    if (r.log2_stride != 0) {
      full = false;
      // strided window: |delta| must be a multiple of 2^s below sw. A tile whose delta interval holds no such multiple
      // attends nothing (conservative: the interval ignores the lattice the sync strides put the coordinates on).
      const int64_t S = int64_t(1) << r.log2_stride;
      {
        const int64_t lo = fa_i64max(int64_t(q.c0_lo) - k.c0_hi, -(sw - 1)), hi = fa_i64min(int64_t(q.c0_hi) - k.c0_lo, sw - 1);
        if (lo > hi || fa_ceil_mult(lo, S) > hi) return FA_TILE_SKIP;
      }
      if (r.dims == 2) {
        const int64_t lo = fa_i64max(int64_t(q.c1_lo) - k.c1_hi, -(sw - 1)), hi = fa_i64min(int64_t(q.c1_hi) - k.c1_lo, sw - 1);
        if (lo > hi || fa_ceil_mult(lo, S) > hi) return FA_TILE_SKIP;
      }
    } else {
      if (fa_max_absdiff(q.c0_lo, q.c0_hi, k.c0_lo, k.c0_hi) >= r.window) full = false;
      if (r.dims == 2 && fa_max_absdiff(q.c1_lo, q.c1_hi, k.c1_lo, k.c1_hi) >= r.window) full = false;
    }
  }
  return full ? FA_TILE_FULL : FA_TILE_PARTIAL;
}

// Range of K tiles [first, last] (inclusive, tiles of tile_k) that can be non-skipped for
// the q rows [q_lo, q_hi]; tiles outside are SKIP for sure. Returns first > last if none.
FA_HD void fa_k_tile_range(const FaRule& r, int32_t q_lo, int32_t q_hi, int32_t tile_k,
                           int32_t* first, int32_t* last) {
  const int32_t nkt = (r.k.total + tile_k - 1) / tile_k;
  int32_t lo = 0, hi = r.k.total - 1;  // candidate local k index interval
  if (r.rule != 0) {
    FaBox q = fa_box(r, r.q, q_lo, q_hi);
    // interval of the outermost K coordinate that can attend
    int64_t c_lo = -(int64_t(1) << 40), c_hi = (int64_t(1) << 40);
    const int32_t qo_lo = r.dims == 1 ? q.c0_lo : q.c1_lo, qo_hi = r.dims == 1 ? q.c0_hi : q.c1_hi;
    if (r.causal) c_hi = qo_hi;  // order(q) >= order(k) implies outer coord of k <= that of q
    if (r.rule == 2) {
      const int64_t sw = int64_t(r.window) << r.log2_stride;
      if (qo_lo - sw + 1 > c_lo) c_lo = qo_lo - sw + 1;
      if (qo_hi + sw - 1 < c_hi) c_hi = qo_hi + sw - 1;
    }
    // outer coordinate -> local index interval
    const int32_t st = r.dims == 1 ? r.k.stride0 : r.k.stride1;
    const int32_t of = r.dims == 1 ? r.k.off0 : r.k.off1;
    const int32_t n = r.dims == 1 ? r.k.n0 : r.k.n1;
    const int32_t base = r.dims == 1 ? r.k.base0 : 0;
    // smallest j with of + (j+base)*st >= c_lo ; largest j with of + (j+base)*st <= c_hi
    int64_t jlo = c_lo - of <= 0 ? 0 : (c_lo - of + st - 1) / st;
    int64_t jhi = c_hi - of < 0 ? -1 : (c_hi - of) / st;
    jlo -= base;
    jhi -= base;
    if (jlo < 0) jlo = 0;
    if (jhi > n - 1) jhi = n - 1;
    if (jlo > jhi) {
      *first = 1;
      *last = 0;
      return;
    }
    if (r.dims == 1) {
      lo = int32_t(jlo);
      hi = int32_t(jhi);
    } else {
      lo = int32_t(jlo) * r.k.n0;
      hi = int32_t(jhi) * r.k.n0 + r.k.n0 - 1;
    }
  }
  *first = lo / tile_k;
  *last = hi / tile_k;
  if (*last > nkt - 1) *last = nkt - 1;
}

// Same, transposed: range of Q tiles that can be non-skipped for the k columns [k_lo, k_hi].
FA_HD void fa_q_tile_range(const FaRule& r, int32_t k_lo, int32_t k_hi, int32_t tile_q,
                           int32_t* first, int32_t* last) {
  const int32_t nqt = (r.q.total + tile_q - 1) / tile_q;
  int32_t lo = 0, hi = r.q.total - 1;
  if (r.rule != 0) {
    FaBox k = fa_box(r, r.k, k_lo, k_hi);
    int64_t c_lo = -(int64_t(1) << 40), c_hi = (int64_t(1) << 40);
    const int32_t ko_lo = r.dims == 1 ? k.c0_lo : k.c1_lo, ko_hi = r.dims == 1 ? k.c0_hi : k.c1_hi;
    if (r.causal) c_lo = ko_lo;
    if (r.rule == 2) {
      const int64_t sw = int64_t(r.window) << r.log2_stride;
      if (ko_lo - sw + 1 > c_lo) c_lo = ko_lo - sw + 1;
      if (ko_hi + sw - 1 < c_hi) c_hi = ko_hi + sw - 1;
    }
    const int32_t st = r.dims == 1 ? r.q.stride0 : r.q.stride1;
    const int32_t of = r.dims == 1 ? r.q.off0 : r.q.off1;
    const int32_t n = r.dims == 1 ? r.q.n0 : r.q.n1;
    const int32_t base = r.dims == 1 ? r.q.base0 : 0;
    int64_t jlo = c_lo - of <= 0 ? 0 : (c_lo - of + st - 1) / st;
    int64_t jhi = c_hi - of < 0 ? -1 : (c_hi - of) / st;
    jlo -= base;
    jhi -= base;
    if (jlo < 0) jlo = 0;
    if (jhi > n - 1) jhi = n - 1;
    if (jlo > jhi) {
      *first = 1;
      *last = 0;
      return;
    }
    if (r.dims == 1) {
      lo = int32_t(jlo);
      hi = int32_t(jhi);
    } else {
      lo = int32_t(jlo) * r.q.n0;
      hi = int32_t(jhi) * r.q.n0 + r.q.n0 - 1;
    }
  }
  *first = lo / tile_q;
  *last = hi / tile_q;
  if (*last > nqt - 1) *last = nqt - 1;
}

// ---- closed-form 32-column masks (used by the tcgen05 kernels; host-testable) --------------------
FA_HD int32_t fa_imin(int32_t a, int32_t b) { return a < b ? a : b; }
FA_HD int32_t fa_imax(int32_t a, int32_t b) { return a > b ? a : b; }
// floor / ceil division by a positive divisor
FA_HD int32_t fa_fdiv(int32_t a, int32_t b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
FA_HD int32_t fa_cdiv(int32_t a, int32_t b) { return a >= 0 ? (a + b - 1) / b : -((-a) / b); }

// bits e in [0,32) with lo <= col0 + e <= hi
FA_HD uint32_t interval_bits32(int32_t lo, int32_t hi, int32_t col0) {
  const int32_t a = fa_imax(lo - col0, 0), b = fa_imin(hi - col0, 31);
  if (a > b) return 0u;
  const uint32_t upto_b = b >= 31 ? 0xffffffffu : ((1u << (b + 1)) - 1u);
  const uint32_t below_a = (1u << a) - 1u;
  return upto_b & ~below_a;
}

// 32-bit attended mask of streamed entries [s0+col0, s0+col0+32) (only the first `nvalid` entries of
// the tile starting at s0 exist) against one fixed resident position: full, causal and local rules (plain windows are an
// interval per grid row; strided windows the multiples of 2^s inside it, fa_strided_row_bits); 1-D and 2-D; any sync
// mode. For one grid row y of the streamed sequence the attended x are
//   { x : |res.c0 - cx| < W and causal(order) },  cx = off0 + x*stride0,  provided |res.c1 - cy| < W,
// an x-interval, so a chunk costs a few integer ops per grid row it touches instead of 32 evaluations
// of the element rule.
// bits of the streamed entries row0 + x, x in [xa, xb] (one grid row; coordinate cx = off0 + (x + xbase) * stride0), whose
// coordinate lies in [clo, chi] and differs from c_res by a multiple of 2^s: enumerates whichever is fewer, the
// multiples inside the interval or the positions of the row segment.
FA_HD uint32_t fa_strided_row_bits(int32_t c_res, int32_t clo, int32_t chi, int32_t log2_stride, int32_t off0,
                                   int32_t stride0, int32_t xbase, int32_t xa, int32_t xb, int32_t bit0) {
  // clip the coordinate interval to the segment
  clo = fa_imax(clo, off0 + (xa + xbase) * stride0);
  chi = fa_imin(chi, off0 + (xb + xbase) * stride0);
  if (clo > chi) return 0u;
  const int32_t S = int32_t(1) << log2_stride;
  uint32_t bits = 0;
  const int32_t m_lo = fa_cdiv(c_res - chi, S), m_hi = fa_fdiv(c_res - clo, S);
  if (m_hi - m_lo <= xb - xa) {
    for (int32_t m = m_lo; m <= m_hi; ++m) {
      const int32_t t = c_res - m * S - off0;
      if (t % stride0 != 0) continue;
      const int32_t x = t / stride0 - xbase;
      if (x >= xa && x <= xb) bits |= 1u << (bit0 + x - xa);
    }
  } else {
    const int32_t rem = S - 1;
    for (int32_t x = xa; x <= xb; ++x) {
      const int32_t cx = off0 + (x + xbase) * stride0;
      if (cx >= clo && cx <= chi && ((c_res - cx) & rem) == 0) bits |= 1u << (bit0 + x - xa);
    }
  }
  return bits;
}

FA_HD uint32_t fa_fast_mask32(const FaRule& rule, bool resident_is_q, const FaPos& res, int32_t s0,
                              int32_t col0, int32_t nvalid) {
  const FaSeqMap& sm = resident_is_q ? rule.k : rule.q;
  const bool strided = rule.rule == 2 && rule.log2_stride != 0;
  // |delta| < W (plain windows) or |delta| < W * 2^s and a multiple of 2^s (strided windows)
  const int64_t w64 = rule.rule == 2 ? (int64_t(rule.window) << rule.log2_stride) : int64_t(0x3fffffff);
  const int32_t W = int32_t(w64 < 0x3fffffff ? w64 : 0x3fffffff);
  const int32_t j0 = s0 + col0;
  const int32_t jend = fa_imin(j0 + 31, s0 + nvalid - 1);
  if (jend < j0) return 0u;
  if (rule.rule == 0) return interval_bits32(0, jend - j0, 0);
  if (rule.dims == 1) {
    int32_t clo = res.c0 - W + 1, chi = res.c0 + W - 1;
    if (rule.causal) {
      if (resident_is_q) chi = fa_imin(chi, res.c0);  // k.c0 <= q.c0
      else clo = fa_imax(clo, res.c0);                // q.c0 >= k.c0
    }
    if (strided)
      return fa_strided_row_bits(res.c0, clo, chi, rule.log2_stride, sm.off0, sm.stride0, sm.base0, j0, jend, 0);
    const int32_t jlo = fa_cdiv(clo - sm.off0, sm.stride0) - sm.base0;
    const int32_t jhi = fa_fdiv(chi - sm.off0, sm.stride0) - sm.base0;
    return interval_bits32(fa_imax(jlo, j0) - j0, fa_imin(jhi, jend) - j0, 0);
  }
  uint32_t bits = 0;
  const int32_t y_first = j0 / sm.n0, y_last = jend / sm.n0;
  for (int32_t y = y_first; y <= y_last; ++y) {
    const int32_t cy = sm.off1 + y * sm.stride1;
    if (fa_iabs(res.c1 - cy) >= W) continue;
    if (strided && (fa_iabs(res.c1 - cy) & ((int32_t(1) << rule.log2_stride) - 1)) != 0) continue;
    int32_t clo = res.c0 - W + 1, chi = res.c0 + W - 1;
    if (rule.causal) {
      if (resident_is_q) {  // order(q) >= order(k)
        if (cy > res.c1) continue;
        if (cy == res.c1) chi = fa_imin(chi, res.c0);
      } else {
        if (cy < res.c1) continue;
        if (cy == res.c1) clo = fa_imax(clo, res.c0);
      }
    }
    const int32_t row0 = y * sm.n0;
    if (strided) {
      const int32_t xa = fa_imax(j0 - row0, 0), xb = fa_imin(jend - row0, sm.n0 - 1);
      if (xa <= xb)
        bits |= fa_strided_row_bits(res.c0, clo, chi, rule.log2_stride, sm.off0, sm.stride0, 0, xa, xb, row0 + xa - j0);
      continue;
    }
    const int32_t xlo = fa_imax(fa_cdiv(clo - sm.off0, sm.stride0), 0);
    const int32_t xhi = fa_imin(fa_fdiv(chi - sm.off0, sm.stride0), sm.n0 - 1);
    if (xlo > xhi) continue;
    bits |= interval_bits32(fa_imax(row0 + xlo, j0) - j0, fa_imin(row0 + xhi, jend) - j0, 0);
  }
  return bits;
}
