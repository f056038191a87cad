// fa_f64_dmma.cu — double precision on the FP64 tensor cores (DMMA, `mma.sync.m8n8k4.f64`).
//
// The reference instantiates every rule for double (flash_attention.cu:2465-2481, macro_util.h:8-25) on its scalar-FMA
// kernels. tcgen05 has no fp64 kind, so double runs as an FA-2 style register kernel on `mma.sync`: one CTA = 4 warps x
// 16 rows (64 resident rows), streamed tiles of 32, accumulators in registers (C fragments), K / V (or Q / dO) tiles
// brought in with 8-byte `cp.async` into a double-buffered shared-memory ring (8-byte copies: rows of a channel-first
// double tensor are only 8-byte aligned when the sequence length is odd; out-of-range positions and padded channels are
// zero-filled by the copy itself). P / dS leave the accumulator layout for the A-operand layout through quad shuffles.
// Channels are padded to 32 or 64 (the reference's fp64 test shapes use 8..32; C4 uses 64); anything larger, and
// `accumulate`, stays on the generic DFMA kernels (fa_generic.cu), which are also the fallback.
//
// Fragment layouts of m8n8k4 (g = lane / 4, t = lane % 4):  A[8x4]: a = A[g][t];  B[4x8]: b = B[t][g];
// C[8x8]: c0 = C[g][2t], c1 = C[g][2t+1].
#include "fa_common.cuh"
#include "fa_launch.h"
#include "fa_plan.h"

namespace fa {
namespace f64 {

constexpr int kRows = 64;       // resident rows per CTA: MT m8 tiles per warp, 64 / (8 MT) warps (MT = 2: 4 warps, MT = 1: 8)
__host__ __device__ constexpr int threads_of(int mt) { return kRows / (8 * mt) * 32; }
constexpr int kTile = 32;       // streamed positions per tile
constexpr int kPitch = 36;      // doubles per shared-memory row of a streamed tile (conflict-free 64-bit reads)
constexpr int kRPitch = 68;     // doubles per row of a resident tile [channel][64 rows]

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* src, bool valid) {
  const uint32_t d = uint32_t(__cvta_generic_to_shared(smem_dst));
  const int n = valid ? 8 : 0;   // src-size 0: the 8 bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// [C_PAD channels][32 positions] tile of a channel-first tensor (row pitch n) starting at x0 -> dst[c][kPitch]
template <int C_PAD>
__device__ __forceinline__ void load_stream_tile(double* dst, const double* __restrict__ src, int channels, int64_t n,
                                                 int64_t x0) {
  for (int idx = threadIdx.x; idx < C_PAD * kTile; idx += blockDim.x) {
    const int c = idx / kTile, x = idx - c * kTile;
    const bool ok = c < channels && x0 + x < n;
    cp_async8(dst + c * kPitch + x, ok ? src + int64_t(c) * n + x0 + x : src, ok);
  }
}
// quad reductions (the 4 lanes that share a row of a C fragment)
__device__ __forceinline__ double quad_max(double v) {
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmax(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ double quad_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
// C-fragment pair (columns 2t, 2t+1 of an 8-wide tile) -> A fragment of the 4 columns [4*half, 4*half+4)
__device__ __forceinline__ double c_to_a(const double (&c)[2], int half, int lane) {
  const int src = (lane & ~3) | (half * 2 + ((lane & 3) >> 1));
  const double v0 = __shfl_sync(0xffffffffu, c[0], src);
  const double v1 = __shfl_sync(0xffffffffu, c[1], src);
  return (lane & 1) ? v1 : v0;
}

struct FwdParams {
  const double *q, *k, *v;
  double *o, *l, *m;
  int32_t d, v_d, nq, nk, n_rtiles;
  int64_t batch;
  FaRule rule;
};

template <int DP, int VP>
struct FwdSmem {
  static constexpr int kStage = (DP + VP) * kPitch;             // doubles per ring stage (K tile then V tile)
  static constexpr int kQ = DP * kRPitch;                       // resident Q tile (only while the fragments are read)
  static constexpr int kDoubles = (2 * kStage > kQ ? 2 * kStage : kQ);
  static constexpr int kBytes = kDoubles * 8;
};

// Classes of 32 consecutive streamed tiles, one per lane, kept as two ballots per warp: fa_classify (two fa_box
// evaluations with integer divisions) costs as much as the DMMAs of a 32-wide tile when every thread repeats it per
// tile; the tile index is uniform over the CTA, so a warp classifies 32 tiles at once and looks the rest up.
struct WarpTileClass {
  uint32_t partial = 0, full = 0;
  int base = -(1 << 30);
  template <typename Classify>
  __device__ __forceinline__ int get(int tile, Classify&& classify) {
    if (tile < base || tile >= base + 32) {   // warp-uniform
      base = tile;
      const int c = classify(tile + int(threadIdx.x & 31));
      partial = __ballot_sync(0xffffffffu, c == FA_TILE_PARTIAL);
      full = __ballot_sync(0xffffffffu, c == FA_TILE_FULL);
    }
    const int b = tile - base;
    return ((full >> b) & 1u) ? FA_TILE_FULL : ((partial >> b) & 1u) ? FA_TILE_PARTIAL : FA_TILE_SKIP;
  }
};

template <int DP, int VP, int MT>
__global__ void __launch_bounds__(threads_of(MT), 2) fwd_kernel(const FwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int rt = p.n_rtiles - 1 - int(blockIdx.x % p.n_rtiles);   // heavy (late) tiles first
  const int64_t b = blockIdx.x / p.n_rtiles;
  const int q0 = rt * kRows;
  const int q_hi = min(q0 + kRows, p.nq) - 1;
  const FaRule& rule = p.rule;
  const double scale = 1.0 / sqrt(double(p.d));

  // Q tile -> shared memory [channel][row] -> A fragments (pre-scaled), two m-tiles per warp
  {
    const double* qg = p.q + b * p.d * int64_t(p.nq);
    for (int idx = threadIdx.x; idx < DP * kRows; idx += blockDim.x) {
      const int c = idx / kRows, r = idx - c * kRows;
      sm[c * kRPitch + r] = (c < p.d && q0 + r < p.nq) ? qg[int64_t(c) * p.nq + q0 + r] * scale : 0.0;
    }
  }
  __syncthreads();
  double qa[MT][DP / 4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int ks = 0; ks < DP / 4; ++ks) qa[mt][ks] = sm[(4 * ks + t) * kRPitch + warp * (8 * MT) + mt * 8 + g];
  __syncthreads();

  int row[MT];
  FaPos qpos[MT];
  double m_i[MT], l_i[MT], o[MT][VP / 8][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    row[mt] = q0 + warp * (8 * MT) + mt * 8 + g;
    qpos[mt] = fa_pos(rule, rule.q, min(row[mt], p.nq - 1));
    m_i[mt] = neg_inf<double>();
    l_i[mt] = 0.0;
#pragma unroll
    for (int vt = 0; vt < VP / 8; ++vt) o[mt][vt][0] = o[mt][vt][1] = 0.0;
  }

  const double* kg = p.k + b * p.d * int64_t(p.nk);
  const double* vg = p.v + b * p.v_d * int64_t(p.nk);
  int kt_first, kt_last;
  fa_k_tile_range(rule, q0, q_hi, kTile, &kt_first, &kt_last);
  WarpTileClass tile_cls;
  auto classify = [&](int kt) {   // beyond the last tile: SKIP
    return kt * kTile < p.nk ? fa_classify(rule, q0, q_hi, kt * kTile, min(kt * kTile + kTile, p.nk) - 1) : int(FA_TILE_SKIP);
  };
  auto next_live = [&](int kt) {   // first tile >= kt that is not skipped (kt_last + 1 if none)
    while (kt <= kt_last && tile_cls.get(kt, classify) == FA_TILE_SKIP) ++kt;
    return kt;
  };
  auto issue = [&](int kt, int stage) {
    double* ks = sm + stage * FwdSmem<DP, VP>::kStage;
    load_stream_tile<DP>(ks, kg, p.d, p.nk, int64_t(kt) * kTile);
    load_stream_tile<VP>(ks + DP * kPitch, vg, p.v_d, p.nk, int64_t(kt) * kTile);
    cp_async_commit();
  };
  int kt = next_live(kt_first), stage = 0;
  if (kt <= kt_last) issue(kt, 0);
  while (kt <= kt_last) {
    const int cls = tile_cls.get(kt, classify);   // before the look-ahead moves the 32-tile window
    const int kn = next_live(kt + 1);
    if (kn <= kt_last) {
      issue(kn, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* Ks = sm + stage * FwdSmem<DP, VP>::kStage;
    const double* Vs = Ks + DP * kPitch;
    const int k0 = kt * kTile;

    // S = (Q scale) K^T : 2 m-tiles x 4 n-tiles
    double s[MT][4][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) s[mt][nt][0] = s[mt][nt][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < DP / 4; ++ks) {
      double bk[4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bk[nt] = Ks[(4 * ks + t) * kPitch + nt * 8 + g];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma(s[mt][nt], qa[mt][ks], bk[nt]);
    }
    // mask: one closed-form 32-bit word per row and tile (fa_fast_mask32), bit = streamed column
    const bool full = cls == FA_TILE_FULL && k0 + kTile <= p.nk;
    if (!full) {
      const int nvalid = min(kTile, p.nk - k0);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const uint32_t bits = fa_fast_mask32(rule, true, qpos[mt], k0, 0, nvalid);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (!((bits >> (nt * 8 + 2 * t + e)) & 1u)) s[mt][nt][e] = neg_inf<double>();
      }
    }
    // online softmax per row (a row lives in the 4 lanes of a quad)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      double mx = s[mt][0][0];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mx = fmax(mx, fmax(s[mt][nt][0], s[mt][nt][1]));
      mx = quad_max(mx);
      const double m_new = fmax(m_i[mt], mx);
      double alpha = 1.0, sum = 0.0;
      if (m_new == neg_inf<double>()) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) s[mt][nt][0] = s[mt][nt][1] = 0.0;
      } else {
        alpha = exp(m_i[mt] - m_new);   // exp(-inf) = 0 on the first live tile
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            s[mt][nt][e] = exp(s[mt][nt][e] - m_new);
            sum += s[mt][nt][e];
          }
      }
      l_i[mt] = l_i[mt] * alpha + sum;   // per-lane partial sums: alpha is the same in the whole quad
      m_i[mt] = m_new;
#pragma unroll
      for (int vt = 0; vt < VP / 8; ++vt) {
        o[mt][vt][0] *= alpha;
        o[mt][vt][1] *= alpha;
      }
    }
    // O += P V : k-steps of 4 keys; A from the S fragments through quad shuffles
#pragma unroll
    for (int kk = 0; kk < kTile / 4; ++kk) {
      double pa[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) pa[mt] = c_to_a(s[mt][kk >> 1], kk & 1, lane);
#pragma unroll
      for (int vt = 0; vt < VP / 8; ++vt) {
        const double bv = Vs[(vt * 8 + g) * kPitch + 4 * kk + t];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dmma(o[mt][vt], pa[mt], bv);
      }
    }
    __syncthreads();   // every warp is done with this stage before it is refilled
    kt = kn;
    stage ^= 1;
  }

  // epilogue: O = acc / l, m, l (channel-first stores: 8 consecutive rows per channel and quad column)
  double* og = p.o + b * p.v_d * int64_t(p.nq);
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const double l_row = quad_sum(l_i[mt]);
    const double inv = l_row > 0.0 ? 1.0 / l_row : 0.0;
    if (row[mt] < p.nq) {
#pragma unroll
      for (int vt = 0; vt < VP / 8; ++vt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ch = vt * 8 + 2 * t + e;
          if (ch < p.v_d) og[int64_t(ch) * p.nq + row[mt]] = o[mt][vt][e] * inv;
        }
      if (t == 0) {
        const int64_t idx = b * p.nq + row[mt];
        if (l_row > 0.0) {
          p.m[idx] = m_i[mt];
          p.l[idx] = l_row;
        } else {
          p.m[idx] = sentinel<double>();
          p.l[idx] = 0.0;
        }
      }
    }
  }
}

template <int DP, int VP, int MT>
static cudaError_t launch_fwd(const LaunchArgs& a, cudaStream_t stream) {
  FwdParams p;
  p.q = (const double*)a.q; p.k = (const double*)a.k; p.v = (const double*)a.v;
  p.o = (double*)a.o; p.l = (double*)a.l; p.m = (double*)a.m;
  p.d = a.d; p.v_d = a.v_d; p.nq = a.rule.q.total; p.nk = a.rule.k.total;
  p.n_rtiles = (p.nq + kRows - 1) / kRows;
  p.batch = a.batch;
  p.rule = a.rule;
  auto kern = fwd_kernel<DP, VP, MT>;
  cudaError_t e = plan::ensure_smem(kern, FwdSmem<DP, VP>::kBytes);
  if (e != cudaSuccess) return e;
  ScopedKernel timed("fwd_f64_dmma", stream);
  kern<<<unsigned(p.batch * p.n_rtiles), threads_of(MT), FwdSmem<DP, VP>::kBytes, stream>>>(p);
  return cudaGetLastError();
}

// =================================================================================================
// backward: statistics, dQ kernel (64 query rows resident, key tiles streamed), dK/dV kernel (64 keys
// resident, query tiles streamed). Formulas as in the reference: P = exp(s*scale - m)/l
// (flash_attention.cu:1838-1841), dS = P (dP - D) scale (:1544-1546), D = rowsum(dO o O) (:1882-1891).
// =================================================================================================
struct BwdParams {
  const double *q, *k, *v, *d_o;
  double *d_q, *d_k, *d_v;
  const double *lse, *dsum;   // [batch, nq]: m + log l (+inf on empty rows), rowsum(dO o O)
  int32_t d, v_d, nq, nk, n_tiles;
  int64_t batch;
  FaRule rule;
};

__global__ void bwd_prep_kernel(const double* __restrict__ o, const double* __restrict__ d_o,
                                const double* __restrict__ l, const double* __restrict__ m, double* __restrict__ lse,
                                double* __restrict__ dsum, int64_t batch, int32_t v_d, int32_t nq, double* zero0,
                                int64_t n0, double* zero1, int64_t n1, double* zero2, int64_t n2) {
  const int64_t total = batch * nq;
  // outputs that the split grids add into with atomics are cleared here (this launch precedes them anyway), not by
  // three memset nodes of their own
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n0 + n1 + n2; i += int64_t(gridDim.x) * blockDim.x) {
    if (i < n0) zero0[i] = 0.0;
    else if (i < n0 + n1) zero1[i - n0] = 0.0;
    else zero2[i - n0 - n1] = 0.0;
  }
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t b = i / nq, r = i - b * nq;
    const double* op = o + b * v_d * int64_t(nq) + r;
    const double* dp = d_o + b * v_d * int64_t(nq) + r;
    double acc = 0.0;
    for (int c = 0; c < v_d; ++c) acc += op[int64_t(c) * nq] * dp[int64_t(c) * nq];
    dsum[i] = acc;
    const double lv = l[i], mv = m[i];
    lse[i] = (lv > 0.0 && !is_sentinel<double>(mv)) ? mv + log(lv) : -neg_inf<double>();
  }
}

// Problems with fewer CTAs than SMs (the reference's fp64 benchmark shapes: 8 heads x 1024 positions = 128 CTAs, each a
// serial loop over 32 tiles) share the streamed range out over gridDim.y CTAs per resident tile; the partial gradients
// are then added into zero-initialised outputs with fp64 atomics (the order of two or four fp64 additions is the only
// thing that varies from run to run: ~1e-16 relative).
__device__ __forceinline__ void split_range(int* first, int* last) {
  if (gridDim.y == 1 || *first > *last) return;
  const int n = *last - *first + 1, per = (n + int(gridDim.y) - 1) / int(gridDim.y);
  const int lo = *first + int(blockIdx.y) * per;
  *first = lo;
  *last = min(lo + per - 1, *last);
}
__device__ __forceinline__ void store_or_add(double* dst, double v) {
  if (gridDim.y == 1) *dst = v;
  else atomicAdd(dst, v);
}

// resident [C_PAD channels][64 rows] tile -> dst[c][kRPitch], optionally scaled
template <int C_PAD>
__device__ __forceinline__ void load_resident_tile(double* dst, const double* __restrict__ src, int channels, int64_t n,
                                                   int64_t x0, double scale) {
  for (int idx = threadIdx.x; idx < C_PAD * kRows; idx += blockDim.x) {
    const int c = idx / kRows, r = idx - c * kRows;
    dst[c * kRPitch + r] = (c < channels && x0 + r < n) ? src[int64_t(c) * n + x0 + r] * scale : 0.0;
  }
}

template <int DP, int VP>
struct BwdSmem {
  static constexpr int kResident = (DP + VP) * kRPitch;          // Q + dO (dQ kernel) or K + V (dK/dV kernel)
  static constexpr int kStage = (DP + VP) * kPitch;              // streamed K + V or Q + dO tile
  static constexpr int kStats = 2 * 2 * kTile;                   // dK/dV kernel: lse + D of the streamed queries, 2 stages
  static constexpr int kBytes = (kResident + 2 * kStage + kStats) * 8;
};

template <int DP, int VP, int MT>
__global__ void __launch_bounds__(threads_of(MT), 1) bwd_dq_kernel(const BwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Qs = reinterpret_cast<double*>(smem_raw);          // [DP][kRPitch], pre-scaled by 1/sqrt(d)
  double* dOs = Qs + DP * kRPitch;                            // [VP][kRPitch]
  double* ring = dOs + VP * kRPitch;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int rt = p.n_tiles - 1 - int(blockIdx.x % p.n_tiles);
  const int64_t b = blockIdx.x / p.n_tiles;
  const int q0 = rt * kRows;
  const int q_hi = min(q0 + kRows, p.nq) - 1;
  const FaRule& rule = p.rule;
  const double scale = 1.0 / sqrt(double(p.d));
  load_resident_tile<DP>(Qs, p.q + b * p.d * int64_t(p.nq), p.d, p.nq, q0, scale);
  load_resident_tile<VP>(dOs, p.d_o + b * p.v_d * int64_t(p.nq), p.v_d, p.nq, q0, 1.0);

  int row[MT];
  FaPos qpos[MT];
  double lse[MT], dsum[MT], dq[MT][DP / 8][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    row[mt] = q0 + warp * (8 * MT) + mt * 8 + g;
    const bool valid = row[mt] < p.nq;
    qpos[mt] = fa_pos(rule, rule.q, min(row[mt], p.nq - 1));
    lse[mt] = valid ? p.lse[b * p.nq + row[mt]] : -neg_inf<double>();
    dsum[mt] = valid ? p.dsum[b * p.nq + row[mt]] : 0.0;
#pragma unroll
    for (int ct = 0; ct < DP / 8; ++ct) dq[mt][ct][0] = dq[mt][ct][1] = 0.0;
  }
  const double* kg = p.k + b * p.d * int64_t(p.nk);
  const double* vg = p.v + b * p.v_d * int64_t(p.nk);
  int kt_first, kt_last;
  fa_k_tile_range(rule, q0, q_hi, kTile, &kt_first, &kt_last);
  split_range(&kt_first, &kt_last);   // small grids: the streamed range is shared out over gridDim.y CTAs
  WarpTileClass tile_cls;
  auto classify = [&](int kt) {   // beyond the last tile: SKIP
    return kt * kTile < p.nk ? fa_classify(rule, q0, q_hi, kt * kTile, min(kt * kTile + kTile, p.nk) - 1) : int(FA_TILE_SKIP);
  };
  auto next_live = [&](int kt) {
    while (kt <= kt_last && tile_cls.get(kt, classify) == FA_TILE_SKIP) ++kt;
    return kt;
  };
  auto issue = [&](int kt, int stage) {
    double* ks = ring + stage * BwdSmem<DP, VP>::kStage;
    load_stream_tile<DP>(ks, kg, p.d, p.nk, int64_t(kt) * kTile);
    load_stream_tile<VP>(ks + DP * kPitch, vg, p.v_d, p.nk, int64_t(kt) * kTile);
    cp_async_commit();
  };
  int kt = next_live(kt_first), stage = 0;
  if (kt <= kt_last) issue(kt, 0);
  while (kt <= kt_last) {
    const int cls = tile_cls.get(kt, classify);   // before the look-ahead moves the 32-tile window
    const int kn = next_live(kt + 1);
    if (kn <= kt_last) {
      issue(kn, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();   // also orders the resident-tile stores of the prologue before the first reads
    const double* Ks = ring + stage * BwdSmem<DP, VP>::kStage;
    const double* Vs = Ks + DP * kPitch;
    const int k0 = kt * kTile;
    double s[MT][4][2], dp[MT][4][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) s[mt][nt][0] = s[mt][nt][1] = dp[mt][nt][0] = dp[mt][nt][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < DP / 4; ++ks) {
      double a[MT], bk[4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a[mt] = Qs[(4 * ks + t) * kRPitch + warp * (8 * MT) + mt * 8 + g];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bk[nt] = Ks[(4 * ks + t) * kPitch + nt * 8 + g];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma(s[mt][nt], a[mt], bk[nt]);
    }
#pragma unroll
    for (int ks = 0; ks < VP / 4; ++ks) {
      double a[MT], bv[4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a[mt] = dOs[(4 * ks + t) * kRPitch + warp * (8 * MT) + mt * 8 + g];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bv[nt] = Vs[(4 * ks + t) * kPitch + nt * 8 + g];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma(dp[mt][nt], a[mt], bv[nt]);
    }
    // dS = P (dP - D), P = exp(s - lse), masked (closed-form 32-bit word per row and tile); written over s
    const bool full = cls == FA_TILE_FULL && k0 + kTile <= p.nk;
    const int nvalid = min(kTile, p.nk - k0);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const uint32_t bits = full ? 0xffffffffu : fa_fast_mask32(rule, true, qpos[mt], k0, 0, nvalid);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const bool ok = (bits >> (nt * 8 + 2 * t + e)) & 1u;
          const double pv = ok ? exp(s[mt][nt][e] - lse[mt]) : 0.0;   // lse = +inf on empty rows -> 0
          s[mt][nt][e] = pv * (dp[mt][nt][e] - dsum[mt]);
        }
    }
    // dQ += dS K
#pragma unroll
    for (int kk = 0; kk < kTile / 4; ++kk) {
      double da[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) da[mt] = c_to_a(s[mt][kk >> 1], kk & 1, lane);
#pragma unroll
      for (int ct = 0; ct < DP / 8; ++ct) {
        const double bk = Ks[(ct * 8 + g) * kPitch + 4 * kk + t];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dmma(dq[mt][ct], da[mt], bk);
      }
    }
    __syncthreads();
    kt = kn;
    stage ^= 1;
  }
  double* dqg = p.d_q + b * p.d * int64_t(p.nq);
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
    if (row[mt] < p.nq) {
#pragma unroll
      for (int ct = 0; ct < DP / 8; ++ct)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ch = ct * 8 + 2 * t + e;
          if (ch < p.d) store_or_add(dqg + int64_t(ch) * p.nq + row[mt], dq[mt][ct][e] * scale);
        }
    }
}

template <int DP, int VP, int MT>
__global__ void __launch_bounds__(threads_of(MT), 1) bwd_dkdv_kernel(const BwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Ks = reinterpret_cast<double*>(smem_raw);          // [DP][kRPitch], pre-scaled by 1/sqrt(d)
  double* Vs = Ks + DP * kRPitch;                             // [VP][kRPitch]
  double* ring = Vs + VP * kRPitch;
  double* stats = ring + 2 * BwdSmem<DP, VP>::kStage;         // [2 stages][lse[32], D[32]]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int kb = int(blockIdx.x % p.n_tiles);                 // early key tiles are the heavy ones under causal rules
  const int64_t b = blockIdx.x / p.n_tiles;
  const int k0 = kb * kRows;
  const int k_hi = min(k0 + kRows, p.nk) - 1;
  const FaRule& rule = p.rule;
  const double scale = 1.0 / sqrt(double(p.d));
  load_resident_tile<DP>(Ks, p.k + b * p.d * int64_t(p.nk), p.d, p.nk, k0, scale);
  load_resident_tile<VP>(Vs, p.v + b * p.v_d * int64_t(p.nk), p.v_d, p.nk, k0, 1.0);

  int row[MT];
  FaPos kpos[MT];
  double dk[MT][DP / 8][2], dv[MT][VP / 8][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    row[mt] = k0 + warp * (8 * MT) + mt * 8 + g;
    kpos[mt] = fa_pos(rule, rule.k, min(row[mt], p.nk - 1));
#pragma unroll
    for (int ct = 0; ct < DP / 8; ++ct) dk[mt][ct][0] = dk[mt][ct][1] = 0.0;
#pragma unroll
    for (int ct = 0; ct < VP / 8; ++ct) dv[mt][ct][0] = dv[mt][ct][1] = 0.0;
  }
  const double* qg = p.q + b * p.d * int64_t(p.nq);
  const double* dog = p.d_o + b * p.v_d * int64_t(p.nq);
  int qt_first, qt_last;
  fa_q_tile_range(rule, k0, k_hi, kTile, &qt_first, &qt_last);
  split_range(&qt_first, &qt_last);
  WarpTileClass tile_cls;
  auto classify = [&](int qt) {   // beyond the last tile: SKIP
    return qt * kTile < p.nq ? fa_classify(rule, qt * kTile, min(qt * kTile + kTile, p.nq) - 1, k0, k_hi) : int(FA_TILE_SKIP);
  };
  auto next_live = [&](int qt) {
    while (qt <= qt_last && tile_cls.get(qt, classify) == FA_TILE_SKIP) ++qt;
    return qt;
  };
  auto issue = [&](int qt, int stage) {
    double* qs = ring + stage * BwdSmem<DP, VP>::kStage;
    load_stream_tile<DP>(qs, qg, p.d, p.nq, int64_t(qt) * kTile);
    load_stream_tile<VP>(qs + DP * kPitch, dog, p.v_d, p.nq, int64_t(qt) * kTile);
    if (threadIdx.x < 2 * kTile) {
      const int x = threadIdx.x % kTile;
      const bool ok = int64_t(qt) * kTile + x < p.nq;
      const double* src = (threadIdx.x < kTile ? p.lse : p.dsum) + b * p.nq;
      cp_async8(stats + stage * 2 * kTile + threadIdx.x, ok ? src + int64_t(qt) * kTile + x : src, ok);
    }
    cp_async_commit();
  };
  int qt = next_live(qt_first), stage = 0;
  if (qt <= qt_last) issue(qt, 0);
  while (qt <= qt_last) {
    const int cls = tile_cls.get(qt, classify);   // before the look-ahead moves the 32-tile window
    const int qn = next_live(qt + 1);
    if (qn <= qt_last) {
      issue(qn, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* Qs = ring + stage * BwdSmem<DP, VP>::kStage;
    const double* dOs = Qs + DP * kPitch;
    const double* lse = stats + stage * 2 * kTile;
    const double* dsum = lse + kTile;
    const int q0 = qt * kTile;
    // S^T = (K scale) Q^T, dP^T = V dO^T : rows = keys, columns = queries
    double s[MT][4][2], dp[MT][4][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) s[mt][nt][0] = s[mt][nt][1] = dp[mt][nt][0] = dp[mt][nt][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < DP / 4; ++ks) {
      double a[MT], bq[4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a[mt] = Ks[(4 * ks + t) * kRPitch + warp * (8 * MT) + mt * 8 + g];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bq[nt] = Qs[(4 * ks + t) * kPitch + nt * 8 + g];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma(s[mt][nt], a[mt], bq[nt]);
    }
#pragma unroll
    for (int ks = 0; ks < VP / 4; ++ks) {
      double a[MT], bo[4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a[mt] = Vs[(4 * ks + t) * kRPitch + warp * (8 * MT) + mt * 8 + g];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bo[nt] = dOs[(4 * ks + t) * kPitch + nt * 8 + g];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma(dp[mt][nt], a[mt], bo[nt]);
    }
    // P^T -> s, dS^T -> dp (column statistics: the queries of this tile); one closed-form mask word per key row
    const bool full = cls == FA_TILE_FULL && q0 + kTile <= p.nq && k0 + kRows <= p.nk;
    const int nvalid = min(kTile, p.nq - q0);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const uint32_t bits = full ? 0xffffffffu
                                 : (row[mt] < p.nk ? fa_fast_mask32(rule, false, kpos[mt], q0, 0, nvalid) : 0u);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int x = nt * 8 + 2 * t + e;
          const bool ok = (bits >> x) & 1u;
          const double pv = ok ? exp(s[mt][nt][e] - lse[x]) : 0.0;   // stats are zero-filled past the sequence
          s[mt][nt][e] = pv;
          dp[mt][nt][e] = pv * (dp[mt][nt][e] - dsum[x]);
        }
    }
    // dV += P^T dO, dK += dS^T Q (k-steps of 4 queries)
#pragma unroll
    for (int kk = 0; kk < kTile / 4; ++kk) {
      double pa[MT], da[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        pa[mt] = c_to_a(s[mt][kk >> 1], kk & 1, lane);
        da[mt] = c_to_a(dp[mt][kk >> 1], kk & 1, lane);
      }
#pragma unroll
      for (int ct = 0; ct < VP / 8; ++ct) {
        const double bo = dOs[(ct * 8 + g) * kPitch + 4 * kk + t];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dmma(dv[mt][ct], pa[mt], bo);
      }
#pragma unroll
      for (int ct = 0; ct < DP / 8; ++ct) {
        const double bq = Qs[(ct * 8 + g) * kPitch + 4 * kk + t];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) dmma(dk[mt][ct], da[mt], bq);
      }
    }
    __syncthreads();
    qt = qn;
    stage ^= 1;
  }
  double* dkg = p.d_k + b * p.d * int64_t(p.nk);
  double* dvg = p.d_v + b * p.v_d * int64_t(p.nk);
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
    if (row[mt] < p.nk) {
#pragma unroll
      for (int ct = 0; ct < DP / 8; ++ct)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ch = ct * 8 + 2 * t + e;
          if (ch < p.d) store_or_add(dkg + int64_t(ch) * p.nk + row[mt], dk[mt][ct][e] * scale);
        }
#pragma unroll
      for (int ct = 0; ct < VP / 8; ++ct)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ch = ct * 8 + 2 * t + e;
          if (ch < p.v_d) store_or_add(dvg + int64_t(ch) * p.nk + row[mt], dv[mt][ct][e]);
        }
    }
}

template <int DP, int VP, int MT>
static cudaError_t launch_bwd(const LaunchArgs& a, cudaStream_t stream) {
  BwdParams p;
  p.q = (const double*)a.q; p.k = (const double*)a.k; p.v = (const double*)a.v; p.d_o = (const double*)a.d_o;
  p.d_q = (double*)a.d_q; p.d_k = (double*)a.d_k; p.d_v = (double*)a.d_v;
  p.d = a.d; p.v_d = a.v_d; p.nq = a.rule.q.total; p.nk = a.rule.k.total;
  p.batch = a.batch;
  p.rule = a.rule;
  double* lse = (double*)a.workspace;
  double* dsum = lse + p.batch * p.nq;
  p.lse = lse;
  p.dsum = dsum;
  auto splits_for = [](int64_t ctas, int64_t streamed_tiles) {   // 1 unless the grid leaves SMs idle
    int n = 1;
    while (ctas * n < 148 && n < 4 && streamed_tiles / (2 * n) >= 4) n *= 2;
    return n;
  };
  const int64_t tiles_q = (p.nq + kRows - 1) / kRows, tiles_k = (p.nk + kRows - 1) / kRows;
  const int ns_q = splits_for(p.batch * tiles_q, (p.nk + kTile - 1) / kTile);
  const int ns_k = splits_for(p.batch * tiles_k, (p.nq + kTile - 1) / kTile);
  {
    const int64_t total = p.batch * p.nq;
    const int64_t z0 = ns_q > 1 ? p.batch * p.d * int64_t(p.nq) : 0;
    const int64_t z1 = ns_k > 1 ? p.batch * p.d * int64_t(p.nk) : 0, z2 = ns_k > 1 ? p.batch * p.v_d * int64_t(p.nk) : 0;
    const int blocks = int(std::min<int64_t>((std::max(total, z0 + z1 + z2) + 255) / 256, 148 * 8));
    ScopedKernel timed("bwd_prep_f64", stream);
    bwd_prep_kernel<<<blocks, 256, 0, stream>>>((const double*)a.o, (const double*)a.d_o, (const double*)a.l,
                                                (const double*)a.m, lse, dsum, p.batch, p.v_d, p.nq, p.d_q, z0, p.d_k, z1,
                                                p.d_v, z2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  {
    auto kern = bwd_dq_kernel<DP, VP, MT>;
    cudaError_t e = plan::ensure_smem(kern, BwdSmem<DP, VP>::kBytes);
    if (e != cudaSuccess) return e;
    p.n_tiles = int(tiles_q);
    ScopedKernel timed("bwd_dq_f64_dmma", stream);
    kern<<<dim3(unsigned(p.batch * p.n_tiles), ns_q), threads_of(MT), BwdSmem<DP, VP>::kBytes, stream>>>(p);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  {
    auto kern = bwd_dkdv_kernel<DP, VP, MT>;
    cudaError_t e = plan::ensure_smem(kern, BwdSmem<DP, VP>::kBytes);
    if (e != cudaSuccess) return e;
    p.n_tiles = int(tiles_k);
    ScopedKernel timed("bwd_dkdv_f64_dmma", stream);
    kern<<<dim3(unsigned(p.batch * p.n_tiles), ns_k), threads_of(MT), BwdSmem<DP, VP>::kBytes, stream>>>(p);
    return cudaGetLastError();
  }
}

}  // namespace f64

bool f64_dmma_forward_supports(const LaunchArgs& a) {
  if (a.dtype != 2 || a.accumulate || a.layout != 0) return false;
  if (a.d < 1 || a.v_d < 1 || a.d > 64 || a.v_d > 64) return false;
  const int64_t tiles = (int64_t(a.rule.q.total) + f64::kRows - 1) / f64::kRows;
  return tiles * a.batch <= 0x7fffffffLL;
}

bool f64_dmma_backward_supports(const LaunchArgs& a) {
  if (a.dtype != 2 || a.layout != 0) return false;
  if (a.d < 1 || a.v_d < 1 || a.d > 64 || a.v_d > 64) return false;
  if (!a.workspace || a.workspace_bytes < size_t(2) * size_t(a.batch) * size_t(a.rule.q.total) * 8) return false;
  const int64_t tq = (int64_t(a.rule.q.total) + f64::kRows - 1) / f64::kRows;
  const int64_t tk = (int64_t(a.rule.k.total) + f64::kRows - 1) / f64::kRows;
  return tq * a.batch <= 0x7fffffffLL && tk * a.batch <= 0x7fffffffLL;
}

cudaError_t f64_dmma_backward(const LaunchArgs& a, cudaStream_t stream) {
  const bool d_small = a.d <= 32, v_small = a.v_d <= 32;
  // 8 warps x one m8 tile each by default; fa_set_path_override(6): 4 warps x two m8 tiles (half the B-fragment loads
  // per DMMA, half the warps to hide latency with)
  if (a.variant == 6) {
    if (d_small && v_small) return f64::launch_bwd<32, 32, 2>(a, stream);
    if (d_small) return f64::launch_bwd<32, 64, 2>(a, stream);
    if (v_small) return f64::launch_bwd<64, 32, 2>(a, stream);
    return f64::launch_bwd<64, 64, 2>(a, stream);
  }
  if (d_small && v_small) return f64::launch_bwd<32, 32, 1>(a, stream);
  if (d_small) return f64::launch_bwd<32, 64, 1>(a, stream);
  if (v_small) return f64::launch_bwd<64, 32, 1>(a, stream);
  return f64::launch_bwd<64, 64, 1>(a, stream);
}

cudaError_t f64_dmma_forward(const LaunchArgs& a, cudaStream_t stream) {
  const bool d_small = a.d <= 32, v_small = a.v_d <= 32;
  // 8 warps x one m8 tile each by default; fa_set_path_override(6): 4 warps x two m8 tiles (half the B-fragment loads
  // per DMMA, half the warps to hide latency with)
  if (a.variant == 6) {
    if (d_small && v_small) return f64::launch_fwd<32, 32, 2>(a, stream);
    if (d_small) return f64::launch_fwd<32, 64, 2>(a, stream);
    if (v_small) return f64::launch_fwd<64, 32, 2>(a, stream);
    return f64::launch_fwd<64, 64, 2>(a, stream);
  }
  if (d_small && v_small) return f64::launch_fwd<32, 32, 1>(a, stream);
  if (d_small) return f64::launch_fwd<32, 64, 1>(a, stream);
  if (v_small) return f64::launch_fwd<64, 32, 1>(a, stream);
  return f64::launch_fwd<64, 64, 1>(a, stream);
}

}  // namespace fa
