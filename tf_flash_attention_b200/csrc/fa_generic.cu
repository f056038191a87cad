// fa_generic.cu — the any-shape / any-dtype kernel family (SIMT, fp32 or fp64 accumulate).
//
// This is the engine's universal path: it accepts every (q, k, d, v_d) the reference
// accepts up to 256 channels, every rule and sync mode, half/float/double, and is what
// the tcgen05 kernels fall back to for shapes TMA cannot address. It is a flash-attention-2
// style design (one CTA per Q tile, online softmax in registers, O written once) instead of
// the reference's K-outer loop with O/l/m read-modify-written in HBM under a spin lock
// (reference: flash_attention/kernel/flash_attention.cu:858-1076, 1725-1950).
//
// Semantics preserved from the reference:
//   scale 1/sqrt(d)                      flash_attention.cu:2162
//   padding rule q<size && k<size        flash_attention.cu:927
//   P = exp(s*scale - m) / l             flash_attention.cu:1838-1841
//   dS = P * (dP - D) * scale            flash_attention.cu:1544-1546
//   fully masked rows: O=0, l=0, m=0xFA… flash_attention_forward.cc:352-365
#include "fa_common.cuh"
#include "fa_launch.h"
#include "fa_plan.h"

namespace fa {
namespace generic {

constexpr int NT = 256;  // 16 x 16 thread grid

template <typename T>
struct FwdParams {
  const T* q; const T* k; const T* v;
  T* o; typename LOf<T>::type* l; T* m;
  int32_t d, v_d, nq, nk, n_rtiles;
  int64_t batch;
  int32_t accumulate;
  FaRule rule;
};

template <typename T>
struct BwdParams {
  const T* q; const T* k; const T* v; const T* d_o;
  T* d_q; T* d_k; T* d_v;
  const typename AccOf<T>::type* lse;   // [batch, nq]  m + log(l)  (+inf on empty rows)
  const typename AccOf<T>::type* dsum;  // [batch, nq]  rowsum(dO * O)
  int32_t d, v_d, nq, nk, n_rtiles;
  int64_t batch;
  FaRule rule;
};

// ---- tile primitives ---------------------------------------------------------------------
// Thread (ty, tx) of the 16x16 grid owns resident rows ty*RM..ty*RM+RM-1 (contiguous) and
// streamed columns tx + 16*j (interleaved -> conflict-free reads of the +1 padded tile).

template <typename A, int BM, int BN>
__device__ __forceinline__ void gemm_s(const A* __restrict__ R, const A* __restrict__ S, int C,
                                       A (&acc)[BM / 16][BN / 16], int ty, int tx) {
  constexpr int RM = BM / 16, CN = BN / 16;
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < CN; ++j) acc[i][j] = A(0);
  for (int c = 0; c < C; ++c) {
    A a[RM], b[CN];
#pragma unroll
    for (int i = 0; i < RM; ++i) a[i] = R[c * BM + ty * RM + i];
#pragma unroll
    for (int j = 0; j < CN; ++j) b[j] = S[c * (BN + 1) + tx + 16 * j];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int j = 0; j < CN; ++j) acc[i][j] += a[i] * b[j];
  }
}

// out[i][s] += sum_kk P[row_i][kk] * B[tx+16s][kk]   (B is [NC][BN+1])
template <typename A, int BM, int BN, int NS>
__device__ __forceinline__ void gemm_pv(const A* __restrict__ P, const A* __restrict__ B, int NC,
                                        A (&out)[BM / 16][NS], int ty, int tx) {
  constexpr int RM = BM / 16;
  const A* brow[NS];
  bool bval[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    int ch = tx + 16 * s;
    bval[s] = ch < NC;
    brow[s] = B + (bval[s] ? ch : 0) * (BN + 1);
  }
  for (int kk = 0; kk < BN; ++kk) {
    A pv[RM];
#pragma unroll
    for (int i = 0; i < RM; ++i) pv[i] = P[(ty * RM + i) * (BN + 1) + kk];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      A b = bval[s] ? brow[s][kk] : A(0);
#pragma unroll
      for (int i = 0; i < RM; ++i) out[i][s] += pv[i] * b;
    }
  }
}

// loads a [C][W] tile of a channel-first tensor (row pitch n) starting at column x0 into
// smem with row pitch `pitch`, zero-filling columns >= n. Coalesced along the sequence.
template <typename T, typename A>
__device__ __forceinline__ void load_tile(A* __restrict__ dst, const T* __restrict__ src, int C, int W,
                                          int pitch, int64_t n, int64_t x0) {
  const int total = C * W;
  for (int idx = threadIdx.x; idx < total; idx += NT) {
    int c = idx / W, x = idx - c * W;
    int64_t gx = x0 + x;
    dst[c * pitch + x] = gx < n ? to_acc<A>(src[int64_t(c) * n + gx]) : A(0);
  }
}

template <typename A>
__device__ __forceinline__ A half_warp_max(A v) {
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) v = acc_max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <typename A>
__device__ __forceinline__ A half_warp_sum(A v) {
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- forward -----------------------------------------------------------------------------
template <typename T, int BM, int BN, int NS>
__global__ void __launch_bounds__(NT) fwd_kernel(const FwdParams<T> p) {
  using A = typename AccOf<T>::type;
  using L = typename LOf<T>::type;
  constexpr int RM = BM / 16, CN = BN / 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  A* Qs = reinterpret_cast<A*>(smem_raw);      // [d][BM]
  A* Ks = Qs + p.d * BM;                       // [d][BN+1]
  A* Vs = Ks + p.d * (BN + 1);                 // [v_d][BN+1]  (also the O staging buffer)
  A* Ps = Vs + p.v_d * (BN + 1);               // [BM][BN+1]

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int rt = p.n_rtiles - 1 - int(blockIdx.x % p.n_rtiles);  // heavy (late) tiles first
  const int64_t b = blockIdx.x / p.n_rtiles;
  const int q0 = rt * BM;
  const int q_hi = min(q0 + BM, p.nq) - 1;
  const FaRule& rule = p.rule;
  const A scale = A(1) / sqrt(A(p.d));

  load_tile<T, A>(Qs, p.q + b * p.d * int64_t(p.nq), p.d, BM, BM, p.nq, q0);

  FaPos qpos[RM];
  A m_i[RM], l_i[RM], acc[RM][NS];
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    int qi = min(q0 + ty * RM + i, p.nq - 1);
    qpos[i] = fa_pos(rule, rule.q, qi);
    m_i[i] = neg_inf<A>();
    l_i[i] = A(0);
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[i][s] = A(0);
  }
  if (p.accumulate) {
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      int qi = q0 + ty * RM + i;
      if (qi < p.nq) {
        T mv = p.m[b * p.nq + qi];
        if (!is_sentinel<T>(mv)) {
          m_i[i] = to_acc<A>(mv);
          l_i[i] = A(p.l[b * p.nq + qi]);
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            int ch = tx + 16 * s;
            if (ch < p.v_d) acc[i][s] = to_acc<A>(p.o[(b * p.v_d + ch) * int64_t(p.nq) + qi]) * l_i[i];
          }
        }
      }
    }
  }

  int kt_first, kt_last;
  fa_k_tile_range(rule, q0, q_hi, BN, &kt_first, &kt_last);
  for (int kt = kt_first; kt <= kt_last; ++kt) {
    const int k0 = kt * BN;
    const int k_hi = min(k0 + BN, p.nk) - 1;
    const int cls = fa_classify(rule, q0, q_hi, k0, k_hi);
    if (cls == FA_TILE_SKIP) continue;
    __syncthreads();
    load_tile<T, A>(Ks, p.k + b * p.d * int64_t(p.nk), p.d, BN, BN + 1, p.nk, k0);
    load_tile<T, A>(Vs, p.v + b * p.v_d * int64_t(p.nk), p.v_d, BN, BN + 1, p.nk, k0);
    __syncthreads();

    A s[RM][CN];
    gemm_s<A, BM, BN>(Qs, Ks, p.d, s, ty, tx);

#pragma unroll
    for (int j = 0; j < CN; ++j) {
      const int kj = k0 + tx + 16 * j;
      const bool kvalid = kj < p.nk;
      const FaPos kpos = fa_pos(rule, rule.k, kvalid ? kj : p.nk - 1);
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        bool ok = kvalid && (cls == FA_TILE_FULL || fa_attend(rule, qpos[i], kpos));
        s[i][j] = ok ? s[i][j] * scale : neg_inf<A>();
      }
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      A mx = s[i][0];
#pragma unroll
      for (int j = 1; j < CN; ++j) mx = acc_max(mx, s[i][j]);
      mx = half_warp_max(mx);
      const A m_new = acc_max(m_i[i], mx);
      A alpha = A(1), sum = A(0);
      if (m_new == neg_inf<A>()) {
#pragma unroll
        for (int j = 0; j < CN; ++j) s[i][j] = A(0);
      } else {
        alpha = acc_exp(m_i[i] - m_new);  // exp(-inf) = 0 on the first live tile
#pragma unroll
        for (int j = 0; j < CN; ++j) {
          s[i][j] = acc_exp(s[i][j] - m_new);
          sum += s[i][j];
        }
      }
      sum = half_warp_sum(sum);
      l_i[i] = l_i[i] * alpha + sum;
      m_i[i] = m_new;
#pragma unroll
      for (int sidx = 0; sidx < NS; ++sidx) acc[i][sidx] *= alpha;
#pragma unroll
      for (int j = 0; j < CN; ++j) Ps[(ty * RM + i) * (BN + 1) + tx + 16 * j] = s[i][j];
    }
    __syncthreads();
    gemm_pv<A, BM, BN, NS>(Ps, Vs, p.v_d, acc, ty, tx);
  }

  // epilogue: O = acc / l staged through smem so that the store is coalesced along q
  __syncthreads();
  A* Os = Vs;  // [v_d][BM+1]  (BM == BN)
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const A inv = l_i[i] > A(0) ? A(1) / l_i[i] : A(0);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      int ch = tx + 16 * s;
      if (ch < p.v_d) Os[ch * (BM + 1) + ty * RM + i] = acc[i][s] * inv;
    }
    const int qi = q0 + ty * RM + i;
    if (tx == 0 && qi < p.nq) {
      if (l_i[i] > A(0)) {
        // m is stored in T; l is re-expressed against the rounded m so that
        // exp(s - m_T) / l_out reproduces the exact probabilities in the backward.
        const T m_t = from_acc<T>(m_i[i]);
        const A m_back = to_acc<A>(m_t);
        p.m[b * p.nq + qi] = m_t;
        p.l[b * p.nq + qi] = L(l_i[i] * acc_exp(m_i[i] - m_back));
      } else {
        p.m[b * p.nq + qi] = sentinel<T>();
        p.l[b * p.nq + qi] = L(0);
      }
    }
  }
  __syncthreads();
  T* og = p.o + b * p.v_d * int64_t(p.nq);
  const int total = p.v_d * BM;
  for (int idx = threadIdx.x; idx < total; idx += NT) {
    int ch = idx / BM, r = idx - ch * BM;
    if (q0 + r < p.nq) og[int64_t(ch) * p.nq + q0 + r] = from_acc<T>(Os[ch * (BM + 1) + r]);
  }
}

// ---- backward preprocess: LSE and D ------------------------------------------------------
template <typename T>
__global__ void bwd_prep_kernel(const T* __restrict__ o, const T* __restrict__ d_o,
                                const typename LOf<T>::type* __restrict__ l, const T* __restrict__ m,
                                typename AccOf<T>::type* __restrict__ lse,
                                typename AccOf<T>::type* __restrict__ dsum, int64_t batch, int32_t v_d,
                                int32_t nq) {
  using A = typename AccOf<T>::type;
  const int64_t total = batch * nq;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total;
       i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t b = i / nq, r = i - b * nq;
    A acc = A(0);
    const T* op = o + b * v_d * int64_t(nq) + r;
    const T* dp = d_o + b * v_d * int64_t(nq) + r;
    for (int c = 0; c < v_d; ++c) acc += to_acc<A>(op[int64_t(c) * nq]) * to_acc<A>(dp[int64_t(c) * nq]);
    dsum[i] = acc;
    const A lv = A(l[i]);
    const T mv = m[i];
    lse[i] = (lv > A(0) && !is_sentinel<T>(mv)) ? to_acc<A>(mv) + acc_log(lv) : -neg_inf<A>();
  }
}

// ---- backward dQ: one CTA per Q tile, streams K/V tiles ------------------------------------
template <typename T, int BM, int BN, int NS>
__global__ void __launch_bounds__(NT) bwd_dq_kernel(const BwdParams<T> p) {
  using A = typename AccOf<T>::type;
  constexpr int RM = BM / 16, CN = BN / 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  A* Qs = reinterpret_cast<A*>(smem_raw);  // [d][BM]
  A* dOs = Qs + p.d * BM;                  // [v_d][BM]
  A* Ks = dOs + p.v_d * BM;                // [d][BN+1]   (also dQ staging)
  A* Vs = Ks + p.d * (BN + 1);             // [v_d][BN+1]
  A* Ds = Vs + p.v_d * (BN + 1);           // [BM][BN+1]  dS tile

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int rt = p.n_rtiles - 1 - int(blockIdx.x % p.n_rtiles);
  const int64_t b = blockIdx.x / p.n_rtiles;
  const int q0 = rt * BM;
  const int q_hi = min(q0 + BM, p.nq) - 1;
  const FaRule& rule = p.rule;
  const A scale = A(1) / sqrt(A(p.d));

  load_tile<T, A>(Qs, p.q + b * p.d * int64_t(p.nq), p.d, BM, BM, p.nq, q0);
  load_tile<T, A>(dOs, p.d_o + b * p.v_d * int64_t(p.nq), p.v_d, BM, BM, p.nq, q0);

  FaPos qpos[RM];
  A lse[RM], dsum[RM], dq[RM][NS];
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    int qi = min(q0 + ty * RM + i, p.nq - 1);
    qpos[i] = fa_pos(rule, rule.q, qi);
    lse[i] = p.lse[b * p.nq + qi];
    dsum[i] = p.dsum[b * p.nq + qi];
#pragma unroll
    for (int s = 0; s < NS; ++s) dq[i][s] = A(0);
  }

  int kt_first, kt_last;
  fa_k_tile_range(rule, q0, q_hi, BN, &kt_first, &kt_last);
  for (int kt = kt_first; kt <= kt_last; ++kt) {
    const int k0 = kt * BN;
    const int k_hi = min(k0 + BN, p.nk) - 1;
    const int cls = fa_classify(rule, q0, q_hi, k0, k_hi);
    if (cls == FA_TILE_SKIP) continue;
    __syncthreads();
    load_tile<T, A>(Ks, p.k + b * p.d * int64_t(p.nk), p.d, BN, BN + 1, p.nk, k0);
    load_tile<T, A>(Vs, p.v + b * p.v_d * int64_t(p.nk), p.v_d, BN, BN + 1, p.nk, k0);
    __syncthreads();
    A s[RM][CN], dp[RM][CN];
    gemm_s<A, BM, BN>(Qs, Ks, p.d, s, ty, tx);
    gemm_s<A, BM, BN>(dOs, Vs, p.v_d, dp, ty, tx);
#pragma unroll
    for (int j = 0; j < CN; ++j) {
      const int kj = k0 + tx + 16 * j;
      const bool kvalid = kj < p.nk;
      const FaPos kpos = fa_pos(rule, rule.k, kvalid ? kj : p.nk - 1);
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        bool ok = kvalid && (cls == FA_TILE_FULL || fa_attend(rule, qpos[i], kpos));
        A pr = ok ? acc_exp(s[i][j] * scale - lse[i]) : A(0);
        Ds[(ty * RM + i) * (BN + 1) + tx + 16 * j] = pr * (dp[i][j] - dsum[i]) * scale;
      }
    }
    __syncthreads();
    gemm_pv<A, BM, BN, NS>(Ds, Ks, p.d, dq, ty, tx);
  }
  __syncthreads();
  A* Os = Ks;  // [d][BM+1]
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      int ch = tx + 16 * s;
      if (ch < p.d) Os[ch * (BM + 1) + ty * RM + i] = dq[i][s];
    }
  __syncthreads();
  T* og = p.d_q + b * p.d * int64_t(p.nq);
  const int total = p.d * BM;
  for (int idx = threadIdx.x; idx < total; idx += NT) {
    int ch = idx / BM, r = idx - ch * BM;
    if (q0 + r < p.nq) og[int64_t(ch) * p.nq + q0 + r] = from_acc<T>(Os[ch * (BM + 1) + r]);
  }
}

// ---- backward dK/dV: one CTA per K tile (resident), streams Q/dO tiles ---------------------
template <typename T, int BM, int BN, int NS>
__global__ void __launch_bounds__(NT) bwd_dkdv_kernel(const BwdParams<T> p) {
  using A = typename AccOf<T>::type;
  constexpr int RM = BM / 16, CN = BN / 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  A* Ks = reinterpret_cast<A*>(smem_raw);  // [d][BM]     resident keys
  A* Vs = Ks + p.d * BM;                   // [v_d][BM]
  A* Qs = Vs + p.v_d * BM;                 // [d][BN+1]   streamed queries (also dK staging)
  A* dOs = Qs + p.d * (BN + 1);            // [v_d][BN+1] (also dV staging)
  A* Ps = dOs + p.v_d * (BN + 1);          // [BM][BN+1]  P^T
  A* Ds = Ps + BM * (BN + 1);              // [BM][BN+1]  dS^T
  A* lse_s = Ds + BM * (BN + 1);           // [BN]
  A* dsum_s = lse_s + BN;                  // [BN]

  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int rt = int(blockIdx.x % p.n_rtiles);  // early key tiles are the heavy ones under causal
  const int64_t b = blockIdx.x / p.n_rtiles;
  const int k0 = rt * BM;
  const int k_hi = min(k0 + BM, p.nk) - 1;
  const FaRule& rule = p.rule;
  const A scale = A(1) / sqrt(A(p.d));

  load_tile<T, A>(Ks, p.k + b * p.d * int64_t(p.nk), p.d, BM, BM, p.nk, k0);
  load_tile<T, A>(Vs, p.v + b * p.v_d * int64_t(p.nk), p.v_d, BM, BM, p.nk, k0);

  FaPos kpos[RM];
  bool kvalid[RM];
  A dk[RM][NS], dv[RM][NS];
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    int ki = k0 + ty * RM + i;
    kvalid[i] = ki < p.nk;
    kpos[i] = fa_pos(rule, rule.k, min(ki, p.nk - 1));
#pragma unroll
    for (int s = 0; s < NS; ++s) dk[i][s] = dv[i][s] = A(0);
  }

  int qt_first, qt_last;
  fa_q_tile_range(rule, k0, k_hi, BN, &qt_first, &qt_last);
  for (int qt = qt_first; qt <= qt_last; ++qt) {
    const int q0 = qt * BN;
    const int q_hi = min(q0 + BN, p.nq) - 1;
    const int cls = fa_classify(rule, q0, q_hi, k0, k_hi);
    if (cls == FA_TILE_SKIP) continue;
    __syncthreads();
    load_tile<T, A>(Qs, p.q + b * p.d * int64_t(p.nq), p.d, BN, BN + 1, p.nq, q0);
    load_tile<T, A>(dOs, p.d_o + b * p.v_d * int64_t(p.nq), p.v_d, BN, BN + 1, p.nq, q0);
    if (threadIdx.x < BN) {
      int qi = q0 + threadIdx.x;
      bool v = qi < p.nq;
      lse_s[threadIdx.x] = v ? p.lse[b * p.nq + qi] : -neg_inf<A>();
      dsum_s[threadIdx.x] = v ? p.dsum[b * p.nq + qi] : A(0);
    }
    __syncthreads();
    A st[RM][CN], dpt[RM][CN];
    gemm_s<A, BM, BN>(Ks, Qs, p.d, st, ty, tx);
    gemm_s<A, BM, BN>(Vs, dOs, p.v_d, dpt, ty, tx);
#pragma unroll
    for (int j = 0; j < CN; ++j) {
      const int col = tx + 16 * j;
      const int qj = q0 + col;
      const bool qvalid = qj < p.nq;
      const FaPos qpos = fa_pos(rule, rule.q, qvalid ? qj : p.nq - 1);
      const A lse = lse_s[col], dsum = dsum_s[col];
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        bool ok = qvalid && kvalid[i] && (cls == FA_TILE_FULL || fa_attend(rule, qpos, kpos[i]));
        A pr = ok ? acc_exp(st[i][j] * scale - lse) : A(0);
        Ps[(ty * RM + i) * (BN + 1) + col] = pr;
        Ds[(ty * RM + i) * (BN + 1) + col] = pr * (dpt[i][j] - dsum) * scale;
      }
    }
    __syncthreads();
    gemm_pv<A, BM, BN, NS>(Ps, dOs, p.v_d, dv, ty, tx);
    gemm_pv<A, BM, BN, NS>(Ds, Qs, p.d, dk, ty, tx);
  }
  __syncthreads();
  A* Ok = Qs;   // [d][BM+1]
  A* Ov = dOs;  // [v_d][BM+1]
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      int ch = tx + 16 * s;
      if (ch < p.d) Ok[ch * (BM + 1) + ty * RM + i] = dk[i][s];
      if (ch < p.v_d) Ov[ch * (BM + 1) + ty * RM + i] = dv[i][s];
    }
  __syncthreads();
  T* gk = p.d_k + b * p.d * int64_t(p.nk);
  T* gv = p.d_v + b * p.v_d * int64_t(p.nk);
  for (int idx = threadIdx.x; idx < p.d * BM; idx += NT) {
    int ch = idx / BM, r = idx - ch * BM;
    if (k0 + r < p.nk) gk[int64_t(ch) * p.nk + k0 + r] = from_acc<T>(Ok[ch * (BM + 1) + r]);
  }
  for (int idx = threadIdx.x; idx < p.v_d * BM; idx += NT) {
    int ch = idx / BM, r = idx - ch * BM;
    if (k0 + r < p.nk) gv[int64_t(ch) * p.nk + k0 + r] = from_acc<T>(Ov[ch * (BM + 1) + r]);
  }
}

// ---- host-side launch helpers --------------------------------------------------------------
template <typename T>
static size_t fwd_smem(int d, int v_d, int BM) {
  using A = typename AccOf<T>::type;
  return sizeof(A) * (size_t(d) * BM + size_t(d) * (BM + 1) + size_t(v_d) * (BM + 1) + size_t(BM) * (BM + 1));
}
template <typename T>
static size_t dq_smem(int d, int v_d, int BM) {
  using A = typename AccOf<T>::type;
  return sizeof(A) * (size_t(d + v_d) * BM + size_t(d + v_d) * (BM + 1) + size_t(BM) * (BM + 1));
}
template <typename T>
static size_t dkdv_smem(int d, int v_d, int BM) {
  using A = typename AccOf<T>::type;
  return sizeof(A) * (size_t(d + v_d) * BM + size_t(d + v_d) * (BM + 1) + 2 * size_t(BM) * (BM + 1) + 2 * BM);
}

template <typename K, typename P>
static cudaError_t launch(const char* name, K kernel, const P& params, int64_t grid, size_t smem,
                          cudaStream_t stream) {
  cudaError_t e = plan::ensure_smem(kernel, int(smem));
  if (e != cudaSuccess) return e;
  ScopedKernel timed(name, stream);
  kernel<<<unsigned(grid), NT, smem, stream>>>(params);
  return cudaGetLastError();
}

constexpr size_t kSmemMax = 227 * 1024;

// Picks (BM, NS). Big tiles when channels <= 128, small tiles up to 256 channels.
template <typename T>
struct Tiles {
  static constexpr int kBig = sizeof(typename AccOf<T>::type) == 8 ? 32 : 64;
  static constexpr int kSmall = kBig / 2;
};

template <typename T>
cudaError_t forward_t(const LaunchArgs& a, cudaStream_t stream) {
  FwdParams<T> p;
  p.q = (const T*)a.q; p.k = (const T*)a.k; p.v = (const T*)a.v;
  p.o = (T*)a.o; p.l = (typename LOf<T>::type*)a.l; p.m = (T*)a.m;
  p.d = a.d; p.v_d = a.v_d; p.nq = a.rule.q.total; p.nk = a.rule.k.total;
  p.batch = a.batch; p.accumulate = a.accumulate; p.rule = a.rule;
  constexpr int BIG = Tiles<T>::kBig, SMALL = Tiles<T>::kSmall;
  const int ns = (a.v_d + 15) / 16;
#define FA_FWD(BM_, NS_)                                                                   \
  do {                                                                                     \
    p.n_rtiles = (p.nq + BM_ - 1) / BM_;                                                   \
    return launch("generic_fwd", fwd_kernel<T, BM_, BM_, NS_>, p, p.batch * p.n_rtiles,                   \
                  fwd_smem<T>(a.d, a.v_d, BM_), stream);                                   \
  } while (0)
  // fp64, up to 64 channels: 64 x 64 tiles (4 x 4 accumulators per thread: one shared-memory load per two DFMAs instead
  // of one per DFMA with the 32 x 32 tiles; C4 fp64 and the reference's fp64 benchmark shapes)
  if (sizeof(typename AccOf<T>::type) == 8 && ns <= 4 && a.d <= 64 && fwd_smem<T>(a.d, a.v_d, 64) <= kSmemMax)
    FA_FWD(64, 4);
  if (ns <= 2 && fwd_smem<T>(a.d, a.v_d, BIG) <= kSmemMax) FA_FWD(BIG, 2);
  if (ns <= 8 && fwd_smem<T>(a.d, a.v_d, BIG) <= kSmemMax) FA_FWD(BIG, 8);
  if (ns <= 16 && fwd_smem<T>(a.d, a.v_d, SMALL) <= kSmemMax) FA_FWD(SMALL, 16);
#undef FA_FWD
  return cudaErrorInvalidValue;
}

template <typename T>
cudaError_t backward_t(const LaunchArgs& a, cudaStream_t stream) {
  using A = typename AccOf<T>::type;
  BwdParams<T> p;
  p.q = (const T*)a.q; p.k = (const T*)a.k; p.v = (const T*)a.v; p.d_o = (const T*)a.d_o;
  p.d_q = (T*)a.d_q; p.d_k = (T*)a.d_k; p.d_v = (T*)a.d_v;
  p.d = a.d; p.v_d = a.v_d; p.nq = a.rule.q.total; p.nk = a.rule.k.total;
  p.batch = a.batch; p.rule = a.rule;
  A* lse = (A*)a.workspace;
  A* dsum = lse + p.batch * p.nq;
  p.lse = lse; p.dsum = dsum;
  {
    int64_t total = p.batch * p.nq;
    int blocks = int(std::min<int64_t>((total + 255) / 256, 148 * 8));
    cudaError_t e;
    {
      ScopedKernel timed("bwd_prep", stream);
      bwd_prep_kernel<T><<<blocks, 256, 0, stream>>>((const T*)a.o, (const T*)a.d_o,
                                                     (const typename LOf<T>::type*)a.l, (const T*)a.m,
                                                     lse, dsum, p.batch, p.v_d, p.nq);
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) return e;
  }
  constexpr int BIG = Tiles<T>::kBig, SMALL = Tiles<T>::kSmall;
  const int ns = (std::max(a.d, a.v_d) + 15) / 16;
#define FA_BWD(BM_, NS_)                                                                      \
  do {                                                                                        \
    p.n_rtiles = (p.nq + BM_ - 1) / BM_;                                                      \
    cudaError_t e = launch("generic_bwd_dq", bwd_dq_kernel<T, BM_, BM_, NS_>, p, p.batch * p.n_rtiles,          \
                           dq_smem<T>(a.d, a.v_d, BM_), stream);                              \
    if (e != cudaSuccess) return e;                                                           \
    p.n_rtiles = (p.nk + BM_ - 1) / BM_;                                                      \
    return launch("generic_bwd_dkdv", bwd_dkdv_kernel<T, BM_, BM_, NS_>, p, p.batch * p.n_rtiles,                 \
                  dkdv_smem<T>(a.d, a.v_d, BM_), stream);                                     \
  } while (0)
  if (sizeof(A) == 8 && ns <= 4 && dkdv_smem<T>(a.d, a.v_d, 64) <= kSmemMax) FA_BWD(64, 4);
  if (ns <= 2 && dkdv_smem<T>(a.d, a.v_d, BIG) <= kSmemMax) FA_BWD(BIG, 2);
  if (ns <= 8 && dkdv_smem<T>(a.d, a.v_d, BIG) <= kSmemMax) FA_BWD(BIG, 8);
  if (ns <= 16 && dkdv_smem<T>(a.d, a.v_d, SMALL) <= kSmemMax) FA_BWD(SMALL, 16);
#undef FA_BWD
  return cudaErrorInvalidValue;
}

}  // namespace generic

bool generic_supports(const LaunchArgs& a) { return a.layout == 0 && a.d <= 256 && a.v_d <= 256; }

size_t generic_workspace_bytes(int dtype, int64_t batch, int64_t nq, bool backward) {
  if (!backward) return 0;
  return size_t(2) * size_t(batch) * size_t(nq) * (dtype == 2 ? 8 : 4);
}

cudaError_t generic_forward(const LaunchArgs& a, cudaStream_t stream) {
  switch (a.dtype) {
    case 0: return generic::forward_t<__half>(a, stream);
    case 1: return generic::forward_t<float>(a, stream);
    default: return generic::forward_t<double>(a, stream);
  }
}

cudaError_t generic_backward(const LaunchArgs& a, cudaStream_t stream) {
  switch (a.dtype) {
    case 0: return generic::backward_t<__half>(a, stream);
    case 1: return generic::backward_t<float>(a, stream);
    default: return generic::backward_t<double>(a, stream);
  }
}

}  // namespace fa
