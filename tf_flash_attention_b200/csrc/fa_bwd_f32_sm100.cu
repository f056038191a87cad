// fa_bwd_f32_sm100.cu — fp32 backward on the tensor cores: split-precision tcgen05 kernels instantiated for head dims
// 64/64, 32/32, 32/16; any channel counts up to 64 and any lengths run on them (the split pass pads).
//
// The forward runs fp32 as 3xTF32 (fa_fwd_f32_sm100.cu). The backward needs every streamed tile in BOTH operand
// orientations (K as the MN-major B operand of S = Q K^T and as the K-major B operand of dQ = dS K, Q likewise in the
// dK/dV kernel). For kind::tf32 that costs two shared-memory copies per tile: MN-major TF32 operands only work in the
// "128B swizzle, 32-byte atom" layout and a K-major descriptor over such a tile faults (measured), and
// with hi/lo copies of everything that does not fit. The same split idea with 16-bit pieces does fit:
//     x = x0 + x1 + x2,  x_i = bf16 pieces (8 + 8 + 8 mantissa bits, fp32 exponent range),
//     a*b ~= a0b0 + (a0b1 + a1b0) + (a0b2 + a1b1 + a2b0)          (dropped terms <= 2^-24 |ab|)
// = 6 kind::f16 (bf16) MMAs per product at full rate, which is the cost of 3 half-rate TF32 MMAs, 6 bytes per
// element instead of 8, and one 128B-swizzled tile per piece that is readable MN-major and K-major exactly like the
// fp16 kernels' tiles. Accumulation is fp32 in TMEM; softmax statistics, exp2f and dS are fp32 in registers; P and dS
// are re-split into three bf16 pieces for the second product of each chain.
//
//   split_bf16x3_all    : Q, K, V, dO -> three bf16 tensors each (workspace), one launch
//   bwd_prep_f32        : LSE2 = (m + log l) log2 e, D = rowsum(dO o O)
//   bwd_dq_f32_kernel   : CTA = 128 query rows, streams 64-key tiles:  S, dP (SS) -> dS pieces (TMEM) -> dQ += dS K (TS)
//   bwd_dkdv_f32_kernel : CTA = 128 keys, streams 64-query tiles: S^T, dP^T (SS) -> P^T, dS^T pieces -> dV, dK (TS)
// Formulas as in the reference (flash_attention.cu:1838-1841, 1544-1546, 1882-1891).
#include <cuda_bf16.h>

#include "fa_common.cuh"
#include "fa_launch.h"
#include "fa_plan.h"
#include "sm100_ptx.cuh"
#include "sm100_tiles.cuh"

namespace fa {
namespace sm100 {

using namespace ptx;

bool make_map_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_cols, int box_rows,
                 bool swizzle128);  // fa_fwd_f16_sm100.cu (2-byte elements; the copy does not interpret them)

constexpr int kXM = 128;        // resident rows of a CTA (TMEM lanes)
constexpr int kXN = 64;         // streamed tile width
constexpr int kXThreads = 384;  // 2 softmax warpgroups (each takes 32 of a tile's 64 columns), TMA / MMA / TMEM warps, spare
constexpr int kXStages = 2;
constexpr int kXStatPad = 64;
constexpr float kXLog2e = 1.4426950408889634f;

// kind::f16 instruction descriptor with BF16 inputs (a/b format 1) and F32 accumulation
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) |
         (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}

// piece pairs of a 3-piece product, smallest terms first (the tensor core truncates when it adds into fp32)
__device__ constexpr int kPairA[6] = {0, 1, 2, 0, 1, 0};
__device__ constexpr int kPairB[6] = {2, 1, 0, 1, 0, 0};

struct alignas(64) BwdF32Params {
  CUtensorMap map_q[3], map_k[3], map_v[3], map_do[3];   // bf16 pieces, 128B swizzle, box 64 x 64 channels
  FaRule rule;
  const float* lse2;    // from the forward's l, m
  const float* dsum;
  float* lse2_refined;  // written by the dQ kernel, read by the dK/dV kernel (see bwd_dq_f32_kernel)
  float *d_q, *d_k, *d_v;
  int32_t nq, nk, n_blocks, batch;
  int32_t d, v_d;        // the tensors' own channel counts (<= the kernel's D, VD; the pieces are zero-padded up to those)
  int32_t stat_pitch;    // floats per batch element in lse2 / dsum / lse2_refined (nq rounded up to 4: 16-byte rows)
  int32_t renormalise;   // 0 when the call sees only a shard of the keys (ring): sum_k P != 1 by construction there
  float scale, scale_log2;
};

// ---- operand split and row statistics --------------------------------------------------------------------
struct PrepJob {
  const float *o, *d_o, *l, *m;
  float *lse2, *dsum, *lse2_refined;
  int64_t batch;
  int32_t v_d, nq, sp;
};
__device__ __forceinline__ void bwd_prep_f32(const float* __restrict__ o, const float* __restrict__ d_o,
                                             const float* __restrict__ l, const float* __restrict__ m,
                                             float* __restrict__ lse2, float* __restrict__ dsum,
                                             float* __restrict__ lse2_refined, int64_t batch, int32_t v_d, int32_t nq,
                                             int32_t sp) {
  const int64_t total = batch * sp;
  // the padding behind the arrays (and behind each batch element's row when nq is not a multiple of 4) is read by the
  // 64-wide bulk copies of the last, ragged query tile: keep it finite
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x < kXStatPad) {
    lse2[total + threadIdx.x] = __int_as_float(0x7f800000);
    lse2_refined[total + threadIdx.x] = __int_as_float(0x7f800000);
    dsum[total + threadIdx.x] = 0.f;
  }
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t b = i / sp, r = i - b * sp;
    if (r >= nq) {
      lse2[i] = lse2_refined[i] = __int_as_float(0x7f800000);
      dsum[i] = 0.f;
      continue;
    }
    const float* op = o + b * v_d * int64_t(nq) + r;
    const float* dp = d_o + b * v_d * int64_t(nq) + r;
    float acc = 0.f;
    for (int c = 0; c < v_d; ++c) acc = fmaf(op[int64_t(c) * nq], dp[int64_t(c) * nq], acc);
    dsum[i] = acc;
    const float lv = l[b * nq + r], mv = m[b * nq + r];
    lse2[i] = (lv > 0.f && !is_sentinel<float>(mv)) ? (mv + logf(lv)) * kXLog2e : __int_as_float(0x7f800000);
  }
}

// three bf16 pieces of two fp32 values, packed pairwise: out[j] = {piece_j(x), piece_j(y)}
__device__ __forceinline__ void split3_pack(float x, float y, uint32_t& o0, uint32_t& o1, uint32_t& o2) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(x, y);
  const float rx = x - __bfloat162float(a.x), ry = y - __bfloat162float(a.y);
  const __nv_bfloat162 b = __floats2bfloat162_rn(rx, ry);
  const __nv_bfloat162 c = __floats2bfloat162_rn(rx - __bfloat162float(b.x), ry - __bfloat162float(b.y));
  o0 = *reinterpret_cast<const uint32_t*>(&a);
  o1 = *reinterpret_cast<const uint32_t*>(&b);
  o2 = *reinterpret_cast<const uint32_t*>(&c);
}

__device__ __forceinline__ void bulk_load_1d_x(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_dst),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(bar)
               : "memory");
}

// D = channels of Q / K, VD = channels of V / dO. One bf16 piece of a tile is [channels][64 positions] boxes of
// 128-byte rows: a resident 128-position tile is two boxes, a streamed tile one.
template <int D, int VD>
struct XCfg {
  static constexpr int kResD = kXM * D * 2;                 // resident piece with D channels (Q or K): 16 KB at D = 64
  static constexpr int kResV = kXM * VD * 2;                // resident piece with VD channels (dO or V)
  static constexpr int kStrD = kXN * D * 2;                 // streamed piece with D channels (K or Q)
  static constexpr int kStrV = kXN * VD * 2;                // streamed piece with VD channels (V or dO)
  static constexpr int kResBytes = 3 * (kResD + kResV);     // two resident tensors x three pieces
  static constexpr int kStageBytes = 3 * (kStrD + kStrV);   // two streamed tensors x three pieces
  static constexpr int kStatBytes = 2 * kXN * 4;            // LSE2[64] + D[64] per stage (dK/dV kernel)
  static constexpr int kRingOffset = kResBytes;
  static constexpr int kStatOffset = kRingOffset + kXStages * kStageBytes;
  static constexpr int kBarOffset = kStatOffset + kXStages * kStatBytes;
  static constexpr int kNumBars = 1 + 2 * kXStages + 1 + 1 + 1;
  static constexpr int kSchedOffset = kBarOffset + kNumBars * 8 + 16;
  static constexpr int kSmemBytes = kSchedOffset + int(sizeof(TileSchedule)) + 1024;
};

// issues the 6 piece products of one 128 x 64 x CH SS GEMM (both operands MN-major, CH = contraction length = channels
// of the tiles), small terms first
template <int CH>
__device__ __forceinline__ void issue_ss6(uint32_t d_tmem, uint32_t a_pieces, uint32_t a_stride, uint32_t b_pieces,
                                          uint32_t b_stride, uint32_t idesc) {
#pragma unroll
  for (int t = 0; t < 6; ++t)
#pragma unroll
    for (int ks = 0; ks < CH / 16; ++ks)
      mma_ss(d_tmem, smem_desc_sw128(a_pieces + kPairA[t] * a_stride + ks * 2048, CH * 128, 1024),
             smem_desc_sw128(b_pieces + kPairB[t] * b_stride + ks * 2048, CH * 128, 1024), idesc, (t | ks) != 0);
}
// 6 piece products of a TS GEMM: A pieces in TMEM (32 columns each), B pieces K-major in shared memory.
// d_corr == d_main: everything into one accumulator. Otherwise the leading product a0*b0 goes to d_main and the five
// small ones to d_corr: an accumulator that lives across many tiles then takes 6x fewer truncating additions.
__device__ __forceinline__ void issue_ts6(uint32_t d_main, uint32_t d_corr, uint32_t a_tmem, uint32_t b_pieces,
                                          uint32_t b_stride, uint32_t idesc, bool accumulate) {
  const bool split = d_main != d_corr;
#pragma unroll
  for (int t = 0; t < 6; ++t)
#pragma unroll
    for (int ks = 0; ks < kXN / 16; ++ks) {
      const bool lead = t == 5;
      const bool first = split ? (lead ? ks == 0 : (t | ks) == 0) : (t | ks) == 0;
      mma_ts(lead ? d_main : d_corr, a_tmem + kPairA[t] * 32 + ks * 8,
             smem_desc_sw128(b_pieces + kPairB[t] * b_stride + ks * 32, 16, 1024), idesc,
             (accumulate || !first) ? 1u : 0u);
    }
}

// =================================================================================================
// dQ kernel
// =================================================================================================
template <int D, int VD>
__global__ void __launch_bounds__(kXThreads, 1) bwd_dq_f32_kernel(const __grid_constant__ BwdF32Params p) {
  using Cfg = XCfg<D, VD>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t q_smem = smem_base;                           // [3][kResD]
  const uint32_t do_smem = smem_base + 3 * Cfg::kResD;         // [3][kResV]
  const uint32_t ring = smem_base + Cfg::kRingOffset;          // stage: K pieces [3][kStrD], V pieces [3][kStrV]
  const uint32_t bars = smem_base + Cfg::kBarOffset;
  const uint32_t bar_q_full = bars;
  const uint32_t bar_kv_full = bars + 8;
  const uint32_t bar_kv_empty = bar_kv_full + 8 * kXStages;
  const uint32_t bar_s_full = bar_kv_empty + 8 * kXStages;
  const uint32_t bar_p_ready = bar_s_full + 8;
  const uint32_t bar_final = bar_p_ready + 8;
  const uint32_t tmem_slot = bar_final + 8;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::kBarOffset + Cfg::kNumBars * 8);

  const int warp = threadIdx.x >> 5;
  const FaRule& rule = p.rule;
  const int b = int(blockIdx.x / p.n_blocks);
  const int blk = p.n_blocks - 1 - int(blockIdx.x % p.n_blocks);
  const int q0 = blk * kXM;
  const int q_hi = min(q0 + kXM, p.nq) - 1;
  int kt_first, kt_last;
  fa_k_tile_range(rule, q0, q_hi, kXN, &kt_first, &kt_last);
  TileSchedule* sched = reinterpret_cast<TileSchedule*>(smem_gen + Cfg::kSchedOffset);
  {
    const int lo[1] = {q0};
    const int hi[1] = {q_hi};
    const bool valid[1] = {true};
    build_schedule(sched, rule, true, lo, hi, valid, 1, kt_first, kt_last, kXN, p.nk, kXThreads / 32);
  }
  if (warp == 8) {
    if (elect_one())
      for (int j = 0; j < 3; ++j) {
        prefetch_tensormap(&p.map_q[j]);
        prefetch_tensormap(&p.map_k[j]);
        prefetch_tensormap(&p.map_v[j]);
        prefetch_tensormap(&p.map_do[j]);
      }
  } else if (warp == 9) {
    if (elect_one()) {
      mbar_init(bar_q_full, 1);
      mbar_init(bar_s_full, 1);
      mbar_init(bar_p_ready, 2 * kXM);
      mbar_init(bar_final, 1);
      for (int s = 0; s < kXStages; ++s) {
        mbar_init(bar_kv_full + 8 * s, 1);
        mbar_init(bar_kv_empty + 8 * s, 1);
      }
      fence_barrier_init();
    }
  } else if (warp == 10) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // TMEM columns: S [0, 64)  dP [64, 128)  dQ [128, 192)  dQ small terms [192, 256)  dS pieces [256 + 32 j, +32)
  constexpr uint32_t kColDp = 64, kColDq = 128, kColDqc = 192, kColDs = 256;

  if (warp >= 8) {
    setmaxnreg_dec<56>();
    if (warp == 8) {
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_q_full, Cfg::kResBytes);
        for (int j = 0; j < 3; ++j)
          for (int h = 0; h < 2; ++h) {
            tma_load_2d(q_smem + j * Cfg::kResD + h * (D * 128), &p.map_q[j], bar_q_full, q0 + h * 64, b * D);
            tma_load_2d(do_smem + j * Cfg::kResV + h * (VD * 128), &p.map_do[j], bar_q_full, q0 + h * 64, b * VD);
          }
        int t = 0;
        TileIter it;
        it.init(sched, 1, kt_first, kt_last);
        int kt, tw, tb;
        while (it.next(&kt, &tw, &tb)) {
          const int s = t % kXStages, u = t / kXStages;
          mbar_wait(bar_kv_empty + 8 * s, (u & 1) ^ 1);
          mbar_arrive_expect_tx(bar_kv_full + 8 * s, Cfg::kStageBytes);
          for (int j = 0; j < 3; ++j) {
            tma_load_2d(ring + s * Cfg::kStageBytes + j * Cfg::kStrD, &p.map_k[j], bar_kv_full + 8 * s, kt * kXN,
                        b * D);
            tma_load_2d(ring + s * Cfg::kStageBytes + 3 * Cfg::kStrD + j * Cfg::kStrV, &p.map_v[j],
                        bar_kv_full + 8 * s, kt * kXN, b * VD);
          }
          ++t;
        }
      }
    } else if (warp == 9) {
      if (elect_one()) {
        TileIter it;
        it.init(sched, 1, kt_first, kt_last);
        const int n = it.count();
        constexpr uint32_t idesc_s = idesc_bf16(kXM, kXN, true, true);
        constexpr uint32_t idesc_dq = idesc_bf16(kXM, D, false, false);
        auto issue_s_dp = [&](int stage) {
          const uint32_t k_s = ring + stage * Cfg::kStageBytes, v_s = k_s + 3 * Cfg::kStrD;
          issue_ss6<D>(tmem_base, q_smem, Cfg::kResD, k_s, Cfg::kStrD, idesc_s);
          issue_ss6<VD>(tmem_base + kColDp, do_smem, Cfg::kResV, v_s, Cfg::kStrV, idesc_s);
        };
        if (n > 0) {
          mbar_wait(bar_kv_full + 0, 0);
          mbar_wait(bar_q_full, 0);
          tc_fence_after();
          issue_s_dp(0);
          mma_commit(bar_s_full);
          for (int j = 0; j < n; ++j) {
            const int sj = j % kXStages, sn = (j + 1) % kXStages;
            mbar_wait(bar_p_ready, j & 1);
            tc_fence_after();
            issue_ts6(tmem_base + kColDq, tmem_base + kColDqc, tmem_base + kColDs, ring + sj * Cfg::kStageBytes,
                      Cfg::kStrD, idesc_dq, j > 0);
            mma_commit(bar_kv_empty + 8 * sj);
            if (j + 1 < n) {
              mbar_wait(bar_kv_full + 8 * sn, ((j + 1) / kXStages) & 1);
              tc_fence_after();
              issue_s_dp(sn);
              mma_commit(bar_s_full);
            } else {
              mma_commit(bar_final);
            }
          }
        }
      }
    }
  } else {
    setmaxnreg_inc<224>();
    // both warpgroups work on every tile: warpgroup x takes key columns [32x, 32x + 32) of the row (no row reduction is
    // needed in the backward), which halves the softmax stretch of the S -> dS -> dQ chain
    const int x = warp >> 2;
    const int r = threadIdx.x & 127;          // query row
    const uint32_t lane_addr = uint32_t((warp & 3) * 32) << 16;
    const uint32_t t_s = tmem_base + lane_addr;
    const int qi = q0 + r;
    const bool q_valid = qi < p.nq;
    const FaPos qpos = fa_pos(rule, rule.q, min(qi, p.nq - 1));
    const float lse2 = q_valid ? p.lse2[int64_t(b) * p.stat_pitch + qi] : __int_as_float(0x7f800000);
    const float dsum = q_valid ? p.dsum[int64_t(b) * p.stat_pitch + qi] : 0.f;
    const float scale_log2 = p.scale_log2;
    // The forward's l, m come from a 3xTF32 S (about 2^-21 relative), this kernel's S from three bf16 pieces (2^-24):
    // P = exp2(S c - LSE2) then misses sum_k P = 1 by a common factor per row of a few 1e-6, which is the whole error
    // budget of the gradients. The row sum of P is free here (one thread owns the row over all key tiles), so the row
    // is renormalised: dQ is divided by it and LSE2 + log2(rowsum) is handed to the dK/dV kernel. (Not when the call
    // covers a shard of the keys only - K/V ring - where the row sum over the shard is not 1 by construction.)
    float row_sum = 0.f;
    int j = 0;
    TileIter it;
    it.init(sched, 1, kt_first, kt_last);
    int kt, tw, tb;
    while (it.next(&kt, &tw, &tb)) {
      const int k0 = kt * kXN;
      const int k_hi = min(k0 + kXN, p.nk) - 1;
      const int cls = it.cls(0, tw, tb);
      const bool ragged = k0 + kXN > p.nk;
      mbar_wait(bar_s_full, j & 1);
      tc_fence_after();
      // masks first: nothing that may move registers between tcgen05.ld and tcgen05.wait::ld
      uint32_t okm = 0xffffffffu;
      if (cls == FA_TILE_PARTIAL || ragged) okm = tile_mask32(rule, true, qpos, k0, 32 * x, k_hi - k0 + 1);
      {
        float s[32], dp[32];
        tmem_ld32f(t_s + x * 32, s);
        tmem_ld32f(t_s + kColDp + x * 32, dp);
        tmem_wait_ld();
        uint32_t c0[16], c1[16], c2[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float p0 = exp2f(fmaf(s[c], scale_log2, -lse2));
          float p1 = exp2f(fmaf(s[c + 1], scale_log2, -lse2));
          p0 = (okm >> c) & 1u ? p0 : 0.f;
          p1 = (okm >> (c + 1)) & 1u ? p1 : 0.f;
          row_sum += p0 + p1;
          split3_pack(p0 * (dp[c] - dsum), p1 * (dp[c + 1] - dsum), c0[c >> 1], c1[c >> 1], c2[c >> 1]);
        }
        tmem_st16(t_s + kColDs + x * 16, c0);
        tmem_st16(t_s + kColDs + 32 + x * 16, c1);
        tmem_st16(t_s + kColDs + 64 + x * 16, c2);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p_ready);
      ++j;
    }
    // epilogue: dQ = scale * acc, fp32, straight to global (lanes = consecutive queries -> coalesced per channel)
    float* out = p.d_q + int64_t(b) * p.d * p.nq + qi;
    {
      // the two halves of the row sum meet in shared memory (the statistics staging area is unused by this kernel)
      float* halves = reinterpret_cast<float*>(smem_gen + Cfg::kStatOffset);
      halves[x * kXM + r] = row_sum;
      named_bar_sync(1, 2 * kXM);
      row_sum = halves[r] + halves[kXM + r];
    }
    const float out_scale = !p.renormalise ? p.scale : (row_sum > 0.f ? p.scale / row_sum : 0.f);
    if (q_valid && x == 0)
      p.lse2_refined[int64_t(b) * p.stat_pitch + qi] =
          !p.renormalise ? lse2 : (row_sum > 0.f ? lse2 + log2f(row_sum) : __int_as_float(0x7f800000));
    if (j > 0) {
      mbar_wait(bar_final, 0);
      tc_fence_after();
      for (int c = x; c < (D + 31) / 32; c += 2) {   // the warpgroups share the channel chunks
        float o[32], oc[32];
        tmem_ld32f(t_s + kColDq + c * 32, o);
        tmem_ld32f(t_s + kColDqc + c * 32, oc);
        tmem_wait_ld();
        if (q_valid) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c * 32 + e < p.d) out[int64_t(c * 32 + e) * p.nq] = (o[e] + oc[e]) * out_scale;
        }
      }
    } else if (q_valid && x == 0) {
      for (int c = 0; c < p.d; ++c) out[int64_t(c) * p.nq] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =================================================================================================
// dK / dV kernel
// =================================================================================================
template <int D, int VD>
__global__ void __launch_bounds__(kXThreads, 1) bwd_dkdv_f32_kernel(const __grid_constant__ BwdF32Params p) {
  using Cfg = XCfg<D, VD>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t k_smem = smem_base;                           // [3][kResD]
  const uint32_t v_smem = smem_base + 3 * Cfg::kResD;          // [3][kResV]
  const uint32_t ring = smem_base + Cfg::kRingOffset;          // stage: Q pieces [3][kStrD], dO pieces [3][kStrV]
  const uint32_t stat_smem = smem_base + Cfg::kStatOffset;
  const uint32_t bars = smem_base + Cfg::kBarOffset;
  const uint32_t bar_kv_res = bars;
  const uint32_t bar_full = bars + 8;
  const uint32_t bar_empty = bar_full + 8 * kXStages;
  const uint32_t bar_s_full = bar_empty + 8 * kXStages;
  const uint32_t bar_p_ready = bar_s_full + 8;
  const uint32_t bar_final = bar_p_ready + 8;
  const uint32_t tmem_slot = bar_final + 8;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::kBarOffset + Cfg::kNumBars * 8);
  const float* stat_gen = reinterpret_cast<const float*>(smem_gen + Cfg::kStatOffset);

  const int warp = threadIdx.x >> 5;
  const FaRule& rule = p.rule;
  const int b = int(blockIdx.x / p.n_blocks);
  const int kblk = int(blockIdx.x % p.n_blocks);
  const int k0 = kblk * kXM;
  const int k_hi = min(k0 + kXM, p.nk) - 1;
  int qt_first, qt_last;
  fa_q_tile_range(rule, k0, k_hi, kXN, &qt_first, &qt_last);
  TileSchedule* sched = reinterpret_cast<TileSchedule*>(smem_gen + Cfg::kSchedOffset);
  {
    const int lo[1] = {k0};
    const int hi[1] = {k_hi};
    const bool valid[1] = {true};
    build_schedule(sched, rule, false, lo, hi, valid, 1, qt_first, qt_last, kXN, p.nq, kXThreads / 32);
  }
  if (warp == 8) {
    if (elect_one())
      for (int j = 0; j < 3; ++j) {
        prefetch_tensormap(&p.map_q[j]);
        prefetch_tensormap(&p.map_k[j]);
        prefetch_tensormap(&p.map_v[j]);
        prefetch_tensormap(&p.map_do[j]);
      }
  } else if (warp == 9) {
    if (elect_one()) {
      mbar_init(bar_kv_res, 1);
      mbar_init(bar_s_full, 1);
      mbar_init(bar_p_ready, 2 * kXM);
      mbar_init(bar_final, 1);
      for (int s = 0; s < kXStages; ++s) {
        mbar_init(bar_full + 8 * s, 1);
        mbar_init(bar_empty + 8 * s, 1);
      }
      fence_barrier_init();
    }
  } else if (warp == 10) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // TMEM columns: S^T [0, 64) and dP^T [64, 128) are read completely into registers by the row's thread, which then
  // writes the P^T pieces over them at [32 j, +32) and the dS^T pieces at [96 + 32 j, +32) (a thread only ever touches
  // its own lane, and the next S^T / dP^T products are issued behind the dV / dK products that consume the pieces).
  // dV [192, 256)  dK [256, 320)  and their small-term accumulators dVc [320, 384)  dKc [384, 448): dV and dK live
  // across all query tiles, and the tensor core truncates on every accumulation - with everything in one accumulator
  // the bias reached 2.7e-5 after 16 tiles; the leading products alone add 6x less often.
  constexpr uint32_t kColDp = 64, kColP = 0, kColDs = 96, kColDv = 192, kColDk = 256, kColDvc = 320, kColDkc = 384;

  if (warp >= 8) {
    setmaxnreg_dec<56>();
    if (warp == 8) {
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_kv_res, Cfg::kResBytes);
        for (int j = 0; j < 3; ++j)
          for (int h = 0; h < 2; ++h) {
            tma_load_2d(k_smem + j * Cfg::kResD + h * (D * 128), &p.map_k[j], bar_kv_res, k0 + h * 64, b * D);
            tma_load_2d(v_smem + j * Cfg::kResV + h * (VD * 128), &p.map_v[j], bar_kv_res, k0 + h * 64, b * VD);
          }
        int t = 0;
        TileIter it;
        it.init(sched, 1, qt_first, qt_last);
        int qt, tw, tb;
        while (it.next(&qt, &tw, &tb)) {
          const int s = t % kXStages, u = t / kXStages;
          mbar_wait(bar_empty + 8 * s, (u & 1) ^ 1);
          mbar_arrive_expect_tx(bar_full + 8 * s, Cfg::kStageBytes + Cfg::kStatBytes);
          for (int j = 0; j < 3; ++j) {
            tma_load_2d(ring + s * Cfg::kStageBytes + j * Cfg::kStrD, &p.map_q[j], bar_full + 8 * s, qt * kXN,
                        b * D);
            tma_load_2d(ring + s * Cfg::kStageBytes + 3 * Cfg::kStrD + j * Cfg::kStrV, &p.map_do[j],
                        bar_full + 8 * s, qt * kXN, b * VD);
          }
          const int64_t off = int64_t(b) * p.stat_pitch + qt * kXN;
          bulk_load_1d_x(stat_smem + s * Cfg::kStatBytes, p.lse2_refined + off, kXN * 4, bar_full + 8 * s);
          bulk_load_1d_x(stat_smem + s * Cfg::kStatBytes + kXN * 4, p.dsum + off, kXN * 4, bar_full + 8 * s);
          ++t;
        }
      }
    } else if (warp == 9) {
      if (elect_one()) {
        TileIter it;
        it.init(sched, 1, qt_first, qt_last);
        const int n = it.count();
        constexpr uint32_t idesc_st = idesc_bf16(kXM, kXN, true, true);
        constexpr uint32_t idesc_dv = idesc_bf16(kXM, VD, false, false);
        constexpr uint32_t idesc_dk = idesc_bf16(kXM, D, false, false);
        auto issue_st_dpt = [&](int stage) {
          const uint32_t q_s = ring + stage * Cfg::kStageBytes, do_s = q_s + 3 * Cfg::kStrD;
          issue_ss6<D>(tmem_base, k_smem, Cfg::kResD, q_s, Cfg::kStrD, idesc_st);
          issue_ss6<VD>(tmem_base + kColDp, v_smem, Cfg::kResV, do_s, Cfg::kStrV, idesc_st);
        };
        if (n > 0) {
          mbar_wait(bar_kv_res, 0);
          mbar_wait(bar_full + 0, 0);
          tc_fence_after();
          issue_st_dpt(0);
          mma_commit(bar_s_full);
          for (int t = 0; t < n; ++t) {
            const int st = t % kXStages, s2 = (t + 1) % kXStages;
            mbar_wait(bar_p_ready, t & 1);
            tc_fence_after();
            const uint32_t q_s = ring + st * Cfg::kStageBytes, do_s = q_s + 3 * Cfg::kStrD;
            issue_ts6(tmem_base + kColDv, tmem_base + kColDvc, tmem_base + kColP, do_s, Cfg::kStrV, idesc_dv, t > 0);
            issue_ts6(tmem_base + kColDk, tmem_base + kColDkc, tmem_base + kColDs, q_s, Cfg::kStrD, idesc_dk, t > 0);
            mma_commit(bar_empty + 8 * st);
            if (t + 1 < n) {
              mbar_wait(bar_full + 8 * s2, ((t + 1) / kXStages) & 1);
              tc_fence_after();
              issue_st_dpt(s2);
              mma_commit(bar_s_full);
            }
          }
          mma_commit(bar_final);
        }
      }
    }
  } else {
    setmaxnreg_inc<224>();
    const int x = warp >> 2;                  // query-column half in the main loop; dV / dK role in the epilogue
    const int r = threadIdx.x & 127;          // key row
    const uint32_t lane_addr = uint32_t((warp & 3) * 32) << 16;
    const uint32_t t_s = tmem_base + lane_addr;
    const int ki = k0 + r;
    const bool k_valid = ki < p.nk;
    const FaPos kpos = fa_pos(rule, rule.k, min(ki, p.nk - 1));
    const float scale_log2 = p.scale_log2;
    int t = 0;
    TileIter it;
    it.init(sched, 1, qt_first, qt_last);
    int qt, tw, tb;
    while (it.next(&qt, &tw, &tb)) {
      const int st = t % kXStages;
      const int q0 = qt * kXN;
      const int q_hi = min(q0 + kXN, p.nq) - 1;
      const int cls = it.cls(0, tw, tb);
      const bool ragged = (q0 + kXN > p.nq) || (k0 + kXM > p.nk);
      mbar_wait(bar_full + 8 * st, (t / kXStages) & 1);   // statistics visible to this thread
      mbar_wait(bar_s_full, t & 1);
      tc_fence_after();
      uint32_t okm = 0xffffffffu;
      if (cls == FA_TILE_PARTIAL || ragged)
        okm = k_valid ? tile_mask32(rule, false, kpos, q0, 32 * x, q_hi - q0 + 1) : 0u;
      const float* lse_s = stat_gen + st * (2 * kXN) + x * 32;
      const float* dsum_s = lse_s + kXN;
      float s[32], dp[32];
      tmem_ld32f(t_s + x * 32, s);
      tmem_ld32f(t_s + kColDp + x * 32, dp);
      tmem_wait_ld();
      // the pieces overwrite S^T / dP^T columns of BOTH halves: nobody stores before both warpgroups have loaded
      named_bar_sync(1, 2 * kXM);
      {
        uint32_t a0[16], a1[16], a2[16], c0[16], c1[16], c2[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float p0 = exp2f(fmaf(s[c], scale_log2, -lse_s[c]));
          float p1 = exp2f(fmaf(s[c + 1], scale_log2, -lse_s[c + 1]));
          p0 = (okm >> c) & 1u ? p0 : 0.f;
          p1 = (okm >> (c + 1)) & 1u ? p1 : 0.f;
          split3_pack(p0, p1, a0[c >> 1], a1[c >> 1], a2[c >> 1]);
          split3_pack(p0 * (dp[c] - dsum_s[c]), p1 * (dp[c + 1] - dsum_s[c + 1]), c0[c >> 1], c1[c >> 1], c2[c >> 1]);
        }
        tmem_st16(t_s + kColP + x * 16, a0);
        tmem_st16(t_s + kColP + 32 + x * 16, a1);
        tmem_st16(t_s + kColP + 64 + x * 16, a2);
        tmem_st16(t_s + kColDs + x * 16, c0);
        tmem_st16(t_s + kColDs + 32 + x * 16, c1);
        tmem_st16(t_s + kColDs + 64 + x * 16, c2);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p_ready);
      ++t;
    }
    // epilogue: dV, dK = scale * acc in fp32, straight to global (lanes = consecutive keys -> coalesced per channel)
    float* out_v = p.d_v + int64_t(b) * p.v_d * p.nk + ki;
    float* out_k = p.d_k + int64_t(b) * p.d * p.nk + ki;
    // warpgroup 0 writes dV, warpgroup 1 dK
    if (t > 0) {
      mbar_wait(bar_final, 0);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < (x == 0 ? (VD + 31) / 32 : 0); ++c) {
        float o[32], oc[32];
        tmem_ld32f(t_s + kColDv + c * 32, o);
        tmem_ld32f(t_s + kColDvc + c * 32, oc);
        tmem_wait_ld();
        if (k_valid) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c * 32 + e < p.v_d) out_v[int64_t(c * 32 + e) * p.nk] = o[e] + oc[e];
        }
      }
#pragma unroll
      for (int c = 0; c < (x == 1 ? (D + 31) / 32 : 0); ++c) {
        float o[32], oc[32];
        tmem_ld32f(t_s + kColDk + c * 32, o);
        tmem_ld32f(t_s + kColDkc + c * 32, oc);
        tmem_wait_ld();
        if (k_valid) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c * 32 + e < p.d) out_k[int64_t(c * 32 + e) * p.nk] = (o[e] + oc[e]) * p.scale;
        }
      }
    } else if (k_valid) {
      if (x == 0)
        for (int c = 0; c < p.v_d; ++c) out_v[int64_t(c) * p.nk] = 0.f;
      else
        for (int c = 0; c < p.d; ++c) out_k[int64_t(c) * p.nk] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side -------------------------------------------------------------------------------------------
static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

static int64_t pad8x(int64_t n) { return (n + 7) & ~int64_t(7); }
static int64_t pad4x(int64_t n) { return (n + 3) & ~int64_t(3); }
struct F32BwdWorkspace {
  size_t stats, q, dout, k, v, total;   // byte sizes: statistics, one bf16 piece of Q, dO, K, V (kernel shape, padded)
  int64_t nqp, nkp, sp;
};
// d, v_d here are the KERNEL's channel counts
static F32BwdWorkspace f32_bwd_layout(int64_t batch, int64_t nq, int64_t nk, int d, int v_d) {
  F32BwdWorkspace w;
  w.nqp = pad8x(nq);
  w.nkp = pad8x(nk);
  w.sp = pad4x(nq);
  w.stats = align256(size_t(3) * (batch * w.sp + kXStatPad) * sizeof(float));
  w.q = align256(size_t(batch) * d * w.nqp * 2);
  w.dout = align256(size_t(batch) * v_d * w.nqp * 2);
  w.k = align256(size_t(batch) * d * w.nkp * 2);
  w.v = align256(size_t(batch) * v_d * w.nkp * 2);
  w.total = w.stats + 3 * (w.q + w.dout + w.k + w.v);
  return w;
}

struct SplitJobs {
  const float* src[4];
  __nv_bfloat16* dst[4][3];
  int64_t batch;
  int32_t c_src[4], c_dst[4], s_src[4], s_dst[4];   // [batch][c_src][s_src] -> [batch][c_dst][s_dst], zero-padded
};
// one launch for the four operands and the row statistics: blockIdx.y selects the tensor, y == 4 the statistics pass.
// The pieces are written in the KERNEL's shape (channels padded to its D / VD, lengths to a multiple of 8), so any
// channel count up to 64 and any length run on the tensor cores.
__global__ void split_bf16x3_all(const SplitJobs jobs, const PrepJob prep) {
  const int t = blockIdx.y;
  if (t == 4) {
    bwd_prep_f32(prep.o, prep.d_o, prep.l, prep.m, prep.lse2, prep.dsum, prep.lse2_refined, prep.batch, prep.v_d,
                 prep.nq, prep.sp);
    return;
  }
  const float* __restrict__ x = jobs.src[t];
  const int64_t cd = jobs.c_dst[t], sd = jobs.s_dst[t], cs = jobs.c_src[t], ss = jobs.s_src[t];
  const int64_t n = jobs.batch * cd * sd;
  const bool same = cd == cs && sd == ss;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    float v;
    if (same) {
      v = x[i];
    } else {
      const int64_t px = i % sd, bc = i / sd, c = bc % cd, b = bc / cd;
      v = (c < cs && px < ss) ? x[(b * cs + c) * ss + px] : 0.f;
    }
    const __nv_bfloat16 a = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(a);
    const __nv_bfloat16 b2 = __float2bfloat16_rn(r1);
    jobs.dst[t][0][i] = a;
    jobs.dst[t][1][i] = b2;
    jobs.dst[t][2][i] = __float2bfloat16_rn(r1 - __bfloat162float(b2));
  }
}

template <int D, int VD>
cudaError_t launch_bwd_f32(const LaunchArgs& a, cudaStream_t stream) {
  using Cfg = XCfg<D, VD>;
  const int nq = a.rule.q.total, nk = a.rule.k.total;
  const F32BwdWorkspace w = f32_bwd_layout(a.batch, nq, nk, D, VD);
  char* ws = reinterpret_cast<char*>(a.workspace);
  float* lse2 = reinterpret_cast<float*>(ws);
  float* dsum = lse2 + (a.batch * w.sp + kXStatPad);
  float* lse2_refined = dsum + (a.batch * w.sp + kXStatPad);
  char* pieces = ws + w.stats;
  __nv_bfloat16 *qp[3], *dop[3], *kp[3], *vp[3];
  for (int j = 0; j < 3; ++j) {
    qp[j] = reinterpret_cast<__nv_bfloat16*>(pieces + j * w.q);
    dop[j] = reinterpret_cast<__nv_bfloat16*>(pieces + 3 * w.q + j * w.dout);
    kp[j] = reinterpret_cast<__nv_bfloat16*>(pieces + 3 * (w.q + w.dout) + j * w.k);
    vp[j] = reinterpret_cast<__nv_bfloat16*>(pieces + 3 * (w.q + w.dout + w.k) + j * w.v);
  }
  BwdF32Params p;
  for (int j = 0; j < 3; ++j)
    if (!make_map_2d(&p.map_q[j], qp[j], a.batch * D, w.nqp, 64, D, true) ||
        !make_map_2d(&p.map_do[j], dop[j], a.batch * VD, w.nqp, 64, VD, true) ||
        !make_map_2d(&p.map_k[j], kp[j], a.batch * D, w.nkp, 64, D, true) ||
        !make_map_2d(&p.map_v[j], vp[j], a.batch * VD, w.nkp, 64, VD, true))
      return cudaErrorInvalidValue;
  p.rule = a.rule;
  p.lse2 = lse2;
  p.dsum = dsum;
  p.lse2_refined = lse2_refined;
  p.d_q = (float*)a.d_q;
  p.d_k = (float*)a.d_k;
  p.d_v = (float*)a.d_v;
  p.nq = nq;
  p.nk = nk;
  p.batch = int32_t(a.batch);
  p.d = a.d;
  p.v_d = a.v_d;
  p.stat_pitch = int32_t(w.sp);
  p.renormalise = a.partial_keys ? 0 : 1;
  p.scale = 1.f / sqrtf(float(a.d));
  p.scale_log2 = p.scale * kXLog2e;
  cudaError_t e;
  {
    SplitJobs jobs;
    const float* src[4] = {(const float*)a.q, (const float*)a.d_o, (const float*)a.k, (const float*)a.v};
    __nv_bfloat16** dst[4] = {qp, dop, kp, vp};
    const int cs[4] = {a.d, a.v_d, a.d, a.v_d}, cd[4] = {D, VD, D, VD};
    const int64_t ss[4] = {nq, nq, nk, nk}, sd[4] = {w.nqp, w.nqp, w.nkp, w.nkp};
    jobs.batch = a.batch;
    int64_t nmax = 0;
    for (int t = 0; t < 4; ++t) {
      jobs.src[t] = src[t];
      jobs.c_src[t] = cs[t];
      jobs.c_dst[t] = cd[t];
      jobs.s_src[t] = int32_t(ss[t]);
      jobs.s_dst[t] = int32_t(sd[t]);
      for (int j = 0; j < 3; ++j) jobs.dst[t][j] = dst[t][j];
      nmax = std::max<int64_t>(nmax, a.batch * cd[t] * sd[t]);
    }
    const int blocks = int(std::min<int64_t>((nmax + 255) / 256, 148 * 8));
    const PrepJob prep{(const float*)a.o, (const float*)a.d_o, (const float*)a.l, (const float*)a.m, lse2, dsum,
                       lse2_refined, a.batch, a.v_d, nq, int32_t(w.sp)};
    ScopedKernel timed("split_bf16x3+prep", stream);
    split_bf16x3_all<<<dim3(blocks, 5), 256, 0, stream>>>(jobs, prep);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  {
    auto kern = bwd_dq_f32_kernel<D, VD>;
    e = plan::ensure_smem(kern, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    p.n_blocks = (nq + kXM - 1) / kXM;
    ScopedKernel timed("bwd_dq_f32_bf16x3_sm100", stream);
    kern<<<unsigned(int64_t(p.n_blocks) * p.batch), kXThreads, Cfg::kSmemBytes, stream>>>(p);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  {
    auto kern = bwd_dkdv_f32_kernel<D, VD>;
    e = plan::ensure_smem(kern, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    p.n_blocks = (nk + kXM - 1) / kXM;
    ScopedKernel timed("bwd_dkdv_f32_bf16x3_sm100", stream);
    kern<<<unsigned(int64_t(p.n_blocks) * p.batch), kXThreads, Cfg::kSmemBytes, stream>>>(p);
    return cudaGetLastError();
  }
}

}  // namespace sm100

static bool aligned16x(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// the instantiation that holds the problem's channel counts with the least padding (same choice as the forward)
static void f32_bwd_pick(const LaunchArgs& a, int* D, int* VD) {
  if (a.d <= 32 && a.v_d <= 16) { *D = 32; *VD = 16; }
  else if (a.d <= 32 && a.v_d <= 32) { *D = 32; *VD = 32; }
  else { *D = 64; *VD = 64; }
}

size_t sm100_f32_backward_workspace_bytes(const LaunchArgs& a) {
  int D, VD;
  f32_bwd_pick(a, &D, &VD);
  return sm100::f32_bwd_layout(a.batch, a.rule.q.total, a.rule.k.total, D, VD).total;
}

// Any channel counts up to 64 and any lengths: the split pass writes the bf16 pieces in the kernel's shape (zero-padded
// channels, lengths padded to 8), the kernels store only the tensors' own channels.
bool sm100_f32_backward_supports(const LaunchArgs& a) {
  if (a.dtype != 1 || a.accumulate || a.layout != 0) return false;
  if (a.d < 1 || a.v_d < 1 || a.d > 64 || a.v_d > 64) return false;
  const int64_t nq = a.rule.q.total, nk = a.rule.k.total;
  if (a.workspace && !aligned16x(a.workspace)) return false;
  if (a.batch * 64 > 0x7fffffffLL) return false;
  if (((nq + 127) / 128) * a.batch > 0x7fffffffLL || ((nk + 127) / 128) * a.batch > 0x7fffffffLL) return false;
  if (a.workspace_bytes < sm100_f32_backward_workspace_bytes(a)) return false;
  if ((nk + 63) / 64 > 32 * sm100::kMaxTileWords || (nq + 63) / 64 > 32 * sm100::kMaxTileWords) return false;
  return true;
}

cudaError_t sm100_f32_backward(const LaunchArgs& a, cudaStream_t stream) {
  int D, VD;
  f32_bwd_pick(a, &D, &VD);
  if (D == 64) return sm100::launch_bwd_f32<64, 64>(a, stream);
  if (VD == 32) return sm100::launch_bwd_f32<32, 32>(a, stream);
  return sm100::launch_bwd_f32<32, 16>(a, stream);
}

}  // namespace fa
