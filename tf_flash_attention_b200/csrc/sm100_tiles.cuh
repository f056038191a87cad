// sm100_tiles.cuh — per-CTA tile schedule shared by the tcgen05 kernels.
//
// The reference decides per (K tile, Q tile) pair whether to skip it inside its hot loop
// (AttentionPolicy::IsSkipped, flash_attention/kernel/flash_attention.h:48-115, called at
// flash_attention.cu:867-871). Here the schedule is "compiled" once per CTA, cooperatively by all
// warps during the set-up phase, into bitmaps in shared memory: for each of up to two row groups of
// the CTA one bitmap of PARTIAL tiles (element mask needed) and one of FULL tiles (no mask) over
// the streamed tiles. SKIP tiles are in neither and are never loaded. Every role (TMA producer, MMA
// issuer, softmax warpgroups) then walks the same bitmaps with a handful of bit operations.
#pragma once
#include <stdint.h>

#include "fa_rules.h"

namespace fa {
namespace sm100 {

constexpr int kMaxTileWords = 64;  // 64 x 32 = 2048 streamed tiles per CTA

struct TileSchedule {
  uint32_t partial[2][kMaxTileWords];
  uint32_t full[2][kMaxTileWords];
};

// Row groups g = 0,1 own resident rows [res_lo[g], res_hi[g]] (valid[g] == false: group is empty).
// Streamed tiles are `tile` wide over [0, stream_total). If `resident_is_q` the resident rows are
// queries and the streamed tiles are keys, else the other way round.
// Must be called by all threads of the CTA (warp-uniform control flow); followed by __syncthreads.
__device__ __forceinline__ void build_schedule(TileSchedule* sch, const FaRule& rule, bool resident_is_q,
                                               const int* res_lo, const int* res_hi, const bool* valid,
                                               int n_groups, int t_first, int t_last, int tile,
                                               int stream_total, int n_warps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (t_first > t_last) return;
  const int t_base = t_first & ~31;
  const int n_words = (t_last - t_base) / 32 + 1;
  for (int w = warp; w < n_words; w += n_warps) {
    const int t = t_base + 32 * w + lane;
    const bool in = t >= t_first && t <= t_last;
    const int s0 = t * tile;
    const int s1 = min(s0 + tile, stream_total) - 1;
    for (int g = 0; g < n_groups; ++g) {
      int cls = FA_TILE_SKIP;
      if (in && valid[g])
        cls = resident_is_q ? fa_classify(rule, res_lo[g], res_hi[g], s0, s1)
                            : fa_classify(rule, s0, s1, res_lo[g], res_hi[g]);
      const uint32_t pw = __ballot_sync(0xffffffffu, cls == FA_TILE_PARTIAL);
      const uint32_t fw = __ballot_sync(0xffffffffu, cls == FA_TILE_FULL);
      if (lane == 0) {
        sch->partial[g][w] = pw;
        sch->full[g][w] = fw;
      }
    }
  }
}

// Iterates the live tiles (any group PARTIAL or FULL) in ascending order.
struct TileIter {
  const TileSchedule* sch;
  int n_groups, n_words, t_base, w;
  uint32_t live;
  __device__ __forceinline__ void init(const TileSchedule* s, int groups, int t_first, int t_last) {
    sch = s;
    n_groups = groups;
    t_base = t_first & ~31;
    n_words = t_first > t_last ? 0 : (t_last - t_base) / 32 + 1;
    w = -1;
    live = 0;
  }
  __device__ __forceinline__ uint32_t word_live(int ww) const {
    uint32_t v = sch->partial[0][ww] | sch->full[0][ww];
    if (n_groups > 1) v |= sch->partial[1][ww] | sch->full[1][ww];
    return v;
  }
  // returns false when exhausted; otherwise sets t (tile index) and bit position
  __device__ __forceinline__ bool next(int* t, int* word, int* bit) {
    while (live == 0) {
      ++w;
      if (w >= n_words) return false;
      live = word_live(w);
    }
    const int b = __ffs(int(live)) - 1;
    live &= live - 1;
    *t = t_base + 32 * w + b;
    *word = w;
    *bit = b;
    return true;
  }
  __device__ __forceinline__ int cls(int g, int word, int bit) const {
    if ((sch->full[g][word] >> bit) & 1u) return FA_TILE_FULL;
    if ((sch->partial[g][word] >> bit) & 1u) return FA_TILE_PARTIAL;
    return FA_TILE_SKIP;
  }
  __device__ __forceinline__ int count() const {
    int n = 0;
    for (int ww = 0; ww < n_words; ++ww) n += __popc(word_live(ww));
    return n;
  }
};

// 1-D full/causal rows: the attended streamed columns form one interval [lo, hi] of the tile.
// resident q: keys attended iff k.c0 <= q.c0  -> prefix      [0, limit]
// resident k: queries attended iff q.c0 >= k.c0 -> suffix    [cmin, nvalid-1]
__device__ __forceinline__ void interval_1d(const FaRule& rule, bool resident_is_q, const FaPos& res, int s0,
                                            int nvalid, int* lo, int* hi) {
  *lo = 0;
  *hi = nvalid - 1;
  if (!rule.causal) return;
  if (resident_is_q) {
    const int num = res.c0 - rule.k.off0;
    const int jmax = num < 0 ? -1 : num / rule.k.stride0;  // last global key index with k.c0 <= q.c0
    *hi = min(*hi, jmax - rule.k.base0 - s0);
  } else {
    const int num = res.c0 - rule.q.off0;
    const int jmin = num <= 0 ? 0 : (num + rule.q.stride0 - 1) / rule.q.stride0;
    *lo = max(0, jmin - rule.q.base0 - s0);
  }
}

// mask builder used by all tcgen05 kernels: closed form whenever the rule allows it
__device__ __forceinline__ uint32_t tile_mask32(const FaRule& rule, bool resident_is_q, const FaPos& res, int s0,
                                               int col0, int nvalid) {
  return fa_fast_mask32(rule, resident_is_q, res, s0, col0, nvalid);
}

}  // namespace sm100
}  // namespace fa
