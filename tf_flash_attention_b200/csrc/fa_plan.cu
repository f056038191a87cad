// fa_plan.cu — see fa_plan.h.
#include "fa_plan.h"

#include <cudaTypedefs.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "../../include/fa_b200.h"

namespace fa {
namespace plan {
namespace {

struct MapKey {
  uint64_t base, gdim[4], gstride[3];
  uint32_t box[4], dtype, rank, swizzle, pad;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) h = (h ^ w[i]) * 0x100000001b3ull;
    return size_t(h ^ (h >> 29));
  }
};
static_assert(sizeof(MapKey) % 8 == 0, "MapKey is hashed as 64-bit words");

std::mutex g_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
struct AttrKey {
  const void* fn;
  int dev;
  bool operator==(const AttrKey& o) const { return fn == o.fn && dev == o.dev; }
};
struct AttrKeyHash {
  size_t operator()(const AttrKey& k) const { return std::hash<const void*>()(k.fn) * 31 + size_t(k.dev); }
};
std::unordered_map<AttrKey, int, AttrKeyHash> g_attrs;
std::atomic<uint64_t> g_stats[4];   // map encodes, map hits, attribute calls, attribute hits

PFN_cuTensorMapEncodeTiled_v12000 encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = []() {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  }();
  return fn;
}

}  // namespace

bool tensor_map(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* gdim,
                const uint64_t* gstride_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  if (rank < 1 || rank > 4) return false;
  MapKey k;
  memset(&k, 0, sizeof(k));
  k.base = reinterpret_cast<uint64_t>(base);
  for (int i = 0; i < rank; ++i) {
    k.gdim[i] = gdim[i];
    k.box[i] = box[i];
    if (i + 1 < rank) k.gstride[i] = gstride_bytes[i];
  }
  k.dtype = uint32_t(dtype);
  k.rank = uint32_t(rank);
  k.swizzle = uint32_t(swizzle);
  {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_maps.find(k);
    if (it != g_maps.end()) {
      *out = it->second;
      g_stats[1].fetch_add(1, std::memory_order_relaxed);
      return true;
    }
  }
  auto enc = encoder();
  if (!enc) return false;
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gd[i] = gdim[i];
    bx[i] = box[i];
    if (i + 1 < rank) gs[i] = gstride_bytes[i];
  }
  CUtensorMap m;
  if (enc(&m, dtype, cuuint32_t(rank), const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  g_stats[0].fetch_add(1, std::memory_order_relaxed);
  {
    std::lock_guard<std::mutex> g(g_mu);
    if (g_maps.size() > 8192) g_maps.clear();   // bounded: descriptors are cheap to rebuild
    g_maps.emplace(k, m);
  }
  *out = m;
  return true;
}

cudaError_t ensure_dynamic_smem(const void* kernel, int bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const AttrKey k{kernel, dev};
  {
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_attrs.find(k);
    if (it != g_attrs.end() && it->second >= bytes) {
      g_stats[3].fetch_add(1, std::memory_order_relaxed);
      return cudaSuccess;
    }
  }
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  g_stats[2].fetch_add(1, std::memory_order_relaxed);
  std::lock_guard<std::mutex> g(g_mu);
  g_attrs[k] = bytes;
  return cudaSuccess;
}

}  // namespace plan
}  // namespace fa

extern "C" int fa_plan_stats(uint64_t* out4, int reset) {
  if (!out4) return FA_EINVAL_NULL;
  for (int i = 0; i < 4; ++i)
    out4[i] = reset ? fa::plan::g_stats[i].exchange(0) : fa::plan::g_stats[i].load();
  return FA_OK;
}
