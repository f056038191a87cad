// fa_plan.h — launch-plan cache shared by the kernel families: everything a launch needs from the driver besides
// the launch itself is computed once and looked up afterwards ("locked plan cache").
//   * tensor maps: a CUtensorMap is a pure function of (base address, extents, strides, box, element type, swizzle),
//     so the encoded 128-byte descriptor is cached under exactly that key (a framework that reuses its buffers —
//     every training loop — hits on every step);
//   * cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is issued once per (kernel, device).
// fa_plan_stats() (include/fa_b200.h) exposes the hit / miss counters to the tests.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fa {
namespace plan {

// rank <= 4; gstride_bytes has rank - 1 entries (strides of dims 1..rank-1)
bool tensor_map(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* gdim,
                const uint64_t* gstride_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle);

cudaError_t ensure_dynamic_smem(const void* kernel, int bytes);

template <typename K>
cudaError_t ensure_smem(K kernel, int bytes) {
  return ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), bytes);
}

}  // namespace plan
}  // namespace fa
