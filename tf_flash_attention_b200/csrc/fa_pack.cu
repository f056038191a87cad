// fa_pack.cu — pitch-padding copies in front of / behind the TMA kernels.
//
// TMA addresses a tensor through strides that must be multiples of 16 bytes, so a channel-first fp16 tensor whose
// sequence length is not a multiple of 8 (the reference accepts any length through predicated loads,
// cute_ext/boundary_check_pred.h:35-137, and its tests draw arbitrary even lengths, tests/test_base.py:144-168)
// cannot be read in place. The tensor-core paths then copy the operand into the workspace with the row pitch rounded up
// to 8 elements (rows = batch x channels; the tensor maps keep the true length, so the padding is never read as data:
// TMA zero-fills past the end) and copy the results back the same way. Pure HBM-bound byte moves:
// algorithmic bytes = 2 x rows x len x sizeof(T) per call.
#include "fa_launch.h"

namespace fa {
namespace {

template <typename T>
__global__ void pack_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t rows, int64_t len,
                                 int64_t src_pitch, int64_t dst_pitch) {
  // one block row per (row chunk); threads run along the sequence: coalesced on both sides
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const T* s = src + r * src_pitch;
    T* d = dst + r * dst_pitch;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < len; i += int64_t(gridDim.x) * blockDim.x)
      d[i] = s[i];
  }
}

// [batch][src_ch][src_pitch] -> [batch][dst_ch][dst_pitch], copying the first `ch` channels and `len` positions of each
template <typename T>
__global__ void pack_channels_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t batch, int64_t ch,
                                     int64_t len, int64_t src_ch, int64_t src_pitch, int64_t dst_ch, int64_t dst_pitch) {
  const int64_t rows = batch * ch;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    const int64_t b = r / ch, c = r - b * ch;
    const T* s = src + (b * src_ch + c) * src_pitch;
    T* d = dst + (b * dst_ch + c) * dst_pitch;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < len; i += int64_t(gridDim.x) * blockDim.x)
      d[i] = s[i];
  }
}

template <typename T>
cudaError_t launch_pack(const void* src, void* dst, int64_t rows, int64_t len, int64_t src_pitch, int64_t dst_pitch,
                        cudaStream_t stream) {
  if (rows <= 0 || len <= 0) return cudaSuccess;
  const int threads = 256;
  const unsigned gx = unsigned(std::min<int64_t>((len + threads - 1) / threads, 64));
  const unsigned gy = unsigned(std::min<int64_t>(rows, 148 * 32));
  ScopedKernel timed("pack_rows", stream);
  pack_rows_kernel<T><<<dim3(gx, gy), threads, 0, stream>>>(static_cast<const T*>(src), static_cast<T*>(dst), rows, len,
                                                             src_pitch, dst_pitch);
  return cudaGetLastError();
}

}  // namespace

cudaError_t pack_channels_f32(const float* src, float* dst, int64_t batch, int64_t ch, int64_t len, int64_t src_ch,
                              int64_t src_pitch, int64_t dst_ch, int64_t dst_pitch, cudaStream_t stream) {
  if (batch <= 0 || ch <= 0 || len <= 0) return cudaSuccess;
  const int threads = 256;
  const unsigned gx = unsigned(std::min<int64_t>((len + threads - 1) / threads, 64));
  const unsigned gy = unsigned(std::min<int64_t>(batch * ch, 148 * 32));
  ScopedKernel timed("pack_channels", stream);
  pack_channels_kernel<float><<<dim3(gx, gy), threads, 0, stream>>>(src, dst, batch, ch, len, src_ch, src_pitch, dst_ch,
                                                                     dst_pitch);
  return cudaGetLastError();
}

cudaError_t pack_rows(int elt_bytes, const void* src, void* dst, int64_t rows, int64_t len, int64_t src_pitch,
                      int64_t dst_pitch, cudaStream_t stream) {
  switch (elt_bytes) {
    case 2: return launch_pack<uint16_t>(src, dst, rows, len, src_pitch, dst_pitch, stream);
    case 4: return launch_pack<uint32_t>(src, dst, rows, len, src_pitch, dst_pitch, stream);
    case 8: return launch_pack<uint64_t>(src, dst, rows, len, src_pitch, dst_pitch, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace fa
