// fa_common.cuh — small shared device/host helpers for all kernel families.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fa_rules.h"

namespace fa {

template <typename T> struct AccOf { using type = float; };
template <> struct AccOf<double> { using type = double; };
// l is float for half, T otherwise (reference: flash_attention.h:181-185)
template <typename T> struct LOf { using type = T; };
template <> struct LOf<__half> { using type = float; };

template <typename T> __device__ __forceinline__ T neg_inf();
template <> __device__ __forceinline__ float neg_inf<float>() { return __int_as_float(0xff800000); }
template <> __device__ __forceinline__ double neg_inf<double>() { return __longlong_as_double(0xfff0000000000000ULL); }

template <typename A, typename T> __device__ __forceinline__ A to_acc(T v) { return static_cast<A>(v); }
template <> __device__ __forceinline__ float to_acc<float, __half>(__half v) { return __half2float(v); }

template <typename T, typename A> __device__ __forceinline__ T from_acc(A v) { return static_cast<T>(v); }
template <> __device__ __forceinline__ __half from_acc<__half, float>(float v) { return __float2half_rn(v); }

// The reference's "-inf approximation": every byte 0xFA (type_util.h:43-45,
// flash_attention_forward.cc:360-365). Observable in `m` on fully masked rows.
template <typename T> __device__ __forceinline__ T sentinel();
template <> __device__ __forceinline__ __half sentinel<__half>() { return __ushort_as_half((unsigned short)0xFAFA); }
template <> __device__ __forceinline__ float sentinel<float>() { return __uint_as_float(0xFAFAFAFAu); }
template <> __device__ __forceinline__ double sentinel<double>() { return __longlong_as_double(0xFAFAFAFAFAFAFAFAULL); }

template <typename T> __device__ __forceinline__ bool is_sentinel(T v);
template <> __device__ __forceinline__ bool is_sentinel<__half>(__half v) { return __half_as_ushort(v) == 0xFAFA; }
template <> __device__ __forceinline__ bool is_sentinel<float>(float v) { return __float_as_uint(v) == 0xFAFAFAFAu; }
template <> __device__ __forceinline__ bool is_sentinel<double>(double v) {
  return (unsigned long long)__double_as_longlong(v) == 0xFAFAFAFAFAFAFAFAULL;
}

__device__ __forceinline__ float acc_exp(float x) { return expf(x); }
__device__ __forceinline__ double acc_exp(double x) { return exp(x); }
__device__ __forceinline__ float acc_log(float x) { return logf(x); }
__device__ __forceinline__ double acc_log(double x) { return log(x); }
__device__ __forceinline__ float acc_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double acc_max(double a, double b) { return fmax(a, b); }

}  // namespace fa
