// TEST / BASELINE INFRASTRUCTURE — a TensorFlow-free harness around the UNMODIFIED reference
// kernel. It is compiled together with /root/reference/flash_attention/kernel/flash_attention.cu
// and sync_methods.cc (sources stay where they lie; nothing is copied) into
// oracle/_ref/libref_fa.so and replaces only the two TF OpKernel files: it performs exactly the
// steps of FlashAttentionForwardBase::Compute (flash_attention_forward.cc:303-385) and
// FlashAttentionBackwardBase::Compute (flash_attention_backward.cc:260-341): sync maps, the four
// cudaMemsetAsync calls, then launcher.Forward / launcher.Backward.
// Used (a) on a B200 to produce tests/golden/refkernel_*.npz that pin the oracle against the real
// reference, (b) by bench.py as the "reference's own CuTe kernel on the same B200" baseline, and
// (c) on the host for EstimateForwardFlops golden values.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <optional>
#include <string>
#include <vector>
#include "tensorflow/core/framework/tensor_shape.h"
#include "flash_attention.h"
#include "sync_methods.h"

using namespace cuda_launch;

namespace {

struct Call {
  int dims; std::string mode; int b, d, v_d;
  std::vector<int64_t> qs, ks;
  const void *Q, *K, *V; void *O, *l, *m;            // forward
  const void* dO; void *dQ, *dK, *dV;                // backward
  uint32_t* occupancy;                               // [b * ceil(q / BR)]
  cudaStream_t stream; int smem_basic, smem_optin;
  float* flops;
  int op;                                            // 0 fwd, 1 bwd, 2 flops
};

template <typename T, int D, typename Policy>
int run(const Call& c, const Policy& policy) {
  std::optional<SyncMethod<D>> sm;
  SyncMethods::Lookup<D>(c.mode, sm);
  if (!sm.has_value()) return -5;
  tensorflow::TensorShape Qs(c.qs), Ks(c.ks);
  const auto [reference_seq_shape, Q_seq_order_map, K_seq_order_map] = (*sm)(Qs, Ks);
  FlashAttentionLauncher<T, std::remove_cv_t<decltype(reference_seq_shape)>,
                         std::remove_cv_t<decltype(Q_seq_order_map)>, Policy> launcher{};
  using L_T = typename decltype(launcher)::L_T;
  int64_t q = 1, k = 1;
  for (auto v : c.qs) q *= v;
  for (auto v : c.ks) k *= v;
  const SharedMemoryDescriptor smd(c.smem_basic, c.smem_optin);
  if (c.op == 2) {
    launcher.EstimateForwardFlops(smd, c.b, int32_t(q), int32_t(k), c.d, c.v_d, reference_seq_shape,
                                  Q_seq_order_map, K_seq_order_map, policy, *c.flops);
    return 0;
  }
  const size_t occ_bytes = size_t(c.b) * launcher.ComputeNumOfBrSections(int(q)) * sizeof(uint32_t);
  cudaError_t e;
  if (c.op == 0) {
    cudaMemsetAsync(c.O, 0, size_t(c.b) * c.v_d * q * sizeof(T), c.stream);
    cudaMemsetAsync(c.l, 0, size_t(c.b) * q * sizeof(L_T), c.stream);
    cudaMemsetAsync(c.m, 0xfa, size_t(c.b) * q * sizeof(T), c.stream);
    cudaMemsetAsync(c.occupancy, 0, occ_bytes, c.stream);
    e = launcher.Forward(c.stream, smd, c.b, int32_t(q), int32_t(k), c.d, c.v_d, (const T*)c.Q,
                         (const T*)c.K, (const T*)c.V, (T*)c.O, (L_T*)c.l, (T*)c.m, c.occupancy,
                         reference_seq_shape, Q_seq_order_map, K_seq_order_map, policy);
  } else {
    cudaMemsetAsync(c.dQ, 0, size_t(c.b) * c.d * q * sizeof(T), c.stream);
    cudaMemsetAsync(c.dK, 0, size_t(c.b) * c.d * k * sizeof(T), c.stream);
    cudaMemsetAsync(c.dV, 0, size_t(c.b) * c.v_d * k * sizeof(T), c.stream);
    cudaMemsetAsync(c.occupancy, 0, occ_bytes, c.stream);
    e = launcher.Backward(c.stream, smd, c.b, int32_t(q), int32_t(k), c.d, c.v_d, (const T*)c.Q,
                          (const T*)c.K, (const T*)c.V, (const T*)c.O, (const L_T*)c.l, (const T*)c.m,
                          (const T*)c.dO, (T*)c.dQ, (T*)c.dK, (T*)c.dV, c.occupancy, reference_seq_shape,
                          Q_seq_order_map, K_seq_order_map, policy);
  }
  return int(e);
}

template <typename T, int D>
int by_rule(const Call& c, int rule, int w, int s, int causal) {
  if (rule == 0) return run<T, D>(c, FullAttentionPolicy{});
  if (rule == 1) return run<T, D>(c, CausalAttentionPolicy{});
  return run<T, D>(c, LocalAttentionPolicy(w, s, causal != 0));
}

template <typename T>
int by_dims(const Call& c, int rule, int w, int s, int causal) {
  return c.dims == 1 ? by_rule<T, 1>(c, rule, w, s, causal) : by_rule<T, 2>(c, rule, w, s, causal);
}

int dispatch(int dtype, const Call& c, int rule, int w, int s, int causal) {
  if (dtype == 0) return by_dims<half>(c, rule, w, s, causal);
  if (dtype == 1) return by_dims<float>(c, rule, w, s, causal);
  return by_dims<double>(c, rule, w, s, causal);
}

Call base(int dims, const char* mode, int b, int d, int v_d, const int64_t* qs, const int64_t* ks,
          void* stream, int smem_optin) {
  Call c{};
  c.dims = dims; c.mode = mode; c.b = b; c.d = d; c.v_d = v_d;
  c.qs.assign(qs, qs + dims); c.ks.assign(ks, ks + dims);
  c.stream = (cudaStream_t)stream;
  c.smem_basic = 48 << 10;
  c.smem_optin = smem_optin;
  return c;
}

}  // namespace

extern "C" {

// number of uint32 the caller must provide as Br_occupancy scratch
int64_t ref_occupancy_elems(int dtype, int b, int64_t q) {
  int br = dtype == 0 ? 64 : 32;  // KernelConfig<T>::BR_SIZE, flash_attention.h:200
  return int64_t(b) * ((q + br - 1) / br);
}

int ref_forward(int dtype, int dims, int rule, const char* mode, int w, int s, int causal, int b, int d,
                int v_d, const int64_t* qs, const int64_t* ks, const void* Q, const void* K, const void* V,
                void* O, void* l, void* m, uint32_t* occupancy, int smem_optin, void* stream) {
  Call c = base(dims, mode, b, d, v_d, qs, ks, stream, smem_optin);
  c.op = 0; c.Q = Q; c.K = K; c.V = V; c.O = O; c.l = l; c.m = m; c.occupancy = occupancy;
  return dispatch(dtype, c, rule, w, s, causal);
}

int ref_backward(int dtype, int dims, int rule, const char* mode, int w, int s, int causal, int b, int d,
                 int v_d, const int64_t* qs, const int64_t* ks, const void* Q, const void* K, const void* V,
                 const void* O, const void* l, const void* m, const void* dO, void* dQ, void* dK, void* dV,
                 uint32_t* occupancy, int smem_optin, void* stream) {
  Call c = base(dims, mode, b, d, v_d, qs, ks, stream, smem_optin);
  c.op = 1; c.Q = Q; c.K = K; c.V = V; c.O = (void*)O; c.l = (void*)l; c.m = (void*)m; c.dO = dO;
  c.dQ = dQ; c.dK = dK; c.dV = dV; c.occupancy = occupancy;
  return dispatch(dtype, c, rule, w, s, causal);
}

int ref_estimate_flops(int dtype, int dims, int rule, const char* mode, int w, int s, int causal, int b,
                       int d, int v_d, const int64_t* qs, const int64_t* ks, int smem_optin, float* flops) {
  Call c = base(dims, mode, b, d, v_d, qs, ks, nullptr, smem_optin);
  c.op = 2; c.flops = flops;
  return dispatch(dtype, c, rule, w, s, causal);
}

}  // extern "C"
