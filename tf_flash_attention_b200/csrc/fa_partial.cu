// fa_partial.cu — merging partial attention results over disjoint key shards (K/V ring support).
// New functionality (the reference is single-GPU): when one long sequence is sharded over GPUs, each
// ring step attends the local queries to the visiting K/V shard with fa_forward (global index bases
// in fa_problem_t) and the partial (O, l, m) is folded into fp32 accumulators with the usual online
// softmax algebra; fa_partial_finalize then emits O, l, m in the reference's output contract.
#include "fa_common.cuh"
#include "fa_launch.h"

namespace fa {

template <typename T, typename A>
__global__ void partial_merge_kernel(const T* __restrict__ o_part, const typename LOf<T>::type* __restrict__ l_part,
                                     const T* __restrict__ m_part, A* __restrict__ o_acc, A* __restrict__ l_acc,
                                     A* __restrict__ m_acc, int64_t batch, int32_t v_d, int32_t nq, int first) {
  const int64_t total = batch * nq;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t b = i / nq, r = i - b * nq;
    const T mp_t = m_part[i];
    const A lp = A(l_part[i]);
    const bool part_empty = is_sentinel<T>(mp_t) || !(lp > A(0));
    const A mp = part_empty ? neg_inf<A>() : to_acc<A>(mp_t);
    const A ma = first ? neg_inf<A>() : m_acc[i];
    const A la = first ? A(0) : l_acc[i];
    const A m_new = acc_max(ma, mp);
    A wa = A(0), wp = A(0);
    if (m_new != neg_inf<A>()) {
      wa = ma == neg_inf<A>() ? A(0) : acc_exp(ma - m_new);
      wp = part_empty ? A(0) : acc_exp(mp - m_new) * lp;
    }
    m_acc[i] = m_new;
    l_acc[i] = la * wa + wp;
    const T* op = o_part + b * v_d * int64_t(nq) + r;
    A* oa = o_acc + b * v_d * int64_t(nq) + r;
    for (int c = 0; c < v_d; ++c) {
      const A prev = first ? A(0) : oa[int64_t(c) * nq];
      oa[int64_t(c) * nq] = prev * wa + to_acc<A>(op[int64_t(c) * nq]) * wp;
    }
  }
}

template <typename T, typename A>
__global__ void partial_finalize_kernel(const A* __restrict__ o_acc, const A* __restrict__ l_acc,
                                        const A* __restrict__ m_acc, T* __restrict__ o,
                                        typename LOf<T>::type* __restrict__ l, T* __restrict__ m, int64_t batch,
                                        int32_t v_d, int32_t nq) {
  using L = typename LOf<T>::type;
  const int64_t total = batch * nq;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t b = i / nq, r = i - b * nq;
    const A la = l_acc[i], ma = m_acc[i];
    const bool empty = !(la > A(0));
    const A inv = empty ? A(0) : A(1) / la;
    if (empty) {
      m[i] = sentinel<T>();
      l[i] = L(0);
    } else {
      const T m_t = from_acc<T>(ma);
      m[i] = m_t;
      l[i] = L(la * acc_exp(ma - to_acc<A>(m_t)));
    }
    const A* oa = o_acc + b * v_d * int64_t(nq) + r;
    T* og = o + b * v_d * int64_t(nq) + r;
    for (int c = 0; c < v_d; ++c) og[int64_t(c) * nq] = from_acc<T>(oa[int64_t(c) * nq] * inv);
  }
}

template <typename T>
static cudaError_t merge_t(const LaunchArgs& a, const void* o_part, const void* l_part, const void* m_part,
                           void* o_acc, void* l_acc, void* m_acc, int first, cudaStream_t stream) {
  using A = typename AccOf<T>::type;
  const int64_t total = a.batch * a.rule.q.total;
  const int blocks = int(std::min<int64_t>((total + 255) / 256, 148 * 16));
  ScopedKernel timed("partial_merge", stream);
  partial_merge_kernel<T, A><<<blocks, 256, 0, stream>>>((const T*)o_part, (const typename LOf<T>::type*)l_part,
                                                         (const T*)m_part, (A*)o_acc, (A*)l_acc, (A*)m_acc, a.batch,
                                                         a.v_d, a.rule.q.total, first);
  return cudaGetLastError();
}

template <typename T>
static cudaError_t finalize_t(const LaunchArgs& a, const void* o_acc, const void* l_acc, const void* m_acc, void* o,
                              void* l, void* m, cudaStream_t stream) {
  using A = typename AccOf<T>::type;
  const int64_t total = a.batch * a.rule.q.total;
  const int blocks = int(std::min<int64_t>((total + 255) / 256, 148 * 16));
  ScopedKernel timed("partial_finalize", stream);
  partial_finalize_kernel<T, A><<<blocks, 256, 0, stream>>>((const A*)o_acc, (const A*)l_acc, (const A*)m_acc, (T*)o,
                                                            (typename LOf<T>::type*)l, (T*)m, a.batch, a.v_d,
                                                            a.rule.q.total);
  return cudaGetLastError();
}

cudaError_t partial_merge(const LaunchArgs& a, const void* o_part, const void* l_part, const void* m_part,
                          void* o_acc, void* l_acc, void* m_acc, int first, cudaStream_t stream) {
  switch (a.dtype) {
    case 0: return merge_t<__half>(a, o_part, l_part, m_part, o_acc, l_acc, m_acc, first, stream);
    case 1: return merge_t<float>(a, o_part, l_part, m_part, o_acc, l_acc, m_acc, first, stream);
    default: return merge_t<double>(a, o_part, l_part, m_part, o_acc, l_acc, m_acc, first, stream);
  }
}

cudaError_t partial_finalize(const LaunchArgs& a, const void* o_acc, const void* l_acc, const void* m_acc, void* o,
                             void* l, void* m, cudaStream_t stream) {
  switch (a.dtype) {
    case 0: return finalize_t<__half>(a, o_acc, l_acc, m_acc, o, l, m, stream);
    case 1: return finalize_t<float>(a, o_acc, l_acc, m_acc, o, l, m, stream);
    default: return finalize_t<double>(a, o_acc, l_acc, m_acc, o, l, m, stream);
  }
}

// ---- gradient shards (ring backward): acc (+)= part, and the final cast ----------------------
template <typename T, typename A>
__global__ void grad_accumulate_kernel(const T* __restrict__ part, A* __restrict__ acc, int64_t n, int first) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    acc[i] = (first ? A(0) : acc[i]) + to_acc<A>(part[i]);
}
template <typename T, typename A>
__global__ void grad_finalize_kernel(const A* __restrict__ acc, T* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    out[i] = from_acc<T>(acc[i]);
}

template <typename T, typename A>
static cudaError_t grad_acc_t(const void* part, void* acc, int64_t n, int first, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int blocks = int(std::min<int64_t>((n + 255) / 256, 148 * 16));
  ScopedKernel timed("grad_accumulate", stream);
  grad_accumulate_kernel<T, A><<<blocks, 256, 0, stream>>>((const T*)part, (A*)acc, n, first);
  return cudaGetLastError();
}
template <typename T, typename A>
static cudaError_t grad_fin_t(const void* acc, void* out, int64_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int blocks = int(std::min<int64_t>((n + 255) / 256, 148 * 16));
  ScopedKernel timed("grad_finalize", stream);
  grad_finalize_kernel<T, A><<<blocks, 256, 0, stream>>>((const A*)acc, (T*)out, n);
  return cudaGetLastError();
}

cudaError_t grad_accumulate(int dtype, const void* part, void* acc, int64_t n, int first, cudaStream_t stream) {
  switch (dtype) {
    case 0: return grad_acc_t<__half, float>(part, acc, n, first, stream);
    case 1: return grad_acc_t<float, float>(part, acc, n, first, stream);
    default: return grad_acc_t<double, double>(part, acc, n, first, stream);
  }
}
cudaError_t grad_finalize(int dtype, const void* acc, void* out, int64_t n, cudaStream_t stream) {
  switch (dtype) {
    case 0: return grad_fin_t<__half, float>(acc, out, n, stream);
    case 1: return grad_fin_t<float, float>(acc, out, n, stream);
    default: return grad_fin_t<double, double>(acc, out, n, stream);
  }
}

}  // namespace fa
