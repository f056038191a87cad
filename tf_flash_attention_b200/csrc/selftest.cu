// selftest.cu — standalone (no TensorFlow, no Python) self-test of libfa_b200.so, the counterpart of the
// reference's kernel/internal_test.cu (`make INTERNAL_TEST=<type>`): full attention on 1-D sequences,
// b=1, h=8, q=k=1024, d=v_d=32 (internal_test.cu:87-96), a naive CPU forward (:135-233) and backward
// (:381-513) in double, the library through its C ABI on the GPU, the error RATE above Precision<T>
// (1e-2 half / 1e-6 float / 1e-9 double, errors normalised by the reduced length, :15-28) and the kernel
// time from CUDA events (:31-66). Unlike the reference it also returns non-zero when the rate is not 0,
// and it additionally runs the tcgen05 shape (d = 128, causal) for half.
//   make -C tf_flash_attention_b200/csrc selftest && ./tf_flash_attention_b200/csrc/build/selftest
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/fa_b200.h"

template <typename T> struct Traits;
template <> struct Traits<__half> { static constexpr int code = FA_F16; static constexpr double prec = 1e-2; using L = float; static const char* name() { return "half"; } };
template <> struct Traits<float> { static constexpr int code = FA_F32; static constexpr double prec = 1e-6; using L = float; static const char* name() { return "float"; } };
template <> struct Traits<double> { static constexpr int code = FA_F64; static constexpr double prec = 1e-9; using L = double; static const char* name() { return "double"; } };

static double to_d(__half v) { return double(__half2float(v)); }
static double to_d(float v) { return v; }
static double to_d(double v) { return v; }
template <typename T> static T from_d(double v) { return T(v); }
template <> __half from_d<__half>(double v) { return __float2half(float(v)); }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

template <typename T>
int run(int B, int d, int vd, int q, int k, int rule) {
  using L = typename Traits<T>::L;
  const size_t nq = size_t(B) * d * q, nk = size_t(B) * d * k, nv = size_t(B) * vd * k, no = size_t(B) * vd * q;
  std::vector<T> Q(nq), K(nk), V(nv), dO(no);
  std::vector<double> Qd(nq), Kd(nk), Vd(nv), dOd(no);
  srand(1234);
  auto fill = [](std::vector<T>& a, std::vector<double>& b) {
    for (size_t i = 0; i < a.size(); ++i) { a[i] = from_d<T>(4.0 * rand() / RAND_MAX - 2.0); b[i] = to_d(a[i]); }
  };
  fill(Q, Qd); fill(K, Kd); fill(V, Vd); fill(dO, dOd);
  // naive CPU forward + backward in double
  std::vector<double> O(no, 0.0), dQ(nq, 0.0), dK(nk, 0.0), dV(nv, 0.0), P(size_t(q) * k), dS(size_t(q) * k);
  const double scale = 1.0 / std::sqrt(double(d));
  for (int b = 0; b < B; ++b) {
    const double *Qb = &Qd[size_t(b) * d * q], *Kb = &Kd[size_t(b) * d * k], *Vb = &Vd[size_t(b) * vd * k], *dOb = &dOd[size_t(b) * vd * q];
    for (int i = 0; i < q; ++i) {
      double mx = -INFINITY, sum = 0;
      for (int j = 0; j < k; ++j) {
        double s = 0;
        for (int c = 0; c < d; ++c) s += Qb[size_t(c) * q + i] * Kb[size_t(c) * k + j];
        s *= scale;
        const bool ok = rule == FA_RULE_FULL || i >= j;
        P[size_t(i) * k + j] = ok ? s : -INFINITY;
        if (ok && s > mx) mx = s;
      }
      for (int j = 0; j < k; ++j) { double e = std::exp(P[size_t(i) * k + j] - mx); P[size_t(i) * k + j] = e; sum += e; }
      for (int j = 0; j < k; ++j) P[size_t(i) * k + j] /= sum;
    }
    for (int c = 0; c < vd; ++c)
      for (int i = 0; i < q; ++i) { double a = 0; for (int j = 0; j < k; ++j) a += P[size_t(i) * k + j] * Vb[size_t(c) * k + j]; O[(size_t(b) * vd + c) * q + i] = a; }
    for (int i = 0; i < q; ++i) {
      double D = 0;
      for (int c = 0; c < vd; ++c) D += O[(size_t(b) * vd + c) * q + i] * dOb[size_t(c) * q + i];
      for (int j = 0; j < k; ++j) { double dp = 0; for (int c = 0; c < vd; ++c) dp += dOb[size_t(c) * q + i] * Vb[size_t(c) * k + j]; dS[size_t(i) * k + j] = P[size_t(i) * k + j] * (dp - D) * scale; }
    }
    for (int c = 0; c < vd; ++c)
      for (int j = 0; j < k; ++j) { double a = 0; for (int i = 0; i < q; ++i) a += P[size_t(i) * k + j] * dOb[size_t(c) * q + i]; dV[(size_t(b) * vd + c) * k + j] = a; }
    for (int c = 0; c < d; ++c) {
      for (int i = 0; i < q; ++i) { double a = 0; for (int j = 0; j < k; ++j) a += dS[size_t(i) * k + j] * Kb[size_t(c) * k + j]; dQ[(size_t(b) * d + c) * q + i] = a; }
      for (int j = 0; j < k; ++j) { double a = 0; for (int i = 0; i < q; ++i) a += dS[size_t(i) * k + j] * Qb[size_t(c) * q + i]; dK[(size_t(b) * d + c) * k + j] = a; }
    }
  }
  // library on the GPU through the C ABI
  fa_problem_t p = {};
  p.dtype = Traits<T>::code; p.seq_dims = 1; p.rule = rule; p.window_size = 1; p.sync_mode = FA_SYNC_NONE_FRONT;
  p.d = d; p.v_d = vd; p.batch = B; p.q_shape[0] = q; p.k_shape[0] = k;
  T *gQ, *gK, *gV, *gO, *gdO, *gdQ, *gdK, *gdV, *gm; L* gl; void* ws;
  CK(cudaMalloc(&gQ, nq * sizeof(T))); CK(cudaMalloc(&gK, nk * sizeof(T))); CK(cudaMalloc(&gV, nv * sizeof(T)));
  CK(cudaMalloc(&gO, no * sizeof(T))); CK(cudaMalloc(&gdO, no * sizeof(T))); CK(cudaMalloc(&gdQ, nq * sizeof(T)));
  CK(cudaMalloc(&gdK, nk * sizeof(T))); CK(cudaMalloc(&gdV, nv * sizeof(T))); CK(cudaMalloc(&gm, size_t(B) * q * sizeof(T)));
  CK(cudaMalloc(&gl, size_t(B) * q * sizeof(L)));
  const size_t wsb = fa_workspace_bytes(&p, 1) + 16;
  CK(cudaMalloc(&ws, wsb));
  CK(cudaMemcpy(gQ, Q.data(), nq * sizeof(T), cudaMemcpyHostToDevice)); CK(cudaMemcpy(gK, K.data(), nk * sizeof(T), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(gV, V.data(), nv * sizeof(T), cudaMemcpyHostToDevice)); CK(cudaMemcpy(gdO, dO.data(), no * sizeof(T), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  int rc = 0;
  for (int it = 0; it < 2; ++it) {  // second pass is the timed one
    cudaEventRecord(e0);
    rc = fa_forward(&p, gQ, gK, gV, gO, gl, gm, ws, wsb, nullptr);
    cudaEventRecord(e1);
    if (!rc) rc = fa_backward(&p, gQ, gK, gV, gO, gl, gm, gdO, gdQ, gdK, gdV, ws, wsb, nullptr);
    cudaEventRecord(e2);
    if (rc) { printf("library error: %s\n", fa_strerror(rc)); return 1; }
    CK(cudaDeviceSynchronize());
  }
  float tf, tb; cudaEventElapsedTime(&tf, e0, e1); cudaEventElapsedTime(&tb, e1, e2);
  const int path = fa_last_path();
  std::vector<T> hO(no), hdQ(nq), hdK(nk), hdV(nv);
  CK(cudaMemcpy(hO.data(), gO, no * sizeof(T), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hdQ.data(), gdQ, nq * sizeof(T), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hdK.data(), gdK, nk * sizeof(T), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hdV.data(), gdV, nv * sizeof(T), cudaMemcpyDeviceToHost));
  auto rate = [](const std::vector<T>& got, const std::vector<double>& ref, double norm, double prec, double* worst) {
    size_t bad = 0; *worst = 0;
    for (size_t i = 0; i < got.size(); ++i) { double e = std::fabs(to_d(got[i]) - ref[i]) / norm; if (e > *worst) *worst = e; if (!(e <= prec)) ++bad; }
    return double(bad) / got.size();
  };
  double w[4];
  // normalisation as in internal_test.cu: forward / dQ by the key length, dK / dV by the query length
  const double r0 = rate(hO, O, 1.0, Traits<T>::prec, &w[0]), r1 = rate(hdQ, dQ, std::sqrt(double(k)), Traits<T>::prec, &w[1]);
  const double r2 = rate(hdK, dK, std::sqrt(double(q)), Traits<T>::prec, &w[2]), r3 = rate(hdV, dV, std::sqrt(double(q)), Traits<T>::prec, &w[3]);
  printf("%-6s %-6s b=%d q=%d k=%d d=%d v_d=%d path=%d | error rate O %.4f dQ %.4f dK %.4f dV %.4f | worst %.2e %.2e %.2e %.2e | fwd %.3f ms bwd %.3f ms\n",
         Traits<T>::name(), rule == FA_RULE_FULL ? "full" : "causal", B, q, k, d, vd, path, r0, r1, r2, r3, w[0], w[1], w[2], w[3], tf, tb);
  cudaFree(gQ); cudaFree(gK); cudaFree(gV); cudaFree(gO); cudaFree(gdO); cudaFree(gdQ); cudaFree(gdK); cudaFree(gdV); cudaFree(gm); cudaFree(gl); cudaFree(ws);
  return (r0 + r1 + r2 + r3) > 0 ? 1 : 0;
}

int main() {
  printf("%s\n", fa_version());
  int fail = 0;
  fail += run<__half>(8, 32, 32, 1024, 1024, FA_RULE_FULL);   // the reference's internal_test shape
  fail += run<float>(8, 32, 32, 1024, 1024, FA_RULE_FULL);
  fail += run<double>(8, 32, 32, 1024, 1024, FA_RULE_FULL);
  fail += run<__half>(4, 128, 128, 1024, 1024, FA_RULE_CAUSAL);  // tcgen05 path
  printf(fail ? "SELFTEST FAILED\n" : "SELFTEST PASSED\n");
  return fail;
}
