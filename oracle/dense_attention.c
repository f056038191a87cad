/* TEST INFRASTRUCTURE ONLY — plain C restatement of the reference's algorithm for the hot path.
 * Nothing here is linked into or called by the product (libfa_b200.so); only tests/ use it, as a
 * second, independent checker next to oracle/dense_attention.py.
 *
 *   oc_pattern   attended-index pattern, tests' formulation:
 *                  locations per sync mode   flash_attention/tests/test_1d.py:9-50, test_2d.py:11-78
 *                  masks                     flash_attention/tests/test_base.py:18-67
 *   oc_forward   dense masked softmax attention  flash_attention/tests/test_1d.py:69-76
 *   oc_backward  closed-form gradients           flash_attention/kernel/internal_test.cu:413-511
 * Tensors: channel-first, double precision, Q [B,d,q] K [B,d,k] V [B,vd,k] dO [B,vd,q]; mask [q,k].
 * Pinning: tests/test_oracle_c.py checks oc_pattern against tests/golden/pattern_golden.json (produced
 * by the reference's own host code) and oc_forward / oc_backward against the NumPy oracle, which is
 * itself pinned against the reference CUDA kernel (tests/golden/refkernel_golden.npz). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static void locations(int dims, int sync, const int32_t* shape, const int32_t* other, int64_t n,
                      int64_t* coord0, int64_t* coord1, int64_t* lin) {
  /* shape in TF order (outer, inner); 1-D uses shape[0] */
  int64_t n_in = dims == 1 ? shape[0] : shape[1];
  int64_t o_in = dims == 1 ? other[0] : other[1];
  int64_t max_in = n_in > o_in ? n_in : o_in;
  int64_t max_out = 1, n_out = 1;
  if (dims == 2) {
    n_out = shape[0];
    max_out = shape[0] > other[0] ? shape[0] : other[0];
  }
  int64_t step_in = sync == 0 ? 1 : max_in / n_in;
  int64_t step_out = (sync == 0 || dims == 1) ? 1 : max_out / n_out;
  for (int64_t i = 0; i < n; ++i) {
    int64_t y = dims == 1 ? 0 : i / n_in, x = dims == 1 ? i : i % n_in;
    int64_t cx = sync == 2 ? (x + 1) * step_in - 1 : x * step_in;
    int64_t cy = dims == 1 ? 0 : (sync == 2 ? (y + 1) * step_out - 1 : y * step_out);
    coord0[i] = cx;
    coord1[i] = cy;
    lin[i] = dims == 1 ? cx : cy * max_in + cx; /* l = y * max_width + x (test_2d.py:22) */
  }
}

int oc_pattern(int dims, int rule, int sync, int window, int log2_stride, int is_causal,
               const int32_t* q_shape, const int32_t* k_shape, uint8_t* mask) {
  if (dims < 1 || dims > 2 || rule < 0 || rule > 2 || sync < 0 || sync > 2) return -1;
  int64_t q = q_shape[0] * (dims == 2 ? (int64_t)q_shape[1] : 1);
  int64_t k = k_shape[0] * (dims == 2 ? (int64_t)k_shape[1] : 1);
  int64_t* buf = (int64_t*)malloc(sizeof(int64_t) * 3 * (q + k));
  if (!buf) return -2;
  int64_t *qx = buf, *qy = buf + q, *ql = buf + 2 * q;
  int64_t *kx = buf + 3 * q, *ky = kx + k, *kl = kx + 2 * k;
  locations(dims, sync, q_shape, k_shape, q, qx, qy, ql);
  locations(dims, sync, k_shape, q_shape, k, kx, ky, kl);
  const int64_t stride = (int64_t)1 << log2_stride;
  for (int64_t i = 0; i < q; ++i)
    for (int64_t j = 0; j < k; ++j) {
      int ok = 1;
      if (rule == 1) ok = ql[i] - kl[j] >= 0;
      if (rule == 2) {
        int64_t dx = llabs(qx[i] - kx[j]), dy = llabs(qy[i] - ky[j]);
        ok = (dx % stride == 0) && (dx / stride < window);
        if (dims == 2) ok = ok && (dy % stride == 0) && (dy / stride < window);
        if (is_causal) ok = ok && (ql[i] - kl[j] >= 0);
      }
      mask[i * k + j] = (uint8_t)ok;
    }
  free(buf);
  return 0;
}

/* P [q,k] row-normalised probabilities of one batch element; l, m per row (m = -inf on empty rows) */
static void softmax_rows(int d, int64_t q, int64_t k, const double* Q, const double* K, const uint8_t* mask,
                         double* P, double* l, double* m) {
  const double scale = 1.0 / sqrt((double)d);
  for (int64_t i = 0; i < q; ++i) {
    double mx = -INFINITY;
    for (int64_t j = 0; j < k; ++j) {
      double s = 0.0;
      for (int c = 0; c < d; ++c) s += Q[c * q + i] * K[c * k + j];
      s *= scale;
      P[i * k + j] = s;
      if (mask[i * k + j] && s > mx) mx = s;
    }
    double sum = 0.0;
    for (int64_t j = 0; j < k; ++j) {
      double e = mask[i * k + j] ? exp(P[i * k + j] - mx) : 0.0;
      P[i * k + j] = e;
      sum += e;
    }
    for (int64_t j = 0; j < k; ++j) P[i * k + j] = sum > 0 ? P[i * k + j] / sum : 0.0;
    l[i] = sum;
    m[i] = mx;
  }
}

int oc_forward(int64_t B, int d, int vd, int64_t q, int64_t k, const double* Q, const double* K, const double* V,
               const uint8_t* mask, double* O, double* l, double* m) {
  double* P = (double*)malloc(sizeof(double) * q * k);
  if (!P) return -2;
  for (int64_t b = 0; b < B; ++b) {
    softmax_rows(d, q, k, Q + b * d * q, K + b * d * k, mask, P, l + b * q, m + b * q);
    for (int c = 0; c < vd; ++c)
      for (int64_t i = 0; i < q; ++i) {
        double acc = 0.0;
        for (int64_t j = 0; j < k; ++j) acc += P[i * k + j] * V[(b * vd + c) * k + j];
        O[(b * vd + c) * q + i] = acc;
      }
  }
  free(P);
  return 0;
}

int oc_backward(int64_t B, int d, int vd, int64_t q, int64_t k, const double* Q, const double* K, const double* V,
                const double* dO, const uint8_t* mask, double* dQ, double* dK, double* dV) {
  double* P = (double*)malloc(sizeof(double) * q * k);
  double* dS = (double*)malloc(sizeof(double) * q * k);
  double* lm = (double*)malloc(sizeof(double) * 3 * q);
  if (!P || !dS || !lm) return -2;
  const double scale = 1.0 / sqrt((double)d);
  for (int64_t b = 0; b < B; ++b) {
    const double *Qb = Q + b * d * q, *Kb = K + b * d * k, *Vb = V + b * vd * k, *dOb = dO + b * vd * q;
    softmax_rows(d, q, k, Qb, Kb, mask, P, lm, lm + q);
    double* D = lm + 2 * q;
    for (int64_t i = 0; i < q; ++i) {       /* D = rowsum(dO o O), O = P V */
      double acc = 0.0;
      for (int c = 0; c < vd; ++c) {
        double o = 0.0;
        for (int64_t j = 0; j < k; ++j) o += P[i * k + j] * Vb[c * k + j];
        acc += o * dOb[c * q + i];
      }
      D[i] = acc;
    }
    for (int c = 0; c < vd; ++c)              /* dV = P^T dO */
      for (int64_t j = 0; j < k; ++j) {
        double acc = 0.0;
        for (int64_t i = 0; i < q; ++i) acc += P[i * k + j] * dOb[c * q + i];
        dV[(b * vd + c) * k + j] = acc;
      }
    for (int64_t i = 0; i < q; ++i)           /* dS = P o (dO V^T - D) / sqrt(d) */
      for (int64_t j = 0; j < k; ++j) {
        double dp = 0.0;
        for (int c = 0; c < vd; ++c) dp += dOb[c * q + i] * Vb[c * k + j];
        dS[i * k + j] = P[i * k + j] * (dp - D[i]) * scale;
      }
    for (int c = 0; c < d; ++c) {
      for (int64_t i = 0; i < q; ++i) {       /* dQ = dS K */
        double acc = 0.0;
        for (int64_t j = 0; j < k; ++j) acc += dS[i * k + j] * Kb[c * k + j];
        dQ[(b * d + c) * q + i] = acc;
      }
      for (int64_t j = 0; j < k; ++j) {       /* dK = dS^T Q */
        double acc = 0.0;
        for (int64_t i = 0; i < q; ++i) acc += dS[i * k + j] * Qb[c * q + i];
        dK[(b * d + c) * k + j] = acc;
      }
    }
  }
  free(P); free(dS); free(lm);
  return 0;
}
