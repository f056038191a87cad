// shim_harness.cc — compiles tf_flash_attention_b200/csrc/tf_ops/fa_tf_ops.cc against the TensorFlow API stub
// (tests/tf_stub/tf_stub.h) and drives its OpKernels the way the TensorFlow runtime would:
//   * registry: the 30 ops / 54 GPU kernels of the reference, names, inputs, outputs, attrs;
//   * shape functions (CPU);
//   * (with a GPU) Forward / Backward / Flops kernels on device buffers, compared BIT-FOR-BIT with direct
//     fa_forward / fa_backward / fa_estimate_forward_flops calls on the same inputs, plus the error path.
// TEST INFRASTRUCTURE (built by `make -C tests/tf_stub`, run by tests/test_tf_shim.py).
#define GOOGLE_CUDA 1
#include "../../tf_flash_attention_b200/csrc/tf_ops/fa_tf_ops.cc"

#include <cuda_fp16.h>
#include <stdio.h>
#include <string.h>

#include <random>

using tensorflow::AttrMap;
using tensorflow::AttrValue;
using tensorflow::DataType;
using tensorflow::KernelDef;
using tensorflow::OpKernel;
using tensorflow::OpKernelConstruction;
using tensorflow::OpKernelContext;

static int g_fail = 0;
#define CHECK_MSG(cond, ...)                    \
  do {                                          \
    if (!(cond)) {                              \
      ++g_fail;                                 \
      printf("FAIL %s:%d: ", __FILE__, __LINE__); \
      printf(__VA_ARGS__);                      \
      printf("\n");                             \
    }                                           \
  } while (0)

static const KernelDef* FindKernel(const std::string& op, const char* attr, DataType t) {
  for (const auto& k : tensorflow::KernelRegistry()) {
    auto it = k.constraints.find(attr);
    if (k.op == op && it != k.constraints.end() && it->second == t) return &k;
  }
  return nullptr;
}

static Tensor DeviceTensor(DataType dt, const std::vector<int64_t>& shape, const void* host) {
  Tensor t(dt, TensorShape(shape));
  cudaMemcpy(const_cast<char*>(t.tensor_data().data()), host, t.bytes(), cudaMemcpyHostToDevice);
  return t;
}
static std::vector<char> ToHost(const Tensor& t) {
  std::vector<char> h(t.bytes());
  cudaMemcpy(h.data(), t.tensor_data().data(), t.bytes(), cudaMemcpyDeviceToHost);
  return h;
}
static std::vector<char> ToHost(const void* dev, size_t n) {
  std::vector<char> h(n);
  cudaMemcpy(h.data(), dev, n, cudaMemcpyDeviceToHost);
  return h;
}

static void TestRegistry() {
  const auto& ops = tensorflow::OpRegistry();
  CHECK_MSG(ops.size() == 30, "expected 30 ops, got %zu", ops.size());
  CHECK_MSG(tensorflow::KernelRegistry().size() == 54, "expected 54 kernels, got %zu", tensorflow::KernelRegistry().size());
  for (const char* fam : {"Full", "Causal", "Local"})
    for (const char* n : {"1", "2"}) {
      const std::string f = std::string(fam) + "AttentionForward" + n + "d", b = std::string(fam) + "AttentionBackward" + n + "d";
      for (const std::string& name : {f, f + "Float16", b, b + "Float16", "Estimate" + f + "Flops"})
        CHECK_MSG(ops.count(name) == 1, "op %s not registered", name.c_str());
      const auto& fd = ops.at(f + "Float16");
      CHECK_MSG(fd.inputs.size() == 3 && fd.outputs.size() == 3 && fd.outputs[1] == "l: float", "forward signature of %s", f.c_str());
      const auto& bd = ops.at(b);
      CHECK_MSG(bd.inputs.size() == 7 && bd.outputs.size() == 3 && bd.inputs[6] == "d_o: T", "backward signature of %s", b.c_str());
      const size_t want_attrs = std::string(fam) == "Local" ? 5 : 2;
      CHECK_MSG(fd.attrs.size() == want_attrs, "%s has %zu attrs", f.c_str(), fd.attrs.size());
    }
  for (const auto& k : tensorflow::KernelRegistry()) CHECK_MSG(k.device == "GPU", "kernel of %s not on GPU", k.op.c_str());
}

static void TestShapeFns() {
  using tensorflow::shape_inference::InferenceContext;
  {  // 2-D forward: Q [2,3,16,8,10] K [2,3,16,4,5] V [2,3,24,4,5] -> O [2,3,24,8,10], l,m [2,3,8,10]
    InferenceContext c;
    c.inputs = {InferenceContext::Make({2, 3, 16, 8, 10}), InferenceContext::Make({2, 3, 16, 4, 5}),
                InferenceContext::Make({2, 3, 24, 4, 5})};
    auto st = tensorflow::OpRegistry().at("LocalAttentionForward2d").shape_fn(&c);
    CHECK_MSG(st.ok(), "shape fn failed: %s", st.message().c_str());
    CHECK_MSG(*c.outputs[0].d == (std::vector<int64_t>{2, 3, 24, 8, 10}), "O shape");
    CHECK_MSG(*c.outputs[1].d == (std::vector<int64_t>{2, 3, 8, 10}) && *c.outputs[2].d == *c.outputs[1].d, "l/m shape");
  }
  {  // rank too small -> InvalidArgument
    InferenceContext c;
    c.inputs = {InferenceContext::Make({16, 8}), InferenceContext::Make({16, 8}), InferenceContext::Make({16, 8})};
    auto st = tensorflow::OpRegistry().at("FullAttentionForward1d").shape_fn(&c);
    CHECK_MSG(!st.ok() && st.code() == tensorflow::Code::INVALID_ARGUMENT, "rank check");
  }
  {
    InferenceContext c;
    c.inputs = {InferenceContext::Make({4, 8, 32}), InferenceContext::Make({4, 8, 64}), InferenceContext::Make({4, 6, 64})};
    auto st = tensorflow::OpRegistry().at("CausalAttentionBackward1d").shape_fn(&c);
    CHECK_MSG(st.ok() && *c.outputs[1].d == (std::vector<int64_t>{4, 8, 64}) && *c.outputs[2].d == (std::vector<int64_t>{4, 6, 64}), "backward shapes");
  }
}

template <typename T> static T FromFloat(float v);
template <> float FromFloat<float>(float v) { return v; }
template <> double FromFloat<double>(float v) { return v; }
template <> __half FromFloat<__half>(float v) { return __float2half(v); }

template <typename T>
static void RunOne(const char* fwd_op, const char* bwd_op, DataType dt, int fa_dtype, int rule, int seq_dims,
                   const AttrMap& attrs, const std::vector<int64_t>& qs, const std::vector<int64_t>& ks,
                   const std::vector<int64_t>& vs) {
  auto numel = [](const std::vector<int64_t>& s) { int64_t n = 1; for (auto v : s) n *= v; return n; };
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> U(-2.f, 2.f);
  auto fill = [&](int64_t n) { std::vector<T> h(n); for (auto& x : h) x = FromFloat<T>(U(rng)); return h; };
  std::vector<int64_t> os = vs;
  const int ch = int(qs.size()) - seq_dims - 1;
  os.resize(ch + 1);
  for (size_t i = ch + 1; i < qs.size(); ++i) os.push_back(qs[i]);
  auto hq = fill(numel(qs)), hk = fill(numel(ks)), hv = fill(numel(vs)), hdo = fill(numel(os));
  const KernelDef* kf = FindKernel(fwd_op, "T", dt);
  const KernelDef* kb = FindKernel(bwd_op, "T", dt);
  CHECK_MSG(kf && kb, "kernels %s / %s not found", fwd_op, bwd_op);
  if (!kf || !kb) return;
  const DataType ldt = dt == tensorflow::DT_HALF ? tensorflow::DT_FLOAT : dt;

  // ---- forward through the OpKernel
  OpKernelConstruction cons(attrs);
  std::unique_ptr<OpKernel> fwd(kf->factory(&cons));
  CHECK_MSG(cons.status().ok(), "construction: %s", cons.status().message().c_str());
  OpKernelContext ctx;
  ctx.inputs = {DeviceTensor(dt, qs, hq.data()), DeviceTensor(dt, ks, hk.data()), DeviceTensor(dt, vs, hv.data())};
  ctx.output_types = {dt, ldt, dt};
  fwd->Compute(&ctx);
  CHECK_MSG(ctx.status().ok(), "%s: %s", fwd_op, ctx.status().message().c_str());
  if (!ctx.status().ok()) return;
  cudaDeviceSynchronize();
  CHECK_MSG(ctx.outputs[0]->shape() == TensorShape(os), "O shape from the kernel");

  // ---- the same call straight through the C ABI
  fa_problem_t p = {};
  p.dtype = fa_dtype;
  p.rule = rule;
  p.sync_mode = SyncCode(attrs.at("sync_mode").s);
  p.window_size = attrs.count("window_size") ? int(attrs.at("window_size").i) : 1;
  p.log2_stride_size = attrs.count("log2_stride_size") ? int(attrs.at("log2_stride_size").i) : 0;
  p.is_causal = attrs.count("is_causal") ? attrs.at("is_causal").b : 0;
  int rc = fa_check_forward_shapes(seq_dims, int(qs.size()), qs.data(), int(ks.size()), ks.data(), int(vs.size()), vs.data(), &p);
  CHECK_MSG(rc == FA_OK, "fa_check_forward_shapes rc=%d", rc);
  std::vector<int64_t> lms(qs);
  lms.erase(lms.begin() + ch);
  Tensor o2(dt, TensorShape(os)), l2(ldt, TensorShape(lms)), m2(dt, TensorShape(lms));
  const size_t wf = fa_workspace_bytes(&p, 0), wb = fa_workspace_bytes(&p, 1);
  Tensor ws(tensorflow::DT_UINT8, TensorShape({int64_t(std::max(wf, wb)) + 16}));
  auto dp = [](const Tensor& t) { return const_cast<char*>(t.tensor_data().data()); };
  rc = fa_forward(&p, dp(ctx.inputs[0]), dp(ctx.inputs[1]), dp(ctx.inputs[2]), dp(o2), dp(l2), dp(m2), dp(ws), wf, nullptr);
  CHECK_MSG(rc == FA_OK, "fa_forward rc=%d", rc);
  cudaDeviceSynchronize();
  CHECK_MSG(ToHost(*ctx.outputs[0]) == ToHost(o2), "%s: O differs from the direct C-ABI call", fwd_op);
  CHECK_MSG(ToHost(*ctx.outputs[1]) == ToHost(l2), "%s: l differs", fwd_op);
  CHECK_MSG(ToHost(*ctx.outputs[2]) == ToHost(m2), "%s: m differs", fwd_op);

  // ---- backward through the OpKernel vs the C ABI (dK, dV bit-equal; dQ of the fused fp16 kernel to rounding)
  OpKernelConstruction cons_b(attrs);
  std::unique_ptr<OpKernel> bwd(kb->factory(&cons_b));
  OpKernelContext bctx;
  bctx.inputs = {ctx.inputs[0], ctx.inputs[1], ctx.inputs[2], *ctx.outputs[0], *ctx.outputs[1], *ctx.outputs[2],
                 DeviceTensor(dt, os, hdo.data())};
  bctx.output_types = {dt, dt, dt};
  bwd->Compute(&bctx);
  CHECK_MSG(bctx.status().ok(), "%s: %s", bwd_op, bctx.status().message().c_str());
  if (!bctx.status().ok()) return;
  Tensor dq2(dt, TensorShape(qs)), dk2(dt, TensorShape(ks)), dv2(dt, TensorShape(vs));
  rc = fa_backward(&p, dp(ctx.inputs[0]), dp(ctx.inputs[1]), dp(ctx.inputs[2]), dp(o2), dp(l2), dp(m2), dp(bctx.inputs[6]),
                   dp(dq2), dp(dk2), dp(dv2), dp(ws), wb, nullptr);
  CHECK_MSG(rc == FA_OK, "fa_backward rc=%d", rc);
  cudaDeviceSynchronize();
  CHECK_MSG(ToHost(*bctx.outputs[1]) == ToHost(dk2), "%s: dK differs", bwd_op);
  CHECK_MSG(ToHost(*bctx.outputs[2]) == ToHost(dv2), "%s: dV differs", bwd_op);
  {
    auto a = ToHost(*bctx.outputs[0]), b = ToHost(dq2);
    double worst = 0;
    const T* pa = reinterpret_cast<const T*>(a.data());
    const T* pb = reinterpret_cast<const T*>(b.data());
    for (int64_t i = 0; i < numel(qs); ++i) {
      const double x = double(float(pa[i])), y = double(float(pb[i]));
      worst = std::max(worst, std::abs(x - y) / std::max(1.0, std::abs(y)));
    }
    CHECK_MSG(worst <= 2e-3, "%s: dQ differs by %g", bwd_op, worst);
  }
  printf("ok   %-34s %-34s q=%lld k=%lld\n", fwd_op, bwd_op, (long long)qs.back(), (long long)ks.back());
}

static void TestKernelsOnGpu() {
  AttrMap full;
  full["sync_mode"].s = "scale_end";
  AttrMap causal;
  causal["sync_mode"].s = "none_front";
  AttrMap local = causal;
  local["sync_mode"].s = "scale_front";
  local["window_size"].i = 5;
  local["log2_stride_size"].i = 1;
  local["is_causal"].b = true;
  RunOne<__half>("CausalAttentionForward1dFloat16", "CausalAttentionBackward1dFloat16", tensorflow::DT_HALF, FA_F16,
                 FA_RULE_CAUSAL, 1, causal, {2, 2, 128, 512}, {2, 2, 128, 512}, {2, 2, 128, 512});
  RunOne<__half>("FullAttentionForward1dFloat16", "FullAttentionBackward1dFloat16", tensorflow::DT_HALF, FA_F16,
                 FA_RULE_FULL, 1, full, {3, 64, 200}, {3, 64, 328}, {3, 64, 328});
  RunOne<float>("LocalAttentionForward2d", "LocalAttentionBackward2d", tensorflow::DT_FLOAT, FA_F32, FA_RULE_LOCAL, 2,
                local, {2, 32, 12, 16}, {2, 32, 24, 16}, {2, 16, 24, 16});
  RunOne<double>("CausalAttentionForward1d", "CausalAttentionBackward1d", tensorflow::DT_DOUBLE, FA_F64,
                 FA_RULE_CAUSAL, 1, causal, {2, 16, 96}, {2, 16, 96}, {2, 24, 96});

  {  // error path: K and V sequence shapes differ -> InvalidArgument with the reference's message class
    const KernelDef* kf = FindKernel("FullAttentionForward1d", "T", tensorflow::DT_FLOAT);
    OpKernelConstruction cons(full);
    std::unique_ptr<OpKernel> k(kf->factory(&cons));
    OpKernelContext ctx;
    ctx.inputs = {Tensor(tensorflow::DT_FLOAT, TensorShape({2, 8, 32})), Tensor(tensorflow::DT_FLOAT, TensorShape({2, 8, 64})),
                  Tensor(tensorflow::DT_FLOAT, TensorShape({2, 8, 48}))};
    ctx.output_types = {tensorflow::DT_FLOAT, tensorflow::DT_FLOAT, tensorflow::DT_FLOAT};
    k->Compute(&ctx);
    CHECK_MSG(!ctx.status().ok() && ctx.status().code() == tensorflow::Code::INVALID_ARGUMENT, "shape mismatch must be InvalidArgument");
    printf("ok   error path: %s\n", ctx.status().message().c_str());
  }
  {  // unsupported sync_mode is rejected at construction (forward.cc:275)
    AttrMap bad;
    bad["sync_mode"].s = "middle";
    const KernelDef* kf = FindKernel("FullAttentionForward1d", "T", tensorflow::DT_FLOAT);
    OpKernelConstruction cons(bad);
    std::unique_ptr<OpKernel> k(kf->factory(&cons));
    CHECK_MSG(!cons.status().ok() && cons.status().message().find("Unsupported sync_mode") != std::string::npos, "sync_mode check");
  }
  {  // flops op: host-memory output equals the direct call
    AttrMap a = causal;
    a["q_shape"].shape = TensorShape({16, 16, 128, 8192});
    a["k_shape"].shape = TensorShape({16, 16, 128, 8192});
    a["v_shape"].shape = TensorShape({16, 16, 128, 8192});
    const KernelDef* kf = FindKernel("EstimateCausalAttentionForward1dFlops", "dtype", tensorflow::DT_HALF);
    CHECK_MSG(kf && kf->host_memory.size() == 1 && kf->host_memory[0] == "flops", "flops kernel def");
    OpKernelConstruction cons(a);
    std::unique_ptr<OpKernel> k(kf->factory(&cons));
    OpKernelContext ctx;
    ctx.output_types = {tensorflow::DT_FLOAT};
    ctx.output_on_host = {true};
    k->Compute(&ctx);
    CHECK_MSG(ctx.status().ok(), "flops op: %s", ctx.status().message().c_str());
    fa_problem_t p = {};
    p.dtype = FA_F16;
    p.rule = FA_RULE_CAUSAL;
    p.window_size = 1;
    const int64_t s[4] = {16, 16, 128, 8192};
    fa_check_forward_shapes(1, 4, s, 4, s, 4, s, &p);
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    float want = 0.f;
    fa_estimate_forward_flops(&p, optin, &want);
    if (ctx.status().ok()) {
      const float got = ctx.outputs[0]->flat<float>()(0);
      CHECK_MSG(got == want && got > 0.f, "flops %g vs %g", got, want);
      printf("ok   EstimateCausalAttentionForward1dFlops = %.6g\n", got);
    }
  }
}

int main(int argc, char** argv) {
  const bool no_gpu = argc > 1 && !strcmp(argv[1], "--no-gpu");
  TestRegistry();
  TestShapeFns();
  printf("registry: %zu ops, %zu kernels; shape functions checked\n", tensorflow::OpRegistry().size(),
         tensorflow::KernelRegistry().size());
  if (!no_gpu) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      printf("no CUDA device: run with --no-gpu for the host-only checks\n");
      return 2;
    }
    TestKernelsOnGpu();
  }
  printf(g_fail ? "TF_SHIM_HARNESS FAIL (%d)\n" : "TF_SHIM_HARNESS PASS\n", g_fail);
  return g_fail ? 1 : 0;
}
