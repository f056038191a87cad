// fa_tf_ops.cc — TensorFlow custom-op shim over the C ABI of include/fa_b200.h.
//
// Registers the SAME 30 ops, attrs, dtypes, shape functions and GPU kernels as the reference
//   forward + flops ops : flash_attention/kernel/flash_attention_forward.cc:144-253, 548-591
//   backward ops        : flash_attention/kernel/flash_attention_backward.cc:51-154, 385-403
// so that the reference's flash_attention/flash_attention.py (tf.load_op_library + RegisterGradient)
// works unchanged on top of libfa_b200.so. The OpKernels only validate, allocate and call
// fa_forward / fa_backward / fa_estimate_forward_flops on TensorFlow's stream; all arithmetic lives
// behind the C ABI. TensorFlow is not installed in the build image, so this file is compiled only by
// `make tf_shim` where `import tensorflow` works; everything it calls is
// exercised by the ctypes tests.
#ifdef GOOGLE_CUDA
#define EIGEN_USE_GPU

#include <string>

#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/framework/tensor.h"
#include "tensorflow/core/framework/tensor_shape.h"
#include "tensorflow/core/lib/core/errors.h"
#include "third_party/eigen3/unsupported/Eigen/CXX11/Tensor"

#include "../../../include/fa_b200.h"

using namespace tensorflow;
using GPUDevice = Eigen::GpuDevice;

namespace {

// ---- shape functions (forward.cc:37-95, backward.cc:36-47) -------------------------------------
template <int N>
Status InferForward(shape_inference::InferenceContext* c) {
  auto Q = c->input(0), K = c->input(1), V = c->input(2);
  const int rq = c->Rank(Q), rk = c->Rank(K), rv = c->Rank(V);
  if (!(rq == rk && rk == rv && rq >= N + 2))
    return errors::InvalidArgument("Failed to infer the shape of outputs as the shape of some inputs might be incorrect");
  const int ch = rq - N - 1;
  shape_inference::ShapeHandle a, b, o, lm;
  TF_RETURN_IF_ERROR(c->Subshape(V, 0, ch + 1, &a));
  TF_RETURN_IF_ERROR(c->Subshape(Q, ch + 1, &b));
  TF_RETURN_IF_ERROR(c->Concatenate(a, b, &o));
  TF_RETURN_IF_ERROR(c->Subshape(Q, 0, ch, &a));
  TF_RETURN_IF_ERROR(c->Concatenate(a, b, &lm));
  c->set_output(0, o);
  c->set_output(1, lm);
  c->set_output(2, lm);
  return OkStatus();
}
Status InferBackward(shape_inference::InferenceContext* c) {
  c->set_output(0, c->input(0));
  c->set_output(1, c->input(1));
  c->set_output(2, c->input(2));
  return OkStatus();
}
Status InferFlops(shape_inference::InferenceContext* c) {
  c->set_output(0, c->Scalar());
  return OkStatus();
}

// ---- helpers ------------------------------------------------------------------------------------
template <typename T> struct DType;
template <> struct DType<Eigen::half> { static constexpr int code = FA_F16; };
template <> struct DType<float> { static constexpr int code = FA_F32; };
template <> struct DType<double> { static constexpr int code = FA_F64; };

int SyncCode(const std::string& s) {
  if (s == "none_front") return FA_SYNC_NONE_FRONT;
  if (s == "scale_front") return FA_SYNC_SCALE_FRONT;
  if (s == "scale_end") return FA_SYNC_SCALE_END;
  return -1;
}

std::vector<int64_t> Dims(const TensorShape& s) {
  std::vector<int64_t> d(s.dims());
  for (int i = 0; i < s.dims(); ++i) d[i] = s.dim_size(i);
  return d;
}

// maps a C-ABI status to the reference's TF status classes and message texts
Status ToStatus(int rc, const char* what) {
  if (rc == FA_OK) return OkStatus();
  switch (rc) {
    case FA_EINVAL_RANK:
      return errors::InvalidArgument("The number of dimensions of the inputs should be equal and >= sequence dims + 2");
    case FA_EINVAL_CHANNEL:
      return errors::InvalidArgument("The channel dimension of Q and K (and of V and O) should be equal");
    case FA_EINVAL_BATCH:
      return errors::InvalidArgument("The batch shape of all inputs should be equal");
    case FA_EINVAL_SEQ_SHAPE:
      return errors::InvalidArgument("The sequence shapes of K and V (and of Q, O, l, m, dO) are expected to be equal");
    case FA_ECUDA:
      return errors::Internal("Failed to launch the ", what, " kernel: ",
                              cudaGetErrorString(static_cast<cudaError_t>(fa_last_cuda_error())), "(",
                              fa_last_cuda_error(), ")");
    default:
      if (rc > -100) return errors::InvalidArgument(what, ": ", fa_strerror(rc));
      return errors::Internal(what, ": ", fa_strerror(rc));
  }
}

struct RuleAttrs {
  int rule = FA_RULE_FULL, window = 1, log2_stride = 0, causal = 0, sync = 0;
};

template <int Rule>
void ReadRuleAttrs(OpKernelConstruction* cons, RuleAttrs* a) {
  a->rule = Rule;
  std::string sync_mode;
  OP_REQUIRES_OK(cons, cons->GetAttr("sync_mode", &sync_mode));
  a->sync = SyncCode(sync_mode);
  OP_REQUIRES(cons, a->sync >= 0, errors::InvalidArgument("Unsupported sync_mode: ", sync_mode));  // forward.cc:275
  if (Rule == FA_RULE_LOCAL) {
    bool is_causal = false;
    OP_REQUIRES_OK(cons, cons->GetAttr("window_size", &a->window));
    OP_REQUIRES_OK(cons, cons->GetAttr("log2_stride_size", &a->log2_stride));
    OP_REQUIRES_OK(cons, cons->GetAttr("is_causal", &is_causal));
    a->causal = is_causal;
  }
}

void FillRule(const RuleAttrs& a, int dtype, fa_problem_t* p) {
  p->dtype = dtype;
  p->rule = a.rule;
  p->window_size = a.window;
  p->log2_stride_size = a.log2_stride;
  p->is_causal = a.causal;
  p->sync_mode = a.sync;
}

// ---- forward (forward.cc:255-387) -----------------------------------------------------------------
template <typename T, int N, int Rule>
class ForwardOp : public OpKernel {
 public:
  explicit ForwardOp(OpKernelConstruction* cons) : OpKernel(cons) { ReadRuleAttrs<Rule>(cons, &attrs_); }

  void Compute(OpKernelContext* ctx) override {
    const Tensor &Q = ctx->input(0), &K = ctx->input(1), &V = ctx->input(2);
    fa_problem_t p = {};
    FillRule(attrs_, DType<T>::code, &p);
    auto qd = Dims(Q.shape()), kd = Dims(K.shape()), vd = Dims(V.shape());
    OP_REQUIRES_OK(ctx, ToStatus(fa_check_forward_shapes(N, Q.dims(), qd.data(), K.dims(), kd.data(), V.dims(),
                                                         vd.data(), &p), "Forward"));
    const int ch = Q.dims() - N - 1;
    TensorShape o_shape = V.shape(), lm_shape = Q.shape();
    o_shape.RemoveDimRange(ch + 1, V.dims());
    lm_shape.RemoveDimRange(ch, ch + 1);
    for (int i = ch + 1; i < Q.dims(); ++i) o_shape.AddDim(Q.dim_size(i));
    Tensor *O = nullptr, *l = nullptr, *m = nullptr, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, o_shape, &O));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, lm_shape, &l));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, lm_shape, &m));
    const size_t ws_bytes = fa_workspace_bytes(&p, 0);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_UINT8, TensorShape({static_cast<int64_t>(ws_bytes) + 16}), &ws));
    auto stream = ctx->eigen_device<GPUDevice>().stream();
    const int rc = fa_forward(&p, Q.tensor_data().data(), K.tensor_data().data(), V.tensor_data().data(),
                              const_cast<char*>(O->tensor_data().data()), const_cast<char*>(l->tensor_data().data()),
                              const_cast<char*>(m->tensor_data().data()), const_cast<char*>(ws.tensor_data().data()),
                              ws_bytes, stream);
    OP_REQUIRES_OK(ctx, ToStatus(rc, "Forward"));
  }

 private:
  RuleAttrs attrs_;
};

// ---- backward (backward.cc:156-345) ---------------------------------------------------------------
template <typename T, int N, int Rule>
class BackwardOp : public OpKernel {
 public:
  explicit BackwardOp(OpKernelConstruction* cons) : OpKernel(cons) { ReadRuleAttrs<Rule>(cons, &attrs_); }

  void Compute(OpKernelContext* ctx) override {
    const Tensor* in[7];
    std::vector<int64_t> dims[7];
    for (int i = 0; i < 7; ++i) {
      in[i] = &ctx->input(i);
      dims[i] = Dims(in[i]->shape());
    }
    fa_problem_t p = {};
    FillRule(attrs_, DType<T>::code, &p);
    OP_REQUIRES_OK(ctx, ToStatus(fa_check_backward_shapes(
                                     N, in[0]->dims(), dims[0].data(), in[1]->dims(), dims[1].data(), in[2]->dims(),
                                     dims[2].data(), in[3]->dims(), dims[3].data(), in[4]->dims(), dims[4].data(),
                                     in[5]->dims(), dims[5].data(), in[6]->dims(), dims[6].data(), &p),
                                 "Backward"));
    Tensor *dQ = nullptr, *dK = nullptr, *dV = nullptr, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, in[0]->shape(), &dQ));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, in[1]->shape(), &dK));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, in[2]->shape(), &dV));
    const size_t ws_bytes = fa_workspace_bytes(&p, 1);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_UINT8, TensorShape({static_cast<int64_t>(ws_bytes) + 16}), &ws));
    auto stream = ctx->eigen_device<GPUDevice>().stream();
    const int rc = fa_backward(&p, in[0]->tensor_data().data(), in[1]->tensor_data().data(),
                               in[2]->tensor_data().data(), in[3]->tensor_data().data(), in[4]->tensor_data().data(),
                               in[5]->tensor_data().data(), in[6]->tensor_data().data(),
                               const_cast<char*>(dQ->tensor_data().data()), const_cast<char*>(dK->tensor_data().data()),
                               const_cast<char*>(dV->tensor_data().data()), const_cast<char*>(ws.tensor_data().data()),
                               ws_bytes, stream);
    OP_REQUIRES_OK(ctx, ToStatus(rc, "Backward"));
  }

 private:
  RuleAttrs attrs_;
};

// ---- flops estimation (forward.cc:390-474), result in host memory ---------------------------------
template <typename T, int N, int Rule>
class FlopsOp : public OpKernel {
 public:
  explicit FlopsOp(OpKernelConstruction* cons) : OpKernel(cons) {
    ReadRuleAttrs<Rule>(cons, &attrs_);
    OP_REQUIRES_OK(cons, cons->GetAttr("q_shape", &q_));
    OP_REQUIRES_OK(cons, cons->GetAttr("k_shape", &k_));
    OP_REQUIRES_OK(cons, cons->GetAttr("v_shape", &v_));
  }
  void Compute(OpKernelContext* ctx) override {
    fa_problem_t p = {};
    FillRule(attrs_, DType<T>::code, &p);
    auto qd = Dims(q_), kd = Dims(k_), vd = Dims(v_);
    OP_REQUIRES_OK(ctx, ToStatus(fa_check_forward_shapes(N, q_.dims(), qd.data(), k_.dims(), kd.data(), v_.dims(),
                                                         vd.data(), &p), "EstimateForwardFlops"));
    Tensor* flops = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({}), &flops));
    int dev = -1, optin = 0;
    OP_REQUIRES(ctx, cudaGetDevice(&dev) == cudaSuccess,
                errors::Internal("Failed to get the current cuda device associated with this thread."));
    if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) optin = 0;
    OP_REQUIRES_OK(ctx, ToStatus(fa_estimate_forward_flops(&p, optin, &flops->flat<float>()(0)), "EstimateForwardFlops"));
  }

 private:
  RuleAttrs attrs_;
  TensorShape q_, k_, v_;
};

}  // namespace

// ---- op registration: names, attrs and dtypes exactly as in the reference ---------------------------


#define FA_REGISTER_FAMILY(NAME, RULE, N, EXTRA)                                                              \
  REGISTER_OP(#NAME "AttentionForward" #N "dFloat16")                                                          \
      .Input("q: T").Input("k: T").Input("v: T").Output("o: T").Output("l: float").Output("m: T")             \
      .Attr("T: {float16}").Attr("sync_mode: string") EXTRA.SetShapeFn(InferForward<N>);                      \
  REGISTER_OP(#NAME "AttentionForward" #N "d")                                                                 \
      .Input("q: T").Input("k: T").Input("v: T").Output("o: T").Output("l: T").Output("m: T")                 \
      .Attr("T: {float, double}").Attr("sync_mode: string") EXTRA.SetShapeFn(InferForward<N>);                \
  REGISTER_OP(#NAME "AttentionBackward" #N "dFloat16")                                                         \
      .Input("q: T").Input("k: T").Input("v: T").Input("o: T").Input("l: float").Input("m: T").Input("d_o: T") \
      .Output("d_q: T").Output("d_k: T").Output("d_v: T")                                                     \
      .Attr("T: {float16}").Attr("sync_mode: string") EXTRA.SetShapeFn(InferBackward);                        \
  REGISTER_OP(#NAME "AttentionBackward" #N "d")                                                                \
      .Input("q: T").Input("k: T").Input("v: T").Input("o: T").Input("l: T").Input("m: T").Input("d_o: T")    \
      .Output("d_q: T").Output("d_k: T").Output("d_v: T")                                                     \
      .Attr("T: {float, double}").Attr("sync_mode: string") EXTRA.SetShapeFn(InferBackward);                  \
  REGISTER_OP("Estimate" #NAME "AttentionForward" #N "dFlops")                                                 \
      .Output("flops: float").Attr("q_shape: shape").Attr("k_shape: shape").Attr("v_shape: shape")            \
      .Attr("dtype: {float16, float, double}").Attr("sync_mode: string") EXTRA.SetShapeFn(InferFlops);        \
  REGISTER_KERNEL_BUILDER(Name(#NAME "AttentionForward" #N "dFloat16").Device(DEVICE_GPU).TypeConstraint<Eigen::half>("T"), \
                          ForwardOp<Eigen::half, N, RULE>);                                                    \
  REGISTER_KERNEL_BUILDER(Name(#NAME "AttentionForward" #N "d").Device(DEVICE_GPU).TypeConstraint<float>("T"), \
                          ForwardOp<float, N, RULE>);                                                          \
  REGISTER_KERNEL_BUILDER(Name(#NAME "AttentionForward" #N "d").Device(DEVICE_GPU).TypeConstraint<double>("T"), \
                          ForwardOp<double, N, RULE>);                                                         \
  REGISTER_KERNEL_BUILDER(Name(#NAME "AttentionBackward" #N "dFloat16").Device(DEVICE_GPU).TypeConstraint<Eigen::half>("T"), \
                          BackwardOp<Eigen::half, N, RULE>);                                                   \
  REGISTER_KERNEL_BUILDER(Name(#NAME "AttentionBackward" #N "d").Device(DEVICE_GPU).TypeConstraint<float>("T"), \
                          BackwardOp<float, N, RULE>);                                                         \
  REGISTER_KERNEL_BUILDER(Name(#NAME "AttentionBackward" #N "d").Device(DEVICE_GPU).TypeConstraint<double>("T"), \
                          BackwardOp<double, N, RULE>);                                                        \
  REGISTER_KERNEL_BUILDER(Name("Estimate" #NAME "AttentionForward" #N "dFlops").Device(DEVICE_GPU)             \
                              .HostMemory("flops").TypeConstraint<Eigen::half>("dtype"), FlopsOp<Eigen::half, N, RULE>); \
  REGISTER_KERNEL_BUILDER(Name("Estimate" #NAME "AttentionForward" #N "dFlops").Device(DEVICE_GPU)             \
                              .HostMemory("flops").TypeConstraint<float>("dtype"), FlopsOp<float, N, RULE>);   \
  REGISTER_KERNEL_BUILDER(Name("Estimate" #NAME "AttentionForward" #N "dFlops").Device(DEVICE_GPU)             \
                              .HostMemory("flops").TypeConstraint<double>("dtype"), FlopsOp<double, N, RULE>);

#define FA_NO_EXTRA
FA_REGISTER_FAMILY(Full, FA_RULE_FULL, 1, FA_NO_EXTRA)
FA_REGISTER_FAMILY(Full, FA_RULE_FULL, 2, FA_NO_EXTRA)
FA_REGISTER_FAMILY(Causal, FA_RULE_CAUSAL, 1, FA_NO_EXTRA)
FA_REGISTER_FAMILY(Causal, FA_RULE_CAUSAL, 2, FA_NO_EXTRA)
#define FA_LOCAL_EXTRA .Attr("window_size: int >= 1").Attr("log2_stride_size: int >= 0").Attr("is_causal: bool")
FA_REGISTER_FAMILY(Local, FA_RULE_LOCAL, 1, FA_LOCAL_EXTRA)
FA_REGISTER_FAMILY(Local, FA_RULE_LOCAL, 2, FA_LOCAL_EXTRA)

#endif  // GOOGLE_CUDA
