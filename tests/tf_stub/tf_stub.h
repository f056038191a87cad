// tf_stub.h — a minimal stand-in for the slice of the TensorFlow C++ API that
// tf_flash_attention_b200/csrc/tf_ops/fa_tf_ops.cc uses. TEST INFRASTRUCTURE ONLY: TensorFlow is not installed in
// the build image, so this lets the OpKernel shim be compiled and EXECUTED (tests/tf_stub/shim_harness.cc) against
// libfa_b200.so: op / kernel registration, attrs, shape functions, OP_REQUIRES error paths, output / temp allocation and
// the stream hand-off. It mirrors names and signatures of the real API (tensorflow/core/framework/{op,op_kernel,
// shape_inference,tensor,tensor_shape}.h); it is not a reimplementation of TensorFlow and ships nowhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace Eigen {
struct half { uint16_t x; };
struct GpuDevice {
  cudaStream_t s = nullptr;
  cudaStream_t stream() const { return s; }
};
}  // namespace Eigen

namespace tensorflow {

enum DataType { DT_INVALID = 0, DT_FLOAT, DT_DOUBLE, DT_HALF, DT_UINT8 };
inline size_t DataTypeSize(DataType t) { return t == DT_FLOAT ? 4 : t == DT_DOUBLE ? 8 : t == DT_HALF ? 2 : 1; }
template <typename T> struct DataTypeToEnum;
template <> struct DataTypeToEnum<float> { static constexpr DataType value = DT_FLOAT; };
template <> struct DataTypeToEnum<double> { static constexpr DataType value = DT_DOUBLE; };
template <> struct DataTypeToEnum<Eigen::half> { static constexpr DataType value = DT_HALF; };

enum class Code { OK = 0, INVALID_ARGUMENT = 3, INTERNAL = 13 };
class Status {
 public:
  Status() = default;
  Status(Code c, std::string m) : code_(c), msg_(std::move(m)) {}
  bool ok() const { return code_ == Code::OK; }
  Code code() const { return code_; }
  const std::string& message() const { return msg_; }

 private:
  Code code_ = Code::OK;
  std::string msg_;
};
inline Status OkStatus() { return Status(); }
namespace errors {
template <typename... A> std::string StrCat(const A&... a) {
  std::ostringstream os;
  (void)std::initializer_list<int>{(os << a, 0)...};
  return os.str();
}
template <typename... A> Status InvalidArgument(const A&... a) { return Status(Code::INVALID_ARGUMENT, StrCat(a...)); }
template <typename... A> Status Internal(const A&... a) { return Status(Code::INTERNAL, StrCat(a...)); }
}  // namespace errors
#define TF_RETURN_IF_ERROR(...)             \
  do {                                      \
    ::tensorflow::Status _s = (__VA_ARGS__); \
    if (!_s.ok()) return _s;                \
  } while (0)

class TensorShape {
 public:
  TensorShape() = default;
  TensorShape(std::initializer_list<int64_t> d) : d_(d) {}
  explicit TensorShape(std::vector<int64_t> d) : d_(std::move(d)) {}
  int dims() const { return int(d_.size()); }
  int64_t dim_size(int i) const { return d_[i]; }
  void AddDim(int64_t v) { d_.push_back(v); }
  void RemoveDimRange(int b, int e) { d_.erase(d_.begin() + b, d_.begin() + e); }
  int64_t num_elements() const {
    int64_t n = 1;
    for (auto v : d_) n *= v;
    return n;
  }
  const std::vector<int64_t>& vec() const { return d_; }
  bool operator==(const TensorShape& o) const { return d_ == o.d_; }

 private:
  std::vector<int64_t> d_;
};

struct StringPiece {
  const char* p;
  size_t n;
  const char* data() const { return p; }
  size_t size() const { return n; }
};

// Tensor: a typed view of a device (or, for HostMemory outputs, host) allocation
class Tensor {
 public:
  Tensor() = default;
  Tensor(DataType dt, TensorShape s, bool host = false) : dt_(dt), shape_(std::move(s)), host_(host) {
    bytes_ = size_t(shape_.num_elements()) * DataTypeSize(dt_);
    void* p = nullptr;
    if (bytes_) {
      if (host_) p = ::operator new(bytes_);
      else if (cudaMalloc(&p, bytes_) != cudaSuccess) p = nullptr;
    }
    const bool h = host_;
    buf_ = std::shared_ptr<void>(p, [h](void* q) { if (!q) return; if (h) ::operator delete(q); else cudaFree(q); });
  }
  DataType dtype() const { return dt_; }
  const TensorShape& shape() const { return shape_; }
  int dims() const { return shape_.dims(); }
  int64_t dim_size(int i) const { return shape_.dim_size(i); }
  StringPiece tensor_data() const { return StringPiece{static_cast<const char*>(buf_.get()), bytes_}; }
  size_t bytes() const { return bytes_; }
  template <typename T> struct Flat {
    T* p;
    T& operator()(int64_t i) const { return p[i]; }
  };
  template <typename T> Flat<T> flat() { return Flat<T>{static_cast<T*>(buf_.get())}; }

 private:
  DataType dt_ = DT_INVALID;
  TensorShape shape_;
  bool host_ = false;
  size_t bytes_ = 0;
  std::shared_ptr<void> buf_;
};

// ---- attrs -------------------------------------------------------------------------------------------
struct AttrValue {
  std::string s;
  int64_t i = 0;
  bool b = false;
  TensorShape shape;
  DataType type = DT_INVALID;
};
using AttrMap = std::map<std::string, AttrValue>;

class OpKernelConstruction {
 public:
  explicit OpKernelConstruction(AttrMap a) : attrs_(std::move(a)) {}
  Status GetAttr(const std::string& n, std::string* v) const { return Get(n, [&](const AttrValue& a) { *v = a.s; }); }
  Status GetAttr(const std::string& n, int* v) const { return Get(n, [&](const AttrValue& a) { *v = int(a.i); }); }
  Status GetAttr(const std::string& n, bool* v) const { return Get(n, [&](const AttrValue& a) { *v = a.b; }); }
  Status GetAttr(const std::string& n, TensorShape* v) const { return Get(n, [&](const AttrValue& a) { *v = a.shape; }); }
  void CtxFailure(const Status& s) { if (status_.ok()) status_ = s; }
  const Status& status() const { return status_; }

 private:
  template <typename F> Status Get(const std::string& n, F f) const {
    auto it = attrs_.find(n);
    if (it == attrs_.end()) return errors::InvalidArgument("No attr named '", n, "'");
    f(it->second);
    return OkStatus();
  }
  AttrMap attrs_;
  Status status_;
};

class OpKernelContext {
 public:
  std::vector<Tensor> inputs;
  std::vector<DataType> output_types;   // from the op definition ("o: T", "l: float", ...)
  std::vector<bool> output_on_host;     // KernelDefBuilder::HostMemory
  std::vector<std::unique_ptr<Tensor>> outputs;
  Eigen::GpuDevice device;

  const Tensor& input(int i) const { return inputs[i]; }
  Status allocate_output(int i, const TensorShape& s, Tensor** t) {
    if (i >= int(output_types.size())) return errors::Internal("output index out of range");
    if (int(outputs.size()) <= i) outputs.resize(i + 1);
    outputs[i].reset(new Tensor(output_types[i], s, i < int(output_on_host.size()) && output_on_host[i]));
    if (outputs[i]->bytes() && !outputs[i]->tensor_data().data()) return errors::Internal("allocation failed");
    *t = outputs[i].get();
    return OkStatus();
  }
  Status allocate_temp(DataType dt, const TensorShape& s, Tensor* t) {
    *t = Tensor(dt, s);
    if (t->bytes() && !t->tensor_data().data()) return errors::Internal("allocation failed");
    temps_.push_back(*t);   // keep alive until the context dies (the stream may still use it)
    return OkStatus();
  }
  template <typename D> const D& eigen_device() const { return device; }
  void CtxFailure(const Status& s) { if (status_.ok()) status_ = s; }
  const Status& status() const { return status_; }

 private:
  Status status_;
  std::vector<Tensor> temps_;
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() = default;
  virtual void Compute(OpKernelContext* ctx) = 0;
};

#define OP_REQUIRES(CTX, EXP, STATUS) \
  do {                                \
    if (!(EXP)) {                     \
      (CTX)->CtxFailure((STATUS));    \
      return;                         \
    }                                 \
  } while (0)
#define OP_REQUIRES_OK(CTX, ...)              \
  do {                                        \
    ::tensorflow::Status _s = (__VA_ARGS__);   \
    if (!_s.ok()) {                           \
      (CTX)->CtxFailure(_s);                  \
      return;                                 \
    }                                         \
  } while (0)

// ---- shape inference ------------------------------------------------------------------------------------
namespace shape_inference {
struct ShapeHandle {
  std::shared_ptr<std::vector<int64_t>> d;
};
class InferenceContext {
 public:
  std::vector<ShapeHandle> inputs, outputs;
  static ShapeHandle Make(std::vector<int64_t> v) { return ShapeHandle{std::make_shared<std::vector<int64_t>>(std::move(v))}; }
  ShapeHandle input(int i) const { return inputs[i]; }
  int Rank(const ShapeHandle& h) const { return int(h.d->size()); }
  Status Subshape(const ShapeHandle& h, int start, ShapeHandle* out) { return Subshape(h, start, Rank(h), out); }
  Status Subshape(const ShapeHandle& h, int start, int end, ShapeHandle* out) {
    if (start < 0 || end > Rank(h) || start > end) return errors::InvalidArgument("Subshape out of range");
    *out = Make(std::vector<int64_t>(h.d->begin() + start, h.d->begin() + end));
    return OkStatus();
  }
  Status Concatenate(const ShapeHandle& a, const ShapeHandle& b, ShapeHandle* out) {
    std::vector<int64_t> v(*a.d);
    v.insert(v.end(), b.d->begin(), b.d->end());
    *out = Make(std::move(v));
    return OkStatus();
  }
  ShapeHandle Scalar() { return Make({}); }
  void set_output(int i, const ShapeHandle& h) {
    if (int(outputs.size()) <= i) outputs.resize(i + 1);
    outputs[i] = h;
  }
};
}  // namespace shape_inference

// ---- registries ------------------------------------------------------------------------------------------
using ShapeFn = std::function<Status(shape_inference::InferenceContext*)>;
struct OpDef {
  std::string name;
  std::vector<std::string> inputs, outputs, attrs;
  ShapeFn shape_fn;
};
class OpDefBuilder {
 public:
  explicit OpDefBuilder(std::string name) { def_.name = std::move(name); }
  OpDefBuilder& Input(std::string s) { def_.inputs.push_back(std::move(s)); return *this; }
  OpDefBuilder& Output(std::string s) { def_.outputs.push_back(std::move(s)); return *this; }
  OpDefBuilder& Attr(std::string s) { def_.attrs.push_back(std::move(s)); return *this; }
  OpDefBuilder& SetShapeFn(ShapeFn f) { def_.shape_fn = std::move(f); return *this; }
  const OpDef& def() const { return def_; }

 private:
  OpDef def_;
};
inline std::map<std::string, OpDef>& OpRegistry() {
  static std::map<std::string, OpDef> r;
  return r;
}
struct OpRegistrar {
  OpRegistrar(const OpDefBuilder& b) { OpRegistry()[b.def().name] = b.def(); }  // NOLINT: implicit on purpose
};

constexpr const char* DEVICE_GPU = "GPU";
struct KernelDef {
  std::string op, device;
  std::map<std::string, DataType> constraints;
  std::vector<std::string> host_memory;
  std::function<OpKernel*(OpKernelConstruction*)> factory;
};
class Name {
 public:
  explicit Name(std::string op) { def_.op = std::move(op); }
  Name& Device(const char* d) { def_.device = d; return *this; }
  template <typename T> Name& TypeConstraint(const char* attr) { def_.constraints[attr] = DataTypeToEnum<T>::value; return *this; }
  Name& HostMemory(const char* arg) { def_.host_memory.push_back(arg); return *this; }
  const KernelDef& def() const { return def_; }

 private:
  KernelDef def_;
};
inline std::vector<KernelDef>& KernelRegistry() {
  static std::vector<KernelDef> r;
  return r;
}
struct KernelRegistrar {
  KernelRegistrar(const Name& n, std::function<OpKernel*(OpKernelConstruction*)> f) {
    KernelDef d = n.def();
    d.factory = std::move(f);
    KernelRegistry().push_back(std::move(d));
  }
};
#define TF_STUB_CONCAT_(a, b) a##b
#define TF_STUB_CONCAT(a, b) TF_STUB_CONCAT_(a, b)
#define REGISTER_OP(name) \
  static ::tensorflow::OpRegistrar TF_STUB_CONCAT(tf_stub_op_, __COUNTER__) = ::tensorflow::OpDefBuilder(name)
#define REGISTER_KERNEL_BUILDER(kb, ...)                                          \
  static ::tensorflow::KernelRegistrar TF_STUB_CONCAT(tf_stub_kernel_, __COUNTER__)( \
      (kb), [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { return new __VA_ARGS__(c); })

}  // namespace tensorflow
