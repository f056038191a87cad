// Test-infrastructure stub (NOT reference code): the smallest stand-in for
// tensorflow::TensorShape that lets the reference's sync_methods.cc compile
// without TensorFlow. sync_methods.cc only calls dims() and dim_size()
// (/root/reference/flash_attention/kernel/sync_methods.cc:13-15).
#pragma once
#include <cstdint>
#include <vector>
#include <initializer_list>
namespace tensorflow {
class TensorShape {
 public:
  TensorShape() = default;
  TensorShape(std::initializer_list<int64_t> d) : d_(d) {}
  explicit TensorShape(const std::vector<int64_t>& d) : d_(d) {}
  int dims() const { return static_cast<int>(d_.size()); }
  int64_t dim_size(int i) const { return d_[i]; }
 private:
  std::vector<int64_t> d_;
};
}  // namespace tensorflow
