// placeholder until the tcgen05 backward lands
#include "fa_launch.h"
namespace fa {
bool sm100_f16_backward_supports(const LaunchArgs&) { return false; }
cudaError_t sm100_f16_backward(const LaunchArgs&, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace fa
