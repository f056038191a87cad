// placeholder until the tcgen05 forward lands (next commit)
#include "fa_launch.h"
namespace fa {
bool sm100_f16_forward_supports(const LaunchArgs&) { return false; }
size_t sm100_f16_workspace_bytes(const LaunchArgs&, bool) { return 0; }
cudaError_t sm100_f16_forward(const LaunchArgs&, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace fa
