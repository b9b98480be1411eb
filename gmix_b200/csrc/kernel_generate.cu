// Batched generation (runner_utils::RunGeneration runner-utils.cpp:158-221): one CTA per prompt, every stream a
// clone of the loaded checkpoint.
#define GMX_OVERLAY 1   // streams of a generation batch are overlays of the shared model (stream_kernel.cuh)
#include "kernels.h"
namespace gmx {
cudaError_t LaunchGenerate(int cfg, const StreamParams& P, unsigned grid, cudaStream_t st) {
  switch (cfg) {
#define X(id, wb, wl, minb, serial, ws) case id: return LaunchStreamKernel<wb, wl, MODE_GENERATE, minb, false, false, ws != 0>(P, grid, st);
    GMX_KERNEL_CONFIGS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace gmx
