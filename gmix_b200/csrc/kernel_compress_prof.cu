// Compress kernels with the per-role cycle counters compiled in (gmx_set_profile). Kept out of the production kernels:
// the per-bit path is instruction-cache bound and the lap timers are ~1000 instructions.
#include "kernels.h"
namespace gmx {
cudaError_t LaunchCompressProf(int cfg, const StreamParams& P, unsigned grid, cudaStream_t st) {
  switch (cfg) {
#define X(id, wb, wl, minb, serial, ws) case id: return LaunchStreamKernel<wb, wl, MODE_COMPRESS, minb, true, serial != 0, ws != 0>(P, grid, st);
    GMX_KERNEL_CONFIGS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace gmx
