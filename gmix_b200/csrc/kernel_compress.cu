#include "kernels.h"
namespace gmx {
KernelConfigInfo KernelConfig(int cfg) {
  switch (cfg) {
#define X(id, wb, wl, minb, serial, ws) case id: return KernelConfigInfo{wb, wl, minb, 32 * (wb + wl + 1), serial, ws};
    GMX_KERNEL_CONFIGS(X)
#undef X
    default: return KernelConfigInfo{0, 0, 0, 0, 0, 0};
  }
}
cudaError_t LaunchCompress(int cfg, const StreamParams& P, unsigned grid, cudaStream_t st) {
  switch (cfg) {
#define X(id, wb, wl, minb, serial, ws) case id: return LaunchStreamKernel<wb, wl, MODE_COMPRESS, minb, false, serial != 0, ws != 0>(P, grid, st);
    GMX_KERNEL_CONFIGS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}
cudaError_t OccupancyCompress(int cfg, int* n) {
  switch (cfg) {
#define X(id, wb, wl, minb, serial, ws) case id: return OccupancyStreamKernel<wb, wl, MODE_COMPRESS, minb, false, serial != 0, ws != 0>(n);
    GMX_KERNEL_CONFIGS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace gmx
