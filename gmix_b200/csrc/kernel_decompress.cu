#include "kernels.h"
namespace gmx {
cudaError_t LaunchDecompress(int cfg, const StreamParams& P, unsigned grid, cudaStream_t st) {
  switch (cfg) {
#define X(id, wb, wl, minb, serial, ws) case id: return LaunchStreamKernel<wb, wl, MODE_DECOMPRESS, minb, false, false, ws != 0>(P, grid, st);
    GMX_KERNEL_CONFIGS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}
cudaError_t OccupancyDecompress(int cfg, int* n) {
  switch (cfg) {
#define X(id, wb, wl, minb, serial, ws) case id: return OccupancyStreamKernel<wb, wl, MODE_DECOMPRESS, minb, false, false, ws != 0>(n);
    GMX_KERNEL_CONFIGS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace gmx
