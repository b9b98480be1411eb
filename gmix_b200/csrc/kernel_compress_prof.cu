// Compress kernel with the per-phase cycle counters compiled in (gmx_set_profile). Kept out of the
// production kernel: the per-bit path is instruction-cache bound and the lap timers are ~1000 instructions.
#include "kernels.h"
namespace gmx {
cudaError_t LaunchCompressProf(const StreamParams& P, unsigned grid, cudaStream_t st) {
  static const cudaError_t carve = cudaFuncSetAttribute(StreamKernel<kStreamThreads, MODE_COMPRESS, kStreamMinBlocks, true>,
                                                        cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (carve != cudaSuccess) return carve;
  StreamKernel<kStreamThreads, MODE_COMPRESS, kStreamMinBlocks, true><<<grid, kStreamThreads, 0, st>>>(P);
  return cudaGetLastError();
}
}  // namespace gmx
