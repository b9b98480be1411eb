#include "kernels.h"
namespace gmx {
cudaError_t LaunchCompress(const StreamParams& P, unsigned grid, cudaStream_t st) {
  StreamKernel<kStreamThreads, MODE_COMPRESS, kStreamMinBlocks><<<grid, kStreamThreads, 0, st>>>(P);
  return cudaGetLastError();
}
cudaError_t OccupancyCompress(int* n) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, StreamKernel<kStreamThreads, MODE_COMPRESS, kStreamMinBlocks>, kStreamThreads, 0);
}
}  // namespace gmx
