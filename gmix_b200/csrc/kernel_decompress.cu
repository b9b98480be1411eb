#include "kernels.h"
namespace gmx {
cudaError_t LaunchDecompress(const StreamParams& P, unsigned grid, cudaStream_t st) {
  StreamKernel<kStreamThreads, MODE_DECOMPRESS, kStreamMinBlocks><<<grid, kStreamThreads, 0, st>>>(P);
  return cudaGetLastError();
}
cudaError_t OccupancyDecompress(int* n) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, StreamKernel<kStreamThreads, MODE_DECOMPRESS, kStreamMinBlocks>, kStreamThreads, 0);
}
}  // namespace gmx
