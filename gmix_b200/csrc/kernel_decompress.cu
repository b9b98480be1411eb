#include "kernels.h"
namespace gmx {
cudaError_t LaunchDecompress(const StreamParams& P, unsigned grid, cudaStream_t st) {
  // all of the SM's unified L1/shared memory as shared memory, so that kStreamMinBlocks CTAs are co-resident
  static const cudaError_t carve = cudaFuncSetAttribute(StreamKernel<kStreamThreads, MODE_DECOMPRESS, kStreamMinBlocks, false>,
                                                        cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (carve != cudaSuccess) return carve;
  StreamKernel<kStreamThreads, MODE_DECOMPRESS, kStreamMinBlocks, false><<<grid, kStreamThreads, 0, st>>>(P);
  return cudaGetLastError();
}
cudaError_t OccupancyDecompress(int* n) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, StreamKernel<kStreamThreads, MODE_DECOMPRESS, kStreamMinBlocks, false>, kStreamThreads, 0);
}
}  // namespace gmx
