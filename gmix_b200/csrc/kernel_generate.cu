// Batched generation (runner_utils::RunGeneration runner-utils.cpp:158-221): one CTA per prompt, every stream a
// clone of the loaded checkpoint.
#include "kernels.h"
namespace gmx {
cudaError_t LaunchGenerate(const StreamParams& P, unsigned grid, cudaStream_t st) {
  static const cudaError_t carve = cudaFuncSetAttribute(StreamKernel<kStreamThreads, MODE_GENERATE, kStreamMinBlocks, false>,
                                                        cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (carve != cudaSuccess) return carve;
  StreamKernel<kStreamThreads, MODE_GENERATE, kStreamMinBlocks, false><<<grid, kStreamThreads, 0, st>>>(P);
  return cudaGetLastError();
}
}  // namespace gmx
