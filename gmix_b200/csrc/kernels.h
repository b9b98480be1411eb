// Launch wrappers of the stream kernels; each mode lives in its own translation unit so the two
// large kernels compile in parallel.
#ifndef GMIX_B200_KERNELS_H_
#define GMIX_B200_KERNELS_H_
#include <cuda_runtime.h>

#include "stream_kernel.cuh"

namespace gmx {
constexpr int kStreamThreads = 128;   // threads per stream CTA
constexpr int kStreamMinBlocks = 8;   // resident CTAs per SM the register budget is capped for
cudaError_t LaunchCompress(const StreamParams& P, unsigned grid, cudaStream_t st);
cudaError_t LaunchCompressProf(const StreamParams& P, unsigned grid, cudaStream_t st);
cudaError_t LaunchDecompress(const StreamParams& P, unsigned grid, cudaStream_t st);
cudaError_t LaunchGenerate(const StreamParams& P, unsigned grid, cudaStream_t st);
cudaError_t LaunchStep(const StepParams& Q, cudaStream_t st);
unsigned StepStateBytes();
cudaError_t OccupancyCompress(int* blocks_per_sm);
cudaError_t OccupancyDecompress(int* blocks_per_sm);
}  // namespace gmx
#endif
