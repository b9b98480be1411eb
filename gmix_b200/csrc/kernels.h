// Launch wrappers of the stream kernels; each mode lives in its own translation unit so the large kernels compile
// in parallel. A kernel CONFIGURATION is the role split of the stream CTA (stream_kernel.cuh): warps of bit role, warps
// of LSTM role (+ one PPMd warp) and the resident CTAs per SM the register budget is capped for. Every configuration
// computes the same bytes; which one is fastest depends on the workload shape (many short streams vs few long ones),
// so the library carries a small set and the host picks (gmx_set_kernel_config).
#ifndef GMIX_B200_KERNELS_H_
#define GMIX_B200_KERNELS_H_
#include <cuda_runtime.h>

#include "stream_kernel.cuh"

namespace gmx {
// SERIAL = 1: compress walks the phases with all threads (no role pipeline); decompress and generation always do.
//            id  WB WL MINB SERIAL
#define GMX_KERNEL_CONFIGS(X) \
  X(0, 2, 1, 8, 1)            \
  X(1, 2, 1, 8, 0)            \
  X(2, 1, 2, 8, 0)            \
  X(3, 2, 2, 6, 0)            \
  X(4, 4, 2, 4, 0)            \
  X(5, 4, 2, 1, 0)            \
  X(6, 3, 0, 8, 0)            \
  X(7, 7, 0, 4, 0)
constexpr int kNumKernelConfigs = 8;
constexpr int kStepWB = 2, kStepWL = 1;   // role split of the single-stream stepping kernel
struct KernelConfigInfo { int wb, wl, minb, threads, serial; };
KernelConfigInfo KernelConfig(int cfg);
cudaError_t LaunchCompress(int cfg, const StreamParams& P, unsigned grid, cudaStream_t st);
cudaError_t LaunchCompressProf(int cfg, const StreamParams& P, unsigned grid, cudaStream_t st);
cudaError_t LaunchDecompress(int cfg, const StreamParams& P, unsigned grid, cudaStream_t st);
cudaError_t LaunchGenerate(int cfg, const StreamParams& P, unsigned grid, cudaStream_t st);
cudaError_t LaunchStep(const StepParams& Q, cudaStream_t st);
unsigned StepStateBytes();
cudaError_t OccupancyCompress(int cfg, int* blocks_per_sm);
cudaError_t OccupancyDecompress(int cfg, int* blocks_per_sm);

// Shared body of the per-mode launchers.
template <int WB, int WL, int MODE, int MINB, bool PROF, bool SERIAL>
inline cudaError_t LaunchStreamKernel(const StreamParams& P, unsigned grid, cudaStream_t st) {
  // all of the SM's unified L1/shared memory as shared memory, so that MINB CTAs are co-resident
  static const cudaError_t carve = cudaFuncSetAttribute(StreamKernel<WB, WL, MODE, MINB, PROF, SERIAL>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                                        cudaSharedmemCarveoutMaxShared);
  if (carve != cudaSuccess) return carve;
  StreamKernel<WB, WL, MODE, MINB, PROF, SERIAL><<<grid, 32 * (WB + WL + 1), 0, st>>>(P);
  return cudaGetLastError();
}
template <int WB, int WL, int MODE, int MINB, bool PROF, bool SERIAL>
inline cudaError_t OccupancyStreamKernel(int* n) {
  cudaError_t e = cudaFuncSetAttribute(StreamKernel<WB, WL, MODE, MINB, PROF, SERIAL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, StreamKernel<WB, WL, MODE, MINB, PROF, SERIAL>, 32 * (WB + WL + 1), 0);
}
}  // namespace gmx
#endif
