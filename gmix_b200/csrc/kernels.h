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
// SERIAL = 1 with WL = 0 (configuration 10): the hybrid order - PPMd warp one byte ahead under the bit path, LSTM phases with all threads.
// WS = 1: the dense LSTM gate weights (184.8 KB) stay resident in shared memory (needs the whole SM: MINB = 1).
//            id  WB WL MINB SERIAL WS
#define GMX_KERNEL_CONFIGS(X)   \
  X(0, 2, 1, 8, 1, 0)           \
  X(1, 2, 1, 8, 0, 0)           \
  X(2, 1, 2, 8, 0, 0)           \
  X(3, 2, 2, 6, 0, 0)           \
  X(4, 4, 2, 4, 0, 0)           \
  X(5, 4, 2, 1, 0, 0)           \
  X(6, 3, 0, 8, 0, 0)           \
  X(7, 7, 0, 4, 0, 0)           \
  X(8, 4, 8, 1, 0, 1)           \
  X(9, 4, 4, 1, 0, 1)           \
  X(10, 3, 0, 8, 1, 0)
constexpr int kNumKernelConfigs = 11;
constexpr int kThroughputConfig = 10, kLatencyConfig = 9;   // what the host picks for a full wave of streams / for at most one stream per SM
constexpr int kStepWB = 2, kStepWL = 1;   // role split of the single-stream stepping kernel
struct KernelConfigInfo { int wb, wl, minb, threads, serial, ws; };
KernelConfigInfo KernelConfig(int cfg);
cudaError_t LaunchCompress(int cfg, const StreamParams& P, unsigned grid, cudaStream_t st);
cudaError_t LaunchCompressProf(int cfg, const StreamParams& P, unsigned grid, cudaStream_t st);
cudaError_t LaunchDecompress(int cfg, const StreamParams& P, unsigned grid, cudaStream_t st);
cudaError_t LaunchGenerate(int cfg, const StreamParams& P, unsigned grid, cudaStream_t st);
cudaError_t LaunchStep(const StepParams& Q, cudaStream_t st);
// lock-step batched generation (gate_gemm.cuh, GenStepKernel)
cudaError_t LaunchGenStep(const GenStepParams& Q, unsigned grid, cudaStream_t st);
cudaError_t LaunchGateExact(const float* W, const float* X, const uint32_t* sym, float* G, uint32_t n_slots, unsigned max_grid, cudaStream_t st);
cudaError_t LaunchGateWeightPrep(const float* W, float* Wt, cudaStream_t st);
cudaError_t LaunchGateTc(const float* Xt, const float* Wt, const float* Wfull, const uint32_t* sym, float* G, uint32_t n_slots, cudaStream_t st);
unsigned GateWtFloats();
unsigned StepStateBytes();
cudaError_t OccupancyCompress(int cfg, int* blocks_per_sm);
cudaError_t OccupancyDecompress(int cfg, int* blocks_per_sm);

// Shared body of the per-mode launchers.
template <bool WS> constexpr size_t DynSmemBytes() { return WS ? (size_t)W_DENSE_BYTES + 16 : 0; }
template <int WB, int WL, int MODE, int MINB, bool PROF, bool SERIAL, bool WS>
inline cudaError_t PrepareStreamKernel() {
  // all of the SM's unified L1/shared memory as shared memory, so that MINB CTAs are co-resident
  cudaError_t e = cudaFuncSetAttribute(StreamKernel<WB, WL, MODE, MINB, PROF, SERIAL, WS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess && WS)
    e = cudaFuncSetAttribute(StreamKernel<WB, WL, MODE, MINB, PROF, SERIAL, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DynSmemBytes<WS>());
  return e;
}
template <int WB, int WL, int MODE, int MINB, bool PROF, bool SERIAL, bool WS>
inline cudaError_t LaunchStreamKernel(const StreamParams& P, unsigned grid, cudaStream_t st) {
  // function attributes are per device (the multi-GPU host launches the same kernel from one thread per GPU): set them
  // for the calling thread's device on every launch, it costs microseconds against kernels that run for seconds
  const cudaError_t prep = PrepareStreamKernel<WB, WL, MODE, MINB, PROF, SERIAL, WS>();
  if (prep != cudaSuccess) return prep;
  StreamKernel<WB, WL, MODE, MINB, PROF, SERIAL, WS><<<grid, 32 * (WB + WL + 1), DynSmemBytes<WS>(), st>>>(P);
  return cudaGetLastError();
}
template <int WB, int WL, int MODE, int MINB, bool PROF, bool SERIAL, bool WS>
inline cudaError_t OccupancyStreamKernel(int* n) {
  cudaError_t e = PrepareStreamKernel<WB, WL, MODE, MINB, PROF, SERIAL, WS>();
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(n, StreamKernel<WB, WL, MODE, MINB, PROF, SERIAL, WS>, 32 * (WB + WL + 1), DynSmemBytes<WS>());
}
}  // namespace gmx
#endif
