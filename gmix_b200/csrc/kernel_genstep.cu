// Lock-step batched generation: one sampled byte of every stream per launch (GenStepKernel, stream_kernel.cuh) around the
// batched gate product (kernel_gate.cu). Streams are overlays of the shared model, as in kernel_generate.cu.
#define GMX_OVERLAY 1
#include "kernels.h"
namespace gmx {
cudaError_t LaunchGenStep(const GenStepParams& Q, unsigned grid, cudaStream_t st) {
  constexpr int WB = 2, WL = 1, MINB = 8;   // the thread count of the throughput configuration (kernels.h)
  cudaError_t e = cudaFuncSetAttribute(GenStepKernel<WB, WL, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  GenStepKernel<WB, WL, MINB><<<grid, 32 * (WB + WL + 1), 0, st>>>(Q);
  return cudaGetLastError();
}
}  // namespace gmx
