// Batched LSTM gate product of lock-step generation (gate_gemm.cuh): exact SIMT variant and the tcgen05 / TMEM / TMA variant.
#include "gate_gemm.cuh"
#include "kernels.h"
namespace gmx {
cudaError_t LaunchGateExact(const float* W, const float* X, const uint32_t* sym, float* G, uint32_t n_slots, unsigned max_grid, cudaStream_t st) {
  unsigned grid = (n_slots + GX_SLOTS - 1) / GX_SLOTS;
  if (grid > max_grid) grid = max_grid;
  if (grid == 0) grid = 1;
  cudaError_t e = cudaFuncSetAttribute(GateDotsExactKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GX_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  GateDotsExactKernel<<<grid, GX_THREADS, GX_SMEM_BYTES, st>>>(W, X, sym, G, n_slots);
  return cudaGetLastError();
}
cudaError_t LaunchGateWeightPrep(const float* W, float* Wt, cudaStream_t st) {
  GateWeightPrepKernel<<<(GG_N * GG_K + 255) / 256, 256, 0, st>>>(W, Wt);
  return cudaGetLastError();
}
cudaError_t LaunchGateTc(const float* Xt, const float* Wt, const float* Wfull, const uint32_t* sym, float* G, uint32_t n_slots, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(GateGemmTcKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GT_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  GateGemmTcKernel<<<(n_slots + GG_M - 1) / GG_M, 128, GT_SMEM_BYTES, st>>>(Xt, Wt, Wfull, sym, G, n_slots);
  return cudaGetLastError();
}
unsigned GateWtFloats() { return (unsigned)(GG_NCHUNK * 2 * GG_B_FLOATS); }
}  // namespace gmx
