// Single-stream stepping kernel behind the gmx_pred_* entry points (Predictor facade).
#include "kernels.h"
namespace gmx {
cudaError_t LaunchStep(const StepParams& Q, cudaStream_t st) {
  StepKernel<kStepWB, kStepWL><<<1, 32 * (kStepWB + kStepWL + 1), 0, st>>>(Q);
  return cudaGetLastError();
}
unsigned StepStateBytes() { return (unsigned)sizeof(StreamSmem); }
}  // namespace gmx
