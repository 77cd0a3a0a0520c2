// K2 (tcgen05 variant) — placeholder until the tensor-core path lands:
// reports "shape not covered" so the dispatcher uses the CUDA-core kernel.
#include "common.cuh"
namespace casmvs {
int conv3d_tc(const float*, const float*, const float*, const float*, float, const float*, float*,
              int, int, int, int, int, int, int, int, int, cudaStream_t) {
  return 1;
}
}  // namespace casmvs
