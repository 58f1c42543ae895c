// ABI version + the optimiser step of find_direction.py:285,339 (SGD, no momentum) fused with the
// analytic L2 term of find_direction.py:190-191 and the unscaling of the (allreduced) gradient.
#include "common.cuh"

namespace smc {
__global__ void sgd_step_kernel(float* __restrict__ delta, const float* __restrict__ grad, long long n, float lr, float grad_scale,
                                float l2_scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = delta[i];
    delta[i] = d - lr * (grad[i] * grad_scale + l2_scale * d);
  }
}
}  // namespace smc

extern "C" int smc_abi_version(void) { return SMC_ABI_VERSION; }

extern "C" int smc_sgd_step(float* delta, const float* grad, int64_t numel, float lr, float grad_scale, float l2_scale, void* stream) {
  if (!delta || !grad || numel < 1) return SMC_EINVAL;
  long long blocks = smc::ceil_div_ll(numel, 256);
  if (blocks > smc::kNumSMs * 8) blocks = smc::kNumSMs * 8;
  smc::sgd_step_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(delta, grad, numel, lr, grad_scale, l2_scale);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
