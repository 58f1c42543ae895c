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
// generate_fromS.py:174-175: (img.permute(0, 2, 3, 1) * 127.5 + 128).clamp(0, 255).to(torch.uint8), written into a canvas
// [N, H, canvas_w, 3] at column offset x_off (the reference concatenates original | edited along the width, :206).
__global__ void __launch_bounds__(256) img_to_uint8_kernel(const float* __restrict__ img, unsigned char* __restrict__ out, int N, int H, int W,
                                                           int canvas_w, int x_off) {
  const long long total = (long long)N * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long t = i / W;
    const int y = (int)(t % H);
    const long long n = t / H;
    unsigned char* o = out + ((n * H + y) * canvas_w + x_off + x) * 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float v = __ldg(img + ((n * 3 + j) * H + y) * (long long)W + x) * 127.5f + 128.f;
      o[j] = (unsigned char)fminf(fmaxf(v, 0.f), 255.f);      // truncation, like Tensor.to(torch.uint8)
    }
  }
}
}  // namespace smc

extern "C" int smc_img_to_uint8(const float* img, unsigned char* out, int n, int h, int w, int canvas_w, int x_off, void* stream) {
  if (!img || !out || n < 1 || h < 1 || w < 1 || x_off < 0 || x_off + w > canvas_w) return SMC_EINVAL;
  long long blocks = smc::ceil_div_ll((long long)n * h * w, 256);
  if (blocks > smc::kNumSMs * 16) blocks = smc::kNumSMs * 16;
  smc::img_to_uint8_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(img, out, n, h, w, canvas_w, x_off);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

extern "C" int smc_abi_version(void) { return SMC_ABI_VERSION; }

extern "C" int smc_sgd_step(float* delta, const float* grad, int64_t numel, float lr, float grad_scale, float l2_scale, void* stream) {
  if (!delta || !grad || numel < 1) return SMC_EINVAL;
  long long blocks = smc::ceil_div_ll(numel, 256);
  if (blocks > smc::kNumSMs * 8) blocks = smc::kNumSMs * 8;
  smc::sgd_step_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(delta, grad, numel, lr, grad_scale, l2_scale);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
