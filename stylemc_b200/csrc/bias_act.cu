// bias_act for sm_100a: y = clamp(act(x + b[c]) * gain), its first and second derivative passes.
//
// Drop-in for the reference plugin entry `_plugin.bias_act(x, b, xref, yref, dy, grad, dim, act, alpha,
// gain, clamp)` (torch_utils/ops/bias_act.cpp:32-90, kernel semantics bias_act.cu:23-147): same nine
// activations, same grad = 0/1/2 meaning, same clamp rule (forward saturates, backward zeroes where
// |yref| >= clamp).  Pure HBM streaming: algorithmic bytes = (1 + #aux inputs + 1) * numel * sizeof(T).
// Design: 16-byte vector loads/stores that bypass L1, one bias lookup per vector when the channel
// stride allows it (NCHW, H*W % vec == 0), grid = a multiple of the 148 SMs with a grid-stride loop.
#include "common.cuh"
#include "stylemc_b200.h"

namespace smc {

template <class T> struct Compute { typedef float type; };
template <> struct Compute<double> { typedef double type; };

template <class S> __device__ __forceinline__ S exp_(S v);
template <> __device__ __forceinline__ float exp_<float>(float v) { return expf(v); }
template <> __device__ __forceinline__ double exp_<double>(double v) { return exp(v); }
template <class S> __device__ __forceinline__ S log_(S v);
template <> __device__ __forceinline__ float log_<float>(float v) { return logf(v); }
template <> __device__ __forceinline__ double log_<double>(double v) { return log(v); }

struct BiasActArgs {
  const void* x; const void* b; const void* xref; const void* yref; const void* dy; void* y;
  long long size_x; int size_b; long long step_b;
  int grad; float alpha, gain, clamp;
};

// One element.  A = activation index (bias_act.py:23-33 cuda_idx), G = derivative order.
template <class S, int A, int G>
__device__ __forceinline__ S bias_act_elem(S x, S b, S xref, S yref, S dy, S alpha, S gain, S clamp) {
  const S one = (S)1, two = (S)2, range = (S)80, half_range = (S)40;
  const S selu_scale = (S)1.0507009873554804934193349852946, selu_alpha = (S)1.6732632423543772848170429916717;
  const S yy = (gain != (S)0) ? yref / gain : (S)0;
  S y = 0;
  if (G == 0) x += b; else xref += b;
  if (A == 1) { y = x; }
  if (A == 2) { y = (G == 0) ? (x > 0 ? x : (S)0) : (yy > 0 ? x : (S)0); }
  if (A == 3) { y = (G == 0) ? (x > 0 ? x : x * alpha) : (yy > 0 ? x : x * alpha); }
  if (A == 4) {
    if (G == 0) { S c = exp_(x), d = one / c; y = (x < -range) ? -one : (x > range) ? one : (c - d) / (c + d); }
    else if (G == 1) y = x * (one - yy * yy);
    else y = x * (one - yy * yy) * (-two * yy);
  }
  if (A == 5) {
    if (G == 0) y = (x < -range) ? (S)0 : one / (exp_(-x) + one);
    else if (G == 1) y = x * yy * (one - yy);
    else y = x * yy * (one - yy) * (one - two * yy);
  }
  if (A == 6) {
    if (G == 0) y = (x >= 0) ? x : exp_(x) - one;
    else if (G == 1) y = (yy >= 0) ? x : x * (yy + one);
    else y = (yy >= 0) ? (S)0 : x * (yy + one);
  }
  if (A == 7) {
    if (G == 0) y = (x >= 0) ? selu_scale * x : (selu_scale * selu_alpha) * (exp_(x) - one);
    else if (G == 1) y = (yy >= 0) ? x * selu_scale : x * (yy + selu_scale * selu_alpha);
    else y = (yy >= 0) ? (S)0 : x * (yy + selu_scale * selu_alpha);
  }
  if (A == 8) {
    if (G == 0) y = (x > range) ? x : log_(exp_(x) + one);
    else if (G == 1) y = x * (one - exp_(-yy));
    else { S c = exp_(-yy); y = x * c * (one - c); }
  }
  if (A == 9) {
    if (G == 0) y = (x < -range) ? (S)0 : x / (exp_(-x) + one);
    else {
      S c = exp_(xref), d = c + one;
      if (G == 1) y = (xref > half_range) ? x : x * c * (xref + d) / (d * d);
      else y = (xref > half_range) ? (S)0 : x * c * (xref * (two - d) + two * d) / (d * d * d);
      yref = (xref < -range) ? (S)0 : xref / (exp_(-xref) + one) * gain;
    }
  }
  y *= gain * dy;
  if (clamp >= 0) {
    if (G == 0) y = (y > -clamp && y < clamp) ? y : (y >= 0 ? clamp : -clamp);
    else y = (yref > -clamp && yref < clamp) ? y : (S)0;
  }
  return y;
}

template <class T> __device__ __forceinline__ typename Compute<T>::type to_s(T v) { return (typename Compute<T>::type)v; }
template <> __device__ __forceinline__ float to_s<__half>(__half v) { return __half2float(v); }
template <class T, class S> __device__ __forceinline__ T from_s(S v) { return (T)v; }
template <> __device__ __forceinline__ __half from_s<__half, float>(float v) { return __float2half_rn(v); }

// VEC elements per thread per iteration; VEC * sizeof(T) == 16 on the fast path, VEC == 1 otherwise.
template <class T, int A, int VEC, int G>
__global__ void __launch_bounds__(256) bias_act_kernel(const BiasActArgs p) {
  typedef typename Compute<T>::type S;
  const S alpha = (S)p.alpha, gain = (S)p.gain, clamp = (S)p.clamp;
  const long long nvec = p.size_x / VEC;
  const T* __restrict__ xb = (const T*)p.x;
  const T* __restrict__ bb = (const T*)p.b;
  const T* __restrict__ xr = (const T*)p.xref;
  const T* __restrict__ yr = (const T*)p.yref;
  const T* __restrict__ dyb = (const T*)p.dy;
  T* __restrict__ yb = (T*)p.y;
  const bool bias_per_vec = (VEC > 1) && (p.step_b % VEC == 0);
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const long long i0 = v * VEC;
    __align__(16) T xv[VEC]; __align__(16) T xrv[VEC]; __align__(16) T yrv[VEC]; __align__(16) T dyv[VEC]; __align__(16) T out[VEC];
    if (VEC > 1) {
      *reinterpret_cast<uint4*>(xv) = ld_stream(xb + i0);
      if (xr) *reinterpret_cast<uint4*>(xrv) = ld_stream(xr + i0);
      if (yr) *reinterpret_cast<uint4*>(yrv) = ld_stream(yr + i0);
      if (dyb) *reinterpret_cast<uint4*>(dyv) = ld_stream(dyb + i0);
    } else {
      xv[0] = xb[i0];
      if (xr) xrv[0] = xr[i0];
      if (yr) yrv[0] = yr[i0];
      if (dyb) dyv[0] = dyb[i0];
    }
    // size_x <= INT_MAX (checked by the entry point, as bias_act.cpp:40 does): 32-bit index arithmetic
    const unsigned i0u = (unsigned)i0, stepu = (unsigned)p.step_b, sizeu = (unsigned)p.size_b;
    S bias0 = 0;
    if (bb && bias_per_vec) bias0 = to_s<T>(bb[(i0u / stepu) % sizeu]);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      S bj = bias0;
      if (bb && !bias_per_vec) bj = to_s<T>(bb[((i0u + (unsigned)j) / stepu) % sizeu]);
      out[j] = from_s<T, S>(bias_act_elem<S, A, G>(to_s<T>(xv[j]), bj, xr ? to_s<T>(xrv[j]) : (S)0, yr ? to_s<T>(yrv[j]) : (S)0,
                                                    dyb ? to_s<T>(dyv[j]) : (S)1, alpha, gain, clamp));
    }
    if (VEC > 1) st_stream(yb + i0, *reinterpret_cast<uint4*>(out));
    else yb[i0] = out[0];
  }
  // tail (size_x % VEC elements), handled by the first threads of block 0
  if (VEC > 1 && blockIdx.x == 0) {
    const long long i = nvec * VEC + threadIdx.x;
    if (i < p.size_x) {
      S bj = bb ? to_s<T>(bb[(i / p.step_b) % p.size_b]) : (S)0;
      yb[i] = from_s<T, S>(bias_act_elem<S, A, G>(to_s<T>(xb[i]), bj, xr ? to_s<T>(xr[i]) : (S)0, yr ? to_s<T>(yr[i]) : (S)0,
                                                   dyb ? to_s<T>(dyb[i]) : (S)1, alpha, gain, clamp));
    }
  }
}

template <class T, int VEC>
static int launch_act(const BiasActArgs& p, int act, cudaStream_t st) {
  const long long nvec = p.size_x / VEC;
  long long blocks = ceil_div_ll(nvec > 0 ? nvec : 1, 256);
  const long long cap = (long long)kNumSMs * 16;  // 16 resident 256-thread CTAs fill the 64 warps/SM twice over
  if (blocks > cap) blocks = cap;
  const int g = (int)blocks;
#define SMC_BA_CASE(A)                                                         \
  case A:                                                                      \
    if (p.grad == 0) bias_act_kernel<T, A, VEC, 0><<<g, 256, 0, st>>>(p);      \
    else if (p.grad == 1) bias_act_kernel<T, A, VEC, 1><<<g, 256, 0, st>>>(p); \
    else bias_act_kernel<T, A, VEC, 2><<<g, 256, 0, st>>>(p);                  \
    break;
  switch (act) {
    SMC_BA_CASE(1) SMC_BA_CASE(2) SMC_BA_CASE(3) SMC_BA_CASE(4) SMC_BA_CASE(5) SMC_BA_CASE(6) SMC_BA_CASE(7) SMC_BA_CASE(8) SMC_BA_CASE(9)
    default: return SMC_EINVAL;
  }
#undef SMC_BA_CASE
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

}  // namespace smc

extern "C" int smc_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                            int dtype, int64_t size_x, int32_t size_b, int64_t step_b, int grad, int act, float alpha,
                            float gain, float clamp, void* stream) {
  using namespace smc;
  if (size_x == 0) return SMC_OK;
  if (!x || !y || size_x < 0) return SMC_EINVAL;
  if (grad < 0 || grad > 2) return SMC_EINVAL;
  if (act < 1 || act > 9) return SMC_EINVAL;
  if (b && (size_b < 1 || step_b < 1)) return SMC_EINVAL;
  BiasActArgs p{x, b, xref, yref, dy, y, (long long)size_x, b ? size_b : 1, b ? (long long)step_b : 1, grad, alpha, gain, clamp};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const uintptr_t all = (uintptr_t)x | (uintptr_t)y | (uintptr_t)xref | (uintptr_t)yref | (uintptr_t)dy;
  const bool aligned = (all & 15) == 0;
  switch (dtype) {
    case SMC_F32: return aligned ? launch_act<float, 4>(p, act, st) : launch_act<float, 1>(p, act, st);
    case SMC_F16: return aligned ? launch_act<__half, 8>(p, act, st) : launch_act<__half, 1>(p, act, st);
    case SMC_F64: return aligned ? launch_act<double, 2>(p, act, st) : launch_act<double, 1>(p, act, st);
    default: return SMC_EUNSUPPORTED;
  }
}
