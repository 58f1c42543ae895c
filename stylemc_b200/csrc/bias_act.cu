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

// VEC elements per vector; VEC * sizeof(T) == 16 on the fast path, VEC == 1 otherwise.  On the fast path a thread keeps UNR
// independent vectors in flight per iteration (all loads issued before the first use): with 16-bit elements the conversion
// and activation math between a load and the next one otherwise leaves too few bytes in flight to cover the HBM latency
// (measured: fp16 lrelu 0.69-0.75 of peak with one vector per iteration while fp32 reached 0.96).
template <class T, int A, int VEC, int G, int UNR>
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
  // size_x <= INT_MAX (checked by the entry point, as bias_act.cpp:40 does): 32-bit index arithmetic
  const unsigned stepu = (unsigned)p.step_b, sizeu = (unsigned)p.size_b;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvec; v0 += stride * UNR) {
    __align__(16) T xv[UNR][VEC]; __align__(16) T xrv[UNR][VEC]; __align__(16) T yrv[UNR][VEC]; __align__(16) T dyv[UNR][VEC];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long v = v0 + u * stride;
      if (v < nvec) {
        const long long i0 = v * VEC;
        if (VEC > 1) {
          *reinterpret_cast<uint4*>(xv[u]) = ld_stream(xb + i0);
          if (xr) *reinterpret_cast<uint4*>(xrv[u]) = ld_stream(xr + i0);
          if (yr) *reinterpret_cast<uint4*>(yrv[u]) = ld_stream(yr + i0);
          if (dyb) *reinterpret_cast<uint4*>(dyv[u]) = ld_stream(dyb + i0);
        } else {
          xv[u][0] = xb[i0];
          if (xr) xrv[u][0] = xr[i0];
          if (yr) yrv[u][0] = yr[i0];
          if (dyb) dyv[u][0] = dyb[i0];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long v = v0 + u * stride;
      if (v < nvec) {
        const long long i0 = v * VEC;
        const unsigned i0u = (unsigned)i0;
        __align__(16) T out[VEC];
        S bias0 = 0;
        if (bb && bias_per_vec) bias0 = to_s<T>(bb[(i0u / stepu) % sizeu]);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          S bj = bias0;
          if (bb && !bias_per_vec) bj = to_s<T>(bb[((i0u + (unsigned)j) / stepu) % sizeu]);
          out[j] = from_s<T, S>(bias_act_elem<S, A, G>(to_s<T>(xv[u][j]), bj, xr ? to_s<T>(xrv[u][j]) : (S)0, yr ? to_s<T>(yrv[u][j]) : (S)0,
                                                        dyb ? to_s<T>(dyv[u][j]) : (S)1, alpha, gain, clamp));
        }
        if (VEC > 1) st_stream(yb + i0, *reinterpret_cast<uint4*>(out));
        else yb[i0] = out[0];
      }
    }
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

// ---- lean fp16 path -----------------------------------------------------------------------------------------------
// The generic kernel above is instruction-issue bound for 16-bit elements: ~130 thread-instructions per 32 bytes moved
// (per-element predicated selects, an IEEE division yref / gain in the backward pass, 64-bit index math, two integer
// divisions for the bias index), i.e. ~75 % of the SM's issue slots at HBM speed -- measured 0.60-0.76 of peak in fp16
// while the same kernel reaches 0.96-1.00 in fp32.  The activations the synthesis network uses (linear, relu, lrelu;
// forward and first derivative) therefore get a kernel with ~6 instructions per element: packed half2 <-> float2
// conversions, packed fp32x2 multiplies, lrelu(x) = max(x, alpha x) (exact for 0 <= alpha <= 1), clamp = min(max()),
// the sign of y instead of y / gain (same predicate for gain > 0), one multiply-shift division per vector for the bias
// channel.  Every output value is bit-identical to the generic kernel's (same fp32 operations in the same order).
struct FastDiv {
  uint32_t mul, shift;   // n / d == (umulhi(n, mul) + n) >> shift for n < 2^31
};
static FastDiv make_fastdiv(uint32_t d) {
  FastDiv f{0u, 0u};
  if (d <= 1) return f;
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.shift = s;
  f.mul = (uint32_t)((((1ull << 32) * ((1ull << s) - d)) / d) + 1ull);
  return f;
}
__device__ __forceinline__ uint32_t fast_div(uint32_t n, FastDiv f) { return (__umulhi(n, f.mul) + n) >> f.shift; }

struct BiasActLeanArgs {
  const __half* x; const __half* b; const __half* yref; __half* y;
  uint32_t nvec;           // 8-element vectors
  uint32_t step_vec;       // step_b / 8
  uint32_t size_b;
  FastDiv div_step, div_size;
  float alpha, gain, clamp;
};

// 8 halves (one uint4) -> 4 float2
__device__ __forceinline__ void h8_to_f2(const uint4& u, float2 (&f)[4]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) f[i] = __half22float2(h[i]);
}

template <int A, int G, bool HAS_B, bool HAS_Y, bool CLAMP>
__global__ void __launch_bounds__(256) bias_act_h_kernel(const BiasActLeanArgs p) {
  constexpr int UNR = 2;
  const uint32_t stride = gridDim.x * blockDim.x;
  const float2 g2 = make_float2(p.gain, p.gain), a2 = make_float2(p.alpha, p.alpha), one2 = make_float2(1.f, 1.f);
  const float cl = p.clamp;
  for (uint32_t v0 = blockIdx.x * blockDim.x + threadIdx.x; v0 < p.nvec; v0 += stride * UNR) {
    uint4 xv[UNR], yv[UNR];
    float bias[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const uint32_t v = v0 + u * stride;
      if (v < p.nvec) {
        xv[u] = ld_stream(p.x + (size_t)v * 8);
        if (HAS_Y) yv[u] = ld_stream(p.yref + (size_t)v * 8);
        if (HAS_B) {
          const uint32_t pl = fast_div(v, p.div_step);                     // (8 v) / step_b
          const uint32_t c = pl - fast_div(pl, p.div_size) * p.size_b;     // % size_b
          bias[u] = __half2float(__ldg(p.b + c));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const uint32_t v = v0 + u * stride;
      if (v < p.nvec) {
        float2 x[4], yr[4];
        h8_to_f2(xv[u], x);
        if (HAS_Y) h8_to_f2(yv[u], yr);
        uint32_t out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 t = x[i];
          if (G == 0) {
            if (HAS_B) t = ffma2(t, one2, make_float2(bias[u], bias[u]));          // x + b (fma(x, 1, b) rounds like the add)
            if (A == 2) t = make_float2(fmaxf(t.x, 0.f), fmaxf(t.y, 0.f));
            if (A == 3) { const float2 ta = fmul2(t, a2); t = make_float2(fmaxf(t.x, ta.x), fmaxf(t.y, ta.y)); }
            t = fmul2(t, g2);
            if (CLAMP) t = make_float2(fminf(fmaxf(t.x, -cl), cl), fminf(fmaxf(t.y, -cl), cl));
          } else {
            if (A == 2) t = make_float2(yr[i].x > 0.f ? t.x : 0.f, yr[i].y > 0.f ? t.y : 0.f);
            if (A == 3) { const float2 ta = fmul2(t, a2); t = make_float2(yr[i].x > 0.f ? t.x : ta.x, yr[i].y > 0.f ? t.y : ta.y); }
            t = fmul2(t, g2);
            if (CLAMP && HAS_Y) t = make_float2((yr[i].x > -cl && yr[i].x < cl) ? t.x : 0.f, (yr[i].y > -cl && yr[i].y < cl) ? t.y : 0.f);
          }
          const __half2 h = __floats2half2_rn(t.x, t.y);
          out[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        st_stream(p.y + (size_t)v * 8, make_uint4(out[0], out[1], out[2], out[3]));
      }
    }
  }
}

template <int A, int G, bool HAS_B>
static void launch_lean_h(const BiasActLeanArgs& a, bool has_y, bool clamp, int grid, cudaStream_t st) {
  if (has_y) {
    if (clamp) bias_act_h_kernel<A, G, HAS_B, true, true><<<grid, 256, 0, st>>>(a);
    else bias_act_h_kernel<A, G, HAS_B, true, false><<<grid, 256, 0, st>>>(a);
  } else {
    if (clamp) bias_act_h_kernel<A, G, HAS_B, false, true><<<grid, 256, 0, st>>>(a);
    else bias_act_h_kernel<A, G, HAS_B, false, false><<<grid, 256, 0, st>>>(a);
  }
}

// Returns SMC_EUNSUPPORTED when the call does not fit the lean path (the caller then takes the generic kernel).
static int try_launch_lean_h(const BiasActArgs& p, int act, cudaStream_t st) {
  if (act < 1 || act > 3 || p.grad > 1 || p.dy != nullptr) return SMC_EUNSUPPORTED;
  if (p.size_x % 8 != 0 || p.size_x / 8 > 0x7fffffffLL) return SMC_EUNSUPPORTED;
  if (!(p.gain > 1e-30f && p.gain < 1e30f)) return SMC_EUNSUPPORTED;
  if (act == 3 && !(p.alpha >= 0.f && p.alpha <= 1.f)) return SMC_EUNSUPPORTED;
  const bool has_b = p.b != nullptr && p.grad == 0;
  if (has_b && (p.step_b % 8 != 0 || p.step_b / 8 > 0x7fffffffLL)) return SMC_EUNSUPPORTED;
  const bool clamp = p.clamp >= 0.f;
  const bool need_y = p.grad == 1 && (act != 1 || clamp);
  if (need_y && p.yref == nullptr) return SMC_EUNSUPPORTED;     // the generic kernel defines what a missing yref means
  BiasActLeanArgs a;
  a.x = (const __half*)p.x; a.b = (const __half*)p.b; a.yref = (const __half*)p.yref; a.y = (__half*)p.y;
  a.nvec = (uint32_t)(p.size_x / 8);
  a.step_vec = has_b ? (uint32_t)(p.step_b / 8) : 1u;
  a.size_b = has_b ? (uint32_t)p.size_b : 1u;
  a.div_step = make_fastdiv(a.step_vec);
  a.div_size = make_fastdiv(a.size_b);
  a.alpha = p.alpha; a.gain = p.gain; a.clamp = p.clamp;
  long long blocks = ceil_div_ll(a.nvec, 256 * 2);
  const long long cap = (long long)kNumSMs * 16;
  const int grid = (int)(blocks > cap ? cap : blocks);
#define SMC_BAH_CASE(A)                                                                          \
  case A:                                                                                        \
    if (p.grad == 0) { if (has_b) launch_lean_h<A, 0, true>(a, false, clamp, grid, st); else launch_lean_h<A, 0, false>(a, false, clamp, grid, st); } \
    else launch_lean_h<A, 1, false>(a, need_y, clamp, grid, st);                                 \
    break;
  switch (act) { SMC_BAH_CASE(1) SMC_BAH_CASE(2) SMC_BAH_CASE(3) }
#undef SMC_BAH_CASE
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

template <class T, int VEC>
static int launch_act(const BiasActArgs& p, int act, cudaStream_t st) {
  const long long nvec = p.size_x / VEC;
  constexpr int UNR = (VEC > 1 && sizeof(T) <= 2) ? 2 : 1;  // 16-bit types: two vectors in flight per thread
  long long blocks = ceil_div_ll(nvec > 0 ? nvec : 1, 256 * UNR);
  const long long cap = (long long)kNumSMs * 16;  // 16 resident 256-thread CTAs fill the 64 warps/SM twice over
  if (blocks > cap) blocks = cap;
  const int g = (int)blocks;
#define SMC_BA_CASE(A)                                                         \
  case A:                                                                      \
    if (p.grad == 0) bias_act_kernel<T, A, VEC, 0, UNR><<<g, 256, 0, st>>>(p);      \
    else if (p.grad == 1) bias_act_kernel<T, A, VEC, 1, UNR><<<g, 256, 0, st>>>(p); \
    else bias_act_kernel<T, A, VEC, 2, UNR><<<g, 256, 0, st>>>(p);                  \
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
    case SMC_F16: {
      if (aligned) {
        const int r = try_launch_lean_h(p, act, st);
        if (r != SMC_EUNSUPPORTED) return r;
      }
      return aligned ? launch_act<__half, 8>(p, act, st) : launch_act<__half, 1>(p, act, st);
    }
    case SMC_F64: return aligned ? launch_act<double, 2>(p, act, st) : launch_act<double, 1>(p, act, st);
    default: return SMC_EUNSUPPORTED;
  }
}
