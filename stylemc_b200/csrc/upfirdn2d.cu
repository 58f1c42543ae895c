// upfirdn2d for sm_100a: pad -> zero-upsample -> 2-D FIR -> decimate, per (n, c) plane.
//
// Drop-in for the reference plugin entry `_plugin.upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1,
// pady0, pady1, flip, gain)` (torch_utils/ops/upfirdn2d.cpp:16-94; kernels upfirdn2d.cu:29-200), with the
// output allocated by the caller.  Semantics (upfirdn2d.py:168-208):
//   y[oy, ox] = gain * sum_{ky,kx} fk[ky,kx] * xup[oy*downy + ky - pady0, ox*downx + kx - padx0]
//   xup[a, b] = x[a/upy, b/upx] when divisible and inside the image, else 0
//   fk = f when flip else f rotated by 180 degrees
// HBM-bound: algorithmic bytes = (in + out) * sizeof(T).  Two kernels:
//   * tiled  (contiguous NCHW, filter <= 8x8): the input footprint of a 16 x 128 output tile is staged in
//     shared memory with zero fill, the filter taps sit in shared memory, each thread produces 8 outputs;
//     global reads are row-contiguous and halo re-reads are served by L2.
//   * direct (any strides / filter size): one thread per output, polyphase tap loop.
#include "common.cuh"
#include "stylemc_b200.h"

namespace smc {

struct UpfirdnArgs {
  const void* x; const float* f; void* y;
  int N, C, inH, inW, outH, outW;
  long long xs_n, xs_c, xs_h, xs_w, ys_n, ys_c, ys_h, ys_w;
  int fH, fW; long long fs_h, fs_w;
  int upx, upy, downx, downy, padx0, pady0, flip;
  float gain;
};

template <class T> struct Acc { typedef float type; };
template <> struct Acc<double> { typedef double type; };
template <class T> __device__ __forceinline__ typename Acc<T>::type ld_as(const T* p) { return (typename Acc<T>::type)(*p); }
template <> __device__ __forceinline__ float ld_as<__half>(const __half* p) { return __half2float(*p); }
template <class T, class S> __device__ __forceinline__ void st_as(T* p, S v) { *p = (T)v; }
template <> __device__ __forceinline__ void st_as<__half, float>(__half* p, float v) { *p = __float2half_rn(v); }

__host__ __device__ __forceinline__ int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
__device__ __forceinline__ int pos_mod(int a, int b) { int r = a % b; return r < 0 ? r + b : r; }

constexpr int kTileH = 16, kTileW = 128, kMaxTaps = 8;

template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_tiled_kernel(const UpfirdnArgs p, int tiles_x, int tiles_y, int in_tile_h, int in_tile_w) {
  typedef typename Acc<T>::type S;
  extern __shared__ __align__(16) unsigned char smem_u8[];
  S* sf = reinterpret_cast<S*>(smem_u8);                 // [fH][fW], already flipped + gain
  S* sx = sf + kMaxTaps * kMaxTaps;                      // [in_tile_h][in_tile_w]
  const int tile = blockIdx.x;
  const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y;
  const long long plane = tile / (tiles_x * tiles_y);    // n * C + c
  const int oy0 = ty * kTileH, ox0 = tx * kTileW;
  for (int i = threadIdx.x; i < p.fH * p.fW; i += blockDim.x) {
    const int ky = i / p.fW, kx = i % p.fW;
    const int sy = p.flip ? ky : p.fH - 1 - ky, sxi = p.flip ? kx : p.fW - 1 - kx;
    sf[ky * p.fW + kx] = (S)(p.f[sy * p.fs_h + sxi * p.fs_w] * p.gain);
  }
  // input footprint of this tile
  const int iy0 = floor_div(oy0 * p.downy - p.pady0 + p.upy - 1, p.upy);   // first input row that can be touched
  const int ix0 = floor_div(ox0 * p.downx - p.padx0 + p.upx - 1, p.upx);
  const T* xp = reinterpret_cast<const T*>(p.x) + plane * (long long)p.inH * p.inW;
  for (int i = threadIdx.x; i < in_tile_h * in_tile_w; i += blockDim.x) {
    const int r = i / in_tile_w, c = i - r * in_tile_w;
    const int iy = iy0 + r, ix = ix0 + c;
    S v = 0;
    if (iy >= 0 && iy < p.inH && ix >= 0 && ix < p.inW) v = ld_as<T>(xp + (long long)iy * p.inW + ix);
    sx[i] = v;
  }
  __syncthreads();
  T* yp = reinterpret_cast<T*>(p.y) + plane * (long long)p.outH * p.outW;
  for (int i = threadIdx.x; i < kTileH * kTileW; i += blockDim.x) {
    const int ry = i / kTileW, rx = i - ry * kTileW;
    const int oy = oy0 + ry, ox = ox0 + rx;
    if (oy >= p.outH || ox >= p.outW) continue;
    const int ay = oy * p.downy - p.pady0, ax = ox * p.downx - p.padx0;   // upsampled coordinate of tap (0, 0)
    const int ky0 = pos_mod(-ay, p.upy), kx0 = pos_mod(-ax, p.upx);
    S acc = 0;
    for (int ky = ky0; ky < p.fH; ky += p.upy) {
      const int sr = (ay + ky) / p.upy - iy0;   // exact division (ay + ky is a multiple of upy)
      const S* row = sx + sr * in_tile_w;
      const S* frow = sf + ky * p.fW;
      for (int kx = kx0; kx < p.fW; kx += p.upx) acc += frow[kx] * row[(ax + kx) / p.upx - ix0];
    }
    st_as<T, S>(yp + (long long)oy * p.outW + ox, acc);
  }
}

template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_direct_kernel(const UpfirdnArgs p, long long total) {
  typedef typename Acc<T>::type S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % p.outW);
    long long t = i / p.outW;
    const int oy = (int)(t % p.outH);
    t /= p.outH;
    const int c = (int)(t % p.C);
    const int n = (int)(t / p.C);
    const T* xp = reinterpret_cast<const T*>(p.x) + n * p.xs_n + c * p.xs_c;
    const int ay = oy * p.downy - p.pady0, ax = ox * p.downx - p.padx0;
    S acc = 0;
    for (int ky = pos_mod(-ay, p.upy); ky < p.fH; ky += p.upy) {
      const int iy = (ay + ky) / p.upy;
      if (iy < 0 || iy >= p.inH) continue;
      const int fy = p.flip ? ky : p.fH - 1 - ky;
      for (int kx = pos_mod(-ax, p.upx); kx < p.fW; kx += p.upx) {
        const int ix = (ax + kx) / p.upx;
        if (ix < 0 || ix >= p.inW) continue;
        const int fx = p.flip ? kx : p.fW - 1 - kx;
        acc += (S)p.f[fy * p.fs_h + fx * p.fs_w] * ld_as<T>(xp + iy * p.xs_h + ix * p.xs_w);
      }
    }
    st_as<T, S>(reinterpret_cast<T*>(p.y) + n * p.ys_n + c * p.ys_c + oy * p.ys_h + ox * p.ys_w, acc * (S)p.gain);
  }
}

template <class T>
static int launch_upfirdn(const UpfirdnArgs& p, cudaStream_t st) {
  typedef typename Acc<T>::type S;
  const bool x_nchw = p.xs_w == 1 && p.xs_h == p.inW && p.xs_c == (long long)p.inH * p.inW && p.xs_n == p.xs_c * p.C;
  const bool y_nchw = p.ys_w == 1 && p.ys_h == p.outW && p.ys_c == (long long)p.outH * p.outW && p.ys_n == p.ys_c * p.C;
  if (x_nchw && y_nchw && p.fH <= kMaxTaps && p.fW <= kMaxTaps) {
    const int in_tile_h = ((kTileH - 1) * p.downy + p.fH - 1) / p.upy + 2;
    const int in_tile_w = ((kTileW - 1) * p.downx + p.fW - 1) / p.upx + 2;
    const size_t smem = (size_t)(kMaxTaps * kMaxTaps + in_tile_h * in_tile_w) * sizeof(S);
    if (smem <= 48 * 1024) {
      const int tiles_x = ceil_div(p.outW, kTileW), tiles_y = ceil_div(p.outH, kTileH);
      const long long grid = (long long)tiles_x * tiles_y * p.N * p.C;
      if (grid > 0x7fffffffLL) return SMC_ETOOLARGE;
      upfirdn2d_tiled_kernel<T><<<(int)grid, 256, smem, st>>>(p, tiles_x, tiles_y, in_tile_h, in_tile_w);
      SMC_LAUNCH_CHECK();
      return SMC_OK;
    }
  }
  const long long total = (long long)p.N * p.C * p.outH * p.outW;
  long long blocks = ceil_div_ll(total, 256);
  if (blocks > (long long)kNumSMs * 32) blocks = (long long)kNumSMs * 32;
  upfirdn2d_direct_kernel<T><<<(int)blocks, 256, 0, st>>>(p, total);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

}  // namespace smc

extern "C" int smc_upfirdn2d(const void* x, const float* f, void* y, int dtype, const smc_upfirdn2d_params* q, void* stream) {
  using namespace smc;
  if (!x || !f || !y || !q) return SMC_EINVAL;
  if (q->N < 1 || q->C < 1 || q->inH < 1 || q->inW < 1) return SMC_EINVAL;
  if (q->fH < 1 || q->fW < 1) return SMC_EINVAL;                                       // upfirdn2d.cpp:26
  if (q->upx < 1 || q->upy < 1 || q->downx < 1 || q->downy < 1) return SMC_EINVAL;     // upfirdn2d.cpp:27-28
  if (q->outH < 1 || q->outW < 1) return SMC_EINVAL;                                   // upfirdn2d.cpp:34
  if ((long long)q->N * q->C * q->inH * q->inW > 0x7fffffffLL || (long long)q->N * q->C * q->outH * q->outW > 0x7fffffffLL)
    return SMC_ETOOLARGE;                                                                // upfirdn2d.cpp:22,36
  UpfirdnArgs p;
  p.x = x; p.f = f; p.y = y;
  p.N = q->N; p.C = q->C; p.inH = q->inH; p.inW = q->inW; p.outH = q->outH; p.outW = q->outW;
  p.xs_n = q->x_stride[0]; p.xs_c = q->x_stride[1]; p.xs_h = q->x_stride[2]; p.xs_w = q->x_stride[3];
  p.ys_n = q->y_stride[0]; p.ys_c = q->y_stride[1]; p.ys_h = q->y_stride[2]; p.ys_w = q->y_stride[3];
  p.fH = q->fH; p.fW = q->fW; p.fs_h = q->f_stride[0]; p.fs_w = q->f_stride[1];
  p.upx = q->upx; p.upy = q->upy; p.downx = q->downx; p.downy = q->downy;
  p.padx0 = q->padx0; p.pady0 = q->pady0; p.flip = q->flip ? 1 : 0; p.gain = q->gain;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dtype) {
    case SMC_F32: return launch_upfirdn<float>(p, st);
    case SMC_F16: return launch_upfirdn<__half>(p, st);
    case SMC_F64: return launch_upfirdn<double>(p, st);
    default: return SMC_EUNSUPPORTED;
  }
}
