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
  int separable; float fsy[4], fsx[4];   // f[i][j] == fsy[i] * fsx[j] when separable (host-side hint)
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

// ---- fast path: 4x4 filter, contiguous NCHW, (up, down) in {(1,1), (2,1), (1,2)}, fp32 / fp16 ----------------------------------
// These are the three cases on the hot path (upfirdn2d.cu:217,252,298 in the reference): the FIR after the
// transposed conv, upsample2d of the skip image, and its transpose.  One thread = one output column x YB consecutive output rows of
// one (n, c) plane; lanes run along x, so every load instruction of a warp is one (or two, down = 2) contiguous 128-byte line and
// the taps' horizontal overlap is served by L1.  The filter (flipped, times gain) sits in registers; trip counts are compile-time.
template <class T> __device__ __forceinline__ float ldg_f(const T* p) { return (float)__ldg(p); }
template <> __device__ __forceinline__ float ldg_f<__half>(const __half* p) { return __half2float(__ldg(p)); }

template <class T, int UP, int DOWN, int YB>
__global__ void __launch_bounds__(256) upfirdn2d_fast_kernel(const UpfirdnArgs p) {
  constexpr int F = 4;
  float fk[F][F];                       // fk[ky][kx] = (flip ? f[ky][kx] : f[F-1-ky][F-1-kx]) * gain
#pragma unroll
  for (int ky = 0; ky < F; ++ky)
#pragma unroll
    for (int kx = 0; kx < F; ++kx) {
      const int sy = p.flip ? ky : F - 1 - ky, sx = p.flip ? kx : F - 1 - kx;
      fk[ky][kx] = __ldg(p.f + sy * p.fs_h + sx * p.fs_w) * p.gain;
    }
  const int ox = blockIdx.x * 128 + (threadIdx.x & 127);
  const int oyb = (blockIdx.y * 2 + (threadIdx.x >> 7)) * YB;
  const long long plane = blockIdx.z;
  if (ox >= p.outW || oyb >= p.outH) return;
  const T* __restrict__ xp = reinterpret_cast<const T*>(p.x) + plane * (long long)p.inH * p.inW;
  T* __restrict__ yp = reinterpret_cast<T*>(p.y) + plane * (long long)p.outH * p.outW;
  float acc[YB];
#pragma unroll
  for (int j = 0; j < YB; ++j) acc[j] = 0.f;
  if (UP == 1) {
    // rows r of the input window: iy = oyb * DOWN - pady0 + r feeds output j with tap ky = r - j * DOWN
    const int ix0 = ox * DOWN - p.padx0, iy0 = oyb * DOWN - p.pady0;
#pragma unroll
    for (int r = 0; r < (YB - 1) * DOWN + F; ++r) {
      const int iy = iy0 + r;
      float v[F];
      const bool rowok = iy >= 0 && iy < p.inH;
#pragma unroll
      for (int kx = 0; kx < F; ++kx) {
        const int ix = ix0 + kx;
        v[kx] = (rowok && ix >= 0 && ix < p.inW) ? ldg_f<T>(xp + (long long)iy * p.inW + ix) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < YB; ++j) {
        const int ky = r - j * DOWN;
        if (ky >= 0 && ky < F) {
#pragma unroll
          for (int kx = 0; kx < F; ++kx) acc[j] += fk[ky][kx] * v[kx];
        }
      }
    }
  } else {
    // zero-insertion up-sampling: only taps with (a + k) % UP == 0 see data; two taps per axis for UP = 2, F = 4
    const int ax = ox - p.padx0;
    const int kx0 = pos_mod(-ax, UP);
#pragma unroll
    for (int j = 0; j < YB; ++j) {
      const int ay = oyb + j - p.pady0;
      const int ky0 = pos_mod(-ay, UP);
#pragma unroll
      for (int a = 0; a < F / UP; ++a) {
        const int iy = (ay + ky0 + a * UP) / UP;       // exact
        const bool rowok = iy >= 0 && iy < p.inH;
#pragma unroll
        for (int b = 0; b < F / UP; ++b) {
          const int ix = (ax + kx0 + b * UP) / UP;
          const float v = (rowok && ix >= 0 && ix < p.inW) ? ldg_f<T>(xp + (long long)iy * p.inW + ix) : 0.f;
          // coefficient fk[ky0 + a UP][kx0 + b UP] with ky0, kx0 in {0, 1}: selected without dynamic indexing
          const float c0 = kx0 ? fk[a * UP][b * UP + 1] : fk[a * UP][b * UP];
          const float c1 = kx0 ? fk[a * UP + 1][b * UP + 1] : fk[a * UP + 1][b * UP];
          acc[j] += (ky0 ? c1 : c0) * v;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < YB; ++j)
    if (oyb + j < p.outH) st_as<T, float>(yp + (long long)(oyb + j) * p.outW + ox, acc[j]);
}

// up = 1 (plain / decimating FIR): the input footprint of a 128 x TOH output tile is staged once in shared memory (coalesced loads,
// zero fill = the padding), then each thread walks YB output rows of one column with a register window: (YB - 1) * DOWN + 4 rows x 4
// conflict-free LDS for YB outputs, no bounds checks or address arithmetic in the FIR loop.
template <class T, int DOWN>
__global__ void __launch_bounds__(256) upfirdn2d_tile_kernel(const UpfirdnArgs p) {
  constexpr int F = 4, TOW = 128, TOH = DOWN == 1 ? 32 : 16, YB = TOH / 2;
  constexpr int IW = (TOW - 1) * DOWN + F, IH = (TOH - 1) * DOWN + F;
  __shared__ float sx[IH * IW];
  float fk[F][F];
#pragma unroll
  for (int ky = 0; ky < F; ++ky)
#pragma unroll
    for (int kx = 0; kx < F; ++kx) {
      const int sy = p.flip ? ky : F - 1 - ky, sxi = p.flip ? kx : F - 1 - kx;
      fk[ky][kx] = __ldg(p.f + sy * p.fs_h + sxi * p.fs_w) * p.gain;
    }
  const int ox0 = blockIdx.x * TOW, oy0 = blockIdx.y * TOH;
  const long long plane = blockIdx.z;
  const T* __restrict__ xp = reinterpret_cast<const T*>(p.x) + plane * (long long)p.inH * p.inW;
  const int ix0 = ox0 * DOWN - p.padx0, iy0 = oy0 * DOWN - p.pady0;
  {
    // one warp per input row (8 rows in flight per pass).  Tiles whose whole footprint lies inside the image (most of them) skip
    // every bounds test; the others test the row once and the column per element.
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const bool interior = iy0 >= 0 && iy0 + IH <= p.inH && ix0 >= 0 && ix0 + IW <= p.inW;
    if (interior) {
      const T* src = xp + (long long)(iy0 + wrp) * p.inW + ix0 + lane;
      float* dst = sx + wrp * IW + lane;
#pragma unroll
      for (int r = 0; r < IH; r += 8) {
        if (r + wrp < IH) {
#pragma unroll
          for (int c = 0; c < IW; c += 32)
            if (c + 32 <= IW || c + lane < IW) dst[r * IW + c] = ldg_f<T>(src + (long long)r * p.inW + c);
        }
      }
    } else {
#pragma unroll 1
      for (int r = wrp; r < IH; r += 8) {
        const int iy = iy0 + r;
        const bool rowok = iy >= 0 && iy < p.inH;
        const T* rowp = xp + (long long)(rowok ? iy : 0) * p.inW;
        float* srow = sx + r * IW;
#pragma unroll
        for (int c = 0; c < IW; c += 32) {
          const int cc = c + lane, ix = ix0 + cc;
          if (cc < IW) srow[cc] = (rowok && ix >= 0 && ix < p.inW) ? ldg_f<T>(rowp + ix) : 0.f;
        }
      }
    }
  }
  __syncthreads();
  const int tx = threadIdx.x & 127, tg = threadIdx.x >> 7;
  float acc[YB];
#pragma unroll
  for (int j = 0; j < YB; ++j) acc[j] = 0.f;
  const float* base = sx + (tg * YB * DOWN) * IW + tx * DOWN;
  // rank-1 filter (host hint; e.g. the [1,3,3,1] resample filter): a window row costs 4 FMAs + one per tap row that uses it instead
  // of 4 per tap row.  fk = flip ? f : rot180(f), times gain: the same index map applies to each factor.
  float fx[F], fy[F];
  const bool separable = p.separable != 0;
#pragma unroll
  for (int k = 0; k < F; ++k) {
    fy[k] = p.fsy[p.flip ? k : F - 1 - k] * p.gain;
    fx[k] = p.fsx[p.flip ? k : F - 1 - k];
  }
  if (separable) {
#pragma unroll
    for (int r = 0; r < (YB - 1) * DOWN + F; ++r) {
      float h = 0.f;
#pragma unroll
      for (int kx = 0; kx < F; ++kx) h += fx[kx] * base[r * IW + kx];
#pragma unroll
      for (int j = 0; j < YB; ++j) {
        const int ky = r - j * DOWN;
        if (ky >= 0 && ky < F) acc[j] += fy[ky] * h;
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < (YB - 1) * DOWN + F; ++r) {
      float v[F];
#pragma unroll
      for (int kx = 0; kx < F; ++kx) v[kx] = base[r * IW + kx];
#pragma unroll
      for (int j = 0; j < YB; ++j) {
        const int ky = r - j * DOWN;
        if (ky >= 0 && ky < F) {
#pragma unroll
          for (int kx = 0; kx < F; ++kx) acc[j] += fk[ky][kx] * v[kx];
        }
      }
    }
  }
  const int ox = ox0 + tx;
  if (ox < p.outW) {
    T* __restrict__ yp = reinterpret_cast<T*>(p.y) + plane * (long long)p.outH * p.outW + ox;
#pragma unroll
    for (int j = 0; j < YB; ++j) {
      const int oy = oy0 + tg * YB + j;
      if (oy < p.outH) st_as<T, float>(yp + (long long)oy * p.outW, acc[j]);
    }
  }
}

// up = 2, down = 1 (upsample2d: zero insertion + 4x4 FIR): polyphase form, 2 x 2 input taps per output.  The (TOH/2 + 2) x (TOW/2 + 2)
// input footprint of a 128 x 32 output tile is staged in shared memory; a thread owns one output column (fixed column phase) and
// 16 rows whose row phase alternates, so its 2 x 4 tap coefficients are selected once.
template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_up2_tile_kernel(const UpfirdnArgs p) {
  constexpr int F = 4, TOW = 128, TOH = 32, YB = TOH / 2, IW = TOW / 2 + 2, IH = TOH / 2 + 2;
  __shared__ float sx[IH * IW];
  float fk[F][F];
#pragma unroll
  for (int ky = 0; ky < F; ++ky)
#pragma unroll
    for (int kx = 0; kx < F; ++kx) {
      const int sy = p.flip ? ky : F - 1 - ky, sxi = p.flip ? kx : F - 1 - kx;
      fk[ky][kx] = __ldg(p.f + sy * p.fs_h + sxi * p.fs_w) * p.gain;
    }
  const int ox0 = blockIdx.x * TOW, oy0 = blockIdx.y * TOH;
  const long long plane = blockIdx.z;
  const T* __restrict__ xp = reinterpret_cast<const T*>(p.x) + plane * (long long)p.inH * p.inW;
  const int ix0 = floor_div(ox0 - p.padx0 + 1, 2), iy0 = floor_div(oy0 - p.pady0 + 1, 2);   // first input column / row a tap can touch
  for (int i = threadIdx.x; i < IH * IW; i += 256) {
    const int r = i / IW, c = i - r * IW;
    const int iy = iy0 + r, ix = ix0 + c;
    sx[i] = (iy >= 0 && iy < p.inH && ix >= 0 && ix < p.inW) ? ldg_f<T>(xp + (long long)iy * p.inW + ix) : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 127, tg = threadIdx.x >> 7;
  const int ox = ox0 + tx;
  const int ax = ox - p.padx0;
  const int kx0 = pos_mod(-ax, 2);
  const int cxb = (ax + kx0) / 2 - ix0;                       // smem column of the first tap (exact division), second tap = +1
  // coefficients of the two column taps for every filter row, column phase folded in
  float c0[F], c1[F];
#pragma unroll
  for (int ky = 0; ky < F; ++ky) {
    c0[ky] = kx0 ? fk[ky][1] : fk[ky][0];
    c1[ky] = kx0 ? fk[ky][3] : fk[ky][2];
  }
  T* __restrict__ yp = reinterpret_cast<T*>(p.y) + plane * (long long)p.outH * p.outW + ox;
#pragma unroll
  for (int j = 0; j < YB; ++j) {
    const int oy = oy0 + tg * YB + j;
    const int ay = oy - p.pady0;
    const int ky0 = pos_mod(-ay, 2);
    const float* r0 = sx + ((ay + ky0) / 2 - iy0) * IW + cxb;   // first row tap; second = next smem row
    const float a00 = ky0 ? c0[1] : c0[0], a01 = ky0 ? c1[1] : c1[0];
    const float a10 = ky0 ? c0[3] : c0[2], a11 = ky0 ? c1[3] : c1[2];
    const float v = a00 * r0[0] + a01 * r0[1] + a10 * r0[IW] + a11 * r0[IW + 1];
    if (ox < p.outW && oy < p.outH) st_as<T, float>(yp + (long long)oy * p.outW, v);
  }
}

template <class T, int DOWN>
static int launch_tile(const UpfirdnArgs& p, cudaStream_t st) {
  constexpr int TOH = DOWN == 1 ? 32 : 16;
  dim3 grid(ceil_div(p.outW, 128), ceil_div(p.outH, TOH), p.N * p.C);
  if (grid.y > 65535 || grid.z > 65535) return SMC_EUNSUPPORTED;
  upfirdn2d_tile_kernel<T, DOWN><<<grid, 256, 0, st>>>(p);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

template <class T, int UP, int DOWN, int YB>
static int launch_fast(const UpfirdnArgs& p, cudaStream_t st) {
  dim3 grid(ceil_div(p.outW, 128), ceil_div(p.outH, 2 * YB), p.N * p.C);
  if (grid.y > 65535 || grid.z > 65535) return SMC_EUNSUPPORTED;
  upfirdn2d_fast_kernel<T, UP, DOWN, YB><<<grid, 256, 0, st>>>(p);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

template <class T> struct FastOk { static constexpr bool value = false; };
template <> struct FastOk<float> { static constexpr bool value = true; };
template <> struct FastOk<__half> { static constexpr bool value = true; };

template <class T>
static int launch_upfirdn(const UpfirdnArgs& p, cudaStream_t st) {
  typedef typename Acc<T>::type S;
  const bool x_nchw = p.xs_w == 1 && p.xs_h == p.inW && p.xs_c == (long long)p.inH * p.inW && p.xs_n == p.xs_c * p.C;
  const bool y_nchw = p.ys_w == 1 && p.ys_h == p.outW && p.ys_c == (long long)p.outH * p.outW && p.ys_n == p.ys_c * p.C;
  if constexpr (FastOk<T>::value) {
    if (x_nchw && y_nchw && p.fH == 4 && p.fW == 4 && p.upx == p.upy && p.downx == p.downy) {
      int r = SMC_EUNSUPPORTED;
      if (p.upx == 1 && p.downx == 1) r = launch_tile<T, 1>(p, st);
      else if (p.upx == 2 && p.downx == 1) {
        dim3 grid(ceil_div(p.outW, 128), ceil_div(p.outH, 32), p.N * p.C);
        if (grid.y <= 65535 && grid.z <= 65535) {
          upfirdn2d_up2_tile_kernel<T><<<grid, 256, 0, st>>>(p);
          SMC_LAUNCH_CHECK();
          r = SMC_OK;
        }
      }
      else if (p.upx == 1 && p.downx == 2) r = launch_tile<T, 2>(p, st);
      if (r != SMC_EUNSUPPORTED) return r;
    }
  }
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
  p.separable = (q->separable != 0 && q->fH == 4 && q->fW == 4) ? 1 : 0;
  for (int i = 0; i < 4; ++i) { p.fsy[i] = q->fsep[i]; p.fsx[i] = q->fsep[4 + i]; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (dtype) {
    case SMC_F32: return launch_upfirdn<float>(p, st);
    case SMC_F16: return launch_upfirdn<__half>(p, st);
    case SMC_F64: return launch_upfirdn<double>(p, st);
    default: return SMC_EUNSUPPORTED;
  }
}
