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

// ---- warp-row streaming kernels: separable 4x4 filter, contiguous NCHW, (up, down) in {(1,1), (2,1), (1,2)} ---------------------------
// One WARP owns a strip of output columns of one (n, c) plane and marches down a segment of rows.  An input row is read once with
// coalesced loads (lane + 32 m), its horizontal neighbours come from warp shuffles (no shared memory, no block barrier, any row
// alignment: the (2H+1)^2 planes of the conv0 FIR have odd widths), the horizontally filtered row enters a 4-row register window and
// every output row is one vertical 4-tap (2-tap for up = 2) combination of window rows.  The next input row is prefetched while the
// current one is filtered.  fk[ky][kx] = fy[ky] * fx[kx] (host hint UpfirdnArgs::separable).
int g_upfirdn_rows = 1;   // 0: keep the tile kernels (A/B diagnostics, smc_synth_config key 3)

template <class T> __device__ __forceinline__ void st2(T* p, float a, float b, bool ok0, bool ok1, bool vec);
template <> __device__ __forceinline__ void st2<float>(float* p, float a, float b, bool ok0, bool ok1, bool vec) {
  if (vec && ok1) { *reinterpret_cast<float2*>(p) = make_float2(a, b); return; }
  if (ok0) p[0] = a;
  if (ok1) p[1] = b;
}
template <> __device__ __forceinline__ void st2<__half>(__half* p, float a, float b, bool ok0, bool ok1, bool vec) {
  if (vec && ok1) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(a, b); return; }
  if (ok0) p[0] = __float2half_rn(a);
  if (ok1) p[1] = __float2half_rn(b);
}

struct RowsFilter { float fx[4], fy[4]; };
__device__ __forceinline__ RowsFilter rows_filter(const UpfirdnArgs& p) {
  RowsFilter f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f.fy[k] = p.fsy[p.flip ? k : 3 - k] * p.gain;
    f.fx[k] = p.fsx[p.flip ? k : 3 - k];
  }
  return f;
}

template <class T, int NJ> struct StN;
template <> struct StN<float, 4> {
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct StN<__half, 8> {
  static __device__ __forceinline__ void st(__half* p, const float (&v)[8]) {
    const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]), c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = make_uint4(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b),
                                              *reinterpret_cast<const uint32_t*>(&c), *reinterpret_cast<const uint32_t*>(&d));
  }
};

// up = down = 1.  A lane owns NJ CONSECUTIVE output columns (one 16-byte store per row).  The input row is read with coalesced scalar
// loads (any row alignment), staged in a per-warp shared-memory row (double buffered, one __syncwarp per row, no block barrier) and
// read back as the lane's NJ + 3 consecutive inputs with 16-byte loads.  (A first version took the neighbours from warp shuffles:
// 27 SHFL + SEL per row made it issue bound at 35 instructions per output.)
template <class T, int RS, int NJ>
__global__ void __launch_bounds__(128) upfirdn2d_rows_up1_kernel(const UpfirdnArgs p, int strips, int segs) {
  constexpr int SW = 32 * NJ, BW = SW + 32;               // strip width, staged row width (NJ + 1 loads per lane)
  __shared__ __align__(16) float buf[4][2][BW];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  long long task = (long long)blockIdx.x * 4 + wrp;
  if (task >= (long long)p.N * p.C * strips * segs) return;
  const int strip = (int)(task % strips); task /= strips;
  const int seg = (int)(task % segs);
  const long long plane = task / segs;
  const int ox0 = strip * SW, oy0 = seg * RS;
  const T* __restrict__ xp = reinterpret_cast<const T*>(p.x) + plane * (long long)p.inH * p.inW;
  T* __restrict__ yp = reinterpret_cast<T*>(p.y) + plane * (long long)p.outH * p.outW;
  const RowsFilter f = rows_filter(p);
  const int ix0 = ox0 - p.padx0 + lane;
  bool cok[NJ + 1];
#pragma unroll
  for (int m = 0; m <= NJ; ++m) cok[m] = ix0 + 32 * m >= 0 && ix0 + 32 * m < p.inW;
  auto load_row = [&](int iy, T (&v)[NJ + 1]) {
    const bool rowok = iy >= 0 && iy < p.inH;
    const T* row = xp + (long long)(rowok ? iy : 0) * p.inW + ix0;
#pragma unroll
    for (int m = 0; m <= NJ; ++m) v[m] = (rowok && cok[m]) ? __ldg(row + 32 * m) : (T)0.f;
  };
  const int rows = (p.outH - oy0 < RS ? p.outH - oy0 : RS) + 3;
  const int iyb = oy0 - p.pady0;
  const int oxl = ox0 + NJ * lane;
  const bool vec = (p.outW % NJ) == 0 && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) && oxl + NJ <= p.outW;
  float hw[4][NJ];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int j = 0; j < NJ; ++j) hw[u][j] = 0.f;
  T pf[2][NJ + 1];                                        // two input rows in flight
  load_row(iyb, pf[0]);
  load_row(iyb + 1, pf[1]);
  for (int r0 = 0; r0 < rows; r0 += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u;
      if (r >= rows) break;
      float* b = buf[wrp][u & 1];
#pragma unroll
      for (int m = 0; m <= NJ; ++m) b[lane + 32 * m] = (float)pf[u & 1][m];
      __syncwarp();
      if (r + 2 < rows) load_row(iyb + r + 2, pf[u & 1]);
      float win[NJ + 4];
#pragma unroll
      for (int q = 0; q < (NJ + 4) / 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(b + NJ * lane + 4 * q);
        win[4 * q] = t.x; win[4 * q + 1] = t.y; win[4 * q + 2] = t.z; win[4 * q + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) hw[u][j] = f.fx[0] * win[j] + f.fx[1] * win[j + 1] + f.fx[2] * win[j + 2] + f.fx[3] * win[j + 3];
      if (r >= 3) {
        const int oy = oy0 + r - 3;
        T* orow = yp + (long long)oy * p.outW + oxl;
        float v[NJ];
#pragma unroll
        for (int j = 0; j < NJ; j += 2) {
          const float2 t = ffma2(make_float2(f.fy[3], f.fy[3]), make_float2(hw[u][j], hw[u][j + 1]),
                                 ffma2(make_float2(f.fy[2], f.fy[2]), make_float2(hw[(u + 3) & 3][j], hw[(u + 3) & 3][j + 1]),
                                       ffma2(make_float2(f.fy[1], f.fy[1]), make_float2(hw[(u + 2) & 3][j], hw[(u + 2) & 3][j + 1]),
                                             fmul2(make_float2(f.fy[0], f.fy[0]), make_float2(hw[(u + 1) & 3][j], hw[(u + 1) & 3][j + 1])))));
          v[j] = t.x; v[j + 1] = t.y;
        }
        if (vec) StN<T, NJ>::st(orow, v);
        else {
#pragma unroll
          for (int j = 0; j < NJ; ++j)
            if (oxl + j < p.outW) st_as<T, float>(orow + j, v[j]);
        }
      }
    }
  }
}

// up = 2 (zero insertion), down = 1: polyphase.  Output column 2b + c reads input columns b + sx[c], b + sx[c] + 1 with taps
// fx[kx0[c]], fx[kx0[c] + 2]; rows alike.  A lane owns b = b0 + lane + 32 m (m = 0, 1) and stores (2b, 2b + 1) pairs.
template <class T, int RA>
__global__ void __launch_bounds__(128) upfirdn2d_rows_up2_kernel(const UpfirdnArgs p, int strips, int segs) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  long long task = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (task >= (long long)p.N * p.C * strips * segs) return;
  // strips fastest: the warps of a block (and of neighbouring blocks) read adjacent chunks of the same input rows (DRAM page locality)
  const int strip = (int)(task % strips); task /= strips;
  const int seg = (int)(task % segs);
  const long long plane = task / segs;
  const int b0 = strip * 64, a0 = seg * RA;
  const T* __restrict__ xp = reinterpret_cast<const T*>(p.x) + plane * (long long)p.inH * p.inW;
  T* __restrict__ yp = reinterpret_cast<T*>(p.y) + plane * (long long)p.outH * p.outW;
  const RowsFilter f = rows_filter(p);
  int sx[2], sy[2];
  float fxa[2], fxb[2], fya[2], fyb[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int kx0 = pos_mod(p.padx0 - c, 2), ky0 = pos_mod(p.pady0 - c, 2);
    sx[c] = floor_div(c - p.padx0 + kx0, 2);
    sy[c] = floor_div(c - p.pady0 + ky0, 2);
    fxa[c] = kx0 ? f.fx[1] : f.fx[0]; fxb[c] = kx0 ? f.fx[3] : f.fx[2];
    fya[c] = ky0 ? f.fy[1] : f.fy[0]; fyb[c] = ky0 ? f.fy[3] : f.fy[2];
  }
  const int msx = sx[0] < sx[1] ? sx[0] : sx[1], msy = sy[0] < sy[1] ? sy[0] : sy[1], Msy = sy[0] < sy[1] ? sy[1] : sy[0];
  const bool d0 = sx[0] != msx, d1 = sx[1] != msx;       // column phase c reads the sequence shifted by (dc, dc + 1)
  const int ix0 = b0 + msx + lane;
  bool cok[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) cok[m] = ix0 + 32 * m >= 0 && ix0 + 32 * m < p.inW;
  auto load_row = [&](int iy, float (&v)[3]) {
    const bool rowok = iy >= 0 && iy < p.inH;
    const T* row = xp + (long long)(rowok ? iy : 0) * p.inW + ix0;
#pragma unroll
    for (int m = 0; m < 3; ++m) v[m] = (rowok && cok[m]) ? ldg_f<T>(row + 32 * m) : 0.f;
  };
  const bool vec = (p.outW & 1) == 0 && ((reinterpret_cast<uintptr_t>(p.y) & 7) == 0);
  float hp[2][2];                       // previous filtered row: [column phase][m]
  hp[0][0] = hp[0][1] = hp[1][0] = hp[1][1] = 0.f;
  const int i_first = a0 + msy, i_last = a0 + RA + Msy;      // input rows a + sy[r], a + sy[r] + 1 for a in [a0, a0 + RA)
  float nxt[3];
  load_row(i_first, nxt);
  for (int i = i_first; i <= i_last; ++i) {
    float in[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) in[m] = nxt[m];
    if (i < i_last) load_row(i + 1, nxt);
    float s1[3], s2[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) { s1[m] = __shfl_sync(full, in[m], (lane + 1) & 31); s2[m] = __shfl_sync(full, in[m], (lane + 2) & 31); }
    float hn[2][2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const float x0 = in[m];
      const float x1 = lane + 1 < 32 ? s1[m] : s1[m + 1];
      const float x2 = lane + 2 < 32 ? s2[m] : s2[m + 1];
      hn[0][m] = fxa[0] * (d0 ? x1 : x0) + fxb[0] * (d0 ? x2 : x1);
      hn[1][m] = fxa[1] * (d1 ? x1 : x0) + fxb[1] * (d1 ? x2 : x1);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int a = i - 1 - sy[r];
      const int oy = 2 * a + r;
      if (a >= a0 && a < a0 + RA && oy < p.outH) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          const int ox = 2 * (b0 + lane + 32 * m);
          const float v0 = fya[r] * hp[0][m] + fyb[r] * hn[0][m], v1 = fya[r] * hp[1][m] + fyb[r] * hn[1][m];
          st2<T>(yp + (long long)oy * p.outW + ox, v0, v1, ox < p.outW, ox + 1 < p.outW, vec);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) { hp[c][0] = hn[c][0]; hp[c][1] = hn[c][1]; }
  }
}

// up = 1, down = 2: output column o reads input columns 2 o - px0 + kx.  With px0 = 2 s + e the lanes load (even, odd) input PAIRS
// q = o - s - e + d (d = 0, 1, 2); e = 0: taps (d0.x, d0.y, d1.x, d1.y), e = 1: (d0.y, d1.x, d1.y, d2.x).  Needs an even input width.
template <class T> struct Pair;
template <> struct Pair<float> { typedef float2 type; static __device__ __forceinline__ float2 ld(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); } };
template <> struct Pair<__half> {
  typedef __half2 type;
  static __device__ __forceinline__ float2 ld(const __half* p) { return __half22float2(__ldg(reinterpret_cast<const __half2*>(p))); }
};

template <class T, int E, int RS>
__global__ void __launch_bounds__(128) upfirdn2d_rows_down2_kernel(const UpfirdnArgs p, int strips, int segs) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  long long task = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (task >= (long long)p.N * p.C * strips * segs) return;
  // strips fastest: the warps of a block (and of neighbouring blocks) read adjacent chunks of the same input rows (DRAM page locality)
  const int strip = (int)(task % strips); task /= strips;
  const int seg = (int)(task % segs);
  const long long plane = task / segs;
  const int ox0 = strip * 128, oy0 = seg * RS;
  const T* __restrict__ xp = reinterpret_cast<const T*>(p.x) + plane * (long long)p.inH * p.inW;
  T* __restrict__ yp = reinterpret_cast<T*>(p.y) + plane * (long long)p.outH * p.outW;
  const RowsFilter f = rows_filter(p);
  const int s = floor_div(p.padx0, 2);                    // padx0 = 2 s + E
  const int q0 = ox0 - s - E + lane;                      // pair index of (lane, m = 0)
  const int npairs = p.inW >> 1;
  bool cok[5];
#pragma unroll
  for (int m = 0; m < 5; ++m) cok[m] = q0 + 32 * m >= 0 && q0 + 32 * m < npairs;
  auto load_row = [&](int iy, float2 (&v)[5]) {
    const bool rowok = iy >= 0 && iy < p.inH;
    const T* row = xp + (long long)(rowok ? iy : 0) * p.inW + 2 * q0;
#pragma unroll
    for (int m = 0; m < 5; ++m) v[m] = (rowok && cok[m]) ? Pair<T>::ld(row + 64 * m) : make_float2(0.f, 0.f);
  };
  float hw[3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) hw[i][j] = 0.f;
  float2 nxt[5];
  const int iy0 = 2 * oy0 - p.pady0;
  load_row(iy0, nxt);
  const int nout = p.outH - oy0 < RS ? p.outH - oy0 : RS;
  const int rows = 2 * nout + 2;                          // input rows 0 .. 2 (nout - 1) + 3
#pragma unroll 2
  for (int r = 0; r < rows; ++r) {
    float2 in[5];
#pragma unroll
    for (int m = 0; m < 5; ++m) in[m] = nxt[m];
    if (r + 1 < rows) load_row(iy0 + r + 1, nxt);
    float x1[5], y1[5], x2[5];
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      x1[m] = __shfl_sync(full, in[m].x, (lane + 1) & 31);
      y1[m] = __shfl_sync(full, in[m].y, (lane + 1) & 31);
      if (E) x2[m] = __shfl_sync(full, in[m].x, (lane + 2) & 31);
    }
    const bool w1 = lane + 1 >= 32, w2 = lane + 2 >= 32;
    float h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d1x = w1 ? x1[j + 1] : x1[j], d1y = w1 ? y1[j + 1] : y1[j];
      if (E) {
        const float d2x = w2 ? x2[j + 1] : x2[j];
        h[j] = f.fx[0] * in[j].y + f.fx[1] * d1x + f.fx[2] * d1y + f.fx[3] * d2x;
      } else {
        h[j] = f.fx[0] * in[j].x + f.fx[1] * in[j].y + f.fx[2] * d1x + f.fx[3] * d1y;
      }
    }
    if (r >= 3 && (r & 1)) {
      const int oy = oy0 + ((r - 3) >> 1);
      T* orow = yp + (long long)oy * p.outW + ox0 + lane;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v = f.fy[0] * hw[0][j] + f.fy[1] * hw[1][j] + f.fy[2] * hw[2][j] + f.fy[3] * h[j];
        if (ox0 + lane + 32 * j < p.outW) st_as<T, float>(orow + 32 * j, v);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { hw[0][j] = hw[1][j]; hw[1][j] = hw[2][j]; hw[2][j] = h[j]; }
  }
}

template <class T>
static int launch_rows(const UpfirdnArgs& p, cudaStream_t st) {
  const long long planes = (long long)p.N * p.C;
  if (p.upx == 1 && p.downx == 1) {
    constexpr int RS = 128, NJ = sizeof(T) == 2 ? 8 : 4;
    const int strips = ceil_div(p.outW, 32 * NJ), segs = ceil_div(p.outH, RS);
    const long long blocks = ceil_div_ll(planes * strips * segs, 4);
    if (blocks > 0x7fffffffLL) return SMC_EUNSUPPORTED;
    upfirdn2d_rows_up1_kernel<T, RS, NJ><<<(int)blocks, 128, 0, st>>>(p, strips, segs);
  } else if (p.upx == 2 && p.downx == 1) {
    constexpr int RA = 32;
    const int strips = ceil_div(ceil_div(p.outW, 2), 64), segs = ceil_div(ceil_div(p.outH, 2), RA);
    const long long blocks = ceil_div_ll(planes * strips * segs, 4);
    if (blocks > 0x7fffffffLL) return SMC_EUNSUPPORTED;
    upfirdn2d_rows_up2_kernel<T, RA><<<(int)blocks, 128, 0, st>>>(p, strips, segs);
  } else if (p.upx == 1 && p.downx == 2) {
    constexpr int RS = 32;
    if ((p.inW & 1) || (reinterpret_cast<uintptr_t>(p.x) & 7) || p.padx0 < 0) return SMC_EUNSUPPORTED;
    const int strips = ceil_div(p.outW, 128), segs = ceil_div(p.outH, RS);
    const long long blocks = ceil_div_ll(planes * strips * segs, 4);
    if (blocks > 0x7fffffffLL) return SMC_EUNSUPPORTED;
    if (p.padx0 & 1) upfirdn2d_rows_down2_kernel<T, 1, RS><<<(int)blocks, 128, 0, st>>>(p, strips, segs);
    else upfirdn2d_rows_down2_kernel<T, 0, RS><<<(int)blocks, 128, 0, st>>>(p, strips, segs);
  } else {
    return SMC_EUNSUPPORTED;
  }
  SMC_LAUNCH_CHECK();
  return SMC_OK;
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
      if (g_upfirdn_rows && p.separable) {
        r = launch_rows<T>(p, st);
        if (r != SMC_EUNSUPPORTED) return r;
      }
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
