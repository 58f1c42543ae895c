// CLIP ViT-B/32 glue kernels for sm_100a: everything between the tcgen05 GEMMs (igemm.cu).
//
// Replaces, for the calls `model.encode_image` / `model.encode_text` made at clip_loss.py:15-16,25-26
// (openai/CLIP clip/model.py: fp32-internal LayerNorm, nn.MultiheadAttention, QuickGELU) and for
// find_direction.py:49-52 (`unprocess`), the ATen elementwise / softmax / resize kernels.  The residual
// stream and all statistics stay fp32; GEMM operands are emitted as fp16 hi (+lo) planes.
#include "common.cuh"
#include "stylemc_b200.h"

namespace smc {

__device__ __forceinline__ void store_split(__half* hi, __half* lo, long long i, float v) {
  const __half h = __float2half_rn(v);
  hi[i] = h;
  if (lo) lo[i] = __float2half_rn(v - __half2float(h));
}

// ---------------------------------------------------------------------------------------------------
// Separable antialiased bicubic resample (F.interpolate(mode='bicubic', antialias=True), i.e. what
// torchvision Resize(224, BICUBIC) does on tensors; find_direction.py:258).  Tables come from the host:
// output index o reads inputs [start[o], start[o] + count[o]) with weights wgt[o * taps + k].
// pass 1 (horizontal): t[b,c,y,ox] = sum_k w * clamp(x[b,c,y,start+k] * 127.5 + 128, 0, 255)
// pass 2 (vertical):   y[b,c,oy,ox] = (sum_k w * t[b,c,start+k,ox] / 255 - mean_c) / std_c
__global__ void __launch_bounds__(256) resample_h_kernel(const float* __restrict__ x, float* __restrict__ t, const int* __restrict__ start,
                                                         const int* __restrict__ count, const float* __restrict__ wgt, int taps, long long rows,
                                                         int in_w, int out_w, int denorm, int norm_rows, float scale, float m0, float m1, float m2,
                                                         float s0, float s1, float s2) {
  const long long total = rows * out_w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % out_w);
    const long long r = i / out_w;
    const float* src = x + r * in_w + start[ox];
    const float* w = wgt + (long long)ox * taps;
    const int cnt = count[ox];
    float acc = 0.f;
    for (int k = 0; k < cnt; ++k) {
      float v = __ldg(src + k);
      if (denorm == 1) v = fminf(fmaxf(v * 127.5f + 128.f, 0.f), 255.f);        // find_direction.py:50
      else if (denorm == 2) v = fmaf(v, 0.5f, 0.5f);                            // clip_loss_nada.py:86-87: Normalize(mean -1, std 2), no clamp
      acc += __ldg(w + k) * v;
    }
    if (norm_rows > 0) {                 // second pass of the vertical-first order: rows per plane = norm_rows, channel = plane % 3
      const int c = (int)((r / norm_rows) % 3);
      const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
      acc = (acc * scale - mean) / sd;
    }
    t[i] = acc;
  }
}

// Vertical pass FIRST (the order used from 2x down-sampling on): a thread owns one image column and marches down the input rows
// once; every load is a coalesced row segment, the image is read exactly once, and the intermediate [planes, out_h, in_w] is
// in_h / out_h times smaller than the horizontal-first one.  The outputs whose window covers the current row (at most ring_mask + 1,
// a window [o_lo, o_hi) that only moves forward because the tables are monotone) accumulate in a per-thread column of a shared-memory
// ring; an output is written when its last row has been added.  Block = 256 columns x one plane x one segment of output rows.
__global__ void __launch_bounds__(256) resample_vfirst_kernel(const float* __restrict__ x, float* __restrict__ t, const int* __restrict__ start,
                                                              const int* __restrict__ count, const float* __restrict__ wgt, int taps, int in_h,
                                                              int out_h, int w, int nseg, int ring_mask, int denorm) {
  extern __shared__ float vf_acc[];      // [ring][256]
  const int tx = threadIdx.x, xcol = blockIdx.x * 256 + tx;
  const int pl = blockIdx.y, seg = blockIdx.z;
  const int o0 = (int)((long long)out_h * seg / nseg), o1 = (int)((long long)out_h * (seg + 1) / nseg);
  if (o0 >= o1) return;
  for (int r = 0; r <= ring_mask; ++r) vf_acc[r * 256 + tx] = 0.f;
  const int y0 = __ldg(start + o0), y1 = __ldg(start + o1 - 1) + __ldg(count + o1 - 1);
  const bool ok = xcol < w;
  const float* src = x + (long long)pl * in_h * w + (ok ? xcol : 0);
  float* dst = t + (long long)pl * out_h * w + xcol;
  int o_lo = o0, o_hi = o0;
  for (int yb = y0; yb < y1; yb += 4) {
    float v4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v4[u] = (ok && yb + u < y1) ? __ldg(src + (long long)(yb + u) * w) : 0.f;    // four rows in flight
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int y = yb + u;
      if (y >= y1) break;
      float v = v4[u];
      if (denorm == 1) v = fminf(fmaxf(v * 127.5f + 128.f, 0.f), 255.f);
      else if (denorm == 2) v = fmaf(v, 0.5f, 0.5f);
      while (o_hi < o1 && __ldg(start + o_hi) <= y) ++o_hi;
      for (int o = o_lo; o < o_hi; ++o) {
        const int k = y - __ldg(start + o);
        if (k < __ldg(count + o)) vf_acc[(o & ring_mask) * 256 + tx] += __ldg(wgt + (long long)o * taps + k) * v;
      }
      while (o_lo < o_hi && __ldg(start + o_lo) + __ldg(count + o_lo) - 1 <= y) {
        float* a = &vf_acc[(o_lo & ring_mask) * 256 + tx];
        if (ok) dst[(long long)o_lo * w] = *a;
        *a = 0.f;
        ++o_lo;
      }
    }
  }
}
__global__ void __launch_bounds__(256) resample_v_kernel(const float* __restrict__ t, float* __restrict__ y, const int* __restrict__ start,
                                                         const int* __restrict__ count, const float* __restrict__ wgt, int taps, int planes,
                                                         int in_h, int out_h, int w, float scale, float m0, float m1, float m2, float s0, float s1,
                                                         float s2, int normalize) {
  const long long total = (long long)planes * out_h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % w);
    long long r = i / w;
    const int oy = (int)(r % out_h);
    const int pl = (int)(r / out_h);
    const float* src = t + ((long long)pl * in_h + start[oy]) * w + ox;
    const float* wt = wgt + (long long)oy * taps;
    const int cnt = count[oy];
    float acc = 0.f;
    for (int k = 0; k < cnt; ++k) acc += __ldg(wt + k) * __ldg(src + (long long)k * w);
    if (normalize) {
      const int c = pl % 3;
      const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
      acc = (acc * scale - mean) / sd;
    }
    y[i] = acc;
  }
}
// Transposed passes (backward): per INPUT index the host lists the outputs that read it.
// gt[b,c,iy,ox] = sum_k wT[iy][k] * (g[b,c,oT[iy][k],ox] / (255 * std_c))
__global__ void __launch_bounds__(256) resample_vT_kernel(const float* __restrict__ g, float* __restrict__ gt, const int* __restrict__ oidx,
                                                          const int* __restrict__ count, const float* __restrict__ wgt, int taps, int planes,
                                                          int in_h, int out_h, int w, float s0, float s1, float s2, const float* __restrict__ x,
                                                          const float* __restrict__ unscale, int mode = 1) {
  // mode 1: unprocess (slope 127.5, clamp mask, 1 / (255 std)); mode 2: the NADA preprocessing (slope 0.5, no clamp, 1 / std)
  const long long total = (long long)planes * in_h * w;
  const float k127 = x ? (mode == 2 ? 0.5f : 127.5f) / (unscale ? __ldg(unscale) : 1.f) : 1.f;
  const float k255 = mode == 2 ? 1.f : 255.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % w);
    long long r = i / w;
    const int iy = (int)(r % in_h);
    const int pl = (int)(r / in_h);
    const int c = pl % 3;
    const float k0 = 1.f / (k255 * (c == 0 ? s0 : (c == 1 ? s1 : s2)));
    const float* src = g + (long long)pl * out_h * w + ox;
    float acc = 0.f;
    const int cnt = count[iy];
    for (int k = 0; k < cnt; ++k) acc += __ldg(wgt + (long long)iy * taps + k) * __ldg(src + (long long)oidx[(long long)iy * taps + k] * w);
    acc *= k0;
    if (x) {                             // last pass of the vertical-first order: the clamp mask and 127.5 / loss scale live here
      const float v = mode == 2 ? 1.f : x[i] * 127.5f + 128.f;
      acc = (v > 0.f && v < 255.f) ? acc * k127 : 0.f;
    }
    gt[i] = acc;
  }
}
// gx[b,c,y,ix] = 127.5 * [0 < x*127.5+128 < 255] * sum_k wT[ix][k] * gt[b,c,y,oT[ix][k]]
__global__ void __launch_bounds__(256) resample_hT_kernel(const float* __restrict__ gt, const float* __restrict__ x, float* __restrict__ gx,
                                                          const int* __restrict__ oidx, const int* __restrict__ count, const float* __restrict__ wgt,
                                                          int taps, long long rows, int in_w, int out_w, const float* __restrict__ unscale,
                                                          int mode = 1) {
  const long long total = rows * in_w;
  const float k127 = x ? (mode == 2 ? 0.5f : 127.5f) / (unscale ? __ldg(unscale) : 1.f) : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(i % in_w);
    const long long r = i / in_w;
    const float v = (x && mode != 2) ? x[i] * 127.5f + 128.f : 1.f;
    float acc = 0.f;
    if (v > 0.f && v < 255.f) {   // torch clamp backward passes gradient on [min, max] inclusive; measure-zero difference
      const float* src = gt + r * out_w;
      const int cnt = count[ix];
      for (int k = 0; k < cnt; ++k) acc += __ldg(wgt + (long long)ix * taps + k) * __ldg(src + oidx[(long long)ix * taps + k]);
      acc *= k127;
    }
    gx[i] = acc;
  }
}

// Row-contraction pass with the lanes of a warp on 32 different ROWS (round 2; replaces resample_h_kernel in the forward pass and resample_hT_kernel
// in the backward pass, which spent ~2 global loads per tap and output: 13-18 % of the HBM peak, LSU-bound).  One CTA owns 32 rows x OBT
// outputs: the input columns those outputs read ([start[ob0], start[last] + count[last]): at most max_span) are loaded once, coalesced along
// the row, into a shared tile with an odd pitch (pre-processing applied on the way in); warp w then computes outputs ob0 + w, w + 8, ...
// with lane = row: the tile reads are conflict-free, the tap weights are warp-uniform loads; the results go through a second shared tile
// so that the global stores are runs of OBT consecutive floats.  Accumulation order per output = the per-thread loop of the old kernels.
//   out[r, ob] = post( sum_k wgt[ob * taps + k] * pre(in[r, start[ob * sstride] + k]) )
// pre: 0 none, 1 clamp(v * 127.5 + 128, 0, 255) (find_direction.py:50), 2 v * 0.5 + 0.5 (clip_loss_nada.py:86-87)
// post: 0 none, 1 * k127 where 0 < xmask * 127.5 + 128 < 255, else 0 (clamp backward), 2 * k127; k127 = (post == 2 ? 0.5 : 127.5) / unscale
template <int OBT, int MAXT>   // MAXT >= taps: the tap loop is fully unrolled (predicated on the window length): 3 instructions per tap, no loop overhead
__global__ void __launch_bounds__(256) resample_rows_kernel(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ start,
                                                            int sstride, const int* __restrict__ count, const float* __restrict__ wgt, int taps,
                                                            long long rows, int in_w, int out_w, int max_span, int pre, int post,
                                                            const float* __restrict__ xmask, const float* __restrict__ unscale) {
  extern __shared__ float rr_smem[];
  const int pitch = max_span | 1;
  float* tile = rr_smem;                      // [32][pitch]
  float* stage = rr_smem + 32 * pitch;        // [32][OBT + 1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // output tiles of one row tile are consecutive blocks: together they read whole rows (DRAM page locality; with the row tiles fastest
  // every block fetched 32 isolated 0.6-KB segments and the forward pass ran at 1.3 TB/s)
  const int n_obt = (out_w + OBT - 1) / OBT;
  const long long r0 = (long long)(blockIdx.x / n_obt) * 32;
  const int ob0 = (int)(blockIdx.x % n_obt) * OBT;
  const int obl = (ob0 + OBT < out_w ? ob0 + OBT : out_w) - 1;
  const int c0 = __ldg(start + (long long)ob0 * sstride);
  const int span = __ldg(start + (long long)obl * sstride) + __ldg(count + obl) - c0;
  if (span > max_span || c0 + span > in_w) __trap();     // tables that are not monotone windows: fail loudly
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int r = warp + 8 * rr;
    const long long row = r0 + r;
    const float* src = in + row * in_w + c0;
    for (int c = lane; c < span; c += 32) {
      float v = 0.f;
      if (row < rows) {
        v = __ldg(src + c);
        if (pre == 1) v = fminf(fmaxf(v * 127.5f + 128.f, 0.f), 255.f);
        else if (pre == 2) v = fmaf(v, 0.5f, 0.5f);
      }
      tile[r * pitch + c] = v;
    }
  }
  __syncthreads();
  for (int j = warp; j < OBT; j += 8) {
    const int ob = ob0 + j;
    float acc = 0.f;
    if (ob < out_w) {
      const int cnt = __ldg(count + ob);
      const float* w = wgt + (long long)ob * taps;
      const float* t = tile + lane * pitch + (__ldg(start + (long long)ob * sstride) - c0);
#pragma unroll
      for (int k = 0; k < MAXT; ++k)
        if (k < cnt) acc += __ldg(w + k) * t[k];            // cnt is warp-uniform; beyond it the tile holds nothing defined
    }
    stage[lane * (OBT + 1) + j] = acc;
  }
  __syncthreads();
  const float k127 = post ? (post == 2 ? 0.5f : 127.5f) / (unscale ? __ldg(unscale) : 1.f) : 1.f;
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int r = warp + 8 * rr;
    const long long row = r0 + r;
    if (row >= rows) continue;
    for (int j = lane; j < OBT; j += 32) {
      const int ob = ob0 + j;
      if (ob >= out_w) break;
      float v = stage[r * (OBT + 1) + j];
      const long long oi = row * out_w + ob;
      if (post == 1) {
        const float xv = __ldg(xmask + oi) * 127.5f + 128.f;
        v = (xv > 0.f && xv < 255.f) ? v * k127 : 0.f;
      } else if (post == 2) {
        v *= k127;
      }
      out[oi] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// [B,3,224,224] fp32 -> patch matrix [B*49, 3072] fp16 (row = b*49 + py*7 + px, col = c*1024 + ky*32 + kx):
// the im2col of clip/model.py `conv1` (kernel = stride = 32), and its transpose for the backward.
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ img, __half* __restrict__ hi, __half* __restrict__ lo, int B,
                                                       int res, int ps) {
  const int grid = res / ps, kk = 3 * ps * ps;
  const long long total = (long long)B * grid * grid * kk;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % kk);
    const long long row = i / kk;
    const int kx = col % ps, ky = (col / ps) % ps, c = col / (ps * ps);
    const int px = (int)(row % grid), py = (int)((row / grid) % grid);
    const long long b = row / (grid * grid);
    store_split(hi, lo, i, img[((b * 3 + c) * res + py * ps + ky) * res + px * ps + kx]);
  }
}
__global__ void __launch_bounds__(256) unpatchify_kernel(const float* __restrict__ gp, float* __restrict__ gimg, int B, int res, int ps) {
  const int grid = res / ps, kk = 3 * ps * ps;
  const long long total = (long long)B * 3 * res * res;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % res), y = (int)((i / res) % res), c = (int)((i / ((long long)res * res)) % 3);
    const long long b = i / (3LL * res * res);
    const long long row = (b * grid + y / ps) * grid + x / ps;
    gimg[i] = gp[row * kk + (c * ps + y % ps) * ps + x % ps];
  }
}

// ---------------------------------------------------------------------------------------------------
// Token assembly: x0[b,0,:] = cls + pos[0];  x0[b,1+p,:] = patch[b*49+p,:] + pos[1+p]   (VisionTransformer.forward)
// Text:           x0[b,t,:] = tok_emb[text[b,t]] + pos[t]                                (CLIP.encode_text)
__global__ void __launch_bounds__(256) assemble_tokens_kernel(const float* __restrict__ patch, const float* __restrict__ cls,
                                                              const float* __restrict__ pos, float* __restrict__ x0, int B, int T, int Wd) {
  const long long total = (long long)B * T * Wd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % Wd), t = (int)((i / Wd) % T);
    const long long b = i / ((long long)Wd * T);
    const float v = (t == 0) ? cls[d] : patch[(b * (T - 1) + (t - 1)) * Wd + d];
    x0[i] = v + pos[(long long)t * Wd + d];
  }
}
__global__ void __launch_bounds__(256) embed_text_kernel(const long long* __restrict__ text, const float* __restrict__ emb,
                                                         const float* __restrict__ pos, float* __restrict__ x0, int B, int T, int Wd) {
  const long long total = (long long)B * T * Wd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % Wd), t = (int)((i / Wd) % T);
    const long long b = i / ((long long)Wd * T);
    x0[i] = emb[text[b * T + t] * Wd + d] + pos[(long long)t * Wd + d];
  }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (eps 1e-5), one warp per row.  Row r of the output reads input row
// r * in_row_stride + in_row_offset (lets ln_post pick the class token).  Outputs: fp32 and/or fp16 hi/lo.
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, long long in_row_stride, long long in_row_offset,
                                                            const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ y32,
                                                            __half* __restrict__ yhi, __half* __restrict__ ylo, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out, long long rows, int Wd) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float* xr = x + (r * in_row_stride + in_row_offset) * Wd;
    float s = 0.f;
    for (int d = lane; d < Wd; d += 32) s += xr[d];
    const float mean = warp_sum(s) / Wd;
    float v = 0.f;
    for (int d = lane; d < Wd; d += 32) { const float t = xr[d] - mean; v += t * t; }
    const float rstd = rsqrtf(warp_sum(v) / Wd + 1e-5f);
    if (lane == 0 && mean_out) { mean_out[r] = mean; rstd_out[r] = rstd; }
    for (int d = lane; d < Wd; d += 32) {
      const float o = (xr[d] - mean) * rstd * w[d] + b[d];
      if (y32) y32[r * Wd + d] = o;
      if (yhi) store_split(yhi, ylo, r * Wd + d, o);
    }
  }
}
// dx = rstd * (w*dy - mean(w*dy) - xhat * mean(w*dy*xhat)); out row = out_row_stride*r + out_row_offset;
// accumulate != 0 adds into dx (residual gradient).
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, long long in_row_stride,
                                                            long long in_row_offset, const float* __restrict__ w, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, float* __restrict__ dx, long long rows, int Wd,
                                                            int accumulate) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const long long xr = (r * in_row_stride + in_row_offset) * Wd;
    const float mu = mean[r], rs = rstd[r];
    float a = 0.f, bsum = 0.f;
    for (int d = lane; d < Wd; d += 32) {
      const float g = w[d] * dy[r * Wd + d];
      const float xh = (x[xr + d] - mu) * rs;
      a += g; bsum += g * xh;
    }
    a = warp_sum(a) / Wd; bsum = warp_sum(bsum) / Wd;
    for (int d = lane; d < Wd; d += 32) {
      const float g = w[d] * dy[r * Wd + d];
      const float xh = (x[xr + d] - mu) * rs;
      const float o = rs * (g - a - xh * bsum);
      if (accumulate) dx[xr + d] += o; else dx[xr + d] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Multi-head attention core, one CTA per (sequence, head): S = (q*scale) k^T (+causal mask), P = softmax(S),
// O = P v.  qkv fp32 [B, T, 3*Wd] (q | k | v blocks, head h at columns h*hd), output [B, T, Wd] fp16 hi/lo.
__global__ void __launch_bounds__(256) attention_fwd_kernel(const float* __restrict__ qkv, __half* __restrict__ ohi, __half* __restrict__ olo,
                                                            float* __restrict__ o32, int T, int Wd, int heads, int causal) {
  extern __shared__ float sm[];
  const int hd = Wd / heads;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  float* q = sm;                  // [T][hd+1]
  float* k = q + T * (hd + 1);
  float* v = k + T * (hd + 1);
  float* S = v + T * (hd + 1);    // [T][T+1]
  const float scale = rsqrtf((float)hd);
  for (int i = threadIdx.x; i < T * hd; i += blockDim.x) {
    const int t = i / hd, d = i % hd;
    const float* row = qkv + ((long long)b * T + t) * 3 * Wd + h * hd + d;
    q[t * (hd + 1) + d] = row[0] * scale;
    k[t * (hd + 1) + d] = row[Wd];
    v[t * (hd + 1) + d] = row[2 * Wd];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    const int r = i / T, c = i % T;
    float acc = 0.f;
    for (int d = 0; d < hd; ++d) acc += q[r * (hd + 1) + d] * k[c * (hd + 1) + d];
    S[r * (T + 1) + c] = (causal && c > r) ? -INFINITY : acc;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < T; r += blockDim.x >> 5) {
    float m = -INFINITY;
    for (int c = lane; c < T; c += 32) m = fmaxf(m, S[r * (T + 1) + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < T; c += 32) { const float e = expf(S[r * (T + 1) + c] - m); S[r * (T + 1) + c] = e; s += e; }
    s = warp_sum(s);
    const float inv = 1.f / s;
    for (int c = lane; c < T; c += 32) S[r * (T + 1) + c] *= inv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * hd; i += blockDim.x) {
    const int t = i / hd, d = i % hd;
    float acc = 0.f;
    for (int c = 0; c < T; ++c) acc += S[t * (T + 1) + c] * v[c * (hd + 1) + d];
    const long long o = ((long long)b * T + t) * Wd + h * hd + d;
    if (ohi) store_split(ohi, olo, o, acc);
    if (o32) o32[o] = acc;
  }
}
// Backward of the core given dO fp32 [B,T,Wd]: recompute P; dV = P^T dO; dP = dO V^T;
// dS = P * (dP - rowsum(dP * P)); dQ = scale * dS K; dK = scale * dS^T Q.  Output dqkv fp16 hi/lo [B,T,3*Wd].
__global__ void __launch_bounds__(256) attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dO, __half* __restrict__ ghi,
                                                            __half* __restrict__ glo, int T, int Wd, int heads, int causal) {
  extern __shared__ float sm[];
  const int hd = Wd / heads;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int ld = hd + 1;
  float* q = sm;
  float* k = q + T * ld;
  float* v = k + T * ld;
  float* go = v + T * ld;
  float* P = go + T * ld;         // [T][T+1]
  float* dS = P + T * (T + 1);    // [T][T+1]
  const float scale = rsqrtf((float)hd);
  for (int i = threadIdx.x; i < T * hd; i += blockDim.x) {
    const int t = i / hd, d = i % hd;
    const float* row = qkv + ((long long)b * T + t) * 3 * Wd + h * hd + d;
    q[t * ld + d] = row[0];
    k[t * ld + d] = row[Wd];
    v[t * ld + d] = row[2 * Wd];
    go[t * ld + d] = dO[((long long)b * T + t) * Wd + h * hd + d];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    const int r = i / T, c = i % T;
    float s = 0.f, dp = 0.f;
    for (int d = 0; d < hd; ++d) { s += q[r * ld + d] * k[c * ld + d]; dp += go[r * ld + d] * v[c * ld + d]; }
    P[r * (T + 1) + c] = (causal && c > r) ? -INFINITY : s * scale;
    dS[r * (T + 1) + c] = dp;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < T; r += blockDim.x >> 5) {
    float m = -INFINITY;
    for (int c = lane; c < T; c += 32) m = fmaxf(m, P[r * (T + 1) + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < T; c += 32) { const float e = expf(P[r * (T + 1) + c] - m); P[r * (T + 1) + c] = e; s += e; }
    const float inv = 1.f / warp_sum(s);
    float dot = 0.f;
    for (int c = lane; c < T; c += 32) { const float p = P[r * (T + 1) + c] * inv; P[r * (T + 1) + c] = p; dot += p * dS[r * (T + 1) + c]; }
    dot = warp_sum(dot);
    for (int c = lane; c < T; c += 32) dS[r * (T + 1) + c] = P[r * (T + 1) + c] * (dS[r * (T + 1) + c] - dot) * scale;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * hd; i += blockDim.x) {
    const int t = i / hd, d = i % hd;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int c = 0; c < T; ++c) {
      dq += dS[t * (T + 1) + c] * k[c * ld + d];
      dk += dS[c * (T + 1) + t] * q[c * ld + d];
      dv += P[c * (T + 1) + t] * go[c * ld + d];
    }
    const long long o = ((long long)b * T + t) * 3 * Wd + h * hd + d;
    store_split(ghi, glo, o, dq);
    store_split(ghi, glo, o + Wd, dk);
    store_split(ghi, glo, o + 2 * Wd, dv);
  }
}

// ---------------------------------------------------------------------------------------------------
// Register-tiled versions of the two attention kernels for head_dim 64 (ncu/launch list: the scalar versions spend their time on
// shared-memory loads, 2 LDS per FMA: 99 us forward / 217 us backward per ViT layer at 64 images).  A thread owns a 4 x 4 tile of
// each small matrix product with INTERLEAVED indices (row tr + TR * i, column tc + TR * j, feature td + 16 * j), so that the lanes of
// a warp read consecutive shared-memory words: 8 LDS per 16 FMA.  Every output is still accumulated in ascending k order with one FMA
// per term, so the results are bit-identical to the scalar kernels.
__global__ void __launch_bounds__(256) attention_fwd2_kernel(const float* __restrict__ qkv, __half* __restrict__ ohi, __half* __restrict__ olo,
                                                             float* __restrict__ o32, int T, int Wd, int heads, int causal) {
  extern __shared__ float sm[];
  constexpr int hd = 64, ld = hd + 1;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  float* q = sm;                  // [T][ld]
  float* k = q + T * ld;
  float* v = k + T * ld;
  float* S = v + T * ld;          // [T][T+1]
  const float scale = rsqrtf((float)hd);
  for (int i = threadIdx.x; i < T * hd; i += blockDim.x) {
    const int t = i / hd, d = i % hd;
    const float* row = qkv + ((long long)b * T + t) * 3 * Wd + h * hd + d;
    q[t * ld + d] = row[0] * scale;
    k[t * ld + d] = row[Wd];
    v[t * ld + d] = row[2 * Wd];
  }
  __syncthreads();
  const int TR = (T + 3) >> 2;
  for (int tile = threadIdx.x; tile < TR * TR; tile += blockDim.x) {
    const int tr = tile / TR, tc = tile - tr * TR;
    const float *qp[4], *kp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = tr + TR * i, c = tc + TR * i;
      qp[i] = q + (r < T ? r : T - 1) * ld;
      kp[i] = k + (c < T ? c : T - 1) * ld;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < hd; ++d) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = qp[i][d]; bb[i] = kp[i][d]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * bb[j];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = tr + TR * i;
      if (r >= T) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = tc + TR * j;
        if (c < T) S[r * (T + 1) + c] = (causal && c > r) ? -INFINITY : acc[i][j];
      }
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < T; r += blockDim.x >> 5) {
    float m = -INFINITY;
    for (int c = lane; c < T; c += 32) m = fmaxf(m, S[r * (T + 1) + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < T; c += 32) { const float e = expf(S[r * (T + 1) + c] - m); S[r * (T + 1) + c] = e; s += e; }
    s = warp_sum(s);
    const float inv = 1.f / s;
    for (int c = lane; c < T; c += 32) S[r * (T + 1) + c] *= inv;
  }
  __syncthreads();
  for (int tile = threadIdx.x; tile < TR * 16; tile += blockDim.x) {
    const int tt = tile >> 4, td = tile & 15;
    const float* sp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = tt + TR * i;
      sp[i] = S + (t < T ? t : T - 1) * (T + 1);
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int c = 0; c < T; ++c) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sp[i][c]; bb[i] = v[c * ld + td + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * bb[j];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = tt + TR * i;
      if (t >= T) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long o = ((long long)b * T + t) * Wd + h * hd + td + 16 * j;
        if (ohi) store_split(ohi, olo, o, acc[i][j]);
        if (o32) o32[o] = acc[i][j];
      }
    }
  }
}

__global__ void __launch_bounds__(256) attention_bwd2_kernel(const float* __restrict__ qkv, const float* __restrict__ dO, __half* __restrict__ ghi,
                                                             __half* __restrict__ glo, int T, int Wd, int heads, int causal) {
  extern __shared__ float sm[];
  constexpr int hd = 64, ld = hd + 1;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  float* q = sm;
  float* k = q + T * ld;
  float* v = k + T * ld;
  float* go = v + T * ld;
  float* P = go + T * ld;         // [T][T+1]
  float* dS = P + T * (T + 1);    // [T][T+1]
  const float scale = rsqrtf((float)hd);
  for (int i = threadIdx.x; i < T * hd; i += blockDim.x) {
    const int t = i / hd, d = i % hd;
    const float* row = qkv + ((long long)b * T + t) * 3 * Wd + h * hd + d;
    q[t * ld + d] = row[0];
    k[t * ld + d] = row[Wd];
    v[t * ld + d] = row[2 * Wd];
    go[t * ld + d] = dO[((long long)b * T + t) * Wd + h * hd + d];
  }
  __syncthreads();
  const int TR = (T + 3) >> 2;
  for (int tile = threadIdx.x; tile < TR * TR; tile += blockDim.x) {
    const int tr = tile / TR, tc = tile - tr * TR;
    int ro[4], co[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = tr + TR * i, c = tc + TR * i;
      ro[i] = (r < T ? r : T - 1) * ld;
      co[i] = (c < T ? c : T - 1) * ld;
    }
    float as[4][4], ap[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { as[i][j] = 0.f; ap[i][j] = 0.f; }
#pragma unroll 2
    for (int d = 0; d < hd; ++d) {
      float a[4], bb[4], g[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = q[ro[i] + d]; bb[i] = k[co[i] + d]; g[i] = go[ro[i] + d]; w[i] = v[co[i] + d]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { as[i][j] += a[i] * bb[j]; ap[i][j] += g[i] * w[j]; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = tr + TR * i;
      if (r >= T) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = tc + TR * j;
        if (c < T) {
          P[r * (T + 1) + c] = (causal && c > r) ? -INFINITY : as[i][j] * scale;
          dS[r * (T + 1) + c] = ap[i][j];
        }
      }
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < T; r += blockDim.x >> 5) {
    float m = -INFINITY;
    for (int c = lane; c < T; c += 32) m = fmaxf(m, P[r * (T + 1) + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < T; c += 32) { const float e = expf(P[r * (T + 1) + c] - m); P[r * (T + 1) + c] = e; s += e; }
    const float inv = 1.f / warp_sum(s);
    float dot = 0.f;
    for (int c = lane; c < T; c += 32) { const float pp = P[r * (T + 1) + c] * inv; P[r * (T + 1) + c] = pp; dot += pp * dS[r * (T + 1) + c]; }
    dot = warp_sum(dot);
    for (int c = lane; c < T; c += 32) dS[r * (T + 1) + c] = P[r * (T + 1) + c] * (dS[r * (T + 1) + c] - dot) * scale;
  }
  __syncthreads();
  for (int tile = threadIdx.x; tile < TR * 16; tile += blockDim.x) {
    const int tt = tile >> 4, td = tile & 15;
    int tv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int t = tt + TR * i; tv[i] = t < T ? t : T - 1; }
    float aq[4][4], ak[4][4], av[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { aq[i][j] = 0.f; ak[i][j] = 0.f; av[i][j] = 0.f; }
    for (int c = 0; c < T; ++c) {
      float s1[4], s2[4], p2[4], kk[4], qq[4], gg[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s1[i] = dS[tv[i] * (T + 1) + c];
        s2[i] = dS[c * (T + 1) + tv[i]];
        p2[i] = P[c * (T + 1) + tv[i]];
        kk[i] = k[c * ld + td + 16 * i];
        qq[i] = q[c * ld + td + 16 * i];
        gg[i] = go[c * ld + td + 16 * i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { aq[i][j] += s1[i] * kk[j]; ak[i][j] += s2[i] * qq[j]; av[i][j] += p2[i] * gg[j]; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = tt + TR * i;
      if (t >= T) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long o = ((long long)b * T + t) * 3 * Wd + h * hd + td + 16 * j;
        store_split(ghi, glo, o, aq[i][j]);
        store_split(ghi, glo, o + Wd, ak[i][j]);
        store_split(ghi, glo, o + 2 * Wd, av[i][j]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Long sequences (ViT-B/16, the second tower of clip_type='double', clip_loss.py:12-13: 14 x 14 + 1 = 197 tokens).  The whole-sequence
// kernels above keep T x (T+1) score words per CTA (310 KB forward, 517 KB backward at T = 197), more than an SM has.  Here a CTA owns a
// block of QB query rows (forward and dQ) or KB key rows (dK, dV) of one (sequence, head); the other operand stays whole in shared memory
// (2 x 51 KB at T = 197) and the score block is [QB][T+1].  Same 4 x 4 interleaved register tiles as above, head_dim 64 only.
//   forward      : S = (q*scale) k^T, P = softmax(S), O = P v                                   for the CTA's query rows
//   backward (q) : recompute P rows; dP = dO v^T; dot = rowsum(P*dP); dS = P*(dP-dot)*scale; dQ = dS k; also stores per row
//                  lse = max + log(sum) and dot in ``stats`` [2][B*heads][T] for the second kernel
//   backward (kv): P^T, dS^T for the CTA's key rows from lse/dot (P = exp(s*scale - lse)); dK = dS^T q, dV = P^T dO
__global__ void __launch_bounds__(256) attention_fwd_rows_kernel(const float* __restrict__ qkv, __half* __restrict__ ohi, __half* __restrict__ olo,
                                                                 float* __restrict__ o32, int T, int Wd, int heads, int causal, int QB, int nqb) {
  extern __shared__ float sm[];
  constexpr int hd = 64, ld = hd + 1;
  const int qb = blockIdx.x % nqb, bh = blockIdx.x / nqb;
  const int b = bh / heads, h = bh % heads;
  const int r0 = qb * QB, R = min(QB, T - r0);
  float* k = sm;                  // [T][ld]
  float* v = k + T * ld;          // [T][ld]
  float* q = v + T * ld;          // [QB][ld]
  float* S = q + QB * ld;         // [QB][T+1]
  const float scale = rsqrtf((float)hd);
  for (int i = threadIdx.x; i < T * hd; i += blockDim.x) {
    const int t = i / hd, d = i % hd;
    const float* row = qkv + ((long long)b * T + t) * 3 * Wd + h * hd + d;
    k[t * ld + d] = row[Wd];
    v[t * ld + d] = row[2 * Wd];
  }
  for (int i = threadIdx.x; i < R * hd; i += blockDim.x) {
    const int t = i / hd, d = i % hd;
    q[t * ld + d] = qkv[((long long)b * T + r0 + t) * 3 * Wd + h * hd + d] * scale;
  }
  __syncthreads();
  const int TRr = (R + 3) >> 2, TRc = (T + 3) >> 2;
  for (int tile = threadIdx.x; tile < TRr * TRc; tile += blockDim.x) {
    const int tr = tile / TRc, tc = tile - tr * TRc;
    const float *qp[4], *kp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = tr + TRr * i, c = tc + TRc * i;
      qp[i] = q + (r < R ? r : R - 1) * ld;
      kp[i] = k + (c < T ? c : T - 1) * ld;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int d = 0; d < hd; ++d) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = qp[i][d]; bb[i] = kp[i][d]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * bb[j];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = tr + TRr * i;
      if (r >= R) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = tc + TRc * j;
        if (c < T) S[r * (T + 1) + c] = (causal && c > r0 + r) ? -INFINITY : acc[i][j];
      }
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < R; r += blockDim.x >> 5) {
    float m = -INFINITY;
    for (int c = lane; c < T; c += 32) m = fmaxf(m, S[r * (T + 1) + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < T; c += 32) { const float e = expf(S[r * (T + 1) + c] - m); S[r * (T + 1) + c] = e; s += e; }
    s = warp_sum(s);
    const float inv = 1.f / s;
    for (int c = lane; c < T; c += 32) S[r * (T + 1) + c] *= inv;
  }
  __syncthreads();
  for (int tile = threadIdx.x; tile < TRr * 16; tile += blockDim.x) {
    const int tt = tile >> 4, td = tile & 15;
    const float* sp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = tt + TRr * i;
      sp[i] = S + (t < R ? t : R - 1) * (T + 1);
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int c = 0; c < T; ++c) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sp[i][c]; bb[i] = v[c * ld + td + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * bb[j];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = tt + TRr * i;
      if (t >= R) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long o = ((long long)b * T + r0 + t) * Wd + h * hd + td + 16 * j;
        if (ohi) store_split(ohi, olo, o, acc[i][j]);
        if (o32) o32[o] = acc[i][j];
      }
    }
  }
}

__global__ void __launch_bounds__(256) attention_bwd_q_kernel(const float* __restrict__ qkv, const float* __restrict__ dO, __half* __restrict__ ghi,
                                                              __half* __restrict__ glo, float* __restrict__ stats, int T, int Wd, int heads, int causal,
                                                              int QB, int nqb) {
  extern __shared__ float sm[];
  constexpr int hd = 64, ld = hd + 1;
  const int qb = blockIdx.x % nqb, bh = blockIdx.x / nqb;
  const int b = bh / heads, h = bh % heads;
  const int r0 = qb * QB, R = min(QB, T - r0);
  float* k = sm;                  // [T][ld]
  float* v = k + T * ld;          // [T][ld]
  float* q = v + T * ld;          // [QB][ld]
  float* go = q + QB * ld;        // [QB][ld]
  float* P = go + QB * ld;        // [QB][T+1]
  float* dS = P + QB * (T + 1);   // [QB][T+1]
  const float scale = rsqrtf((float)hd);
  for (int i = threadIdx.x; i < T * hd; i += blockDim.x) {
    const int t = i / hd, d = i % hd;
    const float* row = qkv + ((long long)b * T + t) * 3 * Wd + h * hd + d;
    k[t * ld + d] = row[Wd];
    v[t * ld + d] = row[2 * Wd];
  }
  for (int i = threadIdx.x; i < R * hd; i += blockDim.x) {
    const int t = i / hd, d = i % hd;
    q[t * ld + d] = qkv[((long long)b * T + r0 + t) * 3 * Wd + h * hd + d];
    go[t * ld + d] = dO[((long long)b * T + r0 + t) * Wd + h * hd + d];
  }
  __syncthreads();
  const int TRr = (R + 3) >> 2, TRc = (T + 3) >> 2;
  for (int tile = threadIdx.x; tile < TRr * TRc; tile += blockDim.x) {
    const int tr = tile / TRc, tc = tile - tr * TRc;
    int ro[4], co[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = tr + TRr * i, c = tc + TRc * i;
      ro[i] = (r < R ? r : R - 1) * ld;
      co[i] = (c < T ? c : T - 1) * ld;
    }
    float as[4][4], ap[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { as[i][j] = 0.f; ap[i][j] = 0.f; }
#pragma unroll 2
    for (int d = 0; d < hd; ++d) {
      float a[4], bb[4], g[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = q[ro[i] + d]; bb[i] = k[co[i] + d]; g[i] = go[ro[i] + d]; w[i] = v[co[i] + d]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { as[i][j] += a[i] * bb[j]; ap[i][j] += g[i] * w[j]; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = tr + TRr * i;
      if (r >= R) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = tc + TRc * j;
        if (c < T) {
          P[r * (T + 1) + c] = (causal && c > r0 + r) ? -INFINITY : as[i][j] * scale;
          dS[r * (T + 1) + c] = ap[i][j];
        }
      }
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_rows = (long long)(gridDim.x / nqb) * T;      // B * heads * T
  for (int r = warp; r < R; r += blockDim.x >> 5) {
    float m = -INFINITY;
    for (int c = lane; c < T; c += 32) m = fmaxf(m, P[r * (T + 1) + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < T; c += 32) { const float e = expf(P[r * (T + 1) + c] - m); P[r * (T + 1) + c] = e; s += e; }
    s = warp_sum(s);
    const float inv = 1.f / s;
    float dot = 0.f;
    for (int c = lane; c < T; c += 32) { const float pp = P[r * (T + 1) + c] * inv; P[r * (T + 1) + c] = pp; dot += pp * dS[r * (T + 1) + c]; }
    dot = warp_sum(dot);
    for (int c = lane; c < T; c += 32) dS[r * (T + 1) + c] = P[r * (T + 1) + c] * (dS[r * (T + 1) + c] - dot) * scale;
    if (lane == 0) {
      stats[(long long)bh * T + r0 + r] = m + logf(s);
      stats[n_rows + (long long)bh * T + r0 + r] = dot;
    }
  }
  __syncthreads();
  for (int tile = threadIdx.x; tile < TRr * 16; tile += blockDim.x) {
    const int tt = tile >> 4, td = tile & 15;
    int tv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int t = tt + TRr * i; tv[i] = (t < R ? t : R - 1) * (T + 1); }
    float aq[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) aq[i][j] = 0.f;
#pragma unroll 2
    for (int c = 0; c < T; ++c) {
      float s1[4], kk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { s1[i] = dS[tv[i] + c]; kk[i] = k[c * ld + td + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) aq[i][j] += s1[i] * kk[j];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = tt + TRr * i;
      if (t >= R) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        store_split(ghi, glo, ((long long)b * T + r0 + t) * 3 * Wd + h * hd + td + 16 * j, aq[i][j]);
    }
  }
}

__global__ void __launch_bounds__(256) attention_bwd_kv_kernel(const float* __restrict__ qkv, const float* __restrict__ dO, __half* __restrict__ ghi,
                                                               __half* __restrict__ glo, const float* __restrict__ stats, int T, int Wd, int heads,
                                                               int causal, int KB, int nkb) {
  extern __shared__ float sm[];
  constexpr int hd = 64, ld = hd + 1;
  const int kb = blockIdx.x % nkb, bh = blockIdx.x / nkb;
  const int b = bh / heads, h = bh % heads;
  const int c0 = kb * KB, C = min(KB, T - c0);
  float* q = sm;                  // [T][ld]
  float* go = q + T * ld;         // [T][ld]
  float* k = go + T * ld;         // [KB][ld]
  float* v = k + KB * ld;         // [KB][ld]
  float* PT = v + KB * ld;        // [KB][T+1]   P transposed: PT[c][r]
  float* dST = PT + KB * (T + 1); // [KB][T+1]
  float* lse = dST + KB * (T + 1);// [T]
  float* dt = lse + T;            // [T]
  const float scale = rsqrtf((float)hd);
  const long long n_rows = (long long)(gridDim.x / nkb) * T;
  for (int i = threadIdx.x; i < T * hd; i += blockDim.x) {
    const int t = i / hd, d = i % hd;
    q[t * ld + d] = qkv[((long long)b * T + t) * 3 * Wd + h * hd + d];
    go[t * ld + d] = dO[((long long)b * T + t) * Wd + h * hd + d];
  }
  for (int i = threadIdx.x; i < C * hd; i += blockDim.x) {
    const int t = i / hd, d = i % hd;
    const float* row = qkv + ((long long)b * T + c0 + t) * 3 * Wd + h * hd + d;
    k[t * ld + d] = row[Wd];
    v[t * ld + d] = row[2 * Wd];
  }
  for (int r = threadIdx.x; r < T; r += blockDim.x) {
    lse[r] = stats[(long long)bh * T + r];
    dt[r] = stats[n_rows + (long long)bh * T + r];
  }
  __syncthreads();
  const int TRr = (C + 3) >> 2, TRc = (T + 3) >> 2;
  for (int tile = threadIdx.x; tile < TRr * TRc; tile += blockDim.x) {
    const int tr = tile / TRc, tc = tile - tr * TRc;
    int co[4], ro[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = tr + TRr * i, r = tc + TRc * i;
      co[i] = (c < C ? c : C - 1) * ld;
      ro[i] = (r < T ? r : T - 1) * ld;
    }
    float as[4][4], ap[4][4];     // [key i][query j]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { as[i][j] = 0.f; ap[i][j] = 0.f; }
#pragma unroll 2
    for (int d = 0; d < hd; ++d) {
      float a[4], bb[4], g[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { bb[i] = k[co[i] + d]; w[i] = v[co[i] + d]; a[i] = q[ro[i] + d]; g[i] = go[ro[i] + d]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { as[i][j] += bb[i] * a[j]; ap[i][j] += w[i] * g[j]; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = tr + TRr * i;
      if (c >= C) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = tc + TRc * j;
        if (r < T) {
          const float p = (causal && c0 + c > r) ? 0.f : expf(as[i][j] * scale - lse[r]);
          PT[c * (T + 1) + r] = p;
          dST[c * (T + 1) + r] = p * (ap[i][j] - dt[r]) * scale;
        }
      }
    }
  }
  __syncthreads();
  for (int tile = threadIdx.x; tile < TRr * 16; tile += blockDim.x) {
    const int tt = tile >> 4, td = tile & 15;
    int cv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int c = tt + TRr * i; cv[i] = (c < C ? c : C - 1) * (T + 1); }
    float ak[4][4], av[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { ak[i][j] = 0.f; av[i][j] = 0.f; }
#pragma unroll 2
    for (int r = 0; r < T; ++r) {
      float s2[4], p2[4], qq[4], gg[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s2[i] = dST[cv[i] + r];
        p2[i] = PT[cv[i] + r];
        qq[i] = q[r * ld + td + 16 * i];
        gg[i] = go[r * ld + td + 16 * i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { ak[i][j] += s2[i] * qq[j]; av[i][j] += p2[i] * gg[j]; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = tt + TRr * i;
      if (c >= C) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long o = ((long long)b * T + c0 + c) * 3 * Wd + h * hd + td + 16 * j;
        store_split(ghi, glo, o + Wd, ak[i][j]);
        store_split(ghi, glo, o + 2 * Wd, av[i][j]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// QuickGELU (x * sigmoid(1.702 x)) forward to fp16 operand planes, and backward dh = dg * gelu'(h).
__global__ void __launch_bounds__(256) quickgelu_fwd_kernel(const float* __restrict__ h, __half* __restrict__ hi, __half* __restrict__ lo, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = h[i];
    store_split(hi, lo, i, x / (1.f + expf(-1.702f * x)));
  }
}
__global__ void __launch_bounds__(256) quickgelu_bwd_kernel(const float* __restrict__ dg, const float* __restrict__ h, __half* __restrict__ hi,
                                                            __half* __restrict__ lo, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float x = h[i];
    const float sg = 1.f / (1.f + expf(-1.702f * x));
    store_split(hi, lo, i, dg[i] * (sg + 1.702f * x * sg * (1.f - sg)));
  }
}
// fp32 rows -> fp16 hi/lo planes; output row r reads input row (r / rows_per_group) * group_stride + group_offset + r % rows_per_group
// (drops the class-token row when feeding the patch-embedding backward GEMM).
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, long long rows,
                                                         int Wd, int rows_per_group, int group_stride, int group_offset) {
  const long long total = rows * Wd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % Wd);
    const long long r = i / Wd;
    const long long src = (r / rows_per_group) * group_stride + group_offset + r % rows_per_group;
    store_split(hi, lo, i, x[src * Wd + d]);
  }
}

// ---------------------------------------------------------------------------------------------------
// Head: E[b,:] = ln[b,:] @ proj   (ln = ln_post(x[:,0,:]) or ln_final(x[b, eot_b]); proj [Wd, E]) and its transpose.
__global__ void __launch_bounds__(256) head_proj_kernel(const float* __restrict__ ln, const float* __restrict__ proj, float* __restrict__ out, int Wd,
                                                        int E) {
  extern __shared__ float row[];
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < Wd; d += blockDim.x) row[d] = ln[(long long)b * Wd + d];
  __syncthreads();
  for (int j = threadIdx.x; j < E; j += blockDim.x) {
    float acc = 0.f;
    for (int d = 0; d < Wd; ++d) acc += row[d] * __ldg(proj + (long long)d * E + j);
    out[(long long)b * E + j] = acc;
  }
}
__global__ void __launch_bounds__(256) head_proj_bwd_kernel(const float* __restrict__ dE, const float* __restrict__ proj, float* __restrict__ dln,
                                                            int Wd, int E) {
  extern __shared__ float row[];
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < E; j += blockDim.x) row[j] = dE[(long long)b * E + j];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int d = warp; d < Wd; d += blockDim.x >> 5) {
    float acc = 0.f;
    for (int j = lane; j < E; j += 32) acc += row[j] * __ldg(proj + (long long)d * E + j);
    acc = warp_sum(acc);
    if (lane == 0) dln[(long long)b * Wd + d] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------
// Directional CLIP loss (clip_loss.py:24-34): e = E_tgt - E_src; cos_n = <e, t> / (|e| |t|);
// loss = coef * (count - sum_n cos_n) / count  with count = global batch (inv_count = 1 / count).
// One block; writes loss_part (this rank's sum of -cos * coef * inv_count; the constant coef is added by
// the caller) and dE_tgt[n,:] = -coef * inv_count * (t/|t| - cos * e/|e|) / |e|.
// normalize != 0: the NADA form (clip_loss_nada.py:162-168,206-218): both embeddings are L2-normalised before the difference,
// e = tgt/|tgt| - src/|src|, and the gradient is taken through the normalisation of tgt: d = (g - (g . a) a) / |tgt| with a = tgt/|tgt|.
__global__ void __launch_bounds__(512) clip_loss_kernel(const float* __restrict__ e_src, const float* __restrict__ e_tgt, const float* __restrict__ text,
                                                        float* __restrict__ loss_part, float* __restrict__ d_tgt, int N, int E, float coef,
                                                        float inv_count, float* __restrict__ gscale_out, float gscale_target, int normalize, int text_stride = 0) {
  __shared__ float red[5][16];
  __shared__ float bc[5];
  float dmax = 0.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float total = 0.f;
  for (int n = 0; n < N; ++n) {
    const float* a = e_tgt + (long long)n * E;
    const float* b = e_src + (long long)n * E;
    const float* tx = text + (long long)n * text_stride;          // text_stride = E: one target vector per sample (identity loss)
    float ia = 1.f, ib = 1.f;                      // 1 / |tgt|, 1 / |src| (normalize), else 1
    if (normalize) {
      float aa = 0.f, bb = 0.f;
      for (int j = threadIdx.x; j < E; j += blockDim.x) { aa += a[j] * a[j]; bb += b[j] * b[j]; }
      aa = warp_sum(aa); bb = warp_sum(bb);
      __syncthreads();
      if (lane == 0) { red[0][warp] = aa; red[1][warp] = bb; }
      __syncthreads();
      if (threadIdx.x == 0) {
        float s0 = 0.f, s1 = 0.f;
        for (int w = 0; w < nw; ++w) { s0 += red[0][w]; s1 += red[1][w]; }
        bc[0] = s0; bc[1] = s1;
      }
      __syncthreads();
      ia = 1.f / fmaxf(sqrtf(bc[0]), 1e-20f);
      ib = 1.f / fmaxf(sqrtf(bc[1]), 1e-20f);
    }
    float ee = 0.f, et = 0.f, tt = 0.f, ta = 0.f, ea = 0.f;
    for (int j = threadIdx.x; j < E; j += blockDim.x) {
      const float ah = a[j] * ia, e = ah - b[j] * ib, t = tx[j];
      ee += e * e; et += e * t; tt += t * t; ta += t * ah; ea += e * ah;
    }
    ee = warp_sum(ee); et = warp_sum(et); tt = warp_sum(tt); ta = warp_sum(ta); ea = warp_sum(ea);
    __syncthreads();
    if (lane == 0) { red[0][warp] = ee; red[1][warp] = et; red[2][warp] = tt; red[3][warp] = ta; red[4][warp] = ea; }
    __syncthreads();
    if (threadIdx.x < 5) {
      float s = 0.f;
      for (int w = 0; w < nw; ++w) s += red[threadIdx.x][w];
      bc[threadIdx.x] = s;
    }
    __syncthreads();
    // e == 0 (edited image identical to the original, i.e. delta == 0): the direction is undefined -- the reference divides
    // 0/0 and turns the whole run into NaN (clip_loss.py:28).  Here that sample contributes cos = 0 and no gradient.
    const bool degenerate = !(bc[0] > 0.f);
    const float ne = fmaxf(sqrtf(bc[0]), 1e-8f), nt = fmaxf(sqrtf(bc[2]), 1e-8f);
    const float cosv = degenerate ? 0.f : bc[1] / (ne * nt);
    total -= cosv;
    if (d_tgt) {
      const float k = -coef * inv_count / ne;
      const float gdot = normalize ? k * (bc[3] / nt - cosv * bc[4] / ne) : 0.f;       // g . a
      for (int j = threadIdx.x; j < E; j += blockDim.x) {
        const float ah = a[j] * ia, e = ah - b[j] * ib;
        float dv = k * (tx[j] / nt - cosv * e / ne);
        if (normalize) dv = (dv - gdot * ah) * ia;
        if (degenerate) dv = 0.f;
        d_tgt[(long long)n * E + j] = dv;
        dmax = fmaxf(dmax, fabsf(dv));
      }
    }
    __syncthreads();                               // bc[] is rewritten by the next sample
  }
  if (threadIdx.x == 0) *loss_part = coef * inv_count * total;
  if (d_tgt && gscale_out) {
    // loss scaling for the fp16 backward GEMMs: d_tgt *= S, S = 2^k with max|d_tgt| * S in [target/2, target)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    __syncthreads();
    if (lane == 0) red[0][warp] = dmax;
    __syncthreads();
    float m = 0.f;
    for (int w = 0; w < nw; ++w) m = fmaxf(m, red[0][w]);
    float S = 1.f;
    if (m > 0.f && isfinite(m)) S = exp2f(floorf(log2f(gscale_target / m)));
    for (int i = threadIdx.x; i < N * E; i += blockDim.x) d_tgt[i] *= S;
    if (threadIdx.x == 0) *gscale_out = S;
  }
}

int g_attention_tiled = 0;   // 0: tiled attention only where the whole-sequence kernel does not fit; 1: always; >= 4: always, rows per CTA capped
                             // at that value (smc_synth_config key 5; tests exercise several row blocks at 50 tokens this way)
int g_resample_vfirst = 0;   // unprocess: vertical pass first for >= 2x down-sampling (smc_synth_config key 4).  Off: under ncu the marching kernel is slower
                             // (1.34 ms at 0.7 TB/s against 0.88 + 0.17 ms, profiles/r02_glue_launches.md); the bench A/B was inside the noise (+0.4 %)

static int grid1d(long long items) {
  long long b = ceil_div_ll(items, 256);
  const long long cap = (long long)kNumSMs * 16;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

int g_resample_rows = 1;     // smc_synth_config key 6: the row-tile kernel for the row-contraction passes of unprocess (0: the per-output kernels)

// resample_rows_kernel when its shared-memory tiles fit; false = not launched (the caller falls back to the per-output kernel)
template <int OBT>
static bool resample_rows_launch(const float* in, float* out, const int* start, int sstride, const int* count, const float* wgt, int taps, long long rows,
                                 int in_w, int out_w, int pre, int post, const float* xmask, const float* unscale, cudaStream_t st) {
  if (!g_resample_rows) return false;
  // OBT consecutive windows: the starts move by at most ceil((OBT - 1) * in / out) + 1, a window has at most `taps` entries
  long long span = ((long long)(OBT - 1) * in_w + out_w - 1) / out_w + taps + 2;
  if (span > in_w) span = in_w;
  const size_t smem = ((size_t)32 * ((size_t)span | 1) + (size_t)32 * (OBT + 1)) * sizeof(float);
  const long long blocks = ceil_div_ll(rows, 32) * ceil_div(out_w, OBT);
  if (smem > 48 * 1024 || blocks > 0x7fffffffLL) return false;
  if (taps <= 8)
    resample_rows_kernel<OBT, 8><<<(unsigned)blocks, 256, smem, st>>>(in, out, start, sstride, count, wgt, taps, rows, in_w, out_w, (int)span, pre, post, xmask, unscale);
  else if (taps <= 12)
    resample_rows_kernel<OBT, 12><<<(unsigned)blocks, 256, smem, st>>>(in, out, start, sstride, count, wgt, taps, rows, in_w, out_w, (int)span, pre, post, xmask, unscale);
  else if (taps <= 24)
    resample_rows_kernel<OBT, 24><<<(unsigned)blocks, 256, smem, st>>>(in, out, start, sstride, count, wgt, taps, rows, in_w, out_w, (int)span, pre, post, xmask, unscale);
  else
    return false;
  return true;
}

}  // namespace smc

using namespace smc;
#define ST ((cudaStream_t)stream)

extern "C" int smc_resample_fwd(const float* x, float* tmp, float* y, const int* start, const int* count, const float* wgt, int taps,
                                int planes, int in_size, int out_size, int denorm_normalize, const float* mean3, const float* std3, void* stream) {
  if (!x || !tmp || !y || !start || !count || !wgt || taps < 1 || planes < 1 || in_size < 1 || out_size < 1 || denorm_normalize < 0 ||
      denorm_normalize > 2)
    return SMC_EINVAL;
  float m[3] = {0, 0, 0}, s[3] = {1, 1, 1};
  if (denorm_normalize) {
    if (!mean3 || !std3) return SMC_EINVAL;
    for (int i = 0; i < 3; ++i) { m[i] = mean3[i]; s[i] = std3[i]; }
  }
  // vertical pass first when down-sampling by 2 or more (the 512 / 1024 px images of the benchmark): see resample_vfirst_kernel
  const int max_active = (int)((long long)taps * out_size / in_size) + 2;         // outputs whose window covers one input row
  if (g_resample_vfirst && in_size >= 2 * out_size && max_active <= 8) {
    const int strips = (in_size + 255) / 256;
    int nseg = (2 * kNumSMs * 4 + strips * planes - 1) / (strips * planes);      // enough blocks for ~8 per SM
    nseg = nseg < 1 ? 1 : (nseg > 8 ? 8 : nseg);
    if (planes <= 65535) {
      resample_vfirst_kernel<<<dim3(strips, planes, nseg), 256, 8 * 256 * sizeof(float), ST>>>(x, tmp, start, count, wgt, taps, in_size, out_size,
                                                                                            in_size, nseg, 7, denorm_normalize);
      const long long rows2 = (long long)planes * out_size;
      resample_h_kernel<<<grid1d(rows2 * out_size), 256, 0, ST>>>(tmp, y, start, count, wgt, taps, rows2, in_size, out_size, 0,
                                                                  denorm_normalize ? out_size : 0, denorm_normalize == 2 ? 1.f : 1.f / 255.f, m[0], m[1], m[2],
                                                                  s[0], s[1], s[2]);
      SMC_LAUNCH_CHECK();
      return SMC_OK;
    }
  }
  const long long rows = (long long)planes * in_size;
  if (!resample_rows_launch<32>(x, tmp, start, 1, count, wgt, taps, rows, in_size, out_size, denorm_normalize, 0, nullptr, nullptr, ST))
    resample_h_kernel<<<grid1d(rows * out_size), 256, 0, ST>>>(x, tmp, start, count, wgt, taps, rows, in_size, out_size, denorm_normalize, 0, 1.f, 0.f,
                                                               0.f, 0.f, 1.f, 1.f, 1.f);
  resample_v_kernel<<<grid1d((long long)planes * out_size * out_size), 256, 0, ST>>>(tmp, y, start, count, wgt, taps, planes, in_size, out_size,
                                                                                      out_size, denorm_normalize == 2 ? 1.f : 1.f / 255.f, m[0], m[1], m[2],
                                                                                      s[0], s[1], s[2], denorm_normalize);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

extern "C" int smc_resample_bwd(const float* g, const float* x, float* tmp, float* gx, const int* oidx, const int* count, const float* wgt,
                                int taps, int planes, int in_size, int out_size, int mode, const float* std3, const float* unscale, void* stream) {
  if (!g || !x || !tmp || !gx || !oidx || !count || !wgt || !std3 || taps < 1 || planes < 1 || mode < 1 || mode > 2) return SMC_EINVAL;
  if (g_resample_vfirst && in_size >= 2 * out_size) {
    // transpose of the vertical-first order: expand the columns on the small [planes, out, .] tensor first, then the rows at full
    // width (coalesced along x) together with the clamp mask and the scale factors
    const long long rows1 = (long long)planes * out_size;
    resample_hT_kernel<<<grid1d(rows1 * in_size), 256, 0, ST>>>(g, nullptr, tmp, oidx, count, wgt, taps, rows1, in_size, out_size, nullptr);
    resample_vT_kernel<<<grid1d((long long)planes * in_size * in_size), 256, 0, ST>>>(tmp, gx, oidx, count, wgt, taps, planes, in_size, out_size,
                                                                                       in_size, std3[0], std3[1], std3[2], x, unscale, mode);
    SMC_LAUNCH_CHECK();
    return SMC_OK;
  }
  resample_vT_kernel<<<grid1d((long long)planes * in_size * out_size), 256, 0, ST>>>(g, tmp, oidx, count, wgt, taps, planes, in_size, out_size,
                                                                                      out_size, std3[0], std3[1], std3[2], nullptr, nullptr, mode);
  const long long rows = (long long)planes * in_size;
  // (the rows of oidx are runs of consecutive output indices -- the transpose of monotone windows: only oidx[ix * taps] is read)
  if (!resample_rows_launch<128>(tmp, gx, oidx, taps, count, wgt, taps, rows, out_size, in_size, 0, mode == 2 ? 2 : 1, x, unscale, ST))
    resample_hT_kernel<<<grid1d(rows * in_size), 256, 0, ST>>>(tmp, x, gx, oidx, count, wgt, taps, rows, in_size, out_size, unscale, mode);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

extern "C" int smc_patchify(const float* img, void* hi, void* lo, int b, int res, int ps, void* stream) {
  if (!img || !hi || b < 1 || res < 1 || ps < 1 || res % ps) return SMC_EINVAL;
  patchify_kernel<<<grid1d((long long)b * 3 * res * res), 256, 0, ST>>>(img, (__half*)hi, (__half*)lo, b, res, ps);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
extern "C" int smc_unpatchify(const float* gp, float* gimg, int b, int res, int ps, void* stream) {
  if (!gp || !gimg || b < 1 || res < 1 || ps < 1 || res % ps) return SMC_EINVAL;
  unpatchify_kernel<<<grid1d((long long)b * 3 * res * res), 256, 0, ST>>>(gp, gimg, b, res, ps);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
extern "C" int smc_assemble_tokens(const float* patch, const float* cls, const float* pos, float* x0, int b, int t, int wd, void* stream) {
  if (!patch || !cls || !pos || !x0 || b < 1 || t < 2 || wd < 1) return SMC_EINVAL;
  assemble_tokens_kernel<<<grid1d((long long)b * t * wd), 256, 0, ST>>>(patch, cls, pos, x0, b, t, wd);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
extern "C" int smc_embed_text(const int64_t* text, const float* emb, const float* pos, float* x0, int b, int t, int wd, void* stream) {
  if (!text || !emb || !pos || !x0 || b < 1 || t < 1 || wd < 1) return SMC_EINVAL;
  embed_text_kernel<<<grid1d((long long)b * t * wd), 256, 0, ST>>>((const long long*)text, emb, pos, x0, b, t, wd);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
extern "C" int smc_layernorm_fwd(const float* x, int64_t in_row_stride, int64_t in_row_offset, const float* w, const float* b, float* y32,
                                 void* yhi, void* ylo, float* mean, float* rstd, int64_t rows, int wd, void* stream) {
  if (!x || !w || !b || rows < 1 || wd < 1 || (!y32 && !yhi) || ((mean == nullptr) != (rstd == nullptr))) return SMC_EINVAL;
  layernorm_fwd_kernel<<<grid1d(rows * 32), 256, 0, ST>>>(x, in_row_stride, in_row_offset, w, b, y32, (__half*)yhi, (__half*)ylo, mean, rstd, rows, wd);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
extern "C" int smc_layernorm_bwd(const float* dy, const float* x, int64_t in_row_stride, int64_t in_row_offset, const float* w,
                                 const float* mean, const float* rstd, float* dx, int64_t rows, int wd, int accumulate, void* stream) {
  if (!dy || !x || !w || !mean || !rstd || !dx || rows < 1 || wd < 1) return SMC_EINVAL;
  layernorm_bwd_kernel<<<grid1d(rows * 32), 256, 0, ST>>>(dy, x, in_row_stride, in_row_offset, w, mean, rstd, dx, rows, wd, accumulate);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
// Rows per CTA of the tiled attention kernels: the fewest blocks whose shared memory (whole[T][65] x 2 + per-row words) stays under 200 KB.
// per_row = words that scale with the block (operand rows + score rows), extra = fixed words.  Returns 0 when even 4 rows do not fit.
static int attention_block_rows(int t, int per_row, int extra, size_t* smem) {
  const int cap = smc::g_attention_tiled >= 4 ? smc::g_attention_tiled : t;
  for (int nb = 1; nb <= t; ++nb) {
    int rows = (((t + nb - 1) / nb) + 3) & ~3;
    if (rows > ((cap + 3) & ~3)) continue;
    *smem = ((size_t)2 * t * 65 + (size_t)rows * per_row + extra) * sizeof(float);
    if (*smem <= 200 * 1024) return rows;
  }
  return 0;
}
extern "C" int smc_attention_fwd(const float* qkv, void* ohi, void* olo, float* o32, int b, int t, int wd, int heads, int causal, void* stream) {
  if (!qkv || (!ohi && !o32) || b < 1 || t < 1 || heads < 1 || wd % heads) return SMC_EINVAL;
  const int hd = wd / heads;
  const size_t smem = (size_t)(3 * t * (hd + 1) + t * (t + 1)) * sizeof(float);
  if (smem > 200 * 1024 || smc::g_attention_tiled) {       // long sequences (ViT-B/16: 197 tokens): query-row blocks
    if (hd != 64) return SMC_EUNSUPPORTED;
    size_t sm2 = 0;
    const int qb = attention_block_rows(t, 65 + t + 1, 0, &sm2);
    if (!qb) return SMC_EUNSUPPORTED;
    const int nqb = (t + qb - 1) / qb;
    cudaError_t e = cudaFuncSetAttribute(attention_fwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    if (e != cudaSuccess) return (int)e;
    attention_fwd_rows_kernel<<<b * heads * nqb, 256, sm2, ST>>>(qkv, (__half*)ohi, (__half*)olo, o32, t, wd, heads, causal, qb, nqb);
    SMC_LAUNCH_CHECK();
    return SMC_OK;
  }
  cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  if (hd == 64) {
    e = cudaFuncSetAttribute(attention_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attention_fwd2_kernel<<<b * heads, 256, smem, ST>>>(qkv, (__half*)ohi, (__half*)olo, o32, t, wd, heads, causal);
    SMC_LAUNCH_CHECK();
    return SMC_OK;
  }
  attention_fwd_kernel<<<b * heads, 256, smem, ST>>>(qkv, (__half*)ohi, (__half*)olo, o32, t, wd, heads, causal);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
extern "C" int smc_attention_bwd(const float* qkv, const float* d_o, void* ghi, void* glo, int b, int t, int wd, int heads, int causal,
                                 void* stream) {
  if (!qkv || !d_o || !ghi || b < 1 || t < 1 || heads < 1 || wd % heads) return SMC_EINVAL;
  const int hd = wd / heads;
  const size_t smem = (size_t)(4 * t * (hd + 1) + 2 * t * (t + 1)) * sizeof(float);
  if (smem > 200 * 1024) return SMC_EUNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  if (hd == 64) {
    e = cudaFuncSetAttribute(attention_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attention_bwd2_kernel<<<b * heads, 256, smem, ST>>>(qkv, d_o, (__half*)ghi, (__half*)glo, t, wd, heads, causal);
    SMC_LAUNCH_CHECK();
    return SMC_OK;
  }
  attention_bwd_kernel<<<b * heads, 256, smem, ST>>>(qkv, d_o, (__half*)ghi, (__half*)glo, t, wd, heads, causal);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
// Backward for sequences the whole-sequence kernel cannot hold (see attention_bwd_q_kernel): two launches, dQ per query-row block, then
// dK/dV per key-row block.  stats: caller-allocated fp32 scratch of 2 * b * heads * t words (row log-sum-exp and rowsum(P * dP)).
extern "C" int smc_attention_bwd_tiled(const float* qkv, const float* d_o, void* ghi, void* glo, float* stats, int b, int t, int wd, int heads,
                                       int causal, void* stream) {
  if (!qkv || !d_o || !ghi || !stats || b < 1 || t < 1 || heads < 1 || wd % heads) return SMC_EINVAL;
  if (wd / heads != 64) return SMC_EUNSUPPORTED;
  size_t smem = 0;
  const int rb = attention_block_rows(t, 2 * 65 + 2 * (t + 1), 2 * t, &smem);
  if (!rb) return SMC_EUNSUPPORTED;
  const int nb = (t + rb - 1) / rb;
  cudaError_t e = cudaFuncSetAttribute(attention_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(attention_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  attention_bwd_q_kernel<<<b * heads * nb, 256, smem, ST>>>(qkv, d_o, (__half*)ghi, (__half*)glo, stats, t, wd, heads, causal, rb, nb);
  SMC_LAUNCH_CHECK();
  attention_bwd_kv_kernel<<<b * heads * nb, 256, smem, ST>>>(qkv, d_o, (__half*)ghi, (__half*)glo, stats, t, wd, heads, causal, rb, nb);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
extern "C" int smc_quickgelu_fwd(const float* h, void* hi, void* lo, int64_t n, void* stream) {
  if (!h || !hi || n < 1) return SMC_EINVAL;
  quickgelu_fwd_kernel<<<grid1d(n), 256, 0, ST>>>(h, (__half*)hi, (__half*)lo, n);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
extern "C" int smc_quickgelu_bwd(const float* dg, const float* h, void* hi, void* lo, int64_t n, void* stream) {
  if (!dg || !h || !hi || n < 1) return SMC_EINVAL;
  quickgelu_bwd_kernel<<<grid1d(n), 256, 0, ST>>>(dg, h, (__half*)hi, (__half*)lo, n);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
extern "C" int smc_split_rows(const float* x, void* hi, void* lo, int64_t rows, int wd, int rows_per_group, int group_stride, int group_offset,
                              void* stream) {
  if (!x || !hi || rows < 1 || wd < 1 || rows_per_group < 1) return SMC_EINVAL;
  split_rows_kernel<<<grid1d(rows * wd), 256, 0, ST>>>(x, (__half*)hi, (__half*)lo, rows, wd, rows_per_group, group_stride, group_offset);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
extern "C" int smc_head_proj(const float* ln, const float* proj, float* out, int b, int wd, int e, void* stream) {
  if (!ln || !proj || !out || b < 1 || wd < 1 || e < 1) return SMC_EINVAL;
  head_proj_kernel<<<b, 256, wd * sizeof(float), ST>>>(ln, proj, out, wd, e);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
extern "C" int smc_head_proj_bwd(const float* d_e, const float* proj, float* dln, int b, int wd, int e, void* stream) {
  if (!d_e || !proj || !dln || b < 1 || wd < 1 || e < 1) return SMC_EINVAL;
  head_proj_bwd_kernel<<<b, 256, e * sizeof(float), ST>>>(d_e, proj, dln, wd, e);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
extern "C" int smc_clip_loss(const float* e_src, const float* e_tgt, const float* text, float* loss_part, float* d_tgt, int n, int e, float coef,
                             float inv_count, float* gscale_out, float gscale_target, int normalize, int text_stride, void* stream) {
  if (!e_src || !e_tgt || !text || !loss_part || n < 1 || e < 1 || (text_stride != 0 && text_stride < e)) return SMC_EINVAL;
  clip_loss_kernel<<<1, 512, 0, ST>>>(e_src, e_tgt, text, loss_part, d_tgt, n, e, coef, inv_count, gscale_out, gscale_target, normalize, text_stride);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
