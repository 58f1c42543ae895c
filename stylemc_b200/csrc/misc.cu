// ABI version + the optimiser step of find_direction.py:285,339 (SGD, no momentum) fused with the
// analytic L2 term of find_direction.py:190-191 and the unscaling of the (allreduced) gradient.
#include "common.cuh"

namespace smc {
__global__ void sgd_step_kernel(float* __restrict__ delta, const float* __restrict__ grad, long long n, float lr, float grad_scale,
                                float l2_scale, const float* __restrict__ lr_dev = nullptr) {
  if (lr_dev) lr = __ldg(lr_dev);          // learning rate from device memory: a captured CUDA graph replays with this step's value
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
// Same, four pixels per thread: one 16-byte load per colour plane, twelve output bytes as three aligned 32-bit stores
// (needs W % 4 == 0, canvas_w % 4 == 0, x_off % 4 == 0 and 16-byte aligned planes; a warp then writes 384 contiguous bytes).
__device__ __forceinline__ unsigned img_u8(float v) { return (unsigned)fminf(fmaxf(fmaf(v, 127.5f, 128.f), 0.f), 255.f); }
__global__ void __launch_bounds__(256) img_to_uint8_v4_kernel(const float* __restrict__ img, unsigned char* __restrict__ out, int N, int H, int W4,
                                                              int canvas_w, int x_off) {
  const long long total = (long long)N * H * W4;
  const long long plane = (long long)H * W4;          // in float4 units
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x4 = (int)(i % W4);
    const long long t = i / W4;
    const int y = (int)(t % H);
    const long long n = t / H;
    const float4* src = reinterpret_cast<const float4*>(img) + (n * 3 * H + y) * (long long)W4 + x4;
    const uint4 ur = ld_stream(src), ug = ld_stream(src + plane), ub = ld_stream(src + 2 * plane);
    const float r[4] = {__uint_as_float(ur.x), __uint_as_float(ur.y), __uint_as_float(ur.z), __uint_as_float(ur.w)};
    const float g[4] = {__uint_as_float(ug.x), __uint_as_float(ug.y), __uint_as_float(ug.z), __uint_as_float(ug.w)};
    const float b[4] = {__uint_as_float(ub.x), __uint_as_float(ub.y), __uint_as_float(ub.z), __uint_as_float(ub.w)};
    unsigned by[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) { by[3 * k] = img_u8(r[k]); by[3 * k + 1] = img_u8(g[k]); by[3 * k + 2] = img_u8(b[k]); }
    unsigned* o = reinterpret_cast<unsigned*>(out + ((n * H + y) * (long long)canvas_w + x_off + 4 * x4) * 3);
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] = by[4 * k] | (by[4 * k + 1] << 8) | (by[4 * k + 2] << 16) | (by[4 * k + 3] << 24);
  }
}

// out = g * mask * (*scale): the ToRGB clamp mask saved by the forward pass (smc_img_finish pass_mask) and the loss scale applied to the
// incoming image gradient in ONE pass (was two ATen launches, the first a type-promoting uint8 x float multiply on the scalar path).
__global__ void __launch_bounds__(256) mask_scale_v4_kernel(const float* __restrict__ g, const unsigned char* __restrict__ mask,
                                                            const float* __restrict__ scale, float* __restrict__ out, long long n4) {
  const float k = scale ? __ldg(scale) : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = ld_stream(reinterpret_cast<const float4*>(g) + i);
    const uchar4 m = __ldg(reinterpret_cast<const uchar4*>(mask) + i);
    st_stream(reinterpret_cast<float4*>(out) + i,
              make_uint4(__float_as_uint(__uint_as_float(u.x) * (float)m.x * k), __float_as_uint(__uint_as_float(u.y) * (float)m.y * k),
                         __float_as_uint(__uint_as_float(u.z) * (float)m.z * k), __float_as_uint(__uint_as_float(u.w) * (float)m.w * k)));
  }
}
__global__ void __launch_bounds__(256) mask_scale_kernel(const float* __restrict__ g, const unsigned char* __restrict__ mask,
                                                         const float* __restrict__ scale, float* __restrict__ out, long long n) {
  const float k = scale ? __ldg(scale) : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = g[i] * (float)mask[i] * k;
}

// ---- one-off weight preparation (frozen G / CLIP): [O, I, T] fp32 conv or linear weights -> the K-major tap matrices of the implicit
// GEMM as fp16 hi / lo planes, forward [T * Op, Ip] (row t * Op + o, column i) and dgrad [T * Ip, Op] (row t * Ip + i, column o), plus
// q[o, i] = sum_t w^2 for the demodulation coefficients.  Replaces ~15 ATen launches per layer (permute / contiguous / half / sub /
// stack / square / sum).  `scale` (device scalar, optional) is a power of two multiplied in before the split (small CLIP weights: keeps the
// lo plane out of the fp16 subnormals; the GEMM undoes it with acc_scale).
struct PrepW {
  const float* w;
  const float* scale;
  __half *fwd_hi, *fwd_lo, *bwd_hi, *bwd_lo;
  float* q;
  int O, I, T, Op, Ip;
  long long n_fwd, n_bwd, n_q;
};
__global__ void __launch_bounds__(256) prepare_weights_kernel(PrepW p) {
  const float k = p.scale ? __ldg(p.scale) : 1.f;
  const long long total = p.n_fwd + p.n_bwd + p.n_q;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    if (idx < p.n_fwd + p.n_bwd) {
      const bool bwd = idx >= p.n_fwd;
      long long r = bwd ? idx - p.n_fwd : idx;
      const int inner = bwd ? p.Op : p.Ip, mid = bwd ? p.Ip : p.Op;
      const int c = (int)(r % inner); r /= inner;
      const int m = (int)(r % mid);
      const int t = (int)(r / mid);
      const int o = bwd ? c : m, i = bwd ? m : c;
      float v = 0.f;
      if (o < p.O && i < p.I) v = __ldg(p.w + ((long long)o * p.I + i) * p.T + t) * k;
      __half hi, lo;
      split_half(v, hi, lo);
      __half* dh = bwd ? p.bwd_hi : p.fwd_hi;
      __half* dl = bwd ? p.bwd_lo : p.fwd_lo;
      const long long at = bwd ? idx - p.n_fwd : idx;
      dh[at] = hi;
      if (dl) dl[at] = lo;
    } else {
      const long long r = idx - p.n_fwd - p.n_bwd;          // (o, i) of the un-padded weight
      const float* src = p.w + r * p.T;
      float acc = 0.f;
      for (int t = 0; t < p.T; ++t) { const float v = __ldg(src + t); acc = fmaf(v, v, acc); }
      p.q[r] = acc;
    }
  }
}

// ---- identity-loss glue (id_loss/id_loss.py:18-24, id_loss/helpers.py): per-channel PReLU and adaptive average pooling of a cropped window
// y = x > 0 ? x : alpha[c] * x on NCHW (channel = (i / hw) % C); grad != 0: y = dy * (x > 0 ? 1 : alpha[c])
__global__ void __launch_bounds__(256) prelu_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ alpha,
                                                    float* __restrict__ y, long long total, int hw, int C) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i], a = __ldg(alpha + (int)((i / hw) % C));
    y[i] = dy ? dy[i] * (v > 0.f ? 1.f : a) : (v > 0.f ? v : a * v);
  }
}
// torch.nn.AdaptiveAvgPool2d over the window [y0, y0 + hc) x [x0, x0 + wc) of every plane: output (oy, ox) averages rows
// floor(oy * hc / OH) .. ceil((oy + 1) * hc / OH) - 1 of the window (and the same for columns).
__device__ __forceinline__ int ap_start(int o, int in, int out) { return (int)(((long long)o * in) / out); }
__device__ __forceinline__ int ap_end(int o, int in, int out) { return (int)((((long long)(o + 1)) * in + out - 1) / out); }
__global__ void __launch_bounds__(256) adaptive_pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long planes, int H, int W,
                                                                int y0, int x0, int hc, int wc, int OH, int OW) {
  const long long total = planes * OH * OW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW), oy = (int)((i / OW) % OH);
    const long long pl = i / ((long long)OW * OH);
    const int ys = ap_start(oy, hc, OH), ye = ap_end(oy, hc, OH), xs = ap_start(ox, wc, OW), xe = ap_end(ox, wc, OW);
    const float* src = x + pl * H * W;
    float acc = 0.f;
    for (int yy = ys; yy < ye; ++yy)
      for (int xx = xs; xx < xe; ++xx) acc += __ldg(src + (long long)(y0 + yy) * W + x0 + xx);
    y[i] = acc / (float)((ye - ys) * (xe - xs));
  }
}
// transpose: every input pixel gathers dy / window size from the (few) outputs whose window contains it; pixels outside the crop get 0
__global__ void __launch_bounds__(256) adaptive_pool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long long planes, int H, int W,
                                                                int y0, int x0, int hc, int wc, int OH, int OW) {
  const long long total = planes * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W) - x0, yy = (int)((i / W) % H) - y0;
    const long long pl = i / ((long long)W * H);
    float acc = 0.f;
    if (yy >= 0 && yy < hc && xx >= 0 && xx < wc) {
      const int oy_lo = max(0, (int)(((long long)yy * OH) / hc) - 1), oy_hi = min(OH - 1, (int)(((long long)(yy + 1) * OH + hc - 1) / hc));
      const int ox_lo = max(0, (int)(((long long)xx * OW) / wc) - 1), ox_hi = min(OW - 1, (int)(((long long)(xx + 1) * OW + wc - 1) / wc));
      const float* g = dy + pl * OH * OW;
      for (int oy = oy_lo; oy <= oy_hi; ++oy) {
        const int ys = ap_start(oy, hc, OH), ye = ap_end(oy, hc, OH);
        if (yy < ys || yy >= ye) continue;
        for (int ox = ox_lo; ox <= ox_hi; ++ox) {
          const int xs = ap_start(ox, wc, OW), xe = ap_end(ox, wc, OW);
          if (xx < xs || xx >= xe) continue;
          acc += __ldg(g + (long long)oy * OW + ox) / (float)((ye - ys) * (xe - xs));
        }
      }
    }
    dx[i] = acc;
  }
}

// ---- latent-mapper glue (latent_mappers.py:12-93, train_latent_mapper.py:131,150-196)
// PixelNorm over dim 1 of x [B, L, C] (encoder4editing/models/stylegan2/model.py:14-15): y = x * rsqrt(mean_l x^2 + 1e-8); one thread per (b, c).
// dy != NULL: the input gradient dx_l = r dy_l - x_l r^3 mean_k(dy_k x_k).
__global__ void __launch_bounds__(256) pixelnorm_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ y, int B, int L, int C) {
  const long long total = (long long)B * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long base = (i / C) * L * C + c;
    float ss = 0.f, dot = 0.f;
    for (int l = 0; l < L; ++l) {
      const float v = x[base + (long long)l * C];
      ss += v * v;
      if (dy) dot += dy[base + (long long)l * C] * v;
    }
    const float r = rsqrtf(ss / L + 1e-8f);
    const float k = dy ? r * r * r * dot / L : 0.f;
    for (int l = 0; l < L; ++l) {
      const long long at = base + (long long)l * C;
      y[at] = dy ? r * dy[at] - x[at] * k : x[at] * r;
    }
  }
}
// y[m, n] = sum_k a[m, k] * b[n, k] (+ bias[n]) with the products accumulated in float64 and rounded once: the mapper's Linear(512, 512)
// layers (latent_mappers.py:16).  Their output becomes the per-image delta S, and the synthesis gradient is so sensitive to S (a 5e-6
// perturbation of delta moves it by 8e-4: leaky-ReLU slope flips) that the split-fp16 tensor-core GEMM (~1e-6) is not accurate enough here;
// the matrices are tiny (B * 4 rows), so the scalar pipe does it.  Element strides: any of x W^T, dy W and dy^T x without a transpose.
__global__ void __launch_bounds__(256) matmul_nt_f64acc_kernel(const float* __restrict__ a, long long sa_m, long long sa_k, const float* __restrict__ b,
                                                               long long sb_n, long long sb_k, const float* __restrict__ bias, float* __restrict__ y,
                                                               int M, int N, int K) {
  __shared__ float ta[16][17], tb[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
  double acc = 0.0;
  for (int k0 = 0; k0 < K; k0 += 16) {
    // tile loads: thread (ty, tx) fetches a[m0 + ty, k0 + tx] and b[n0 + ty, k0 + tx]
    const int am = blockIdx.y * 16 + ty, bn = blockIdx.x * 16 + ty, kk = k0 + tx;
    ta[ty][tx] = (am < M && kk < K) ? __ldg(a + am * sa_m + kk * sa_k) : 0.f;
    tb[ty][tx] = (bn < N && kk < K) ? __ldg(b + bn * sb_n + kk * sb_k) : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += (double)ta[ty][k] * (double)tb[tx][k];
    __syncthreads();
  }
  if (m < M && n < N) y[(long long)m * N + n] = (float)(acc + (bias ? (double)__ldg(bias + n) : 0.0));
}
// torch.optim.Adam (no weight decay, no amsgrad; train_latent_mapper.py:131): bc1 = 1 - beta1^t, bc2_sqrt = sqrt(1 - beta2^t) from the host
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                        long long n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

// ---- fma.py:15-58 as stand-alone kernels: out = a * b + c over a broadcast 4-D index space, and the "un-broadcast" of its
// backward (sum of x * y over the axes broadcasting expanded).  Element strides; 0 marks a broadcast / reduced axis.
struct FmaDims {
  long long size[4];
  long long sa[4], sb[4], sc[4];
};
template <typename T, typename ACC>
__global__ void __launch_bounds__(256) fma_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c, T* __restrict__ out,
                                                      FmaDims d, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const long long i3 = t % d.size[3]; t /= d.size[3];
    const long long i2 = t % d.size[2]; t /= d.size[2];
    const long long i1 = t % d.size[1];
    const long long i0 = t / d.size[1];
    const ACC va = (ACC)a[i0 * d.sa[0] + i1 * d.sa[1] + i2 * d.sa[2] + i3 * d.sa[3]];
    const ACC vb = (ACC)b[i0 * d.sb[0] + i1 * d.sb[1] + i2 * d.sb[2] + i3 * d.sb[3]];
    const ACC vc = (ACC)c[i0 * d.sc[0] + i1 * d.sc[1] + i2 * d.sc[2] + i3 * d.sc[3]];
    out[i] = (T)(va * vb + vc);           // contracted to one FMA (torch.addcmul rounds the product for fp32; the difference is < 1 ulp)
  }
}
// out[kept index] = sum over the reduced axes of x * y (y optional).  size = full shape, sa = x strides, sb = y strides,
// sc = out strides (0 on reduced axes).  One block per output element when the reduction is long, else one thread.
template <typename T, typename ACC, bool BLOCK>
__global__ void __launch_bounds__(256) fma_reduce_kernel(const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ out, FmaDims d,
                                                         long long n_out, long long n_red) {
  __shared__ ACC red[8];
  long long ksz[4], rsz[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool reduced = d.sc[k] == 0 && d.size[k] > 1;
    ksz[k] = reduced ? 1 : d.size[k];
    rsz[k] = reduced ? d.size[k] : 1;
  }
  const long long first = BLOCK ? (long long)blockIdx.x : (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = BLOCK ? (long long)gridDim.x : (long long)gridDim.x * blockDim.x;
  for (long long o = first; o < n_out; o += stride) {
    long long t = o;
    const long long k3 = t % ksz[3]; t /= ksz[3];
    const long long k2 = t % ksz[2]; t /= ksz[2];
    const long long k1 = t % ksz[1];
    const long long k0 = t / ksz[1];
    const long long xo = k0 * d.sa[0] + k1 * d.sa[1] + k2 * d.sa[2] + k3 * d.sa[3];
    const long long yo = k0 * d.sb[0] + k1 * d.sb[1] + k2 * d.sb[2] + k3 * d.sb[3];
    ACC acc = (ACC)0;
    for (long long r = BLOCK ? threadIdx.x : 0; r < n_red; r += BLOCK ? blockDim.x : 1) {
      long long u = r;
      const long long r3 = u % rsz[3]; u /= rsz[3];
      const long long r2 = u % rsz[2]; u /= rsz[2];
      const long long r1 = u % rsz[1];
      const long long r0 = u / rsz[1];
      const ACC vx = (ACC)x[xo + r0 * d.sa[0] + r1 * d.sa[1] + r2 * d.sa[2] + r3 * d.sa[3]];
      acc += y ? vx * (ACC)y[yo + r0 * d.sb[0] + r1 * d.sb[1] + r2 * d.sb[2] + r3 * d.sb[3]] : vx;
    }
    if (BLOCK) {
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
      __syncthreads();                                   // red[] of the previous output has been read
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
      __syncthreads();
      if (threadIdx.x == 0) {
        ACC tot = (ACC)0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
        out[k0 * d.sc[0] + k1 * d.sc[1] + k2 * d.sc[2] + k3 * d.sc[3]] = (T)tot;
      }
    } else {
      out[k0 * d.sc[0] + k1 * d.sc[1] + k2 * d.sc[2] + k3 * d.sc[3]] = (T)acc;
    }
  }
}

static bool fma_dims(FmaDims& d, const int64_t* shape, const int64_t* sa, const int64_t* sb, const int64_t* sc, long long* total) {
  long long n = 1;
  for (int k = 0; k < 4; ++k) {
    if (shape[k] < 1) return false;
    d.size[k] = shape[k];
    d.sa[k] = sa ? sa[k] : 0; d.sb[k] = sb ? sb[k] : 0; d.sc[k] = sc ? sc[k] : 0;
    n *= shape[k];
    if (n > 0x7fffffffLL) return false;
  }
  *total = n;
  return true;
}
template <typename T, typename ACC>
static int fma_launch(const void* a, const void* b, const void* c, void* out, const FmaDims& d, long long total, cudaStream_t st) {
  long long blocks = ceil_div_ll(total, 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  fma_fwd_kernel<T, ACC><<<(int)blocks, 256, 0, st>>>((const T*)a, (const T*)b, (const T*)c, (T*)out, d, total);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
template <typename T, typename ACC>
static int fma_reduce_launch(const void* x, const void* y, void* out, const FmaDims& d, long long n_out, long long n_red, cudaStream_t st) {
  if (n_red >= 128) {
    const long long blocks = n_out < (long long)kNumSMs * 16 ? n_out : (long long)kNumSMs * 16;
    fma_reduce_kernel<T, ACC, true><<<(int)blocks, 256, 0, st>>>((const T*)x, (const T*)y, (T*)out, d, n_out, n_red);
  } else {
    long long blocks = ceil_div_ll(n_out, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    fma_reduce_kernel<T, ACC, false><<<(int)blocks, 256, 0, st>>>((const T*)x, (const T*)y, (T*)out, d, n_out, n_red);
  }
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
}  // namespace smc

extern "C" int smc_prepare_weights(const float* w, int n_out, int n_in, int ntaps, int n_out_padded, int n_in_padded, const float* scale,
                                   void* fwd_hi, void* fwd_lo, void* bwd_hi, void* bwd_lo, float* q, void* stream) {
  if (!w || n_out < 1 || n_in < 1 || ntaps < 1 || n_out_padded < n_out || n_in_padded < n_in) return SMC_EINVAL;
  if ((!fwd_hi && fwd_lo) || (!bwd_hi && bwd_lo) || (!fwd_hi && !bwd_hi && !q)) return SMC_EINVAL;
  smc::PrepW p;
  p.w = w; p.scale = scale;
  p.fwd_hi = (__half*)fwd_hi; p.fwd_lo = (__half*)fwd_lo; p.bwd_hi = (__half*)bwd_hi; p.bwd_lo = (__half*)bwd_lo; p.q = q;
  p.O = n_out; p.I = n_in; p.T = ntaps; p.Op = n_out_padded; p.Ip = n_in_padded;
  const long long plane = (long long)ntaps * n_out_padded * n_in_padded;
  if (plane > 0x7fffffffLL) return SMC_ETOOLARGE;
  p.n_fwd = fwd_hi ? plane : 0;
  p.n_bwd = bwd_hi ? plane : 0;
  p.n_q = q ? (long long)n_out * n_in : 0;
  long long blocks = smc::ceil_div_ll(p.n_fwd + p.n_bwd + p.n_q, 256);
  if (blocks > smc::kNumSMs * 16) blocks = smc::kNumSMs * 16;
  smc::prepare_weights_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

extern "C" int smc_prelu(const float* x, const float* dy, const float* alpha, float* y, int64_t numel, int hw, int c, void* stream) {
  if (!x || !alpha || !y || numel < 1 || hw < 1 || c < 1) return SMC_EINVAL;
  long long blocks = smc::ceil_div_ll(numel, 256);
  if (blocks > smc::kNumSMs * 16) blocks = smc::kNumSMs * 16;
  smc::prelu_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, dy, alpha, y, numel, hw, c);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

extern "C" int smc_adaptive_avg_pool(const float* x, float* y, int64_t planes, int h, int w, int y0, int x0, int hc, int wc, int oh, int ow,
                                     int backward, void* stream) {
  if (!x || !y || planes < 1 || h < 1 || w < 1 || oh < 1 || ow < 1 || y0 < 0 || x0 < 0 || hc < 1 || wc < 1 || y0 + hc > h || x0 + wc > w)
    return SMC_EINVAL;
  const long long total = planes * (backward ? (long long)h * w : (long long)oh * ow);
  if (total > 0x7fffffffLL * 4) return SMC_ETOOLARGE;
  long long blocks = smc::ceil_div_ll(total, 256);
  if (blocks > smc::kNumSMs * 16) blocks = smc::kNumSMs * 16;
  if (backward) smc::adaptive_pool_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, planes, h, w, y0, x0, hc, wc, oh, ow);
  else smc::adaptive_pool_fwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, planes, h, w, y0, x0, hc, wc, oh, ow);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

extern "C" int smc_pixelnorm(const float* x, const float* dy, float* y, int b, int l, int c, void* stream) {
  if (!x || !y || b < 1 || l < 1 || c < 1) return SMC_EINVAL;
  long long blocks = smc::ceil_div_ll((long long)b * c, 256);
  if (blocks > smc::kNumSMs * 16) blocks = smc::kNumSMs * 16;
  smc::pixelnorm_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, dy, y, b, l, c);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

extern "C" int smc_matmul_nt_f64acc(const float* a, int64_t sa_m, int64_t sa_k, const float* b, int64_t sb_n, int64_t sb_k, const float* bias, float* y,
                                    int m, int n, int k, void* stream) {
  if (!a || !b || !y || m < 1 || n < 1 || k < 1) return SMC_EINVAL;
  if (smc::ceil_div(m, 16) > 65535) return SMC_ETOOLARGE;
  smc::matmul_nt_f64acc_kernel<<<dim3(smc::ceil_div(n, 16), smc::ceil_div(m, 16)), 256, 0, (cudaStream_t)stream>>>(a, sa_m, sa_k, b, sb_n, sb_k, bias, y, m, n,
                                                                                                               k);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

extern "C" int smc_adam_step(float* p, const float* g, float* m, float* v, int64_t numel, float lr, float beta1, float beta2, float eps, float bc1,
                             float bc2_sqrt, void* stream) {
  if (!p || !g || !m || !v || numel < 1 || !(bc1 > 0.f) || !(bc2_sqrt > 0.f)) return SMC_EINVAL;
  long long blocks = smc::ceil_div_ll(numel, 256);
  if (blocks > smc::kNumSMs * 16) blocks = smc::kNumSMs * 16;
  smc::adam_step_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, numel, lr, beta1, beta2, eps, bc1, bc2_sqrt);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

extern "C" int smc_fma(const void* a, const void* b, const void* c, void* out, int dtype, const int64_t* shape, const int64_t* stride_a,
                       const int64_t* stride_b, const int64_t* stride_c, void* stream) {
  if (!a || !b || !c || !out || !shape || !stride_a || !stride_b || !stride_c) return SMC_EINVAL;
  smc::FmaDims d;
  long long total = 0;
  if (!smc::fma_dims(d, shape, stride_a, stride_b, stride_c, &total)) return SMC_ETOOLARGE;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SMC_F32) return smc::fma_launch<float, float>(a, b, c, out, d, total, st);
  if (dtype == SMC_F16) return smc::fma_launch<__half, float>(a, b, c, out, d, total, st);
  if (dtype == SMC_F64) return smc::fma_launch<double, double>(a, b, c, out, d, total, st);
  return SMC_EINVAL;
}

extern "C" int smc_fma_reduce(const void* x, const void* y, void* out, int dtype, const int64_t* shape, const int64_t* stride_x,
                              const int64_t* stride_y, const int64_t* stride_out, void* stream) {
  if (!x || !out || !shape || !stride_x || !stride_out || (y && !stride_y)) return SMC_EINVAL;
  smc::FmaDims d;
  long long total = 0;
  if (!smc::fma_dims(d, shape, stride_x, stride_y, stride_out, &total)) return SMC_ETOOLARGE;
  long long n_red = 1;
  for (int k = 0; k < 4; ++k)
    if (stride_out[k] == 0 && shape[k] > 1) n_red *= shape[k];
  const long long n_out = total / n_red;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SMC_F32) return smc::fma_reduce_launch<float, float>(x, y, out, d, n_out, n_red, st);
  if (dtype == SMC_F16) return smc::fma_reduce_launch<__half, float>(x, y, out, d, n_out, n_red, st);
  if (dtype == SMC_F64) return smc::fma_reduce_launch<double, double>(x, y, out, d, n_out, n_red, st);
  return SMC_EINVAL;
}

extern "C" int smc_mask_scale(const float* g, const unsigned char* mask, const float* scale, float* out, int64_t numel, void* stream) {
  if (!g || !mask || !out || numel < 1) return SMC_EINVAL;
  const bool v4 = numel % 4 == 0 && (((uintptr_t)g | (uintptr_t)out) & 15) == 0 && ((uintptr_t)mask & 3) == 0;
  long long blocks = smc::ceil_div_ll(v4 ? numel / 4 : numel, 256);
  if (blocks > smc::kNumSMs * 16) blocks = smc::kNumSMs * 16;
  if (v4) smc::mask_scale_v4_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(g, mask, scale, out, numel / 4);
  else smc::mask_scale_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(g, mask, scale, out, numel);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

extern "C" int smc_img_to_uint8(const float* img, unsigned char* out, int n, int h, int w, int canvas_w, int x_off, void* stream) {
  if (!img || !out || n < 1 || h < 1 || w < 1 || x_off < 0 || x_off + w > canvas_w) return SMC_EINVAL;
  const bool v4 = w % 4 == 0 && canvas_w % 4 == 0 && x_off % 4 == 0 && ((uintptr_t)img & 15) == 0 && ((uintptr_t)out & 3) == 0;
  long long blocks = smc::ceil_div_ll((long long)n * h * (v4 ? w / 4 : w), 256);
  if (blocks > smc::kNumSMs * 16) blocks = smc::kNumSMs * 16;
  if (v4)
    smc::img_to_uint8_v4_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(img, out, n, h, w / 4, canvas_w, x_off);
  else
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

extern "C" int smc_sgd_step_dev(float* delta, const float* grad, int64_t numel, const float* lr_dev, float grad_scale, float l2_scale, void* stream) {
  if (!delta || !grad || !lr_dev || numel < 1) return SMC_EINVAL;
  long long blocks = smc::ceil_div_ll(numel, 256);
  if (blocks > smc::kNumSMs * 8) blocks = smc::kNumSMs * 8;
  smc::sgd_step_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(delta, grad, numel, 0.f, grad_scale, l2_scale, lr_dev);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
