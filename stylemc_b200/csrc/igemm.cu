// tcgen05 / TMEM / TMA implicit-GEMM for sm_100a: the dense contraction behind modulated_conv2d
// (forward, dgrad) and the CLIP ViT linears.  Replaces the cuDNN/cuBLAS calls the reference reaches
// through torch_utils/ops/conv2d_gradfix.py:35-43 and clip/model.py's nn.Linear / nn.MultiheadAttention.
//
// One CTA = one 128 x BN output tile.  Warp roles (192 threads):
//   warp 0     TMA producer   : per k-step one 4-D box of A (128 pixels x KC channels, shifted by the tap)
//                               and one 2-D box of B (BN rows x KC) into a STAGES-deep smem ring
//   warp 1     MMA issuer     : tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16), fp32 accum in TMEM;
//                               tcgen05.commit releases smem stages and finally signals the epilogue
//   warps 2-5  epilogue       : tcgen05.ld 32x32b (one TMEM lane = one output pixel per thread),
//                               demod * noise + bias -> lrelu*gain -> clamp -> style of next layer,
//                               fp16 hi (+lo) / fp32 stores
// Two CTAs fit per SM (<= 3 x 32 KB stages, <= 256 TMEM columns each), so one tile's epilogue overlaps
// the other's main loop without a persistent scheduler.
#include "tc.cuh"

namespace smc {

struct IgParams {
  int n_img, H, W, C, n_out, ntaps;
  int tw, th, tn, tiles_w, tiles_h;
  int chunk_iters;   // ACC mode: smem stages per accumulation chunk
  smc_igemm_tap taps[SMC_IGEMM_MAX_TAPS];
  smc_igemm_epilogue epi;
};

// ---- kernel ---------------------------------------------------------------------------------------
// ACC = true ("promoted accumulation"): the tensor core adds into its fp32 accumulator with truncation, a bias toward zero that
// grows with the length of the accumulation chain (measured on B200: 2.3e-5 relative after 864 K=16 steps, 2.7e-6 after 96).
// In this mode the K loop is cut into chunks that alternate between two TMEM accumulators; the epilogue warps drain each
// finished chunk into fp32 registers (round-to-nearest adds) while the next chunk is being accumulated.
template <int BN, int KC, int STAGES, bool ACC>
__global__ void __launch_bounds__(192, ACC ? 1 : 2) igemm_kernel(const __grid_constant__ CUtensorMap mapA,
                                                    const __grid_constant__ CUtensorMap mapB,
                                                    const __grid_constant__ IgParams p) {
  constexpr int A_BYTES = 128 * KC * 2;
  constexpr int B_BYTES = BN * KC * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int SWZ = KC * 2;
  constexpr uint32_t TMEM_COLS = (ACC ? 2 : 1) * (BN < 32 ? 32 : BN);
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=f16 [7,10)=0, B=f16 [10,13)=0,
  // A,B K-major [15],[16]=0, N>>3 at [17,23), M>>4 at [24,29)
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;          // [2] accumulator full
  uint64_t* drained_bar = accum_bar + 2;               // [2] accumulator drained by the epilogue warps (ACC mode)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(drained_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_tiles_n = p.n_out / BN;
  const int nt = blockIdx.x % n_tiles_n;
  const int mt = blockIdx.x / n_tiles_n;
  const int w0 = (mt % p.tiles_w) * p.tw;
  const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.th;
  const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.tn;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar[0], 1);
    mbar_init(&accum_bar[1], 1);
    mbar_init(&drained_bar[0], 128);
    mbar_init(&drained_bar[1], 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kchunks = p.C / KC;
  const int total = p.ntaps * kchunks;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < total; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        const int t = it / kchunks;
        const int kc = (it - t * kchunks) * KC;
        const smc_igemm_tap tap = p.taps[t];
        uint8_t* sa = smem + s * STAGE_BYTES;
        tma_load_4d(sa, &mapA, &full_bar[s], kc, w0 + tap.dx, h0 + tap.dy, n0 + tap.dn);
        tma_load_2d(sa + A_BYTES, &mapB, &full_bar[s], kc, tap.brow + nt * BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const int chunk_iters = ACC ? p.chunk_iters : total;
      int in_chunk = 0, chunk = 0;
      for (int it = 0; it < total; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        const int buf = ACC ? (chunk & 1) : 0;
        if (ACC && in_chunk == 0 && chunk >= 2) {   // the epilogue must have drained this accumulator (chunk - 2)
          mbar_wait(&drained_bar[buf], (uint32_t)((chunk >> 1) - 1) & 1u);
          tcgen05_fence_after();
        }
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t da = make_kmajor_desc<SWZ>(a_addr);
        const uint64_t db = make_kmajor_desc<SWZ>(a_addr + A_BYTES);
#pragma unroll
        for (int k = 0; k < KC / 16; ++k)  // advance 16 fp16 = 32 B along K inside the swizzle atom: +2 in 16-B units
          umma_f16(tmem_base + (uint32_t)(buf * BN), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, (uint32_t)((in_chunk | k) != 0));
        tcgen05_commit(&empty_bar[s]);
        if (++in_chunk == chunk_iters || it == total - 1) {
          tcgen05_commit(&accum_bar[buf]);
          in_chunk = 0;
          ++chunk;
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: thread <-> TMEM lane <-> output pixel ----------------
    const int q = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int m = q * 32 + lane;
    const int wl = m % p.tw;
    const int hl = (m / p.tw) % p.th;
    const int nl = m / (p.tw * p.th);
    const int n = n0 + nl, h = h0 + hl, w = w0 + wl;
    const bool valid = (n < p.n_img) && (h < p.H) && (w < p.W);
    const smc_igemm_epilogue& e = p.epi;
    const long long opix = e.o_off + (long long)n * e.o_sn + (long long)h * e.o_sh + (long long)w * e.o_sw;
    float nz = 0.f;
    if (e.noise != nullptr && valid) nz = __ldg(e.noise + (long long)h * e.noise_sh + (long long)w * e.noise_sw);
    const float* rs = e.row_scale ? e.row_scale + (long long)(valid ? n : 0) * p.n_out : nullptr;
    const float* ps = e.post_scale ? e.post_scale + (long long)(valid ? n : 0) * p.n_out : nullptr;
    const float acc_scale = e.acc_scale != 0.f ? e.acc_scale : 1.f;

    float accr[ACC ? BN : 1];
    if (ACC) {
#pragma unroll
      for (int j = 0; j < (ACC ? BN : 1); ++j) accr[j] = 0.f;
      const int nchunks = (total + p.chunk_iters - 1) / p.chunk_iters;
      for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int buf = chunk & 1;
        mbar_wait(&accum_bar[buf], (uint32_t)(chunk >> 1) & 1u);
        tcgen05_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0), r);
#pragma unroll
          for (int j = 0; j < 32; ++j) accr[(ACC ? c0 : 0) + (ACC ? j : 0)] += __uint_as_float(r[j]);
        }
        tcgen05_fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&drained_bar[buf])) : "memory");
      }
    } else {
      mbar_wait(&accum_bar[0], 0);
      tcgen05_fence_after();
    }
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      if (ACC) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(accr[(ACC ? c0 : 0) + (ACC ? j : 0)]);
      } else {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      }
      if (valid) {
        const int o0 = nt * BN + c0;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(r[j]) * acc_scale;
          if (rs) x *= __ldg(rs + o0 + j);
          x += nz;
          if (e.bias) x += __ldg(e.bias + o0 + j);
          if (e.act == 1) x = x > 0.f ? x : x * e.alpha;
          x *= e.gain;
          if (e.clamp >= 0.f) x = fminf(fmaxf(x, -e.clamp), e.clamp);
          v[j] = x;
        }
        if (e.out_raw) {
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out_raw) + opix + o0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __half2 h0_ = __floats2half2_rn(v[8 * j + 0], v[8 * j + 1]), h1_ = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
            __half2 h2_ = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]), h3_ = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
            dst[j] = make_uint4(*reinterpret_cast<uint32_t*>(&h0_), *reinterpret_cast<uint32_t*>(&h1_),
                                *reinterpret_cast<uint32_t*>(&h2_), *reinterpret_cast<uint32_t*>(&h3_));
          }
        }
        if (ps) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= __ldg(ps + o0 + j);
        }
        if (e.residual) {
          const float4* res = reinterpret_cast<const float4*>(e.residual + opix + o0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(res + j);
            v[4 * j + 0] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
          }
        }
        if (e.out_f32) {
          float4* dst = reinterpret_cast<float4*>(e.out_f32 + opix + o0);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (e.out_hi) {
          uint4* dh = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out_hi) + opix + o0);
          uint4* dl = e.out_lo ? reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out_lo) + opix + o0) : nullptr;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float a = v[8 * j + 2 * u], b = v[8 * j + 2 * u + 1];
              const __half2 hh = __floats2half2_rn(a, b);
              const float2 hf = __half22float2(hh);
              const __half2 ll = __floats2half2_rn(a - hf.x, b - hf.y);
              hi[u] = *reinterpret_cast<const uint32_t*>(&hh);
              lo[u] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            dh[j] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (dl) dl[j] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------------
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int BN, int KC, int STAGES, bool ACC>
static int launch_cfg(const CUtensorMap& ma, const CUtensorMap& mb, const IgParams& p, int grid, cudaStream_t st) {
  constexpr int SMEM = STAGES * (128 * KC * 2 + BN * KC * 2) + 1024 + 256;
  static SmemOptIn opt_in;
  if (cudaError_t e = smem_opt_in(opt_in, igemm_kernel<BN, KC, STAGES, ACC>, SMEM); e != cudaSuccess) return (int)e;
  igemm_kernel<BN, KC, STAGES, ACC><<<grid, 192, SMEM, st>>>(ma, mb, p);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

int hconv_try_launch(const smc_igemm_desc* d, cudaStream_t st);   // hconv.cu

int igemm_launch(const smc_igemm_desc* d, cudaStream_t st) {
  if (!d || !d->A || !d->B) return SMC_EINVAL;
  if (d->ntaps < 1 || d->ntaps > SMC_IGEMM_MAX_TAPS) return SMC_EINVAL;
  if (d->n_img < 1 || d->H < 1 || d->W < 1 || d->n_out < 1 || d->C < 1) return SMC_EINVAL;
  if (((uintptr_t)d->A & 15) || ((uintptr_t)d->B & 15)) return SMC_EINVAL;
  if (d->nprob < 0 || d->nprob > 4) return SMC_EINVAL;
  {
    const int r = hconv_try_launch(d, st);   // halo-tile kernel for the spatial convolutions it supports
    if (r != SMC_EUNSUPPORTED) return r;
  }
  if (d->nprob > 1) {
    // problem group not taken as one launch: run its GEMMs one by one (same results, the input is read once per problem)
    int sum = 0;
    for (int q = 0; q < d->nprob; ++q) {
      if (d->prob_ntaps[q] < 1) return SMC_EINVAL;
      sum += d->prob_ntaps[q];
    }
    const int reps = d->ntaps == sum ? 1 : (d->ntaps == 3 * sum ? 3 : 0);
    if (reps == 0) return SMC_EINVAL;
    int t0 = 0;
    for (int q = 0; q < d->nprob; ++q) {
      smc_igemm_desc s = *d;
      s.nprob = 0;
      s.ntaps = reps * d->prob_ntaps[q];
      for (int r = 0; r < reps; ++r)
        for (int i = 0; i < d->prob_ntaps[q]; ++i) s.taps[r * d->prob_ntaps[q] + i] = d->taps[r * sum + t0 + i];
      s.epi.o_off = d->epi.o_off + d->prob_o_off[q];
      const int rc = igemm_launch(&s, st);
      if (rc != SMC_OK) return rc;
      t0 += d->prob_ntaps[q];
    }
    return SMC_OK;
  }
  if (d->C % 32 != 0 || d->lda % 8 != 0 || d->ldb % 8 != 0) return SMC_EUNSUPPORTED;
  if (d->epi.out_raw_lo || d->epi.rgb_acc || d->epi.rgb_w) return SMC_EUNSUPPORTED;   // fused ToRGB lives in hconv.cu only
  if (d->epi.mask_y || d->epi.mask_y_lo || d->epi.mask_grgb) return SMC_EUNSUPPORTED;    // so does the fused activation backward
  if (((uintptr_t)d->A & 15) || ((uintptr_t)d->B & 15)) return SMC_EINVAL;
  const int KC = (d->C % 64 == 0) ? 64 : 32;
  int BN = 0;
  if (d->n_out % 128 == 0) BN = 128;
  else if (d->n_out % 64 == 0) BN = 64;
  else if (d->n_out % 32 == 0) BN = 32;
  else return SMC_EUNSUPPORTED;
  if (d->acc_chunk_k > 0 && BN == 128) BN = 64;   // 64 register accumulators per epilogue thread (128 would spill)

  IgParams p;
  p.n_img = d->n_img; p.H = d->H; p.W = d->W; p.C = d->C; p.n_out = d->n_out; p.ntaps = d->ntaps;
  if (d->tw > 0) {
    p.tw = d->tw; p.th = d->th; p.tn = d->tn;
  } else if (d->H == 1) {
    p.tw = d->W >= 128 ? 128 : pow2_ceil(d->W);
    p.th = 1;
    p.tn = 128 / p.tw;
  } else {
    p.tw = d->W >= 16 ? 16 : pow2_ceil(d->W);
    const int hmax = 128 / p.tw;
    p.th = d->H >= hmax ? hmax : pow2_ceil(d->H);
    p.tn = 128 / (p.tw * p.th);
  }
  if (p.tw * p.th * p.tn != 128 || p.tw > 256 || p.th > 256 || p.tn > 256) return SMC_EINVAL;
  p.tiles_w = ceil_div(d->W, p.tw);
  p.tiles_h = ceil_div(d->H, p.th);
  const long long tiles_m = (long long)p.tiles_w * p.tiles_h * ceil_div(d->n_img, p.tn);
  const long long grid = tiles_m * (d->n_out / BN);
  if (grid > 0x7fffffffLL) return SMC_ETOOLARGE;
  for (int t = 0; t < d->ntaps; ++t) {
    p.taps[t] = d->taps[t];
    if (d->taps[t].brow < 0 || d->taps[t].brow + d->n_out > d->rowsB) return SMC_EINVAL;
  }
  p.epi = d->epi;
  if (!p.epi.out_f32 && !p.epi.out_hi && !p.epi.out_raw) return SMC_EINVAL;
  if ((p.epi.o_sn | p.epi.o_sh | p.epi.o_sw | p.epi.o_off) & 7) return SMC_EUNSUPPORTED;  // 16-B vector stores

  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return SMC_EDRIVER;
  const CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap ma, mb;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->C, (cuuint64_t)d->WA, (cuuint64_t)d->HA, (cuuint64_t)d->NA};
    cuuint64_t strides[3] = {(cuuint64_t)d->lda * 2, (cuuint64_t)d->lda * 2 * d->WA, (cuuint64_t)d->lda * 2 * d->WA * d->HA};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)p.tw, (cuuint32_t)p.th, (cuuint32_t)p.tn};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(d->A), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return SMC_EDRIVER;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->C, (cuuint64_t)d->rowsB};
    cuuint64_t strides[1] = {(cuuint64_t)d->ldb * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->B), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return SMC_EDRIVER;
  }
  const int g = (int)grid;
  if (d->acc_chunk_k > 0) {   // promoted accumulation: chunks of about acc_chunk_k K-elements
    p.chunk_iters = d->acc_chunk_k / KC < 1 ? 1 : d->acc_chunk_k / KC;
    if (KC == 64) {
      if (BN == 64) return launch_cfg<64, 64, 6, true>(ma, mb, p, g, st);
      return launch_cfg<32, 64, 6, true>(ma, mb, p, g, st);
    }
    if (BN == 64) return launch_cfg<64, 32, 6, true>(ma, mb, p, g, st);
    return launch_cfg<32, 32, 6, true>(ma, mb, p, g, st);
  }
  p.chunk_iters = 0;
  if (KC == 64) {
    if (BN == 128) return launch_cfg<128, 64, 3, false>(ma, mb, p, g, st);
    if (BN == 64) return launch_cfg<64, 64, 4, false>(ma, mb, p, g, st);
    return launch_cfg<32, 64, 4, false>(ma, mb, p, g, st);
  }
  if (BN == 128) return launch_cfg<128, 32, 4, false>(ma, mb, p, g, st);
  if (BN == 64) return launch_cfg<64, 32, 4, false>(ma, mb, p, g, st);
  return launch_cfg<32, 32, 4, false>(ma, mb, p, g, st);
}

}  // namespace smc

extern "C" int smc_igemm(const smc_igemm_desc* desc, void* stream) {
  return smc::igemm_launch(desc, reinterpret_cast<cudaStream_t>(stream));
}
