// Halo-tile implicit GEMM for the 3x3 (and polyphase) modulated convolutions, sm_100a: tcgen05 + TMEM + TMA.
//
// Same contract as igemm.cu (smc_igemm_desc: D[m, o] = sum_t sum_k A_t[m, k] * B_t[o, k]) and the same reference call sites
// (torch_utils/ops/conv2d_gradfix.py:35-43 via conv2d_resample.py:125-147), but built around on-chip reuse of the
// activation tile: igemm.cu re-loads the A tile once per filter tap and per hi/lo term (27 loads of the same pixels for a
// split-precision 3x3 conv), which makes it L2->SM bound at ~25% of the tensor pipe.  Here
//
//   * one CTA owns 256 consecutive "positions" of a halo-padded row-major window of one image (pitch Wp = Wt + halo):
//     the TMA box [RB rows][Wp cols][64 channels] (128-byte swizzled rows) is loaded ONCE per 64-channel slab and every
//     tap is the same smem tile read through a UMMA descriptor whose start address is shifted by (dy * Wp + dx) rows
//     (matrix-descriptor base_offset = (addr >> 7) & 7 keeps the swizzle phase right); positions that fall in the halo
//     columns are computed and discarded (Wt / Wp efficiency);
//   * the weights of one tap (BN x 64, hi or lo plane) are a small ring stage; a B_hi stage is used by the A_hi and the
//     A_lo term before it is released, so the split costs 2 B loads per tap, not 3;
//   * two 128-row accumulator blocks share every B stage (M = 256 per CTA), halving weight traffic per FLOP again;
//   * persistent CTAs (grid = #SMs) walk the tile list; TMEM holds two accumulator sets so a chunk can be drained into
//     fp32 registers by the 16 epilogue warps ("promoted accumulation", see igemm.cu) while the next chunk accumulates,
//     and the stores of tile i overlap the MMAs of tile i+1.
//
// Warp roles (608 threads): warp 0 = A-tile TMA producer, warp 1 = weight-stage TMA producer, warp 2 = MMA issuer
// (one elected thread, tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16), warps 3..18 = epilogue
// (TMEM lane quadrant = warp % 4; block = bit 0, column half = bit 1 of the per-quadrant index).
#include "tc.cuh"

namespace smc {

constexpr int HC_MB = 2;            // accumulator blocks of 128 positions per CTA
constexpr int HC_MT = 128 * HC_MB;  // positions per tile
constexpr int HC_MAX_GROUPS = 4;
constexpr int HC_MAX_TAPS = 16;
constexpr int HC_THREADS = 32 * (3 + 16);
constexpr int HC_EPI_THREADS = 32 * 16;
constexpr int HC_BAR_BYTES = 1024;                 // mbarriers + TMEM slot
#ifndef HC_STAGE_MIN
#define HC_STAGE_MIN 32            // smallest N tile whose epilogue parameter vectors are staged in shared memory (hconv_kernel::STAGE)
#endif
constexpr int HC_PSTAGE_VECS = 6;                  // staged epilogue vectors per tile (hconv_kernel::pstage)
static constexpr int hc_pstage_bytes(int bn) { return 2 * HC_PSTAGE_VECS * bn * 4; }

struct HcTap {
  int32_t posoff;            // (dy + padT) * Wp + dx + padL
  int32_t brow_hi, brow_lo;  // first weight row of this tap in the hi / lo plane
};
struct HcGroup {             // one A source (image-axis offset of its hi and lo plane) and the taps that read it
  int32_t dn_hi, dn_lo;
  int32_t tap_begin, tap_end;
};
// A slab (KC channels) is processed as a list of segments = runs of taps of one group.  x3: a segment is a B_hi pass
// (A_hi*B_hi -> main, A_lo*B_hi -> cross) followed by a B_lo pass (A_hi*B_lo -> cross); the main accumulator may be committed
// for draining after a segment's B_hi pass, so that a main accumulation chain is at most `max_chain` MMAs long and is drained
// while the segment's B_lo pass runs.
constexpr int HC_MAX_SEGS = 16;
enum : int32_t { HC_SEG_FIRST = 1, HC_SEG_LAST = 2, HC_SEG_COMMIT = 4, HC_SEG_SLABEND = 8 };
struct HcSeg {
  int32_t g, tb, te, flags;   // group, tap range, FIRST / LAST segment of its group, COMMIT main after it, last segment of the slab
};
// A launch may hold up to HC_MAX_PROBS "problems" that share the A operand and the iteration space but have their own taps and
// output plane (the four output parities of a stride-2 transposed conv, gemm.up2_parity_taps).  A CTA walks the problems of one
// position tile back to back, so the A tile is fetched from DRAM once and served from L2 for the other problems; separate launches
// read the whole input once each.
constexpr int HC_MAX_PROBS = 4;
// n / d for n < 2^31 as one multiply-high, one add and one shift (round-up method): the tile decode runs in the single-thread
// roles between two tiles, and a hardware-less integer division is ~25 dependent instructions there.
struct HcDiv {
  uint32_t mul, shift;
};
static HcDiv hc_make_div(uint32_t d) {
  HcDiv f{0u, 0u};
  if (d <= 1) return f;
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.shift = s;
  f.mul = (uint32_t)((((1ull << 32) * ((1ull << s) - d)) / d) + 1ull);
  return f;
}
__device__ __forceinline__ int hc_div(int n, HcDiv f) { return (int)((__umulhi((uint32_t)n, f.mul) + (uint32_t)n) >> f.shift); }
struct HcProb {
  int32_t seg_begin, nsegs;     // its segments in HcParams::segs (per slab)
  int32_t kcs_per_drain, ndrains;   // slabs per main-accumulator chunk; main-accumulator drains per tile
  int32_t stage_base;           // first resident weight stage (b_resident)
  int32_t pad_;
  int64_t o_off;                // element offset of its output plane, added to epi.o_off
};
struct HcParams {
  int n_img, H, W, C, n_out;
  int Wt, Wp, RB, padL, padT;
  int col_tiles, tiles_per_col, n_tiles_n;
  int total_tiles;              // nprob * super_tiles
  int nprob, super_tiles;
  int a_lo_term;                // split precision: 1 = A_lo * B_hi is computed (three terms), 0 = the A operand has a hi plane only (two terms)
  int pair;                     // CTA-pair launch (cta_group::2): super tiles are pairs of images (2 q + cluster rank)
  HcDiv div_ntn, div_tpc, div_ct, div_wp;   // n_tiles_n, tiles_per_col, col_tiles, Wp
  int ngroups, kchunks, nb;
  HcProb probs[HC_MAX_PROBS];
  int na_hi, na_lo;  // A buffers: a ring for the hi tiles (read by both passes) and one for the lo tiles (released after the
                     // B_hi pass); x1 uses the hi ring only.  Small-C layers get deeper rings: their tiles are short.
  int b_resident;    // every weight stage of a tile has its own smem slot and is loaded once per CTA (n_tiles_n == 1)
  int a_share;       // problem group with a single slab: the A tile loaded for problem 0 stays in shared memory for the other problems
  uint32_t a_box_bytes, a_buf_bytes;
  HcGroup groups[HC_MAX_GROUPS];
  HcSeg segs[HC_MAX_SEGS];
  HcTap taps[HC_MAX_TAPS];
  smc_igemm_epilogue epi;
};

// K-major operand tile with rows of KC * 2 bytes (128: SWIZZLE_128B, 64: SWIZZLE_64B), 8-row groups contiguous.  The start
// address may sit on ANY row of the tile: measured on B200, the swizzle XOR is taken from the absolute shared-memory address
// bits (what TMA wrote), so a shifted start needs no matrix-descriptor base_offset (setting it per the PTX formula gives
// wrong results; tools/hconv_probe.py, profiles/r01b_hconv.md).  The upper descriptor word is a constant; the lower word is
// the start address in 16-byte units, so stepping along K or to the second 128-row block is one 32-bit add.
template <int KC>
struct HcDesc {
  static constexpr uint32_t HI = (uint32_t)((KC * 2 * 8) >> 4)      // stride byte offset (8 rows), bits [32,46)
                                 | (1u << 14)                        // descriptor version (Blackwell), bit 46
                                 | ((KC == 64 ? 2u : 4u) << 29);     // SWIZZLE_128B / SWIZZLE_64B, bits [61,64)
  static __device__ __forceinline__ uint64_t make(uint32_t lo16) { return ((uint64_t)HI << 32) | (uint64_t)lo16; }
};
// D[tmem] += A * B (accumulate always)
__device__ __forceinline__ void umma_f16_acc(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.u32 p, 1, 1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc) : "memory");
}
// exactly one lane of a converged warp: lets ptxas keep the MMA operands in uniform registers (no per-lane R2UR loop)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}" : "+r"(pred));
  return pred != 0;
}
// one tap of one pass for both 128-row blocks: 2 * KC/16 MMAs; `first` = 0 makes the first MMA of each block overwrite
// (BLK = TMEM column stride between the two blocks; the MMA's N is in idesc)
template <int BLK, int KC, bool PAIR = false>
__device__ __forceinline__ void hc_issue_tap(uint32_t tm, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t accumulate_first) {
  constexpr uint32_t MBOFF = (uint32_t)(128 * KC * 2) >> 4;
#pragma unroll
  for (int mb = 0; mb < HC_MB; ++mb) {
    if (PAIR) {
      umma_f16_2sm(tm + (uint32_t)(mb * BLK), HcDesc<KC>::make(alo + mb * MBOFF), HcDesc<KC>::make(blo), idesc, accumulate_first);
#pragma unroll
      for (int k = 1; k < KC / 16; ++k)
        umma_f16_2sm(tm + (uint32_t)(mb * BLK), HcDesc<KC>::make(alo + mb * MBOFF + 2 * k), HcDesc<KC>::make(blo + 2 * k), idesc, 1u);
    } else {
      umma_f16(tm + (uint32_t)(mb * BLK), HcDesc<KC>::make(alo + mb * MBOFF), HcDesc<KC>::make(blo), idesc, accumulate_first);
#pragma unroll
      for (int k = 1; k < KC / 16; ++k)
        umma_f16_acc(tm + (uint32_t)(mb * BLK), HcDesc<KC>::make(alo + mb * MBOFF + 2 * k), HcDesc<KC>::make(blo + 2 * k), idesc);
    }
  }
}

template <bool PAIR>
__device__ __forceinline__ void hc_commit(uint64_t* bar) {
  if (PAIR) tcgen05_commit_2sm(bar);
  else tcgen05_commit(bar);
}

struct HcTile {
  int prob, nt, n, w0, q0, hfirst;
};
// Tile walk of one CTA: (position, N) tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the nprob problems of a tile are consecutive
// iterations of the SAME CTA (balanced work per CTA although the problems differ in size; the A tile of problems 1.. hits in L2 or
// stays in shared memory).  Counters only: the single-thread roles run this between two tiles, on the MMA issue path (a version
// with two more integer divisions per tile slowed the 1024-px conv1 down by 12 %).
struct HcWalk {
  int sup, prob;
};
// a CTA pair walks the list together: cluster id and cluster count take the place of block id and grid size (compile-time switch: the
// walk runs on the MMA issue path between two tiles)
template <bool PAIR>
__device__ __forceinline__ HcWalk hc_walk_begin() { return HcWalk{(int)(PAIR ? blockIdx.x >> 1 : blockIdx.x), 0}; }
template <bool PAIR>
__device__ __forceinline__ void hc_walk_next(const HcParams& p, HcWalk& w) {
  if (++w.prob == p.nprob) { w.prob = 0; w.sup += (int)(PAIR ? gridDim.x >> 1 : gridDim.x); }
}
template <bool PAIR>
__device__ __forceinline__ HcTile hc_tile(const HcParams& p, const HcWalk& w) {
  HcTile r;
  int t = w.sup;
  r.prob = w.prob;
  int q = hc_div(t, p.div_ntn);
  r.nt = t - q * p.n_tiles_n;
  t = q;
  q = hc_div(t, p.div_tpc);
  const int ti = t - q * p.tiles_per_col;
  t = q;
  q = hc_div(t, p.div_ct);
  const int ct = t - q * p.col_tiles;
  r.n = PAIR ? 2 * q + (int)(blockIdx.x & 1u) : q;       // pair: the even CTA takes image 2q, the odd one 2q + 1 (same tile geometry)
  r.w0 = ct * p.Wt;
  r.q0 = ti * HC_MT;
  r.hfirst = hc_div(r.q0, p.div_wp);
  return r;
}

// Accumulators in TMEM (columns), per set: x1: {block 0, block 1}, the two sets alternate per accumulation chunk.
// x3: {main 0, main 1, cross 0, cross 1}: "main" receives only the hi*hi products and is drained every chunk (short chains:
// the tensor core adds with truncation, igemm.cu); "cross" receives hi*lo and lo*hi, which are 2^-11 of the result, so its
// chain may run over the whole tile.  main is drained while the B_lo pass (cross only) runs.  With BN <= 64 two sets fit and
// alternate per tile; with BN = 128 there is one set and cross is drained while the next tile's first main MMAs wait.
// fp32 x8 -> fp16 hi (and lo = rn(v - hi)) 16-byte vectors
__device__ __forceinline__ void hc_split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const __half2 hh = __floats2half2_rn(v[2 * u], v[2 * u + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(v[2 * u] - hf.x, v[2 * u + 1] - hf.y);
    h[u] = *reinterpret_cast<const uint32_t*>(&hh);
    l[u] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void hc_ld8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// 32-byte store (sm_100 256-bit vector store): one full sector per thread and instruction instead of two half-sector writes
__device__ __forceinline__ void hc_st32(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y),
               "r"(b.z), "r"(b.w) : "memory");
}

// 8 consecutive floats of a per-tile parameter vector staged in shared memory (32-bit shared address): every lane of a warp reads the
// same address (one broadcast wavefront), latency ~25 clk instead of a global-load round trip per 8 channels
__device__ __forceinline__ void hc_lds8(uint32_t saddr, float (&f)[8]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(saddr));
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "r"(saddr + 16u));
}

// Per-channel parameter vector of the fused epilogues: staged in shared memory (N tiles of 64 channels and more, where a thread reads ~100
// vectors per tile) or read from global memory as before (32-channel tiles: their tiles last ~5k clocks and a thread owns 16 channels, so
// the staging pass and its barrier cost more than they save -- measured +9..17 % on the 1024-px convs).
template <bool STAGE>
struct HcVec;
template <>
struct HcVec<true> {
  uint32_t a;                                     // shared address of this thread's first channel, 0 = absent
  __device__ __forceinline__ bool ok() const { return a != 0u; }
  __device__ __forceinline__ void ld8(int off, float (&f)[8]) const { hc_lds8(a + 4u * (uint32_t)off, f); }
  __device__ __forceinline__ HcVec row(int k, uint32_t stride_bytes, int) const { return HcVec{a + (uint32_t)k * stride_bytes}; }
};
template <>
struct HcVec<false> {
  const float* p;
  __device__ __forceinline__ bool ok() const { return p != nullptr; }
  __device__ __forceinline__ void ld8(int off, float (&f)[8]) const { hc_ld8(p + off, f); }
  __device__ __forceinline__ HcVec row(int k, uint32_t, int n_out) const { return HcVec{p + (long long)k * n_out}; }
};

// The modulated-conv epilogue (demodulation, noise, bias, leaky ReLU * gain, clamp all present) with the instruction count that
// matters when a thread owns 64 channels of four output planes: vector parameter loads, one FMA for demod + noise + bias,
// lrelu(x) * g = max(x g, x g alpha), the hi/lo splits and 32-byte stores; the ToRGB partial sums ride along.
// rs / bs / ps / rw: this thread's slice of the vectors (staged: rs already multiplied by acc_scale, rs_scale = 1; ps / rw may be absent; the
// three rgb rows are rw_stride bytes (staged) or n_out floats (global) apart).
template <int CW, bool STAGE>
__device__ __forceinline__ void hc_epilogue_modconv(const float (&acc)[CW], const smc_igemm_epilogue& e, float nz, float rs_scale, HcVec<STAGE> rs,
                                                    HcVec<STAGE> bs, HcVec<STAGE> ps, HcVec<STAGE> rw, uint32_t rw_stride, int n_out, long long opix,
                                                    float& rgb0, float& rgb1, float& rgb2) {
  const float g = e.gain, ga = e.gain * e.alpha, cl = e.clamp;
  static_assert(CW % 16 == 0, "16 channels (32 bytes of fp16) per step");
#pragma unroll
  for (int c0 = 0; c0 < CW; c0 += 16) {
    float v[16];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float r8[8], b8[8];
      rs.ld8(c0 + 8 * hh, r8);
      bs.ld8(c0 + 8 * hh, b8);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = fmaf(acc[c0 + 8 * hh + i], STAGE ? r8[i] : r8[i] * rs_scale, nz + b8[i]);
        v[8 * hh + i] = fminf(fmaxf(fmaxf(x * g, x * ga), -cl), cl);
      }
    }
    if (e.out_raw) {
      uint4 h0, l0, h1, l1;
      hc_split8(reinterpret_cast<const float(&)[8]>(v[0]), h0, l0);
      hc_split8(reinterpret_cast<const float(&)[8]>(v[8]), h1, l1);
      hc_st32(reinterpret_cast<__half*>(e.out_raw) + opix + c0, h0, h1);
      if (e.out_raw_lo) hc_st32(reinterpret_cast<__half*>(e.out_raw_lo) + opix + c0, l0, l1);
    }
    if (rw.ok()) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float w0[8], w1[8], w2[8];
        rw.ld8(c0 + 8 * hh, w0);
        rw.row(1, rw_stride, n_out).ld8(c0 + 8 * hh, w1);
        rw.row(2, rw_stride, n_out).ld8(c0 + 8 * hh, w2);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          rgb0 = fmaf(w0[i], v[8 * hh + i], rgb0);
          rgb1 = fmaf(w1[i], v[8 * hh + i], rgb1);
          rgb2 = fmaf(w2[i], v[8 * hh + i], rgb2);
        }
      }
    }
    if (e.out_hi) {
      if (ps.ok()) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float p8[8];
          ps.ld8(c0 + 8 * hh, p8);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[8 * hh + i] *= p8[i];
        }
      }
      uint4 h0, l0, h1, l1;
      hc_split8(reinterpret_cast<const float(&)[8]>(v[0]), h0, l0);
      hc_split8(reinterpret_cast<const float(&)[8]>(v[8]), h1, l1);
      hc_st32(reinterpret_cast<__half*>(e.out_hi) + opix + c0, h0, h1);
      if (e.out_lo) hc_st32(reinterpret_cast<__half*>(e.out_lo) + opix + c0, l0, l1);
    }
  }
}

// 32-byte read-only load (sm_100 256-bit vector load): the whole sector in one request
__device__ __forceinline__ void hc_ld32(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void hc_h8_to_f(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}

// Activation backward fused into a dgrad GEMM (smc_igemm_epilogue::mask_y): the accumulator is the gradient w.r.t. the modulated input
// of the consumer conv; multiplied by post = (consumer style) * (demodulation of the layer below) it becomes the gradient w.r.t. that
// layer's pre-activation once the leaky-ReLU slope and the clamp mask of the SAVED output y are applied (bias_act.cu:71-72,136-142).
// The optional ToRGB branch adds sum_j rw[j][c] * g_j (g = masked dL/drgb of this pixel) before the slope.
template <int CW, bool STAGE>
__device__ __forceinline__ void hc_epilogue_actbwd(const float (&acc)[CW], const smc_igemm_epilogue& e, float acc_scale, HcVec<STAGE> ps, HcVec<STAGE> rw,
                                                   uint32_t rw_stride, int n_out, long long opix, float g0, float g1, float g2) {
  const float g = e.gain, ga = e.gain * e.alpha, cl = e.clamp;
  const __half* yh = reinterpret_cast<const __half*>(e.mask_y) + opix;
  const __half* yl = e.mask_y_lo ? reinterpret_cast<const __half*>(e.mask_y_lo) + opix : nullptr;
#pragma unroll
  for (int c0 = 0; c0 < CW; c0 += 16) {
    float y[16], v[16];
    {
      uint4 a, b;
      hc_ld32(yh + c0, a, b);
      hc_h8_to_f(a, y); hc_h8_to_f(b, y + 8);
      if (yl) {
        float l[16];
        hc_ld32(yl + c0, a, b);
        hc_h8_to_f(a, l); hc_h8_to_f(b, l + 8);
#pragma unroll
        for (int i = 0; i < 16; ++i) y[i] += l[i];
      }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      float p8[8];
      ps.ld8(c0 + 8 * hh, p8);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[8 * hh + i] = acc[c0 + 8 * hh + i] * (p8[i] * acc_scale);
      if (rw.ok()) {
        float w0[8], w1[8], w2[8];
        rw.ld8(c0 + 8 * hh, w0);
        rw.row(1, rw_stride, n_out).ld8(c0 + 8 * hh, w1);
        rw.row(2, rw_stride, n_out).ld8(c0 + 8 * hh, w2);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[8 * hh + i] += w0[i] * g0 + w1[i] * g1 + w2[i] * g2;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float yy = y[i];
      const bool pass = (cl < 0.f) || (yy > -cl && yy < cl);
      v[i] = pass ? v[i] * (yy > 0.f ? g : ga) : 0.f;
    }
    uint4 h0, l0, h1, l1;
    hc_split8(reinterpret_cast<const float(&)[8]>(v[0]), h0, l0);
    hc_split8(reinterpret_cast<const float(&)[8]>(v[8]), h1, l1);
    hc_st32(reinterpret_cast<__half*>(e.out_hi) + opix + c0, h0, h1);
    if (e.out_lo) hc_st32(reinterpret_cast<__half*>(e.out_lo) + opix + c0, l0, l1);
  }
}

// MODE 2 (x3, BN <= 64): with both operands in shared memory an MMA costs ~64 clk for its 128 x 16 A slab whatever N is, so narrow
// layers are issue-bound on the NUMBER of MMAs.  There the stage of a tap holds [B_hi; B_lo] (2 BN rows) and ONE MMA of N = 2 BN
// computes A_hi*B_hi (columns [0, BN) = main) and A_hi*B_lo (columns [BN, 2 BN) = cross) together; A_lo*B_hi is a second MMA of
// N = BN into the cross columns: 2 MMAs per tap and k-step instead of 3.  Both halves are drained per chunk (the first MMA of a
// chunk overwrites all 2 BN columns); two sets alternate per chunk.
enum { HC_X1 = 0, HC_X3_TWO_PASS = 1, HC_X3_MERGED = 2 };
// ALO: the A operand has a lo plane (three-term split); false = two-term split (hi plane only).  A template parameter, not a field of
// HcParams: the flag sits inside the MMA issue loop, and a run-time branch there cost the 32-channel layers 10-15 %.
// EPI: 0 = every fused epilogue; 1 = the plain one only (fp32 output = acc * acc_scale [* row_scale] [+ bias] [+ residual]: CLIP linears, conv0's
// parity planes, fp32 dgrad outputs).  A separate instantiation because the 128-wide epilogue threads (64 accumulators each) sit at the register
// limit: every path compiled into the kernel costs all of them spills (adding the lean path to EPI 0 made the convolutions 4-7 % slower).
template <int BN, int KC, int MODE, bool PAIR = false, bool ALO = true, int EPI = 0>
__global__ void __launch_bounds__(HC_THREADS, 1) hconv_kernel(const __grid_constant__ CUtensorMap mapA,
                                                              const __grid_constant__ CUtensorMap mapB,
                                                              const __grid_constant__ HcParams p) {
  constexpr bool X3 = MODE != HC_X1;
  constexpr bool TWO_PASS = MODE == HC_X3_TWO_PASS;
  constexpr bool MERGED = MODE == HC_X3_MERGED;
  constexpr int ROWB = KC * 2;                                 // bytes per position row
  constexpr int B_TILE = (PAIR ? BN / 2 : BN) * ROWB;          // one weight tile (hi or lo plane of a tap); a pair holds half each
  constexpr int B_BYTES = (MERGED ? 2 : 1) * B_TILE;           // one ring stage (per CTA)
  constexpr uint32_t B_TX = (PAIR ? 2u : 1u) * (uint32_t)B_BYTES;   // bytes arriving on the (leader's) full barrier of a stage
  static_assert(!PAIR || !MERGED, "merged weight stages put B_hi and B_lo in different halves of N: not split across a CTA pair");
  constexpr int CW = BN / 2;                                   // accumulator columns per epilogue thread
  constexpr int BLK = (MERGED ? 2 : 1) * BN;                   // TMEM columns of one 128-row block
  constexpr int SETCOLS = (TWO_PASS ? 2 : 1) * HC_MB * BLK;
  constexpr int SETS = (2 * SETCOLS <= 512) ? 2 : 1;
  constexpr uint32_t TMEM_COLS = (SETS * SETCOLS) < 32 ? 32 : (SETS * SETCOLS);
  constexpr bool STAGE = BN >= HC_STAGE_MIN;                             // epilogue parameter vectors staged in shared memory (HcVec)
  constexpr uint32_t MMA_M = PAIR ? 256 : 128;                 // cta_group::2: 128 rows from each CTA of the pair
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(MMA_M >> 4) << 24);
  constexpr uint32_t IDESC2 = (1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(MMA_M >> 4) << 24);
  constexpr int MAXB = 32;
  static_assert(TWO_PASS || SETS == 2, "x1 / merged alternate two sets per chunk");
  static_assert(!MERGED || BN <= 64, "merged B needs N = 2 BN <= 128 to pay off");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* a_buf = smem;                                        // na buffers of a_buf_bytes
  uint8_t* b_buf = smem + (size_t)(p.na_hi + p.na_lo) * p.a_buf_bytes;         // nb stages of B_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + (size_t)p.nb * B_BYTES);
  uint64_t* a_full = bars;            // [8]
  uint64_t* a_empty = bars + 8;       // [8]
  uint64_t* b_full = bars + 16;       // [MAXB]
  uint64_t* b_empty = b_full + MAXB;  // [MAXB]
  uint64_t* main_full = b_empty + MAXB;    // [2] per set
  uint64_t* main_drained = main_full + 2;  // [2]
  uint64_t* cross_full = main_drained + 2; // [2]
  uint64_t* cross_drained = cross_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cross_drained + 2);
  // per-tile parameter vectors of the fused epilogues, staged by the epilogue warps (two buffers, alternating per tile):
  // [row_scale * acc_scale | bias | post_scale | rgb_w row 0 | row 1 | row 2] x BN floats
  float* pstage = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + HC_BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int i = 0; i < 8; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < MAXB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&main_full[i], 1);
      mbar_init(&main_drained[i], PAIR ? 32 : 16);       // the leader's copy collects the epilogue warps of both CTAs
      mbar_init(&cross_full[i], 1);
      mbar_init(&cross_drained[i], PAIR ? 32 : 16);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (PAIR) {          // the same warp of both CTAs allocates; both receive the same column address
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all();          // barriers of BOTH CTAs are initialised before any remote arrival / TMA transaction
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool leader = !PAIR || (blockIdx.x & 1u) == 0;     // cluster dims (2, 1, 1): rank in the pair = blockIdx.x & 1

  // A buffer of (step, plane): hi tiles cycle through buffers [0, na_hi), lo tiles through [na_hi, na_hi + na_lo)
  const uint32_t na_hi = (uint32_t)p.na_hi, na_lo = (uint32_t)p.na_lo;
  auto hi_buf = [&](uint32_t step) -> int { return (int)(step % na_hi); };
  auto lo_buf = [&](uint32_t step) -> int { return X3 ? (int)(na_hi + step % na_lo) : 0; };

  if (warp == 0) {
    // ---------------- A producer: one halo tile per (slab, group, plane) ----------------
    if (elect_one()) {
      uint32_t eph = 0;                 // bit b: number of loads into buffer b so far, mod 2
      uint32_t step = 0;
      for (HcWalk wk = hc_walk_begin<PAIR>(); wk.sup < p.super_tiles; hc_walk_next<PAIR>(p, wk)) {
        if (p.a_share && wk.prob > 0) continue;            // same position tile as problem 0: its buffers are reused
        const HcTile tl = hc_tile<PAIR>(p, wk);
        const int wbox = tl.w0 - p.padL, hbox = tl.hfirst - p.padT;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int g = 0; g < p.ngroups; ++g, ++step) {
            // pair: each CTA loads the tile of its own image into its own buffer; both transactions land on the LEADER's full barrier
            const int hb = hi_buf(step);
            mbar_wait(&a_empty[hb], ((eph >> hb) & 1u) ^ 1u);
            eph ^= 1u << hb;
            if (leader) mbar_expect_tx(&a_full[hb], (PAIR ? 2u : 1u) * p.a_box_bytes);
            if (PAIR) tma_load_4d_2sm(a_buf + (size_t)hb * p.a_buf_bytes, &mapA, &a_full[hb], kc * KC, wbox, hbox, tl.n + p.groups[g].dn_hi);
            else tma_load_4d(a_buf + (size_t)hb * p.a_buf_bytes, &mapA, &a_full[hb], kc * KC, wbox, hbox, tl.n + p.groups[g].dn_hi);
            if (X3 && ALO) {
              const int lb = lo_buf(step);
              mbar_wait(&a_empty[lb], ((eph >> lb) & 1u) ^ 1u);
              eph ^= 1u << lb;
              if (leader) mbar_expect_tx(&a_full[lb], (PAIR ? 2u : 1u) * p.a_box_bytes);
              if (PAIR) tma_load_4d_2sm(a_buf + (size_t)lb * p.a_buf_bytes, &mapA, &a_full[lb], kc * KC, wbox, hbox, tl.n + p.groups[g].dn_lo);
              else tma_load_4d(a_buf + (size_t)lb * p.a_buf_bytes, &mapA, &a_full[lb], kc * KC, wbox, hbox, tl.n + p.groups[g].dn_lo);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- B producer: one weight stage per (slab, group, pass, tap) ----------------
    if (elect_one()) {
      uint32_t bs = 0, bph = 0;
      const uint32_t nb = (uint32_t)p.nb;
      // resident weights: the stages of ALL problems are loaded once, problem-major (slot = HcProb::stage_base + running index)
      for (HcWalk wk = hc_walk_begin<PAIR>(); wk.sup < p.super_tiles; hc_walk_next<PAIR>(p, wk)) {
        if (p.b_resident && wk.sup != (int)blockIdx.x) break;   // (never combined with pair launches)       // resident: one pass over the problems (n_tiles_n == 1)
        const int nt = wk.sup - hc_div(wk.sup, p.div_ntn) * p.n_tiles_n;
        const HcProb& pr = p.probs[wk.prob];
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int si = pr.seg_begin; si < pr.seg_begin + pr.nsegs; ++si) {
            const int tb = p.segs[si].tb, te = p.segs[si].te;
            for (int pass = 0; pass < (TWO_PASS ? 2 : 1); ++pass) {
              for (int tp = tb; tp < te; ++tp) {
                mbar_wait(&b_empty[bs], bph ^ 1u);
                if (leader) mbar_expect_tx(&b_full[bs], B_TX);
                if (PAIR)       // each CTA fetches its half of the tile's BN weight rows
                  tma_load_2d_2sm(b_buf + (size_t)bs * B_BYTES, &mapB, &b_full[bs], kc * KC,
                                  (pass == 0 ? p.taps[tp].brow_hi : p.taps[tp].brow_lo) + nt * BN + (int)(blockIdx.x & 1u) * (BN / 2));
                else
                tma_load_2d(b_buf + (size_t)bs * B_BYTES, &mapB, &b_full[bs], kc * KC,
                            (pass == 0 ? p.taps[tp].brow_hi : p.taps[tp].brow_lo) + nt * BN);
                if (MERGED)
                  tma_load_2d(b_buf + (size_t)bs * B_BYTES + B_TILE, &mapB, &b_full[bs], kc * KC, p.taps[tp].brow_lo + nt * BN);
                if (++bs == nb) { bs = 0; bph ^= 1u; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ---------------- MMA issuer: ONE elected lane runs the whole role.  elect.sync (rather than lane == 0) tells ptxas
    // that a single thread is active, so the tcgen05 operands go to uniform registers with one R2UR each instead of a
    // per-lane waterfall loop (15 instructions per MMA), and the issue loop stays far below the MMA duration.
    if (leader && elect_one()) {
      uint32_t aph = 0;                 // bit b: number of tiles consumed from A buffer b so far, mod 2
      uint32_t step = 0, bs = 0, bph = 0, chunk_ctr = 0, tile_ctr = 0;
      uint32_t nm0 = 0, nm1 = 0, nc0 = 0, nc1 = 0;     // chunks committed so far per set (main / cross)
      const uint32_t nb = (uint32_t)p.nb;
      const uint32_t a_base = smem_u32(a_buf), b_base = smem_u32(b_buf);
      const bool resident = p.b_resident != 0;
      constexpr bool alo = X3 && ALO;                      // the A operand has a lo plane (A_lo * B_hi is computed)
      uint32_t a_hi = 0, a_lo = 0;                         // live across the problems of a tile when the A tile is shared
      int hb = 0, lb = 0;
      for (HcWalk wk = hc_walk_begin<PAIR>(); wk.sup < p.super_tiles; hc_walk_next<PAIR>(p, wk), ++tile_ctr) {
        const HcTile tl = hc_tile<PAIR>(p, wk);
        const HcProb& pr = p.probs[tl.prob];
        const bool a_load = !(p.a_share && tl.prob > 0);              // this tile waits for its own A buffers
        const bool a_free = !(p.a_share && tl.prob + 1 < p.nprob);    // and releases them
        if (resident) bs = (uint32_t)pr.stage_base;
        const int rel0 = tl.q0 - tl.hfirst * p.Wp;         // position of tile row 0 inside the box (before the tap offset)
        const uint32_t set_t = (TWO_PASS && SETS == 2) ? (tile_ctr & 1u) : 0u;
        int in_chunk = 0;
        bool fresh = true;                                 // next main MMA starts an accumulation chunk (overwrites)
        bool fresh_cross = true;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const bool chunk_ends = (in_chunk + 1 == pr.kcs_per_drain) || (kc == p.kchunks - 1);
          for (int si = pr.seg_begin; si < pr.seg_begin + pr.nsegs; ++si) {
            const int tb = p.segs[si].tb, te = p.segs[si].te, sflags = p.segs[si].flags;
            // main chain complete after this segment's B_hi pass (mid-slab commit, or the slab-end commit of a chunk)
            const bool main_done = (sflags & HC_SEG_COMMIT) && (!(sflags & HC_SEG_SLABEND) || chunk_ends);
            if ((sflags & HC_SEG_FIRST) && a_load) {
              hb = hi_buf(step);
              mbar_wait(&a_full[hb], (aph >> hb) & 1u);
              aph ^= 1u << hb;
              if (alo) {
                lb = lo_buf(step);
                mbar_wait(&a_full[lb], (aph >> lb) & 1u);
                aph ^= 1u << lb;
                a_lo = a_base + (uint32_t)lb * p.a_buf_bytes + (uint32_t)rel0 * (uint32_t)ROWB;
              }
              a_hi = a_base + (uint32_t)hb * p.a_buf_bytes + (uint32_t)rel0 * (uint32_t)ROWB;
            }
            for (int pass = 0; pass < (TWO_PASS ? 2 : 1); ++pass) {
              for (int tp = tb; tp < te; ++tp) {
                const uint32_t set = TWO_PASS ? set_t : (chunk_ctr & 1u);
                if (pass == 0 && fresh) {                  // this set's main accumulator must have been drained
                  const uint32_t nm = set ? nm1 : nm0;
                  if (nm >= 1) mbar_wait(&main_drained[set], (nm - 1u) & 1u);
                }
                if (TWO_PASS && fresh_cross) {
                  const uint32_t nc = set_t ? nc1 : nc0;
                  if (nc >= 1) mbar_wait(&cross_drained[set_t], (nc - 1u) & 1u);
                }
                mbar_wait(&b_full[bs], resident ? 0u : bph);
                tcgen05_fence_after();
                const uint32_t blo = (b_base + bs * (uint32_t)B_BYTES) >> 4;
                const uint32_t roff = (uint32_t)p.taps[tp].posoff * (uint32_t)ROWB;
                const uint32_t tm = tmem_base + set * (uint32_t)SETCOLS;
                if (MERGED) {
                  // A_hi * [B_hi; B_lo] -> (main | cross), then A_lo * B_hi -> cross
                  hc_issue_tap<BLK, KC, PAIR>(tm, (a_hi + roff) >> 4, blo, IDESC2, fresh ? 0u : 1u);
                  if (alo) hc_issue_tap<BLK, KC, PAIR>(tm + (uint32_t)BN, (a_lo + roff) >> 4, blo, IDESC, 1u);
                } else {
                  if (pass == 0) hc_issue_tap<BLK, KC, PAIR>(tm, (a_hi + roff) >> 4, blo, IDESC, fresh ? 0u : 1u);
                  if (TWO_PASS && (pass == 1 || alo)) {    // pass 0: A_lo * B_hi (three-term split only), pass 1: A_hi * B_lo -> cross
                    hc_issue_tap<BLK, KC, PAIR>(tm + (uint32_t)(HC_MB * BN), ((pass == 0 ? a_lo : a_hi) + roff) >> 4, blo, IDESC, fresh_cross ? 0u : 1u);
                    fresh_cross = false;
                  }
                }
                if (!resident) hc_commit<PAIR>(&b_empty[bs]);
                if (pass == 0) fresh = false;
                if (++bs == nb) { bs = 0; bph ^= 1u; }
              }
              if (TWO_PASS && pass == 0) {
                if ((sflags & HC_SEG_LAST) && alo) hc_commit<PAIR>(&a_empty[lb]);   // the lo tile is only read by B_hi passes
                if (main_done) {                                          // drained while the B_lo pass runs
                  hc_commit<PAIR>(&main_full[set_t]);
                  if (set_t) ++nm1; else ++nm0;
                  fresh = true;
                }
              }
            }
            if (!TWO_PASS && main_done) {                  // x1 / merged: the two sets alternate per chunk
              const uint32_t set = chunk_ctr & 1u;
              hc_commit<PAIR>(&main_full[set]);
              if (set) ++nm1; else ++nm0;
              ++chunk_ctr;
              fresh = true;
            }
            if ((sflags & HC_SEG_LAST) && a_free) {
              if (MERGED && alo) hc_commit<PAIR>(&a_empty[lb]);
              hc_commit<PAIR>(&a_empty[hb]);
              ++step;
            }
          }
          if (chunk_ends) in_chunk = 0; else ++in_chunk;
        }
        if (TWO_PASS) {
          hc_commit<PAIR>(&cross_full[set_t]);
          if (set_t) ++nc1; else ++nc0;
        }
      }
    }
  } else {
    // ---------------- epilogue: thread <-> TMEM lane <-> position ----------------
    const int ew = warp - 3;
    const int q = warp & 3;                 // a warp may only touch TMEM lanes [32 * (warp % 4), +32)
    const int j = ew >> 2;                  // 0..3 within the quadrant
    const int mb = j & 1, ch = j >> 1;
    const int m = mb * 128 + q * 32 + lane;
    const smc_igemm_epilogue& e = p.epi;
    const float acc_scale = e.acc_scale != 0.f ? e.acc_scale : 1.f;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * BLK + ch * CW);
    uint32_t em0 = 0, em1 = 0, ec0 = 0, ec1 = 0, chunk_ctr = 0, tile_ctr = 0;
    int st_n = -1, st_nt = -1;               // (image, N tile) whose parameter vectors are staged in pstage[st_buf]
    uint32_t st_buf = 0;
    float acc[CW];
    for (HcWalk wk = hc_walk_begin<PAIR>(); wk.sup < p.super_tiles; hc_walk_next<PAIR>(p, wk), ++tile_ctr) {
      const HcTile tl = hc_tile<PAIR>(p, wk);
      const int ndrains = p.probs[tl.prob].ndrains;
      const uint32_t set_t = (TWO_PASS && SETS == 2) ? (tile_ctr & 1u) : 0u;
      // ---- this thread's position, the lean-path decision and the per-pixel operands.  STAGE (N tiles of 64+ channels): all of it runs
      // here, while the tensor core works on the tile, together with the staging of the per-tile parameter vectors in shared memory (one
      // coalesced pass of the 512 epilogue threads instead of ~100 global loads per thread after the drain) and an L2 prefetch of the saved
      // activation the fused activation backward reads.  32-channel tiles (short, 16 channels per thread) keep the lean order: after the drain.
      int h = 0, w = 0, o0 = 0;
      bool valid = false, modconv = false;
      long long opix = 0;
      float nz = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
      auto locate = [&]() {
        const int qpos = tl.q0 + m;
        h = hc_div(qpos, p.div_wp);
        const int wr = qpos - h * p.Wp;
        w = tl.w0 + wr;
        valid = (wr < p.Wt) && (w < p.W) && (h < p.H);
        o0 = tl.nt * BN + ch * CW;
        opix = e.o_off + p.probs[tl.prob].o_off + (long long)tl.n * e.o_sn + (long long)h * e.o_sh + (long long)w * e.o_sw + o0;
        const bool out32 = ((((uintptr_t)e.out_raw | (uintptr_t)e.out_raw_lo | (uintptr_t)e.out_hi | (uintptr_t)e.out_lo) & 31) == 0) &&
                           ((((e.o_sn | e.o_sh | e.o_sw | e.o_off) * 2) & 31) == 0) && (p.n_out % 16 == 0);
        // modulated-conv layers (the bulk of the epilogue work): lean path; alpha < 1 makes max(x, alpha x) the leaky ReLU
        modconv = e.row_scale && e.bias && e.act == 1 && e.clamp >= 0.f && e.alpha >= 0.f && e.alpha <= 1.f && e.gain > 0.f && !e.residual &&
                  !e.out_f32 && !e.mask_y && out32 &&
                  (STAGE || (((uintptr_t)e.row_scale | (uintptr_t)e.bias | (uintptr_t)e.post_scale | (uintptr_t)e.rgb_w) & 15) == 0);
        if (valid) {
          if (e.noise != nullptr) nz = __ldg(e.noise + (long long)h * e.noise_sh + (long long)w * e.noise_sw);
          if (e.mask_y && e.mask_grgb) {
            const float* gp = e.mask_grgb + (long long)tl.n * e.rgb_sn + (long long)h * e.rgb_sh + w;
            g0 = __ldg(gp); g1 = __ldg(gp + e.rgb_sj); g2 = __ldg(gp + 2 * e.rgb_sj);
          }
        }
      };
      if (STAGE && EPI != 1) {
        locate();
        // the vectors depend on (image, N tile) only: consecutive tiles of a CTA mostly share them (always at 512 / 1024 px), so they are
        // re-staged -- into the other buffer, followed by one barrier of the 16 epilogue warps -- only when that pair changes; every
        // epilogue warp walks the same tile list, so all of them take this branch together
        if ((modconv || e.mask_y != nullptr) && (tl.n != st_n || tl.nt != st_nt)) {
          st_n = tl.n; st_nt = tl.nt; st_buf ^= 1u;
          float* pst = pstage + st_buf * (uint32_t)(HC_PSTAGE_VECS * BN);
          const int et = (int)threadIdx.x - 96;
          const long long nb_off = (long long)tl.n * p.n_out + tl.nt * BN;
          for (int i = et; i < HC_PSTAGE_VECS * BN; i += HC_EPI_THREADS) {
            const int which = i / BN, c = i - which * BN;
            float v = 0.f;
            if (which == 0) { if (e.row_scale) v = __ldg(e.row_scale + nb_off + c) * acc_scale; }
            else if (which == 1) { if (e.bias) v = __ldg(e.bias + tl.nt * BN + c); }
            else if (which == 2) { if (e.post_scale) v = __ldg(e.post_scale + nb_off + c); }
            else if (e.rgb_w) v = __ldg(e.rgb_w + ((long long)tl.n * 3 + (which - 3)) * p.n_out + tl.nt * BN + c);
            pst[i] = v;
          }
          // all 16 warps have left the tiles that used the buffer being replaced next time, and this buffer is visible to all of them
          asm volatile("bar.sync 1, %0;" ::"n"(HC_EPI_THREADS) : "memory");
        }
        if (valid && e.mask_y) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const __half*>(e.mask_y) + opix));
          if (e.mask_y_lo) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const __half*>(e.mask_y_lo) + opix));
        }
      }
#pragma unroll
      for (int i = 0; i < CW; ++i) acc[i] = 0.f;
      for (int dch = 0; dch < ndrains + (TWO_PASS ? 1 : 0); ++dch) {
        const bool cross = TWO_PASS && dch == ndrains;
        const uint32_t set = TWO_PASS ? set_t : (chunk_ctr & 1u);
        uint64_t* full = cross ? &cross_full[set] : &main_full[set];
        uint64_t* drained = cross ? &cross_drained[set] : &main_drained[set];
        uint32_t cnt;
        if (cross) { cnt = set ? ec1 : ec0; if (set) ++ec1; else ++ec0; }
        else { cnt = set ? em1 : em0; if (set) ++em1; else ++em0; }
        if (!TWO_PASS) ++chunk_ctr;
        const uint32_t col = set * (uint32_t)SETCOLS + (cross ? (uint32_t)(HC_MB * BN) : 0u);
        mbar_wait(full, cnt & 1u);
        tcgen05_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < CW; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_addr + col + (uint32_t)c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[c0 + i] += __uint_as_float(r[i]);
          if (MERGED) {                                     // the cross half of the same block
            tmem_ld16(lane_addr + col + (uint32_t)(BN + c0), r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[c0 + i] += __uint_as_float(r[i]);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(drained); else mbar_arrive(drained); }
      }
      // ---- fused epilogue + stores for this thread's position
      if constexpr (EPI == 1) {
        const int qpos = tl.q0 + m;
        const int hh = hc_div(qpos, p.div_wp);
        const int wr = qpos - hh * p.Wp, ww = tl.w0 + wr;
        if ((wr < p.Wt) && (ww < p.W) && (hh < p.H)) {
          const int oc = tl.nt * BN + ch * CW;
          const long long op = e.o_off + p.probs[tl.prob].o_off + (long long)tl.n * e.o_sn + (long long)hh * e.o_sh + (long long)ww * e.o_sw + oc;
          const float* rs = e.row_scale ? e.row_scale + (long long)tl.n * p.n_out + oc : nullptr;
          const float* bs = e.bias ? e.bias + oc : nullptr;
#pragma unroll
          for (int c0 = 0; c0 < CW; c0 += 8) {
            float v[8], t8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = acc[c0 + i] * acc_scale;                   // the operation order of the generic path
            if (rs) {
              hc_ld8(rs + c0, t8);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] *= t8[i];
            }
            if (bs) {
              hc_ld8(bs + c0, t8);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += t8[i];
            }
            if (e.residual) {
              hc_ld8(e.residual + op + c0, t8);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += t8[i];
            }
            hc_st32(e.out_f32 + op + c0, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])),
                    make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
          }
        }
        continue;
      }
      if (!STAGE) locate();
      if (valid) {
        const int n = tl.n;
        const float* rs = e.row_scale ? e.row_scale + (long long)n * p.n_out + o0 : nullptr;
        const float* ps = e.post_scale ? e.post_scale + (long long)n * p.n_out + o0 : nullptr;
        const float* bs = e.bias ? e.bias + o0 : nullptr;
        const float* rw = e.rgb_acc ? e.rgb_w + (long long)n * 3 * p.n_out + o0 : nullptr;
        float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
        const bool f32_aligned32 = (((uintptr_t)e.out_f32) & 31) == 0;     // element offsets are multiples of 8 floats (checked on the host)
        constexpr uint32_t VEC = 4u * (uint32_t)BN;                        // bytes between two staged vectors
        const uint32_t pst_s = STAGE ? smem_u32(pstage + st_buf * (uint32_t)(HC_PSTAGE_VECS * BN)) + 4u * (uint32_t)(ch * CW) : 0u;
        if constexpr (EPI == 2 && STAGE) {            // the modulated-conv epilogue only (the host checked what `modconv` checks)
          hc_epilogue_modconv<CW, true>(acc, e, nz, 1.f, HcVec<true>{pst_s}, HcVec<true>{pst_s + VEC}, HcVec<true>{e.post_scale ? pst_s + 2u * VEC : 0u},
                                        HcVec<true>{e.rgb_acc ? pst_s + 3u * VEC : 0u}, VEC, p.n_out, opix, rgb0, rgb1, rgb2);
        } else if constexpr (EPI == 3 && STAGE) {     // the fused activation backward only
          hc_epilogue_actbwd<CW, true>(acc, e, acc_scale, HcVec<true>{pst_s + 2u * VEC}, HcVec<true>{e.mask_grgb ? pst_s + 3u * VEC : 0u}, VEC,
                                       p.n_out, opix, g0, g1, g2);
        } else
        if (e.mask_y) {
          if constexpr (STAGE) {
            hc_epilogue_actbwd<CW, true>(acc, e, acc_scale, HcVec<true>{pst_s + 2u * VEC}, HcVec<true>{e.mask_grgb ? pst_s + 3u * VEC : 0u}, VEC,
                                         p.n_out, opix, g0, g1, g2);
          } else {
            hc_epilogue_actbwd<CW, false>(acc, e, acc_scale, HcVec<false>{ps},
                                          HcVec<false>{e.mask_grgb ? e.rgb_w + (long long)n * 3 * p.n_out + o0 : nullptr}, 0u, p.n_out, opix, g0, g1, g2);
          }
        } else if (modconv) {
          if constexpr (STAGE) {
            hc_epilogue_modconv<CW, true>(acc, e, nz, 1.f, HcVec<true>{pst_s}, HcVec<true>{pst_s + VEC}, HcVec<true>{e.post_scale ? pst_s + 2u * VEC : 0u},
                                          HcVec<true>{e.rgb_acc ? pst_s + 3u * VEC : 0u}, VEC, p.n_out, opix, rgb0, rgb1, rgb2);
          } else {
            hc_epilogue_modconv<CW, false>(acc, e, nz, acc_scale, HcVec<false>{rs}, HcVec<false>{bs}, HcVec<false>{ps}, HcVec<false>{rw}, 0u, p.n_out, opix,
                                           rgb0, rgb1, rgb2);
          }
        } else if (BN <= 64 && e.out_f32 && f32_aligned32 && !e.out_hi && !e.out_raw && !e.rgb_acc && !ps && !bs && !e.residual && e.noise == nullptr &&
                   e.act == 0 && e.gain == 1.f && e.clamp < 0.f && (((uintptr_t)e.row_scale) & 15) == 0) {
          // (not in the 128-wide kernels: their epilogue threads already spill at 96 registers, and the extra path cost them 4-7 %)
          // conv0's parity planes and the fp32 dgrad outputs: v = acc * acc_scale (* row_scale), one 32-byte store per 8 channels.
          // The generic loop below spends ~10 instructions per element on this (scalar parameter loads, dead predicated branches); with
          // tiles of 16-64 MMAs per problem the 16 epilogue warps, not the tensor pipe, set the pace of these launches.
#pragma unroll
          for (int c0 = 0; c0 < CW; c0 += 8) {
            float r8[8], v[8];
            if (rs) hc_ld8(rs + c0, r8);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = rs ? (acc[c0 + i] * acc_scale) * r8[i] : acc[c0 + i] * acc_scale;     // the generic path's operation order
            hc_st32(e.out_f32 + opix + c0, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])),
                    make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
          }
        } else {
#pragma unroll
        for (int c0 = 0; c0 < CW; c0 += 8) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float x = acc[c0 + i] * acc_scale;
            if (rs) x *= __ldg(rs + c0 + i);
            x += nz;
            if (bs) x += __ldg(bs + c0 + i);
            if (e.act == 1) x = x > 0.f ? x : x * e.alpha;
            x *= e.gain;
            if (e.clamp >= 0.f) x = fminf(fmaxf(x, -e.clamp), e.clamp);
            v[i] = x;
          }
          if (e.out_raw) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const __half2 hh = __floats2half2_rn(v[2 * u], v[2 * u + 1]);
              const float2 hf = __half22float2(hh);
              const __half2 ll = __floats2half2_rn(v[2 * u] - hf.x, v[2 * u + 1] - hf.y);
              hi[u] = *reinterpret_cast<const uint32_t*>(&hh);
              lo[u] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out_raw) + opix + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (e.out_raw_lo)
              *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out_raw_lo) + opix + c0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          if (rw) {                                         // fused ToRGB: partial dot products over this thread's channels
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              rgb0 += __ldg(rw + c0 + i) * v[i];
              rgb1 += __ldg(rw + p.n_out + c0 + i) * v[i];
              rgb2 += __ldg(rw + 2 * p.n_out + c0 + i) * v[i];
            }
          }
          if (ps) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= __ldg(ps + c0 + i);
          }
          if (e.residual) {
            const float4 r0 = __ldg(reinterpret_cast<const float4*>(e.residual + opix + c0));
            const float4 r1 = __ldg(reinterpret_cast<const float4*>(e.residual + opix + c0 + 4));
            v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
            v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
          }
          if (e.out_f32) {
            if (f32_aligned32) {
              hc_st32(e.out_f32 + opix + c0, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])),
                      make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
            } else {
              float4* dst = reinterpret_cast<float4*>(e.out_f32 + opix + c0);
              dst[0] = make_float4(v[0], v[1], v[2], v[3]);
              dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
          }
          if (e.out_hi) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const __half2 hh = __floats2half2_rn(v[2 * u], v[2 * u + 1]);
              const float2 hf = __half22float2(hh);
              const __half2 ll = __floats2half2_rn(v[2 * u] - hf.x, v[2 * u + 1] - hf.y);
              hi[u] = *reinterpret_cast<const uint32_t*>(&hh);
              lo[u] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out_hi) + opix + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (e.out_lo)
              *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(e.out_lo) + opix + c0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        }   // generic epilogue
        if (e.rgb_acc) {
          float* ra = e.rgb_acc + (long long)tl.nt * e.rgb_snt + (long long)n * e.rgb_sn + (long long)h * e.rgb_sh + w;
          atomicAdd(ra, rgb0);
          atomicAdd(ra + e.rgb_sj, rgb1);
          atomicAdd(ra + 2 * e.rgb_sj, rgb2);
        }
      }
    }
  }
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all();          // neither CTA leaves while the other may still read its operands or signal its barriers
  else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------------
static int g_hconv_mode = 1;          // 0: never use this kernel, 1: auto, 2: use it whenever the shape is supported
static int g_hconv_nb = 0;   // 0: by stage size
static int g_hconv_wt = 0;   // 0: widest that fits (<= 64)
static int g_hconv_grid = 0;
static int g_hconv_minpos = 64;     // auto mode: smallest H * W routed to this kernel
static int g_hconv_mask = 7;   // bit 0: single-source convs with > 4 taps, bit 1: <= 4 taps, bit 2: multi-source (up2 dgrad)
static int g_hconv_pair = 1;   // CTA-pair (cta_group::2) launches for 128-wide N tiles over an even number of images
static int g_hconv_plain = 1;  // plain-epilogue instantiation (EPI = 1) for fp32-output GEMMs on 128-wide tiles

static smc_igemm_plan_info* g_plan_out = nullptr;   // set only inside smc_igemm_plan (diagnostics, not thread-safe)

void hconv_config(int key, int value) {
  if (key == 0) g_hconv_mode = value;
  if (key == 2) g_hconv_nb = value;
  if (key == 3) g_hconv_wt = value;
  if (key == 4) g_hconv_grid = value;
  if (key == 5) g_hconv_mask = value;
  if (key == 6) g_hconv_minpos = value;
  if (key == 7) g_hconv_pair = value;
  if (key == 8) g_hconv_plain = value;
}

template <int BN, int KC, int MODE, bool ALO, int EPI = 0>
static int hc_launch_a(const CUtensorMap& ma, const CUtensorMap& mb, const HcParams& p, int grid, size_t smem, cudaStream_t st) {
  static SmemOptIn opt_in;
  if (cudaError_t e = smem_opt_in(opt_in, hconv_kernel<BN, KC, MODE, false, ALO, EPI>, smem); e != cudaSuccess) return (int)e;
  hconv_kernel<BN, KC, MODE, false, ALO, EPI><<<grid, HC_THREADS, smem, st>>>(ma, mb, p);
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}
template <int BN, int KC, int MODE, int EPI = 0>
static int hc_launch(const CUtensorMap& ma, const CUtensorMap& mb, const HcParams& p, int grid, size_t smem, cudaStream_t st) {
  if constexpr (MODE != HC_X1) {
    if (!p.a_lo_term) return hc_launch_a<BN, KC, MODE, false, EPI>(ma, mb, p, grid, smem, st);
  }
  return hc_launch_a<BN, KC, MODE, true, EPI>(ma, mb, p, grid, smem, st);
}
// CTA-pair launch: clusters of two CTAs (the two SMs of a TPC), grid = 2 x clusters
template <int BN, int KC, int MODE, bool ALO = true, int EPI = 0>
static int hc_launch_pair(const CUtensorMap& ma, const CUtensorMap& mb, const HcParams& p, int grid, size_t smem, cudaStream_t st) {
  if constexpr (MODE != HC_X1 && ALO) {
    if (!p.a_lo_term) return hc_launch_pair<BN, KC, MODE, false, EPI>(ma, mb, p, grid, smem, st);
  }
  static SmemOptIn opt_in;
  if (cudaError_t e = smem_opt_in(opt_in, hconv_kernel<BN, KC, MODE, true, ALO, EPI>, smem); e != cudaSuccess) return (int)e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(HC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, hconv_kernel<BN, KC, MODE, true, ALO, EPI>, ma, mb, p);
  if (e != cudaSuccess) return (int)e;
  SMC_LAUNCH_CHECK();
  return SMC_OK;
}

template <int BN, int KC>
static int hc_launch_x(int mode, const CUtensorMap& ma, const CUtensorMap& mb, const HcParams& p, int grid, size_t smem, cudaStream_t st) {
  if (mode == HC_X1) return hc_launch<BN, KC, HC_X1>(ma, mb, p, grid, smem, st);
  if constexpr (BN <= 64) {
    if (mode == HC_X3_MERGED) return hc_launch<BN, KC, HC_X3_MERGED>(ma, mb, p, grid, smem, st);
  }
  return hc_launch<BN, KC, HC_X3_TWO_PASS>(ma, mb, p, grid, smem, st);
}

static int hconv_try_launch_impl(const smc_igemm_desc* d, cudaStream_t st);

// A plain GEMM (one tap, H = 1, M = W rows) is folded into an (M / 64) x 64 "image" so that the same position tiling applies:
// 256 consecutive rows per tile, no halo.
int hconv_try_launch(const smc_igemm_desc* d, cudaStream_t st) {
  if (g_hconv_mode == 0) return SMC_EUNSUPPORTED;
  if (d->H == 1 && d->n_img == 1 && d->HA == 1 && d->WA == d->W && (d->ntaps == 1 || d->ntaps == 3) && !d->epi.noise && !d->epi.rgb_acc &&
      d->tw == 0 && (g_hconv_mask & 128) == 0) {
    bool plain = true;
    for (int i = 0; i < d->ntaps; ++i) plain = plain && d->taps[i].dy == 0 && d->taps[i].dx == 0;
    int wf = 0;
    for (int c = 64; c >= 16 && !wf; --c)
      if (d->W % c == 0) wf = c;
    if (plain && wf && d->W / wf >= 8 && (g_hconv_mode == 2 || d->W >= 1024)) {
      smc_igemm_desc f = *d;
      f.H = f.HA = d->W / wf;
      f.W = f.WA = wf;
      f.epi.o_sh = (int64_t)wf * d->epi.o_sw;
      return hconv_try_launch_impl(&f, st);
    }
  }
  return hconv_try_launch_impl(d, st);
}

static int hconv_try_launch_impl(const smc_igemm_desc* d, cudaStream_t st) {
  if (d->C % 32 != 0 || d->n_out % 32 != 0 || d->lda % 8 != 0 || d->ldb % 8 != 0) return SMC_EUNSUPPORTED;
  const int KC = d->C % 64 == 0 ? 64 : 32;
  if (d->tw > 0) return SMC_EUNSUPPORTED;          // caller pinned the igemm.cu tile shape
  if (g_hconv_mode == 1 && ((long long)d->H * d->W < g_hconv_minpos || d->H < 8)) return SMC_EUNSUPPORTED;
  const int BN = d->n_out % 128 == 0 ? 128 : (d->n_out % 64 == 0 ? 64 : 32);

  // split-precision pattern of gemm.igemm: [T base taps] [same, B rows + b_lo] [same, A planes + a_lo]
  int T = d->ntaps;
  bool x3 = false;
  int a_lo = 0, b_lo = 0;
  if (d->ntaps % 3 == 0 && d->ntaps >= 3) {
    const int t3 = d->ntaps / 3;
    a_lo = d->taps[2 * t3].dn - d->taps[0].dn;
    b_lo = d->taps[t3].brow - d->taps[0].brow;
    bool ok = a_lo > 0 && b_lo > 0;
    for (int i = 0; i < t3 && ok; ++i) {
      const smc_igemm_tap &a = d->taps[i], &b = d->taps[t3 + i], &c = d->taps[2 * t3 + i];
      ok = b.dn == a.dn && b.dy == a.dy && b.dx == a.dx && b.brow == a.brow + b_lo && c.dn == a.dn + a_lo && c.dy == a.dy &&
           c.dx == a.dx && c.brow == a.brow;
    }
    if (ok) {
      x3 = true;
      T = t3;
    }
  }
  // two-term split (gemm.igemm precision 'x2'): [T base taps] [same, B rows + b_lo] -- the A operand has a hi plane only
  bool a_lo_term = x3;
  if (!x3 && d->ntaps % 2 == 0 && d->ntaps >= 2) {
    const int t2 = d->ntaps / 2;
    b_lo = d->taps[t2].brow - d->taps[0].brow;
    bool ok = b_lo > 0;
    for (int i = 0; i < t2 && ok; ++i) {
      const smc_igemm_tap &a = d->taps[i], &b = d->taps[t2 + i];
      ok = b.dn == a.dn && b.dy == a.dy && b.dx == a.dx && b.brow == a.brow + b_lo;
    }
    if (ok) {
      x3 = true;
      T = t2;
      a_lo = 0;
    } else {
      b_lo = 0;
    }
  }
  if (T > HC_MAX_TAPS) return SMC_EUNSUPPORTED;

  HcParams p;
  p.a_lo_term = a_lo_term ? 1 : 0;
  p.n_img = d->n_img; p.H = d->H; p.W = d->W; p.C = d->C; p.n_out = d->n_out;
  int min_dx = 0, max_dx = 0, min_dy = 0, max_dy = 0;
  for (int i = 0; i < T; ++i) {
    min_dx = d->taps[i].dx < min_dx ? d->taps[i].dx : min_dx;
    max_dx = d->taps[i].dx > max_dx ? d->taps[i].dx : max_dx;
    min_dy = d->taps[i].dy < min_dy ? d->taps[i].dy : min_dy;
    max_dy = d->taps[i].dy > max_dy ? d->taps[i].dy : max_dy;
    if (d->taps[i].brow < 0 || d->taps[i].brow + (x3 ? b_lo : 0) + d->n_out > d->rowsB) return SMC_EINVAL;
  }
  if (max_dx - min_dx > 4 || max_dy - min_dy > 4) return SMC_EUNSUPPORTED;
  p.padL = -min_dx; p.padT = -min_dy;
  const int padW = max_dx - min_dx, padH = max_dy - min_dy;
  // smem plan: weight stages (resident when a tile's whole weight set is small and there is one N tile), then the widest
  // tile whose A buffers fit (wider = fewer discarded halo positions, but more halo rows fetched per tile)
  const int mode = !x3 ? HC_X1 : ((BN <= 64 && !(g_hconv_mask & 256)) ? HC_X3_MERGED : HC_X3_TWO_PASS);
  const int passes = mode == HC_X3_TWO_PASS ? 2 : 1;
  const int stages_per_tile = (d->C / KC) * T * passes;      // all problems together
  const bool can_reside = d->n_out == BN && stages_per_tile <= 32 && stages_per_tile * (BN * KC * 2 * (mode == HC_X3_MERGED ? 2 : 1)) <= 96 * 1024;
  int b_bytes = 0, saved_b_bytes = 0;
  size_t smem_fixed = 0, saved_fixed = 0;
  HcParams saved_plan;
  // attempt 0: resident weights where they fit; attempt 1 (only when the preferred tile width did not fit beside them): a weight ring
  for (int attempt = 0; attempt < 2; ++attempt) {
  p.b_resident = (can_reside && attempt == 0) ? 1 : 0;
  // CTA pair (cta_group::2): the 128-wide N tiles are bound by the shared-memory operand fetch of the MMAs (8 KB per 128 x 128 x 16
  // instruction); a pair reads the weight tile once for 2 x 128 rows.  The two CTAs take the same tile of two consecutive images.
  p.pair = (g_hconv_pair && BN == 128 && mode != HC_X3_MERGED && d->n_img % 2 == 0 && !p.b_resident) ? 1 : 0;
  b_bytes = BN * KC * 2 * (mode == HC_X3_MERGED ? 2 : 1) / (p.pair ? 2 : 1);        // one ring stage (per CTA)
  if (p.b_resident) {
    p.nb = stages_per_tile;
  } else {
    p.nb = g_hconv_nb > 0 ? g_hconv_nb : (b_bytes >= 16384 ? 4 : (b_bytes >= 8192 ? 6 : 8));
    p.nb = p.nb < 2 ? 2 : (p.nb > 32 ? 32 : p.nb);
  }
  smem_fixed = 1024 + (size_t)p.nb * b_bytes + HC_BAR_BYTES + (size_t)hc_pstage_bytes(BN);
  // Widths just above a multiple of 64 (the parity planes of the stride-2 transposed conv are 2^k + 1 wide) get column tiles of
  // ceil(W / floor(W / 64)) <= 80 pixels: with 64-wide tiles a 65-pixel row became a 64-wide and a 1-wide column tile, and the second one
  // cost as many position tiles as the first (conv0 at 128 / 256 / 512 / 1024 px ran 2.0 / 1.5 / 1.25 / 1.12 times the useful MMAs).
  int wt_first = 64;
  if (g_hconv_wt > 0) {
    wt_first = g_hconv_wt;
  } else if (d->W > 64 && d->W % 64 != 0) {
    const int bal = ceil_div(d->W, d->W / 64);
    if (bal <= 80) wt_first = bal;
  }
  const int wt_cands[5] = {wt_first, 64, 32, 16, 8};
  const bool short_tiles = d->C / KC <= 2;            // few MMAs per tile: prefetch A tiles further ahead
  p.Wt = 0;
  int chosen = -1;
  for (int ci = 0; ci < 5 && p.Wt == 0; ++ci) {
    const int wt = d->W < wt_cands[ci] ? d->W : wt_cands[ci];
    const int wp = wt + padW;
    const int rb = (HC_MT % wp == 0) ? HC_MT / wp + padH : ceil_div(HC_MT, wp) + 1 + padH;
    const size_t abuf = ((size_t)(rb * wp + 8) * (size_t)(KC * 2) + 1023u) & ~(size_t)1023u;
    if (wp > 256 || rb > 256) continue;
    const int nmax = (int)((227 * 1024 - smem_fixed) / abuf);
    int nh, nl;
    if (x3 && a_lo_term) {
      if (nmax < 3) continue;
      // short slabs (few taps): the lo buffer is only free during the short B_lo pass, so it needs a second buffer
      const bool short_slabs = T * (KC / 16) < 16 || d->nprob > 1;   // problems of a parity group have 1-4 taps each
      const int want_h = short_tiles ? 4 : (short_slabs ? 3 : 2);
      // two-pass: the lo tile is released after the B_hi pass; merged: it lives as long as the hi tile
      const int want_l = mode == HC_X3_MERGED ? want_h : (short_tiles ? 4 : (short_slabs ? 2 : 1));
      nl = want_l < (nmax / 2 < 1 ? 1 : nmax / 2) ? want_l : (nmax / 2 < 1 ? 1 : nmax / 2);
      nh = nmax - nl > want_h ? want_h : nmax - nl;
      if (nh < 2) { nh = 2; nl = nmax - 2; }
    } else {
      if (nmax < 2) continue;
      nh = nmax > 4 ? 4 : nmax;
      nl = 0;
    }
    p.Wt = wt; p.Wp = wp; p.RB = rb; p.na_hi = nh; p.na_lo = nl;
    chosen = ci;
  }
  if (attempt == 1) {
    if (chosen != 0) { p = saved_plan; b_bytes = saved_b_bytes; smem_fixed = saved_fixed; }    // the ring did not help: keep the resident plan
    break;
  }
  if (chosen == 0 || !p.b_resident || wt_first == 64) break;      // else: retry with a weight ring, which leaves room for the preferred width
  saved_plan = p; saved_b_bytes = b_bytes; saved_fixed = smem_fixed;
  }
  if (p.Wt == 0) return SMC_EUNSUPPORTED;
  p.col_tiles = ceil_div(d->W, p.Wt);
  p.tiles_per_col = (int)ceil_div_ll((long long)d->H * p.Wp, HC_MT);
  p.n_tiles_n = d->n_out / BN;
  p.nprob = d->nprob > 1 ? d->nprob : 1;
  if (p.nprob > HC_MAX_PROBS) return SMC_EINVAL;
  {
    const long long tt = (long long)(p.pair ? d->n_img / 2 : d->n_img) * p.col_tiles * p.tiles_per_col * p.n_tiles_n;
    if (tt * p.nprob > 0x7fffffffLL) return SMC_ETOOLARGE;
    p.super_tiles = (int)tt;
    p.total_tiles = (int)tt * p.nprob;
  }
  p.kchunks = d->C / KC;
  p.div_ntn = hc_make_div((uint32_t)p.n_tiles_n);
  p.div_tpc = hc_make_div((uint32_t)p.tiles_per_col);
  p.div_ct = hc_make_div((uint32_t)p.col_tiles);
  p.div_wp = hc_make_div((uint32_t)p.Wp);

  // group the base taps by A source
  p.ngroups = 0;
  int ntap = 0;
  bool used[SMC_IGEMM_MAX_TAPS] = {false};
  for (int i = 0; i < T; ++i) {
    if (used[i]) continue;
    if (p.ngroups == HC_MAX_GROUPS) return SMC_EUNSUPPORTED;
    HcGroup& g = p.groups[p.ngroups++];
    g.dn_hi = d->taps[i].dn;
    g.dn_lo = d->taps[i].dn + a_lo;
    g.tap_begin = ntap;
    for (int k = i; k < T; ++k) {
      if (used[k] || d->taps[k].dn != d->taps[i].dn) continue;
      used[k] = true;
      HcTap& tp = p.taps[ntap++];
      tp.posoff = (d->taps[k].dy + p.padT) * p.Wp + d->taps[k].dx + p.padL;
      tp.brow_hi = d->taps[k].brow;
      tp.brow_lo = d->taps[k].brow + b_lo;
    }
    g.tap_end = ntap;
  }
  {
    const int kind = p.ngroups > 1 ? 4 : (T > 4 ? 1 : 2);
    if (!(g_hconv_mask & kind)) return SMC_EUNSUPPORTED;
    if ((g_hconv_mask & 8) && d->epi.act == 1) return SMC_EUNSUPPORTED;     // diagnostics: keep activated (forward) convs on igemm.cu
    if ((g_hconv_mask & 16) && d->epi.act != 1) return SMC_EUNSUPPORTED;    // diagnostics: keep linear (dgrad, parity) convs on igemm.cu
    if ((g_hconv_mask & 32) && d->C % 64 != 0) return SMC_EUNSUPPORTED;
    if ((g_hconv_mask & 64) && d->C % 64 == 0) return SMC_EUNSUPPORTED;
  }
  // accumulation chains: at most max_chain MMAs (K = 16 each) go into the main accumulator between two drains
  const int mma_per_tap = KC / 16;
  const int max_chain = d->acc_chunk_k > 0 ? (d->acc_chunk_k / 16 < mma_per_tap ? mma_per_tap : d->acc_chunk_k / 16) : (1 << 30);
  // problems: one (all groups) by default; with nprob > 1 the base taps of the single group are split into consecutive runs
  if (p.nprob > 1) {
    if (p.ngroups != 1) return SMC_EUNSUPPORTED;
    int sum = 0;
    for (int q = 0; q < p.nprob; ++q) {
      if (d->prob_ntaps[q] < 1) return SMC_EINVAL;
      sum += d->prob_ntaps[q];
    }
    if (sum != T) return SMC_EINVAL;
  }
  int nsegs_total = 0, tap_cursor = 0, stage_cursor = 0;
  for (int q = 0; q < p.nprob; ++q) {
    HcProb& pr = p.probs[q];
    // the tap ranges this problem covers: (group, begin, end)
    int rg[HC_MAX_GROUPS], rb[HC_MAX_GROUPS], re[HC_MAX_GROUPS], nr = 0;
    if (p.nprob > 1) {
      rg[0] = 0; rb[0] = tap_cursor; re[0] = tap_cursor + d->prob_ntaps[q]; nr = 1;
      tap_cursor = re[0];
    } else {
      for (int g = 0; g < p.ngroups; ++g) { rg[nr] = g; rb[nr] = p.groups[g].tap_begin; re[nr] = p.groups[g].tap_end; ++nr; }
    }
    int Tq = 0;
    for (int i = 0; i < nr; ++i) Tq += re[i] - rb[i];
    const int chain_per_slab = Tq * mma_per_tap;
    pr.seg_begin = nsegs_total;
    pr.stage_base = stage_cursor;
    pr.pad_ = 0;
    pr.o_off = p.nprob > 1 ? d->prob_o_off[q] : 0;
    stage_cursor += (d->C / KC) * Tq * passes;
    int mid_commits = 0;
    if (x3 && chain_per_slab > max_chain) {
      // split inside the slab: segments of at most max_chain / mma_per_tap taps, greedy commits
      const int seg_taps = max_chain / mma_per_tap < 1 ? 1 : max_chain / mma_per_tap;
      int chain = 0;
      for (int i = 0; i < nr; ++i) {
        const int nt_g = re[i] - rb[i];
        const int nseg = ceil_div(nt_g, seg_taps);
        const int len = ceil_div(nt_g, nseg);
        for (int qs = 0; qs < nseg; ++qs) {
          if (nsegs_total == HC_MAX_SEGS) return SMC_EUNSUPPORTED;
          HcSeg& sg = p.segs[nsegs_total++];
          sg.g = rg[i];
          sg.tb = rb[i] + qs * len;
          sg.te = sg.tb + len < re[i] ? sg.tb + len : re[i];
          sg.flags = (qs == 0 ? HC_SEG_FIRST : 0) | (qs == nseg - 1 ? HC_SEG_LAST : 0);
          const int add = (sg.te - sg.tb) * mma_per_tap;
          if (chain > 0 && chain + add > max_chain) {          // commit after the previous segment
            p.segs[nsegs_total - 2].flags |= HC_SEG_COMMIT;
            ++mid_commits;
            chain = 0;
          }
          chain += add;
        }
      }
      pr.kcs_per_drain = 1;
    } else {
      for (int i = 0; i < nr; ++i) {
        if (nsegs_total == HC_MAX_SEGS) return SMC_EUNSUPPORTED;
        HcSeg& sg = p.segs[nsegs_total++];
        sg.g = rg[i]; sg.tb = rb[i]; sg.te = re[i]; sg.flags = HC_SEG_FIRST | HC_SEG_LAST;
      }
      pr.kcs_per_drain = d->acc_chunk_k > 0 ? (max_chain / chain_per_slab < 1 ? 1 : max_chain / chain_per_slab) : p.kchunks;
    }
    pr.nsegs = nsegs_total - pr.seg_begin;
    p.segs[nsegs_total - 1].flags |= HC_SEG_COMMIT | HC_SEG_SLABEND;
    pr.ndrains = p.kchunks * mid_commits + ceil_div(p.kchunks, pr.kcs_per_drain);
  }
  for (int q = p.nprob; q < HC_MAX_PROBS; ++q) p.probs[q] = p.probs[0];
  // (two-pass mode releases the lo tile after each B_hi pass, so it keeps one load per problem)
  p.a_share = (p.nprob > 1 && p.kchunks == 1 && mode != HC_X3_TWO_PASS && !(g_hconv_mask & 512) && !p.pair) ? 1 : 0;
  p.a_box_bytes = (uint32_t)(p.RB * p.Wp) * (uint32_t)(KC * 2);
  p.a_buf_bytes = ((uint32_t)(p.RB * p.Wp + 8) * (uint32_t)(KC * 2) + 1023u) & ~1023u;
  p.epi = d->epi;
  if (!p.epi.out_f32 && !p.epi.out_hi && !p.epi.out_raw && !p.epi.rgb_acc) return SMC_EINVAL;
  if (p.epi.mask_y) {
    // fused activation backward: the lean path needs 32-byte aligned fp16 planes and float4-aligned parameter vectors
    const smc_igemm_epilogue& e = p.epi;
    if (!e.post_scale || !e.out_hi || e.row_scale || e.bias || e.noise || e.residual || e.out_f32 || e.out_raw || e.out_raw_lo || e.rgb_acc)
      return SMC_EINVAL;
    if ((e.mask_grgb != nullptr) != (e.rgb_w != nullptr)) return SMC_EINVAL;
    if ((((uintptr_t)e.mask_y | (uintptr_t)e.mask_y_lo | (uintptr_t)e.out_hi | (uintptr_t)e.out_lo) & 31) || (((uintptr_t)e.post_scale | (uintptr_t)e.rgb_w) & 15) ||
        (((e.o_sn | e.o_sh | e.o_sw | e.o_off) * 2) & 31) || (d->n_out % 16 != 0))
      return SMC_EUNSUPPORTED;
  } else if (p.epi.mask_y_lo || p.epi.mask_grgb) {
    return SMC_EINVAL;
  } else
  if ((p.epi.rgb_acc != nullptr) != (p.epi.rgb_w != nullptr) || (p.epi.out_raw_lo && !p.epi.out_raw)) return SMC_EINVAL;
  if ((p.epi.o_sn | p.epi.o_sh | p.epi.o_sw | p.epi.o_off) & 7) return SMC_EUNSUPPORTED;
  const size_t smem = smem_fixed + (size_t)(p.na_hi + p.na_lo) * p.a_buf_bytes;
  if (smem > 227 * 1024) return SMC_EUNSUPPORTED;

  if (g_plan_out) {                    // smc_igemm_plan: report the launch plan instead of launching (host only, no CUDA call)
    smc_igemm_plan_info& o = *g_plan_out;
    o.kernel = 1; o.bn = BN; o.kc = KC; o.mode = mode;
    o.Wt = p.Wt; o.Wp = p.Wp; o.RB = p.RB; o.na_hi = p.na_hi; o.na_lo = p.na_lo; o.nb = p.nb;
    o.b_resident = p.b_resident; o.a_share = p.a_share; o.nprob = p.nprob; o.kchunks = p.kchunks;
    o.super_tiles = p.super_tiles;
    o.pair = p.pair;
    o.grid = p.super_tiles < (g_hconv_grid > 0 ? g_hconv_grid : kNumSMs) ? p.super_tiles : (g_hconv_grid > 0 ? g_hconv_grid : kNumSMs);
    if (p.pair) o.grid = 2 * (p.super_tiles < kNumSMs / 2 ? p.super_tiles : kNumSMs / 2);
    o.smem_bytes = (int32_t)smem;
    {                                  // TMEM columns exactly as hconv_kernel derives them from <BN, MODE>
      const int blk = (mode == HC_X3_MERGED ? 2 : 1) * BN;
      const int setcols = (mode == HC_X3_TWO_PASS ? 2 : 1) * HC_MB * blk;
      const int sets = 2 * setcols <= 512 ? 2 : 1;
      o.tmem_cols = sets * setcols < 32 ? 32 : sets * setcols;
    }
    for (int q = 0; q < HC_MAX_PROBS; ++q) {
      const HcProb& pr = p.probs[q < p.nprob ? q : 0];
      o.prob_nsegs[q] = q < p.nprob ? pr.nsegs : 0;
      o.prob_ndrains[q] = q < p.nprob ? pr.ndrains : 0;
      o.prob_stages[q] = 0;
      // main-accumulator commits per tile as the MMA issuer's loop makes them (must equal the drains the epilogue waits for)
      int commits = 0, in_chunk = 0, ntap = 0;
      for (int kc = 0; kc < p.kchunks && q < p.nprob; ++kc) {
        const bool chunk_ends = (in_chunk + 1 == pr.kcs_per_drain) || (kc == p.kchunks - 1);
        for (int si = pr.seg_begin; si < pr.seg_begin + pr.nsegs; ++si) {
          const int f = p.segs[si].flags;
          if ((f & HC_SEG_COMMIT) && (!(f & HC_SEG_SLABEND) || chunk_ends)) ++commits;
          if (kc == 0) ntap += p.segs[si].te - p.segs[si].tb;
        }
        in_chunk = chunk_ends ? 0 : in_chunk + 1;
      }
      o.prob_commits[q] = commits;
      o.prob_ntaps[q] = ntap;
      if (q < p.nprob) o.prob_stages[q] = pr.stage_base;
    }
    return SMC_OK;
  }

  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return SMC_EDRIVER;
  const CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMap ma, mb;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->C, (cuuint64_t)d->WA, (cuuint64_t)d->HA, (cuuint64_t)d->NA};
    cuuint64_t strides[3] = {(cuuint64_t)d->lda * 2, (cuuint64_t)d->lda * 2 * d->WA, (cuuint64_t)d->lda * 2 * d->WA * d->HA};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)p.Wp, (cuuint32_t)p.RB, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(d->A), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return SMC_EDRIVER;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->C, (cuuint64_t)d->rowsB};
    cuuint64_t strides[1] = {(cuuint64_t)d->ldb * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)(p.pair ? BN / 2 : BN)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->B), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return SMC_EDRIVER;
  }
  const int max_grid = g_hconv_grid > 0 ? g_hconv_grid : sm_count();
  // plain epilogue (hconv_kernel EPI = 1): fp32 output = acc * acc_scale [* row_scale] [+ bias] [+ residual] and nothing else
  const smc_igemm_epilogue& pe = p.epi;
  const bool plain = g_hconv_plain && BN == 128 && KC == 64 && pe.out_f32 && !pe.out_hi && !pe.out_lo && !pe.out_raw && !pe.out_raw_lo && !pe.rgb_acc && !pe.rgb_w &&
                     !pe.post_scale && !pe.noise && !pe.mask_y && pe.act == 0 && pe.gain == 1.f && pe.clamp < 0.f &&
                     ((((uintptr_t)pe.out_f32 | (uintptr_t)pe.residual) & 31) == 0) && ((((uintptr_t)pe.row_scale | (uintptr_t)pe.bias) & 15) == 0);
  // one epilogue family per instantiation for the 128-wide kernels (EPI = 2 modulated conv, EPI = 3 fused activation backward): what
  // hconv_kernel's `modconv` test and the mask_y validation above establish, decided here so that each kernel compiles one path only
  const bool out32 = ((((uintptr_t)pe.out_raw | (uintptr_t)pe.out_raw_lo | (uintptr_t)pe.out_hi | (uintptr_t)pe.out_lo) & 31) == 0) &&
                     ((((pe.o_sn | pe.o_sh | pe.o_sw | pe.o_off) * 2) & 31) == 0) && (d->n_out % 16 == 0);
  const bool fam_modconv = pe.row_scale && pe.bias && pe.act == 1 && pe.clamp >= 0.f && pe.alpha >= 0.f && pe.alpha <= 1.f && pe.gain > 0.f &&
                           !pe.residual && !pe.out_f32 && !pe.mask_y && out32;
  const int family = (g_hconv_plain && BN == 128 && KC == 64 && mode == HC_X3_TWO_PASS) ? (pe.mask_y ? 3 : (fam_modconv ? 2 : 0)) : 0;
  if (family) {
    if (p.pair) {
      const int max_clusters = max_grid / 2;
      const int grid2 = 2 * (p.super_tiles < max_clusters ? (int)p.super_tiles : max_clusters);
      return family == 2 ? hc_launch_pair<128, 64, HC_X3_TWO_PASS, true, 2>(ma, mb, p, grid2, smem, st)
                         : hc_launch_pair<128, 64, HC_X3_TWO_PASS, true, 3>(ma, mb, p, grid2, smem, st);
    }
    const int grid1 = p.super_tiles < max_grid ? (int)p.super_tiles : max_grid;
    return family == 2 ? hc_launch<128, 64, HC_X3_TWO_PASS, 2>(ma, mb, p, grid1, smem, st) : hc_launch<128, 64, HC_X3_TWO_PASS, 3>(ma, mb, p, grid1, smem, st);
  }
  if (plain) {
    if (p.pair) {
      const int max_clusters = max_grid / 2;
      const int grid2 = 2 * (p.super_tiles < max_clusters ? (int)p.super_tiles : max_clusters);
      return mode == HC_X1 ? hc_launch_pair<128, 64, HC_X1, true, 1>(ma, mb, p, grid2, smem, st)
                           : hc_launch_pair<128, 64, HC_X3_TWO_PASS, true, 1>(ma, mb, p, grid2, smem, st);
    }
    const int grid1 = p.super_tiles < max_grid ? (int)p.super_tiles : max_grid;
    return mode == HC_X1 ? hc_launch<128, 64, HC_X1, 1>(ma, mb, p, grid1, smem, st) : hc_launch<128, 64, HC_X3_TWO_PASS, 1>(ma, mb, p, grid1, smem, st);
  }
  if (p.pair) {
    const int max_clusters = max_grid / 2;
    const int grid2 = 2 * (p.super_tiles < max_clusters ? (int)p.super_tiles : max_clusters);
    if (KC == 64) return mode == HC_X1 ? hc_launch_pair<128, 64, HC_X1>(ma, mb, p, grid2, smem, st)
                                       : hc_launch_pair<128, 64, HC_X3_TWO_PASS>(ma, mb, p, grid2, smem, st);
    return mode == HC_X1 ? hc_launch_pair<128, 32, HC_X1>(ma, mb, p, grid2, smem, st)
                         : hc_launch_pair<128, 32, HC_X3_TWO_PASS>(ma, mb, p, grid2, smem, st);
  }
  const int grid = p.super_tiles < max_grid ? (int)p.super_tiles : max_grid;
  if (KC == 64) {
    if (BN == 128) return hc_launch_x<128, 64>(mode, ma, mb, p, grid, smem, st);
    if (BN == 64) return hc_launch_x<64, 64>(mode, ma, mb, p, grid, smem, st);
    return hc_launch_x<32, 64>(mode, ma, mb, p, grid, smem, st);
  }
  if (BN == 128) return hc_launch_x<128, 32>(mode, ma, mb, p, grid, smem, st);
  if (BN == 64) return hc_launch_x<64, 32>(mode, ma, mb, p, grid, smem, st);
  return hc_launch_x<32, 32>(mode, ma, mb, p, grid, smem, st);
}

}  // namespace smc

// The launch plan the halo-tile kernel would use for a descriptor (tile shape, shared-memory and TMEM budget, problem table), without
// touching the GPU: lets the CPU test-suite check the planner on every layer shape of the benchmark networks.
extern "C" int smc_igemm_plan(const smc_igemm_desc* desc, smc_igemm_plan_info* out) {
  if (!desc || !out) return SMC_EINVAL;
  if (desc->nprob < 0 || desc->nprob > 4 || desc->ntaps < 1 || desc->ntaps > SMC_IGEMM_MAX_TAPS) return SMC_EINVAL;
  *out = smc_igemm_plan_info{};
  smc::g_plan_out = out;
  const int r = smc::hconv_try_launch(desc, nullptr);
  smc::g_plan_out = nullptr;
  return r;               // SMC_EUNSUPPORTED: the call would go to the per-tap kernel (out->kernel stays 0)
}

extern "C" int smc_igemm_config(int key, int value) {
  smc::hconv_config(key, value);
  return SMC_OK;
}
