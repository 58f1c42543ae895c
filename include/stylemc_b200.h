/* stylemc_b200.h -- C ABI of libstylemc_b200.so (sm_100a kernels for the StyleMC hot path).
 *
 * Plain pointers and sizes only; no torch types.  Every device pointer is borrowed, every output is
 * allocated by the caller, every call is asynchronous on the `stream` argument (a cudaStream_t passed as
 * void*; NULL = legacy default stream) and re-entrant.  Return value: 0 ok; < 0 argument error
 * (SMC_E*); > 0 a cudaError_t.  The host-side Python mirror (stylemc_b200/ops, networks, clip) raises
 * RuntimeError on any non-zero status, like TORCH_CHECK does in the reference plugins.
 *
 * Reference interfaces replaced (file:line under the reference tree):
 *   smc_bias_act       torch_utils/ops/bias_act.cpp:32  `bias_act(x,b,xref,yref,dy,grad,dim,act,alpha,gain,clamp)`
 *                      (pybind, bias_act.cpp:94-97; kernel params bias_act.h:12-31; called from bias_act.py:153,182,201)
 *   smc_upfirdn2d      torch_utils/ops/upfirdn2d.cpp:16 `upfirdn2d(x,f,upx,upy,downx,downy,padx0,padx1,pady0,pady1,flip,gain)`
 *                      (pybind, upfirdn2d.cpp:98-101; kernel params upfirdn2d.h:14-40; called from upfirdn2d.py:237-240)
 *   smc_igemm          the cuDNN / cuBLAS calls reached through torch_utils/ops/conv2d_gradfix.py:35-43
 *                      (F.conv2d / F.conv_transpose2d from conv2d_resample.py:29-54,138,147) and CLIP's nn.Linear
 *   smc_demod_coefs, smc_pack_nhwc, smc_unpack_nchw, smc_fir_act, smc_torgb, smc_act_bwd, smc_fir_bwd,
 *   smc_sgrad_finish, smc_grad_scale
 *                      the ATen passes of [UPSTREAM] modulated_conv2d / SynthesisLayer / ToRGBLayer as driven by
 *                      utils.py:13-53 (block_forward), plus fma.py:15-58 and the autograd of all of them
 *   smc_resample_*     find_direction.py:49-52 `unprocess` (torchvision Resize(224, BICUBIC) + normalise)
 *   smc_patchify .. smc_clip_loss
 *                      openai/CLIP clip/model.py VisionTransformer / Transformer pieces behind
 *                      clip_loss.py:15-16,25-26 (`encode_text`, `encode_image`) and clip_loss.py:24-34 (the loss)
 *   smc_sgd_step       find_direction.py:285,298-301,339 (SGD without momentum, cosine LR computed by the host)
 */
#pragma once
#include <stdint.h>

#define SMC_OK 0
#define SMC_EINVAL (-1)
#define SMC_EUNSUPPORTED (-2)
#define SMC_ETOOLARGE (-3)
#define SMC_EDRIVER (-4)

#define SMC_F32 0
#define SMC_F16 1
#define SMC_F64 2

#define SMC_ABI_VERSION 1

#ifdef __cplusplus
extern "C" {
#endif

int smc_abi_version(void);

/* ---- bias_act ------------------------------------------------------------------------------------
 * y = clamp(act(x + b[(i / step_b) % size_b]) * gain)  (grad = 0), or the 1st / 2nd derivative pass
 * (grad = 1 / 2) with the reference plugin's operand meaning (x = dy-like input, xref, yref, dy).
 * b, xref, yref, dy may be NULL.  act = 1..9 = linear, relu, lrelu, tanh, sigmoid, elu, selu, softplus, swish
 * (bias_act.py:23-33 cuda_idx).  clamp < 0 disables clamping.  All tensors dense, same layout, dtype `dtype`. */
int smc_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                 int dtype, int64_t size_x, int32_t size_b, int64_t step_b, int grad, int act, float alpha,
                 float gain, float clamp, void* stream);

/* ---- upfirdn2d ----------------------------------------------------------------------------------- */
typedef struct smc_upfirdn2d_params {
  int32_t N, C, inH, inW, outH, outW;   /* outH/outW as computed at upfirdn2d.cpp:32-33 */
  int64_t x_stride[4], y_stride[4];     /* element strides (n, c, h, w) */
  int32_t fH, fW;
  int64_t f_stride[2];                  /* fp32 filter strides (h, w) */
  int32_t upx, upy, downx, downy, padx0, pady0, flip;
  float gain;
  /* optional hint for 4x4 filters: f[i][j] == fsep[i] * fsep[4 + j] (rank-1 filter, e.g. setup_filter([1,3,3,1])); the kernels then
   * filter rows and columns separately.  0 = not known / not separable. */
  int32_t separable;
  float fsep[8];
} smc_upfirdn2d_params;
int smc_upfirdn2d(const void* x, const float* f, void* y, int dtype, const smc_upfirdn2d_params* p, void* stream);

/* ---- implicit GEMM (tcgen05) --------------------------------------------------------------------- */
/*   D[m, o] = sum_t sum_k A_t[m, k] * B_t[o, k]          fp16 operands, fp32 accumulate in TMEM
 * A is an NHWC fp16 tensor [NA, HA, WA, C] read through ONE 4-D TMA map; a GEMM row m is a point (n, h, w)
 * of the iteration space [n_img, H, W] and tap t reads A at (n + dn_t, h + dy_t, w + dx_t); out-of-range
 * coordinates are zero-filled by TMA (that is the conv padding).  B is a K-major fp16 matrix [rowsB, C];
 * tap t uses rows brow_t .. brow_t + n_out.  A plain GEMM is H = 1 with one tap.  Split-precision ("x3")
 * runs are extra taps: (A_hi,B_hi), (A_hi,B_lo), (A_lo,B_hi), the lo planes stacked behind the hi planes. */
#define SMC_IGEMM_MAX_TAPS 32

typedef struct smc_igemm_tap {
  int32_t dn, dy, dx;   // A coordinate offsets (image/plane, row, column)
  int32_t brow;         // first B row of this tap
} smc_igemm_tap;

// Epilogue: v = acc * acc_scale; v *= row_scale[n, o]; v += noise[h, w]; v += bias[o]; v = act(v) * gain;
// v = clamp(v); out_raw = fp16(v); v *= post_scale[n, o]; v += residual; out = v.
typedef struct smc_igemm_epilogue {
  const float* row_scale;    // [n_img, n_out] or NULL   (demodulation coefficients)
  const float* post_scale;   // [n_img, n_out] or NULL   (next layer's styles)
  const float* bias;         // [n_out] or NULL
  const float* noise;        // fp32 plane or NULL; element (h, w) at noise[h * noise_sh + w * noise_sw]
  int64_t noise_sh, noise_sw;
  int32_t act;               // 0 linear, 1 leaky relu
  float alpha, gain, clamp;  // clamp < 0: none
  const float* residual;     // fp32 or NULL, addressed like out_f32
  float* out_f32;            // any subset of the four outputs may be NULL
  void* out_hi;              // fp16
  void* out_lo;              // fp16: rn(v - hi)
  void* out_raw;             // fp16 value before post_scale / residual
  int64_t o_sn, o_sh, o_sw;  // element strides of the output address for (n, h, w); channel stride 1
  int64_t o_off;             // element offset of (0, 0, 0, channel 0)
  float acc_scale;           // applied to the raw accumulator first (undoes a power-of-two pre-scaling of B); 0 means 1
  // Fused ToRGB (halo-tile conv kernel only; the per-tap kernel returns SMC_EUNSUPPORTED when these are set):
  void* out_raw_lo;          // fp16: rn(v - out_raw) of the value before post_scale (second plane of the saved activation)
  const float* rgb_w;        // [n_img, 3, n_out] modulated ToRGB weights (w[j, c] * style[n, c] * weight_gain) or NULL
  float* rgb_acc;            // fp32, atomically accumulated: rgb_acc[n * rgb_sn + j * rgb_sj + h * rgb_sh + w] += sum_c rgb_w * v
  int64_t rgb_sn, rgb_sj, rgb_sh;
  // Fused activation backward (halo-tile conv kernel only; used by the dgrad GEMMs, replaces a separate smc_act_bwd pass where no
  // style-gradient reduction is needed): with y = mask_y (+ mask_y_lo), the saved output of the layer BELOW,
  //   v = acc * acc_scale * post_scale[n, o] * (y > 0 ? gain : gain * alpha), v = 0 where |y| >= clamp (clamp >= 0);
  // out_hi / out_lo receive v (bias_act.cu:71-72,136-142 semantics: the gradient uses the saved output, not the input).
  // row_scale, bias, noise, act, residual, out_f32, out_raw and the ToRGB fields must be unset.  Addressed like the outputs.
  const void* mask_y;        // fp16 or NULL
  const void* mask_y_lo;     // fp16 or NULL
  // ToRGB branch of the same backward step (optional, with mask_y): v gets, before the activation slope,
  //   + sum_j rgb_w[n, j, o] * mask_grgb[n * rgb_sn + j * rgb_sj + h * rgb_sh + w]   (mask_grgb = masked, loss-scaled dL/drgb, fp32)
  const float* mask_grgb;
  // Fused ToRGB over several N tiles (n_out > the kernel's N tile): N tile t adds into rgb_acc + t * rgb_snt (one partial-sum image per N
  // tile, summed in index order by smc_img_finish: deterministic).  0: every N tile adds into the same image (sum order not fixed).
  int64_t rgb_snt;
} smc_igemm_epilogue;

typedef struct smc_igemm_desc {
  const void* A;             // fp16 [NA, HA, WA, C], channel stride 1, pixel stride lda elements
  int32_t NA, HA, WA, C;
  int64_t lda;
  const void* B;             // fp16 [rowsB, C] row-major (K contiguous), row stride ldb elements
  int32_t rowsB;
  int64_t ldb;
  int32_t n_img, H, W;       // iteration space (GEMM M = n_img * H * W)
  int32_t n_out;             // GEMM N
  int32_t tw, th, tn;        // M-tile box (tw * th * tn == 128); 0 = choose automatically
  int32_t ntaps;
  smc_igemm_tap taps[SMC_IGEMM_MAX_TAPS];
  smc_igemm_epilogue epi;
  int32_t acc_chunk_k;       // 0: one TMEM accumulation chain; > 0: drain the accumulator into fp32 registers every
                             // ~acc_chunk_k K-elements (removes the tensor core's truncation bias on long chains)
  // Problem group (optional): nprob = 2..4 GEMMs that share A, B, the iteration space and the epilogue but own a run of the base
  // taps each (prob_ntaps[q] consecutive taps; in a split-precision tap list the runs refer to the first third) and write to
  // epi.o_off + prob_o_off[q].  The four output parities of the stride-2 transposed conv (conv2d_resample.py:132-139 with up = 2)
  // are such a group: one launch then fetches every input tile from DRAM once instead of once per parity.  nprob = 0 or 1: one GEMM.
  int32_t nprob;
  int32_t prob_ntaps[4];
  int64_t prob_o_off[4];
} smc_igemm_desc;

int smc_igemm(const smc_igemm_desc* desc, void* stream);
/* Diagnostics (host only, no CUDA call, not thread-safe): the plan csrc/hconv.cu would launch `desc` with.  Returns SMC_EUNSUPPORTED
 * (and kernel = 0) when the call would be served by the per-tap kernel csrc/igemm.cu instead. */
typedef struct smc_igemm_plan_info {
  int32_t kernel;                 /* 1: halo-tile kernel */
  int32_t bn, kc, mode;           /* template arguments: N tile, K slab, 0 x1 / 1 split two-pass / 2 split merged-B */
  int32_t Wt, Wp, RB;             /* tile width, pitch with halo, rows of the TMA box */
  int32_t na_hi, na_lo, nb;       /* A buffers (hi / lo ring), weight stages */
  int32_t b_resident, a_share, nprob, kchunks;
  int32_t super_tiles, grid;      /* (position, N) tiles; CTAs */
  int32_t smem_bytes, tmem_cols;
  int32_t prob_ntaps[4], prob_nsegs[4], prob_ndrains[4], prob_commits[4], prob_stages[4];
  int32_t pair;                   /* 1: CTA-pair launch (cta_group::2, clusters of 2): super_tiles counts PAIRS of images, grid = 2 x clusters */
} smc_igemm_plan_info;
int smc_igemm_plan(const smc_igemm_desc* desc, smc_igemm_plan_info* out);
/* Tuning / diagnostics knobs of the convolution path (process-global; set before launching, not thread-safe):
 *   key 0: halo-tile kernel (hconv.cu) use: 0 never, 1 auto (default), 2 whenever the shape is supported
 *   key 2: weight-stage ring depth (0 = by stage size)   key 3: tile width Wt in pixels (0 = widest that fits, <= 64)
 *   key 4: persistent grid size (0 = one CTA per SM)      key 5: bit mask of conv kinds routed to hconv.cu (diagnostics)
 *   key 6: smallest H * W the auto mode routes to hconv.cu (default 64)
 *   key 7: CTA-pair launches (tcgen05 cta_group::2) for 128-wide N tiles over an even number of images: 1 on (default), 0 off
 *   key 8: the plain-epilogue instantiation for fp32-output GEMMs on 128-wide tiles (CLIP linears, conv0 planes): 1 on (default), 0 off */
int smc_igemm_config(int key, int value);

/* ---- synthesis glue (synth.cu) -------------------------------------------------------------------
 * Activations are NHWC fp16 ("hi" plane, optional "lo" plane = rn(v - hi)); styles are rows of the
 * [N, 26, 512] S tensor addressed as base pointer + n * stride. */
/* A/B diagnostics (process-global, not thread-safe): key 0 / 1 / 2 = use the newer smc_fir_act / smc_fir_bwd / smc_act_bwd kernels,
 * key 3 = the warp-row streaming smc_upfirdn2d kernels (all default on; 0 selects the older kernel);
 * key 4 = smc_resample_fwd/bwd run the vertical pass first for >= 2x down-sampling;
 * key 6 = the row-tile kernel (32 rows x a run of outputs per CTA, shared-memory staged) for the row-contraction passes of smc_resample_fwd/bwd (default on). */
int smc_synth_config(int key, int value);
int smc_demod_coefs(const float* q, const float* s, int64_t s_stride, float* d, int n, int cin, int cout, void* stream);
/* c_pitch >= c is the channel pitch of the NHWC side (channels c .. c_pitch-1 are left untouched: zero them once). */
int smc_pack_nhwc(const float* x, int64_t x_stride_n, const float* s, int64_t s_stride, void* hi, void* lo, int n, int c,
                  int hw, int c_pitch, void* stream);
int smc_unpack_nchw(const void* x, int x_is_half, float* y, const float* noise, int n, int c, int hw, int c_pitch, void* stream);
/* fsep_host: optional HOST pointer to 8 floats {fy[0..3], fx[0..3]} with fk[fy][fx] == fy[fy] * fx[fx] (separable filter, as the
 * [1,3,3,1] resample filter is): selects the register-window kernels; NULL = generic 4x4 kernel. */
int smc_fir_act(const void* planes, int planes_is_half, int n, int h, int w, int c, const float* fk, const float* fsep_host, const float* noise,
                const float* bias, float alpha, float gain, float clamp, const float* post, int64_t post_stride,
                void* out_raw, void* out_raw_lo, void* out_hi, void* out_lo, void* stream);
/* img[n, j, y, x] = clamp(img[n, j, y, x] + b[j]) + upsample2d(img_prev)[n, j, y, x]   (in place; img holds the fused-ToRGB sums of
 * smc_igemm's rgb_acc; ToRGBLayer bias/clamp and utils.py:45-49; img_prev [N, 3, H/2, W/2] or NULL for the first block).
 * pass_mask (optional, [N, 3, H, W] bytes): 1 where the clamp passes the gradient (|img + b| < clamp, bias_act.cu:136-142), kept for
 * the backward pass so that it need not recompute the ToRGB output.
 * parts > 1: img + q * part_stride (q < parts) are the per-N-tile partial sums of smc_igemm_epilogue::rgb_snt; they are added first, in index order. */
int smc_img_finish(float* img, const float* img_prev, const float* b_rgb, float clamp, const float* fk_up, int n, int h, int w,
                   unsigned char* pass_mask, int parts, int64_t part_stride, void* stream);
int smc_torgb(const void* x_hi, const void* x_lo, int n, int h, int w, int c, const float* w_rgb, const float* s_t,
              int64_t st_stride, float wgain, const float* b_rgb, float clamp, const float* img_prev, const float* fk_up,
              float* img, const float* s_next, int64_t sn_stride, void* xs_hi, void* xs_lo, void* stream);
/* y_lo / gd_lo: optional lo planes (split precision); g_up is fp16, or fp32 when g_up_is_f32 != 0. */
int smc_act_bwd(const void* y, const void* y_lo, int n, int h, int w, int c, const void* g_up, int g_up_is_f32, const float* s_next,
                int64_t sn_stride, const float* g_img, const float* w_rgb, const float* s_t, int64_t st_stride, float wgain,
                const float* b_rgb, float rgb_clamp, const float* gscale, const float* dcoef, const float* noise, const float* bias,
                float alpha, float gain, float clamp, void* gd, void* gd_lo, float* t1, float* r, void* stream);
int smc_fir_bwd(const void* gd, const void* gd_lo, int n, int h, int w, int c, const float* fk, const float* fsep_host, void* planes,
                void* planes_lo, void* stream);
/* t1 [n, cin] is read (the T1 sums of smc_act_bwd) and OVERWRITTEN with the loss-scaled per-sample gradient ds[n, i]; grad_row receives their sum over
 * n in index order.  grad_samples (optional): also writes ds[n, 0..cin) / gscale at grad_samples + n * gs_stride (latent mapper). */
int smc_sgrad_finish(float* t1, const float* r, const float* q, const float* d, const float* s, int64_t s_stride,
                     const float* gscale, float* grad_row, int n, int cin, int cout, float* grad_samples, int64_t gs_stride, void* stream);
int smc_grad_scale(const float* g, int64_t numel, float target, uint32_t* amax_scratch, float* gscale, void* stream);
/* out[i] = g[i] * mask[i] * (*scale): the ToRGB clamp mask kept by smc_img_finish (bytes, 0 / 1) and the loss scale (device scalar, NULL = 1)
 * applied to the incoming image gradient in one pass; feeds smc_act_bwd's g_img / smc_igemm_epilogue::mask_grgb. */
int smc_mask_scale(const float* g, const unsigned char* mask, const float* scale, float* out, int64_t numel, void* stream);

/* ---- unprocess + CLIP glue (vit.cu) --------------------------------------------------------------
 * mean3 / std3 are HOST arrays of 3 floats; all other pointers are device pointers. */
/* denorm_normalize: 0 plain resample; 1 `unprocess` of find_direction.py:49-52 (clamp(x * 127.5 + 128, 0, 255) -> resize -> / 255 -> normalise);
 * 2 the NADA preprocessing of clip_loss_nada.py:86-89 ((x + 1) / 2 without a clamp -> resize -> normalise). */
int smc_resample_fwd(const float* x, float* tmp, float* y, const int* start, const int* count, const float* wgt, int taps,
                     int planes, int in_size, int out_size, int denorm_normalize, const float* mean3, const float* std3, void* stream);
/* mode: 1 / 2 as denorm_normalize above.  unscale: optional DEVICE pointer to the loss scale S carried by g (see smc_clip_loss);
 * gx = d/dx of the unscaled loss.  Tables (device): start / count / wgt[out, taps] = the window of each output index, monotone in the index;
 * oidx / count / wgt[in, taps] = their transpose, every row of oidx a run of CONSECUTIVE output indices (oidx[i][k] = oidx[i][0] + k). */
int smc_resample_bwd(const float* g, const float* x, float* tmp, float* gx, const int* oidx, const int* count, const float* wgt,
                     int taps, int planes, int in_size, int out_size, int mode, const float* std3, const float* unscale, void* stream);
int smc_patchify(const float* img, void* hi, void* lo, int b, int res, int ps, void* stream);
int smc_unpatchify(const float* gp, float* gimg, int b, int res, int ps, void* stream);
int smc_assemble_tokens(const float* patch, const float* cls, const float* pos, float* x0, int b, int t, int wd, void* stream);
int smc_embed_text(const int64_t* text, const float* emb, const float* pos, float* x0, int b, int t, int wd, void* stream);
int smc_layernorm_fwd(const float* x, int64_t in_row_stride, int64_t in_row_offset, const float* w, const float* b, float* y32,
                      void* yhi, void* ylo, float* mean, float* rstd, int64_t rows, int wd, void* stream);
int smc_layernorm_bwd(const float* dy, const float* x, int64_t in_row_stride, int64_t in_row_offset, const float* w,
                      const float* mean, const float* rstd, float* dx, int64_t rows, int wd, int accumulate, void* stream);
int smc_attention_fwd(const float* qkv, void* ohi, void* olo, float* o32, int b, int t, int wd, int heads, int causal, void* stream);
int smc_attention_bwd(const float* qkv, const float* d_o, void* ghi, void* glo, int b, int t, int wd, int heads, int causal,
                      void* stream);
/* Sequences longer than the whole-sequence kernels hold in shared memory (ViT-B/16, clip_loss.py:12-13: 197 tokens): smc_attention_fwd
   switches to query-row blocks by itself; the backward pass is this entry point (two launches: dQ per query-row block, dK/dV per key-row
   block).  stats: caller-allocated fp32 scratch, 2 * b * heads * t words.  head_dim 64 only. */
int smc_attention_bwd_tiled(const float* qkv, const float* d_o, void* ghi, void* glo, float* stats, int b, int t, int wd, int heads,
                            int causal, void* stream);
int smc_quickgelu_fwd(const float* h, void* hi, void* lo, int64_t n, void* stream);
int smc_quickgelu_bwd(const float* dg, const float* h, void* hi, void* lo, int64_t n, void* stream);
int smc_split_rows(const float* x, void* hi, void* lo, int64_t rows, int wd, int rows_per_group, int group_stride, int group_offset,
                   void* stream);
int smc_head_proj(const float* ln, const float* proj, float* out, int b, int wd, int e, void* stream);
int smc_head_proj_bwd(const float* d_e, const float* proj, float* dln, int b, int wd, int e, void* stream);
/* gscale_out (optional, device): d_tgt is multiplied by S = 2^k, max|d_tgt| * S in [target/2, target), and S is stored there
 * (loss scaling for the fp16-operand backward GEMMs; smc_resample_bwd divides it out again). */
/* normalize != 0: the NADA directional form (clip_loss_nada.py:162-168,206-218): e = tgt/|tgt| - src/|src| instead of tgt - src, gradient taken
 * through the normalisation of tgt.  The "global" NADA term (clip_loss_nada.py:220-229) is this entry point with e_src = 0, text = the prompt
 * embedding and coef multiplied by exp(logit_scale) / 100.  text_stride: 0 = one text vector for the batch; e = one target vector PER SAMPLE
 * (the identity loss 1 - <f(edited) / |.|, f(original) / |.|>, id_loss.py:26-39, is this entry point with e_src = 0 and text = f(original)). */
int smc_clip_loss(const float* e_src, const float* e_tgt, const float* text, float* loss_part, float* d_tgt, int n, int e, float coef,
                  float inv_count, float* gscale_out, float gscale_target, int normalize, int text_stride, void* stream);

/* ---- generate_fromS output stage ---------------------------------------------------------------
 * out[n, y, x_off + x, j] = uint8(clamp(img[n, j, y, x] * 127.5 + 128, 0, 255))   (generate_fromS.py:174-175; canvas [N, H, canvas_w, 3],
 * the reference concatenates original | edited along the width, :206). */
int smc_img_to_uint8(const float* img, unsigned char* out, int n, int h, int w, int canvas_w, int x_off, void* stream);

/* ---- one-off preparation of frozen weights (what the reference does implicitly per call in modulated_conv2d: w.flip / transpose /
 * square().sum(), [UPSTREAM] training/networks.py; conv2d_resample.py:125-147) ------------------------------------------------
 * w [n_out, n_in, ntaps] fp32 -> K-major tap matrices as fp16 hi (+ lo = rn(v - hi)) planes: forward [ntaps * n_out_padded, n_in_padded]
 * (row t * n_out_padded + o, column i), dgrad [ntaps * n_in_padded, n_out_padded] (row t * n_in_padded + i, column o), zero padded; and
 * q[o, i] = sum_t w[o, i, t]^2 (un-padded, un-scaled) for smc_demod_coefs.  Any of the three outputs may be NULL.  scale: optional DEVICE
 * scalar multiplied into the planes (a power of two from smc_grad_scale). */
int smc_prepare_weights(const float* w, int n_out, int n_in, int ntaps, int n_out_padded, int n_in_padded, const float* scale,
                        void* fwd_hi, void* fwd_lo, void* bwd_hi, void* bwd_lo, float* q, void* stream);

/* ---- identity-loss glue (id_loss/id_loss.py:18-24, id_loss/helpers.py:58-119) ------------------------------------
 * smc_prelu: y = x > 0 ? x : alpha[c] * x on NCHW fp32 (torch.nn.PReLU(C)); with dy != NULL the input gradient dy * (x > 0 ? 1 : alpha[c]).
 * smc_adaptive_avg_pool: torch.nn.AdaptiveAvgPool2d((oh, ow)) of the window [y0, y0 + hc) x [x0, x0 + wc) of every [h, w] plane
 * (id_loss.py:12-13,19-22: pool to 256, crop [35:223, 32:220], pool to 112); backward != 0: x is dy [planes, oh, ow], y is dx [planes, h, w]. */
int smc_prelu(const float* x, const float* dy, const float* alpha, float* y, int64_t numel, int hw, int c, void* stream);
int smc_adaptive_avg_pool(const float* x, float* y, int64_t planes, int h, int w, int y0, int x0, int hc, int wc, int oh, int ow,
                          int backward, void* stream);

/* ---- latent-mapper glue (latent_mappers.py:12-93, train_latent_mapper.py:131) ----------------------------------------------------------------
 * smc_pixelnorm: PixelNorm over dim 1 of x [b, l, c] (encoder4editing/models/stylegan2/model.py:14-15); dy != NULL: the input gradient.
 * smc_adam_step: torch.optim.Adam without weight decay; bc1 = 1 - beta1^t and bc2_sqrt = sqrt(1 - beta2^t) are computed by the host. */
int smc_pixelnorm(const float* x, const float* dy, float* y, int b, int l, int c, void* stream);
/* y[m, n] = sum_k a[m * sa_m + k * sa_k] * b[n * sb_n + k * sb_k] (+ bias[n]); fp32 in / out, float64 accumulation (one rounding): torch.nn.Linear of
 * the mapper (latent_mappers.py:16) and its three backward products through strides.  y dense [m, n]. */
int smc_matmul_nt_f64acc(const float* a, int64_t sa_m, int64_t sa_k, const float* b, int64_t sb_n, int64_t sb_k, const float* bias, float* y,
                         int m, int n, int k, void* stream);
int smc_adam_step(float* p, const float* g, float* m, float* v, int64_t numel, float lr, float beta1, float beta2, float eps, float bc1,
                  float bc2_sqrt, void* stream);

/* ---- fma.py:15-58 as a stand-alone op (on the fused path the multiply-add is the GEMM epilogue) -----------------------
 * smc_fma: out = a * b + c over the broadcast index space `shape` (4 sizes, leading 1s for lower ranks); stride_* are ELEMENT strides
 * of each operand viewed in that space, 0 on the axes it is broadcast along (fma.py:22, torch.addcmul); out is contiguous.
 * smc_fma_reduce: the "un-broadcast" of the backward (fma.py:36-43,49-58): out[kept] = sum over the reduced axes of x * y (y may be
 * NULL: plain sum); `shape` is the full space, stride_out is 0 exactly on the axes that are summed.  Host pointers for the arrays. */
int smc_fma(const void* a, const void* b, const void* c, void* out, int dtype, const int64_t* shape, const int64_t* stride_a,
            const int64_t* stride_b, const int64_t* stride_c, void* stream);
int smc_fma_reduce(const void* x, const void* y, void* out, int dtype, const int64_t* shape, const int64_t* stride_x,
                   const int64_t* stride_y, const int64_t* stride_out, void* stream);

/* ---- optimiser -----------------------------------------------------------------------------------
 * delta -= lr * (grad * grad_scale + l2_scale * delta)   (SGD, no momentum; L2 term of find_direction.py:190-191) */
int smc_sgd_step(float* delta, const float* grad, int64_t numel, float lr, float grad_scale, float l2_scale, void* stream);
/* the same update with the learning rate read from DEVICE memory: the step can be captured once in a CUDA graph and replayed under the cosine
 * schedule of find_direction.py:298-301 */
int smc_sgd_step_dev(float* delta, const float* grad, int64_t numel, const float* lr_dev, float grad_scale, float l2_scale, void* stream);

#ifdef __cplusplus
}  /* extern "C" */
#endif
