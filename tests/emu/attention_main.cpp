// Driver of the emulated attention kernels (see cuda_emu.h).  kernels_extracted.inc is cut out of stylemc_b200/csrc/vit.cu by
// tests/test_kernels_emu.py (store_split, warp_sum, the attention kernels and attention_block_rows, verbatim except for the
// `extern __shared__` line).  Compares against a float64 restatement of softmax attention and its gradient.
#include "cuda_emu.h"
namespace smc { int g_attention_tiled = 0; }
#include "kernels_extracted.inc"
using namespace smc;

struct Case { int B, T, heads, causal, tiled; };

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }
static float joined(const std::vector<__half>& hi, const std::vector<__half>& lo, size_t i) { return (float)hi[i] + (float)lo[i]; }

static int run(const Case& c) {
  const int B = c.B, T = c.T, H = c.heads, hd = 64, Wd = H * hd;
  std::vector<float> qkv((size_t)B * T * 3 * Wd), dO((size_t)B * T * Wd);
  for (auto& v : qkv) v = (float)(1.5 * frand());
  for (auto& v : dO) v = (float)frand();
  // float64 reference
  std::vector<double> O((size_t)B * T * Wd), dQKV((size_t)B * T * 3 * Wd, 0.0);
  const double scale = 1.0 / std::sqrt((double)hd);
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h) {
      auto at = [&](int t, int which, int d) { return (double)qkv[((size_t)b * T + t) * 3 * Wd + which * Wd + h * hd + d]; };
      std::vector<double> P((size_t)T * T), dP((size_t)T * T);
      for (int r = 0; r < T; ++r) {
        double m = -INFINITY;
        for (int cc = 0; cc < T; ++cc) {
          double s = 0;
          for (int d = 0; d < hd; ++d) s += at(r, 0, d) * at(cc, 1, d);
          s = (c.causal && cc > r) ? -INFINITY : s * scale;
          P[(size_t)r * T + cc] = s;
          m = std::max(m, s);
        }
        double sum = 0;
        for (int cc = 0; cc < T; ++cc) { P[(size_t)r * T + cc] = std::exp(P[(size_t)r * T + cc] - m); sum += P[(size_t)r * T + cc]; }
        for (int cc = 0; cc < T; ++cc) P[(size_t)r * T + cc] /= sum;
        for (int d = 0; d < hd; ++d) {
          double o = 0;
          for (int cc = 0; cc < T; ++cc) o += P[(size_t)r * T + cc] * at(cc, 2, d);
          O[((size_t)b * T + r) * Wd + h * hd + d] = o;
        }
        double dot = 0;
        for (int cc = 0; cc < T; ++cc) {
          double dp = 0;
          for (int d = 0; d < hd; ++d) dp += (double)dO[((size_t)b * T + r) * Wd + h * hd + d] * at(cc, 2, d);
          dP[(size_t)r * T + cc] = dp;
          dot += dp * P[(size_t)r * T + cc];
        }
        for (int cc = 0; cc < T; ++cc) dP[(size_t)r * T + cc] = P[(size_t)r * T + cc] * (dP[(size_t)r * T + cc] - dot) * scale;   // dS
      }
      for (int t = 0; t < T; ++t)
        for (int d = 0; d < hd; ++d) {
          double dq = 0, dk = 0, dv = 0;
          for (int cc = 0; cc < T; ++cc) {
            dq += dP[(size_t)t * T + cc] * at(cc, 1, d);
            dk += dP[(size_t)cc * T + t] * at(cc, 0, d);
            dv += P[(size_t)cc * T + t] * (double)dO[((size_t)b * T + cc) * Wd + h * hd + d];
          }
          const size_t o = ((size_t)b * T + t) * 3 * Wd + h * hd + d;
          dQKV[o] = dq; dQKV[o + Wd] = dk; dQKV[o + 2 * Wd] = dv;
        }
    }
  // emulated kernels, launched the way smc_attention_fwd / smc_attention_bwd(_tiled) launch them
  std::vector<__half> ohi(O.size()), olo(O.size()), ghi(dQKV.size()), glo(dQKV.size());
  std::vector<float> o32(O.size(), NAN), stats((size_t)2 * B * H * T, NAN);
  for (auto& v : ohi) v = (__half)NAN;
  for (auto& v : ghi) v = (__half)NAN;
  const float* qp = qkv.data();
  const float* gp = dO.data();
  int blocks = 1;
  if (c.tiled) {
    g_attention_tiled = c.tiled;
    size_t smem = 0;
    const int qb = attention_block_rows(T, 65 + T + 1, 0, &smem);
    if (!qb || smem > 200 * 1024) { printf("FAIL no forward block size\n"); return 1; }
    const int nqb = (T + qb - 1) / qb;
    emu_launch(B * H * nqb, 256, smem, [&] { attention_fwd_rows_kernel(qp, ohi.data(), olo.data(), o32.data(), T, Wd, H, c.causal, qb, nqb); });
    const int rb = attention_block_rows(T, 2 * 65 + 2 * (T + 1), 2 * T, &smem);
    if (!rb || smem > 200 * 1024) { printf("FAIL no backward block size\n"); return 1; }
    const int nb = (T + rb - 1) / rb;
    blocks = nb;
    emu_launch(B * H * nb, 256, smem, [&] { attention_bwd_q_kernel(qp, gp, ghi.data(), glo.data(), stats.data(), T, Wd, H, c.causal, rb, nb); });
    emu_launch(B * H * nb, 256, smem, [&] { attention_bwd_kv_kernel(qp, gp, ghi.data(), glo.data(), stats.data(), T, Wd, H, c.causal, rb, nb); });
  } else {   // the whole-sequence kernels that are verified on the GPU: validates the emulator itself
    emu_launch(B * H, 256, (size_t)(3 * T * 65 + T * (T + 1)) * 4, [&] { attention_fwd2_kernel(qp, ohi.data(), olo.data(), o32.data(), T, Wd, H, c.causal); });
    emu_launch(B * H, 256, (size_t)(4 * T * 65 + 2 * T * (T + 1)) * 4, [&] { attention_bwd2_kernel(qp, gp, ghi.data(), glo.data(), T, Wd, H, c.causal); });
  }
  double eo = 0, eo32 = 0, eg = 0, mo = 0, mg = 0;
  for (size_t i = 0; i < O.size(); ++i) {
    eo = std::max(eo, std::fabs((double)joined(ohi, olo, i) - O[i]));
    eo32 = std::max(eo32, std::fabs((double)o32[i] - O[i]));
    mo = std::max(mo, std::fabs(O[i]));
  }
  for (size_t i = 0; i < dQKV.size(); ++i) {
    eg = std::max(eg, std::fabs((double)joined(ghi, glo, i) - dQKV[i]));
    mg = std::max(mg, std::fabs(dQKV[i]));
  }
  const bool ok = eo <= 2e-6 * mo + 1e-7 && eo32 <= 2e-6 * mo && eg <= 4e-6 * mg + 1e-7;     // NaN (an element never written) fails too
  printf("%s B=%d T=%d heads=%d causal=%d tiled=%d blocks=%d  O err %.2e (max %.2f)  o32 err %.2e  dQKV err %.2e (max %.2f)\n", ok ? "ok  " : "FAIL",
         B, T, H, c.causal, c.tiled, blocks, eo, mo, eo32, eg, mg);
  return ok ? 0 : 1;
}

int main() {
  srand(7);
  const Case cases[] = {{1, 50, 2, 0, 0},      // emulator check on the GPU-verified kernels
                        {1, 50, 2, 0, 16},     // 4 row blocks (16, 16, 16, 2)
                        {1, 77, 1, 1, 32},     // causal, 3 blocks (32, 32, 13)
                        {1, 37, 1, 0, 1},      // one block, rows not a multiple of 4
                        {1, 197, 1, 0, 1}};     // ViT-B/16: block size chosen by the shared-memory budget, as in production
  int bad = 0;
  for (const Case& c : cases) bad += run(c);
  return bad ? 1 : 0;
}
