// Driver of the emulated act_bwd1_kernel and act_bwd_kernel of stylemc_b200/csrc/synth.cu (see cuda_emu.h): the activation backward of a modulated-conv layer
// on the fused synthesis path, which also accumulates the two reductions of the style gradient (SURVEY.md 8a algebra) with shared-memory
// and global atomics -- the kernel where a race would hide.  Reference (float64), per pixel p of image n and channel c, y = saved activation:
//   m_j[c]  = w_rgb[j,c] * s_t[n,c] * wgain                      ToRGB weights times its styles
//   rgb_j   = sum_c m_j[c] y[c] + b_rgb[j];  grgb_j = |rgb_j| < rgb_clamp ? gscale * g_img[n,j,p] : 0
//   g_y[c]  = g_up[c] * s_next[n,c] + sum_j m_j[c] grgb_j
//   slope   = (y > 0 ? 1 : alpha) * gain;   g_z = |y| < clamp ? g_y * slope : 0;   gd = g_z * dcoef[n,c]
//   T1[n,c] += g_up[c] * y[c];   R[n,c] += g_z * (y / slope - noise[p] - bias[c])
#include "cuda_emu.h"
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
#include "kernels_extracted.inc"
using namespace smc;

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }
static std::vector<float> rnd(size_t n, double scale = 1.0, double shift = 0.0) {
  std::vector<float> v(n);
  for (auto& x : v) x = (float)(scale * frand() + shift);
  return v;
}

template <class TG, bool GENERIC = false>
static int run(int N, int H, int W, int C, bool with_rgb, bool with_up, bool gd_lo_plane, float clamp, float rgb_clamp) {
  const long long hw = (long long)H * W;
  const size_t ne = (size_t)N * hw * C;
  const int sn_stride = 512, st_stride = 512;
  std::vector<__half> yh(ne), yl(ne), gd(ne, (__half)NAN), gdl(ne, (__half)NAN);
  std::vector<double> y(ne), gup(ne, 0.0);
  std::vector<TG> g_up(ne);
  for (size_t i = 0; i < ne; ++i) {
    const float v = (float)(1.5 * frand());
    yh[i] = (__half)v; yl[i] = (__half)(v - (float)yh[i]);
    y[i] = (double)(float)yh[i] + (double)(float)yl[i];
    const float g = (float)frand();
    g_up[i] = (TG)g;
    gup[i] = with_up ? (double)(float)g_up[i] : 0.0;
  }
  auto s_next = rnd((size_t)N * sn_stride, 0.5, 1.0), s_t = rnd((size_t)N * st_stride, 0.5, 1.0), w_rgb = rnd((size_t)3 * C, 0.3), b_rgb = rnd(3, 0.1);
  auto g_img = rnd((size_t)N * 3 * hw), dcoef = rnd((size_t)N * C, 0.3, 0.7), noise = rnd(hw, 0.2), bias = rnd(C, 0.2);
  const float wgain = 1.0f / std::sqrt((float)C), gscale = 4.0f, alpha = 0.2f, gain = 1.41421356f;
  std::vector<float> T1 = rnd((size_t)N * C), R = rnd((size_t)N * C);
  const std::vector<float> T10 = T1, R0 = R;
  // launch configuration of smc_act_bwd
  int lpp = C >> 3;
  if (lpp > 32) lpp = 32;
  int pix_per_block = 8 * (32 / lpp) * 16;
  while (pix_per_block > 8 * (32 / lpp) && ceil_div_ll(hw, pix_per_block) * N < 2 * 148) pix_per_block >>= 1;
  const int blocks = (int)(ceil_div_ll(hw, pix_per_block) * N);
  emu_launch(blocks, 256, 2 * C * sizeof(float), [&] {
    auto kernel = GENERIC ? act_bwd_kernel<TG> : act_bwd1_kernel<TG>;      // same arguments; smc_act_bwd picks by C / 8 being a power of two <= 32
    kernel(yh.data(), yl.data(), N, H, W, C, with_up ? g_up.data() : nullptr, s_next.data(), sn_stride, with_rgb ? g_img.data() : nullptr, w_rgb.data(),
           s_t.data(), st_stride, wgain, b_rgb.data(), rgb_clamp, &gscale, dcoef.data(), noise.data(), bias.data(), alpha, gain, clamp, gd.data(),
           gd_lo_plane ? gdl.data() : nullptr, T1.data(), R.data(), lpp, pix_per_block);
  });
  double e_gd = 0, m_gd = 0, e_t1 = 0, e_r = 0, m_t1 = 0, m_r = 0;
  int masked = 0, rgb_masked = 0;
  std::vector<double> t1((size_t)N * C, 0.0), r((size_t)N * C, 0.0);
  for (int n = 0; n < N; ++n)
    for (long long p = 0; p < hw; ++p) {
      const size_t base = ((size_t)n * hw + p) * C;
      double grgb[3] = {0, 0, 0};
      if (with_rgb)
        for (int j = 0; j < 3; ++j) {
          double rgb = b_rgb[j];
          for (int c = 0; c < C; ++c) rgb += (double)w_rgb[(size_t)j * C + c] * s_t[(size_t)n * st_stride + c] * wgain * y[base + c];
          const bool pass = rgb_clamp < 0 || std::fabs(rgb) < rgb_clamp;
          rgb_masked += !pass;
          grgb[j] = pass ? (double)gscale * g_img[((size_t)n * 3 + j) * hw + p] : 0.0;
        }
      for (int c = 0; c < C; ++c) {
        double gy = gup[base + c] * s_next[(size_t)n * sn_stride + c];
        if (with_rgb)
          for (int j = 0; j < 3; ++j) gy += (double)w_rgb[(size_t)j * C + c] * s_t[(size_t)n * st_stride + c] * wgain * grgb[j];
        const double yy = y[base + c], slope = (yy > 0 ? 1.0 : alpha) * gain;
        const bool pass = clamp < 0 || std::fabs(yy) < clamp;
        masked += !pass;
        double gz = pass ? gy * slope : 0.0;
        const double want = gz * dcoef[(size_t)n * C + c];
        const double got = (double)(float)gd[base + c] + (gd_lo_plane ? (double)(float)gdl[base + c] : 0.0);
        e_gd = std::max(e_gd, std::fabs(got - want)); m_gd = std::max(m_gd, std::fabs(want));
        if (!gd_lo_plane) gz = (double)(float)(__half)(float)want / dcoef[(size_t)n * C + c];      // the kernel reduces what the dgrad GEMM will see
        r[(size_t)n * C + c] += gz * (yy / slope - noise[p] - bias[c]);
        t1[(size_t)n * C + c] += gup[base + c] * yy;
      }
    }
  for (size_t i = 0; i < t1.size(); ++i) {
    e_t1 = std::max(e_t1, std::fabs((double)T1[i] - (T10[i] + t1[i]))); m_t1 = std::max(m_t1, std::fabs(t1[i]));
    e_r = std::max(e_r, std::fabs((double)R[i] - (R0[i] + r[i]))); m_r = std::max(m_r, std::fabs(r[i]));
  }
  const double tol_gd = gd_lo_plane ? 3e-6 : 6e-4, tol_r = gd_lo_plane ? 2e-5 : 2e-3;
  const bool ok = e_gd <= tol_gd * m_gd && e_t1 <= 2e-5 * std::max(m_t1, 1.0) && e_r <= tol_r * std::max(m_r, 1.0) && (clamp < 0 || clamp > 2 || masked > 0) &&
                  (!with_rgb || rgb_clamp < 0 || rgb_masked > 0);      // a small clamp must bite
  printf("%s act_bwd%s<%s> N=%d HW=%lld C=%d rgb=%d up=%d lo=%d blocks=%d lpp=%d: gd err %.2e (max %.2f), T1 err %.2e (max %.1f), R err %.2e (max %.1f), %d + %d masked\n",
         ok ? "ok  " : "FAIL", GENERIC ? "" : "1", sizeof(TG) == 2 ? "half" : "float", N, hw, C, (int)with_rgb, (int)with_up, (int)gd_lo_plane, blocks, lpp, e_gd, m_gd, e_t1, m_t1, e_r,
         m_r, masked, rgb_masked);
  return ok ? 0 : 1;
}

int main() {
  srand(23);
  int bad = 0;
  bad += run<float>(2, 13, 11, 64, true, true, true, 1.0f, 0.05f);     // both gradient sources, both clamps bite, ragged pixel count
  bad += run<__half>(1, 9, 7, 128, false, true, false, -1.0f, -1.0f);  // fp16 incoming gradient, hi-only gd, no ToRGB branch, no clamp
  bad += run<float>(1, 8, 8, 32, true, false, true, 256.0f, -1.0f);    // ToRGB gradient only (g_up absent), 4 lanes per pixel
  bad += run<float, true>(1, 5, 7, 512, true, true, true, 1.0f, 0.1f);  // the 512-channel layers: generic kernel, two channel groups per lane
  bad += run<__half, true>(1, 6, 6, 320, false, true, false, -1.0f, -1.0f); // C / 8 = 40: not a power of two, lane groups still tile a warp (lpp = 32)
  return bad ? 1 : 0;
}
