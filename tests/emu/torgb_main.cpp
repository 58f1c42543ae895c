// Driver of the emulated ToRGB kernels of stylemc_b200/csrc/synth.cu (see cuda_emu.h): torgb1_kernel (C / 8 a power of two <= 32) and the
// generic torgb_kernel (the 512-channel blocks), launched as smc_torgb launches them.  Reference (float64):
//   rgb_j = clamp(sum_c w_rgb[j,c] * s_t[n,c] * wgain * x[n,p,c] + b_rgb[j]);   img = upsample2d(img_prev) + rgb   (utils.py:45-49;
//   upsample2d = zero-insert x2, pad [2,1,2,1], correlate with the 4x4 taps fk_up);   xs = x * s_next[n,c] as fp16 hi/lo.
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

static int run(int N, int H, int W, int C, bool with_prev, bool with_xs, bool xs_lo_plane, float clamp) {
  const int hw = H * W, st_stride = 512, sn_stride = 512;
  const size_t ne = (size_t)N * hw * C;
  std::vector<__half> xh(ne), xl(ne), xsh(ne, (__half)NAN), xsl(ne, (__half)NAN);
  std::vector<double> x(ne);
  for (size_t i = 0; i < ne; ++i) {
    const float v = (float)(1.5 * frand());
    xh[i] = (__half)v; xl[i] = (__half)(v - (float)xh[i]);
    x[i] = (double)(float)xh[i] + (double)(float)xl[i];
  }
  auto w_rgb = rnd((size_t)3 * C, 0.6), s_t = rnd((size_t)N * st_stride, 0.5, 1.0), s_next = rnd((size_t)N * sn_stride, 0.5, 1.0), b_rgb = rnd(3, 0.1);
  auto prev = rnd((size_t)N * 3 * (H / 2) * (W / 2)), fk = rnd(16, 0.5, 0.5);
  const float wgain = 1.0f / std::sqrt((float)C);
  std::vector<float> img((size_t)N * 3 * hw, NAN);
  // launch configuration of smc_torgb
  int lpp = C >> 3;
  if (lpp > 32) lpp = 32;
  const int cgn = C >> 3;
  const bool one_group_per_lane = cgn <= 32 && (cgn & (cgn - 1)) == 0;
  int blocks;
  if (one_group_per_lane) {
    int pix_per_block = 8 * (32 / lpp) * 16;
    while (pix_per_block > 8 * (32 / lpp) && (long long)ceil_div(hw, pix_per_block) * N < 4 * 148) pix_per_block >>= 1;
    blocks = ceil_div(hw, pix_per_block) * N;
    emu_launch(blocks, 256, 0, [&] {
      torgb1_kernel(xh.data(), xl.data(), N, H, W, C, w_rgb.data(), s_t.data(), st_stride, wgain, b_rgb.data(), clamp, with_prev ? prev.data() : nullptr, fk.data(),
                    img.data(), s_next.data(), sn_stride, with_xs ? xsh.data() : nullptr, with_xs && xs_lo_plane ? xsl.data() : nullptr, lpp, pix_per_block);
    });
  } else {
    blocks = (int)std::min<long long>(ceil_div_ll((long long)N * hw * lpp, 256), 148LL * 16);
    emu_launch(blocks, 256, 0, [&] {
      torgb_kernel(xh.data(), xl.data(), N, H, W, C, w_rgb.data(), s_t.data(), st_stride, wgain, b_rgb.data(), clamp, with_prev ? prev.data() : nullptr, fk.data(),
                   img.data(), s_next.data(), sn_stride, with_xs ? xsh.data() : nullptr, with_xs && xs_lo_plane ? xsl.data() : nullptr, lpp);
    });
  }
  double e_img = 0, m_img = 0, e_xs = 0;
  int clamped = 0;
  const int h2 = H / 2, w2 = W / 2;
  for (int n = 0; n < N; ++n)
    for (int yy = 0; yy < H; ++yy)
      for (int xx = 0; xx < W; ++xx) {
        const size_t base = ((size_t)n * hw + (size_t)yy * W + xx) * C;
        for (int j = 0; j < 3; ++j) {
          double r = b_rgb[j];
          for (int c = 0; c < C; ++c) r += (double)w_rgb[(size_t)j * C + c] * s_t[(size_t)n * st_stride + c] * wgain * x[base + c];
          if (clamp >= 0) { clamped += std::fabs(r) > clamp; r = std::min(std::max(r, -(double)clamp), (double)clamp); }
          if (with_prev)
            for (int fy = 0; fy < 4; ++fy)
              for (int fx = 0; fx < 4; ++fx) {
                const int ay = yy + fy - 2, ax = xx + fx - 2;       // coordinate in the zero-inserted image
                if (ay < 0 || ax < 0 || (ay & 1) || (ax & 1) || ay / 2 >= h2 || ax / 2 >= w2) continue;
                r += (double)fk[fy * 4 + fx] * prev[(((size_t)n * 3 + j) * h2 + ay / 2) * w2 + ax / 2];
              }
          const double got = img[(((size_t)n * 3 + j) * H + yy) * W + xx];
          e_img = std::max(e_img, std::fabs(got - r)); m_img = std::max(m_img, std::fabs(r));
        }
        if (with_xs)
          for (int c = 0; c < C; ++c) {
            const double want = x[base + c] * s_next[(size_t)n * sn_stride + c];
            const double got = (double)(float)xsh[base + c] + (xs_lo_plane ? (double)(float)xsl[base + c] : 0.0);
            e_xs = std::max(e_xs, std::fabs(got - want) / (xs_lo_plane ? 1.0 : 200.0));      // hi-only: one fp16 rounding
          }
      }
  const bool ok = e_img <= 4e-6 * std::max(m_img, 1.0) && e_xs <= 3e-6 * 2.3 && (clamp < 0 || clamp > 2 || clamped > 0);   // a small clamp must bite
  printf("%s %s N=%d %dx%d C=%d prev=%d xs=%d lo=%d blocks=%d lpp=%d: img err %.2e (max %.2f), xs err %.2e, %d clamped\n", ok ? "ok  " : "FAIL",
         one_group_per_lane ? "torgb1" : "torgb ", N, H, W, C, (int)with_prev, (int)with_xs, (int)xs_lo_plane, blocks, lpp, e_img, m_img, e_xs, clamped);
  return ok ? 0 : 1;
}

int main() {
  srand(29);
  int bad = 0;
  bad += run(2, 10, 14, 64, true, true, true, 0.8f);        // skip image + next styles, the clamp bites, ragged pixel count
  bad += run(1, 6, 6, 256, false, false, false, -1.0f);     // first block: no previous image, last-block style: no xs
  bad += run(1, 8, 4, 32, true, true, false, 256.0f);       // 4 lanes per pixel, hi-only xs
  bad += run(1, 4, 6, 512, true, true, true, 0.8f);         // generic kernel (two channel groups per lane)
  bad += run(2, 4, 4, 320, true, false, false, -1.0f);      // generic kernel, C / 8 = 40
  return bad ? 1 : 0;
}
