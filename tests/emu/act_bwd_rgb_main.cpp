// Driver of the emulated act_bwd_rgb_kernel of stylemc_b200/csrc/synth.cu (see cuda_emu.h): the activation backward of the LAST block, whose
// only consumer is ToRGB (1024-px conv1 in the benchmark: the largest activation of the step).  Reference (float64):
//   gd[n,p,c] = dcoef[n,c] * slope(y) * sum_j w_rgb[j,c] * s_t[n,c] * wgain * g[n,j,p],  slope = (y > 0 ? 1 : alpha) * gain,  0 where |y| >= clamp.
#include "cuda_emu.h"
static inline void hsubf2(uint32_t hh, float v0, float v1, float& d0, float& d1) {
  __half h[2];
  __builtin_memcpy(h, &hh, 4);
  d0 = (float)h[0] - v0;
  d1 = (float)h[1] - v1;
}
#include "kernels_extracted.inc"
using namespace smc;

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }
static std::vector<float> rnd(size_t n, double scale = 1.0, double shift = 0.0) {
  std::vector<float> v(n);
  for (auto& x : v) x = (float)(scale * frand() + shift);
  return v;
}

template <int C, bool YLO, bool LO>
static int run(int N, long long HW, float clamp) {
  const size_t ne = (size_t)N * HW * C;
  const int st_stride = 512;
  std::vector<__half> yh(ne), yl(ne), gd(ne, (__half)NAN), gdl(ne, (__half)NAN);
  std::vector<double> y(ne);
  for (size_t i = 0; i < ne; ++i) {
    const float v = (float)(1.5 * frand());
    yh[i] = (__half)v; yl[i] = (__half)(v - (float)yh[i]);
    y[i] = (double)(float)yh[i] + (YLO ? (double)(float)yl[i] : 0.0);
  }
  auto g = rnd((size_t)N * 3 * HW), w_rgb = rnd((size_t)3 * C, 0.6), s_t = rnd((size_t)N * st_stride, 0.5, 1.0), dcoef = rnd((size_t)N * C, 0.3, 0.7);
  const float wgain = 1.0f / std::sqrt((float)C), alpha = 0.2f, gain = 1.41421356f;
  const int pix_per_block = 8 * (32 / (C / 8)) * 4 * 8;       // as launch_act_bwd_rgb
  const int blocks = (int)(((HW + pix_per_block - 1) / pix_per_block) * N);
  emu_launch(blocks, 256, 0, [&] {
    act_bwd_rgb_kernel<C, YLO, LO>(yh.data(), YLO ? yl.data() : nullptr, HW, g.data(), w_rgb.data(), s_t.data(), st_stride, wgain, dcoef.data(), alpha, gain, clamp,
                                   gd.data(), LO ? gdl.data() : nullptr, pix_per_block);
  });
  double err = 0, m = 0;
  int masked = 0;
  for (int n = 0; n < N; ++n)
    for (long long p = 0; p < HW; ++p)
      for (int c = 0; c < C; ++c) {
        const size_t i = ((size_t)n * HW + p) * C + c;
        double s = 0;
        for (int j = 0; j < 3; ++j) s += (double)w_rgb[(size_t)j * C + c] * s_t[(size_t)n * st_stride + c] * wgain * g[((size_t)n * 3 + j) * HW + p];
        const bool pass = clamp < 0 || std::fabs(y[i]) < clamp;
        masked += !pass;
        const double want = pass ? dcoef[(size_t)n * C + c] * (y[i] > 0 ? 1.0 : alpha) * gain * s : 0.0;
        const double got = (double)(float)gd[i] + (LO ? (double)(float)gdl[i] : 0.0);
        err = std::max(err, std::fabs(got - want)); m = std::max(m, std::fabs(want));
      }
  const bool ok = err <= (LO ? 3e-6 : 6e-4) * m && (clamp < 0 || clamp > 2 || masked > 0);
  printf("%s act_bwd_rgb<C=%d, YLO=%d, LO=%d> N=%d HW=%lld blocks=%d: max err %.2e (max |ref| %.2f), %d masked\n", ok ? "ok  " : "FAIL", C, (int)YLO, (int)LO, N, HW,
         blocks, err, m, masked);
  return ok ? 0 : 1;
}

int main() {
  srand(31);
  int bad = 0;
  bad += run<32, true, true>(2, 2500, 1.0f);      // the benchmark case (C = 32, split planes): two blocks per image, ragged tail, the clamp bites
  bad += run<64, false, false>(1, 300, 256.0f);   // hi-only activation and gradient
  bad += run<128, true, false>(1, 77, -1.0f);     // fewer pixels than one pass of the pipeline
  return bad ? 1 : 0;
}
