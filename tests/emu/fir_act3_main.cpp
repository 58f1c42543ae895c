// Driver of the emulated fir_act3_kernel of stylemc_b200/csrc/synth.cu (see cuda_emu.h): the conv0 tail of the fused synthesis path
// (4x4 separable FIR over the four parity planes of the stride-2 transposed conv + noise + bias + lrelu + gain + clamp, written as fp16
// hi/lo planes raw and multiplied by the next layer's styles).  It is the heaviest HBM-bound kernel of a step.  Reference: float64
// upfirdn2d(pad 1) + bias_act restated over the (2H+1) x (2W+1) transposed-conv output t, t[2a+r, 2b+c] = P[r][c][n][a][b][:].
// Plane entries that lie outside t (row 2H+1, column 2W+1) are NaN here: a kernel that read them would poison its output.
#include "cuda_emu.h"
// the two inline-PTX helpers of synth.cu, restated for the host
static inline float fmax3(float a, float b, float c) { return std::fmax(std::fmax(a, b), c); }
static inline void hsubf2(uint32_t hh, float v0, float v1, float& d0, float& d1) {
  __half h[2];
  __builtin_memcpy(h, &hh, 4);
  d0 = (float)h[0] - v0;
  d1 = (float)h[1] - v1;
}
#include "kernels_extracted.inc"
using namespace smc;

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

template <int C, int SAVE, bool NOISE>
static int run(int N, int H, int W, float clamp) {
  constexpr int JT = 16, KCOLS = 256 / (C / 4);
  const size_t plane_sz = (size_t)N * (H + 1) * (W + 1) * C;
  std::vector<float> planes(4 * plane_sz), noise((size_t)4 * H * W), bias(C), post((size_t)N * 64);
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 2; ++c)
      for (int n = 0; n < N; ++n)
        for (int a = 0; a <= H; ++a)
          for (int b = 0; b <= W; ++b)
            for (int ch = 0; ch < C; ++ch) {
              const bool inside = 2 * a + r <= 2 * H && 2 * b + c <= 2 * W;
              planes[(size_t)(r * 2 + c) * plane_sz + (((size_t)n * (H + 1) + a) * (W + 1) + b) * C + ch] = inside ? (float)frand() : NAN;
            }
  for (auto& v : noise) v = (float)(0.3 * frand());
  for (auto& v : bias) v = (float)(0.2 * frand());
  for (auto& v : post) v = (float)(1.0 + 0.5 * frand());
  const float4 fy{0.125f, 0.375f, 0.375f, 0.125f}, fx{0.25f, 0.75f, 0.70f, 0.30f};      // asymmetric on purpose: catches a flipped tap order
  const float alpha = 0.2f, gain = 1.41421356f;
  const size_t out_n = (size_t)N * 2 * H * 2 * W * C;
  std::vector<__half> raw(out_n, (__half)NAN), raw_lo(out_n, (__half)NAN), hi(out_n, (__half)NAN), lo(out_n, (__half)NAN);
  emu_dim3 grid;
  grid.x = (W + KCOLS - 1) / KCOLS; grid.y = H / JT; grid.z = N;
  emu_launch(grid, 256, 0, [&] {
    fir_act3_kernel<C, JT, SAVE, NOISE, 1>(planes.data(), N, H, W, fy, fx, NOISE ? noise.data() : nullptr, bias.data(), alpha, gain, clamp, post.data(), 64,
                                           SAVE ? raw.data() : nullptr, SAVE == 2 ? raw_lo.data() : nullptr, hi.data(), lo.data());
  });
  const double fyd[4] = {fy.x, fy.y, fy.z, fy.w}, fxd[4] = {fx.x, fx.y, fx.z, fx.w};
  auto t = [&](int n, int ty, int tx, int ch) -> double {
    if (ty < 0 || tx < 0 || ty > 2 * H || tx > 2 * W) return 0.0;
    return planes[(size_t)((ty & 1) * 2 + (tx & 1)) * plane_sz + (((size_t)n * (H + 1) + ty / 2) * (W + 1) + tx / 2) * C + ch];
  };
  double e_raw = 0, e_post = 0, m = 0;
  int clamped = 0;
  for (int n = 0; n < N; ++n)
    for (int oy = 0; oy < 2 * H; ++oy)
      for (int ox = 0; ox < 2 * W; ++ox)
        for (int ch = 0; ch < C; ++ch) {
          double z = 0;
          for (int dy = 0; dy < 4; ++dy)
            for (int dx = 0; dx < 4; ++dx) z += fyd[dy] * fxd[dx] * t(n, oy - 1 + dy, ox - 1 + dx, ch);
          z += (NOISE ? noise[(size_t)oy * 2 * W + ox] : 0.0) + bias[ch];
          double y = (z > 0 ? z : z * alpha) * gain;
          if (clamp >= 0) { if (std::fabs(y) >= clamp) ++clamped; y = std::min(std::max(y, -(double)clamp), (double)clamp); }
          const size_t o = (((size_t)n * 2 * H + oy) * 2 * W + ox) * C + ch;
          if (SAVE == 2) e_raw = std::max(e_raw, std::fabs((double)(float)raw[o] + (float)raw_lo[o] - y));
          if (SAVE == 1) e_raw = std::max(e_raw, std::fabs((double)(float)raw[o] - (double)(float)(__half)(float)y) > 1e-3 * std::fabs(y) + 1e-6 ? 1.0 : 0.0);
          e_post = std::max(e_post, std::fabs((double)(float)hi[o] + (float)lo[o] - y * post[(size_t)n * 64 + ch]));
          m = std::max(m, std::fabs(y));
        }
  const bool ok = e_raw <= 3e-6 * m && e_post <= 3e-6 * 1.5 * m && (clamp < 0 || clamp > 2 || clamped > 0);   // a small clamp must bite
  printf("%s fir_act3<C=%d, SAVE=%d, NOISE=%d> N=%d H=%d W=%d clamp=%g: raw err %.2e, styled err %.2e (max |y| %.2f, %d clamped)\n", ok ? "ok  " : "FAIL", C,
         SAVE, (int)NOISE, N, H, W, clamp, e_raw, e_post, m, clamped);
  return ok ? 0 : 1;
}

int main() {
  srand(17);
  int bad = 0;
  bad += run<32, 2, true>(2, 16, 40, 1.0f);      // two column blocks (KCOLS = 32), ragged in W; the clamp bites
  bad += run<64, 1, false>(1, 32, 16, -1.0f);    // two row blocks, hi-only save, no noise, no clamp
  bad += run<32, 0, true>(1, 16, 5, 256.0f);     // narrower than one column block, no raw output
  return bad ? 1 : 0;
}
