// Driver of the emulated fir_bwd3_kernel of stylemc_b200/csrc/synth.cu (see cuda_emu.h): the transpose of the conv0 FIR on the fused
// synthesis path -- gradient planes dt[2a+r, 2b+q] (four parity planes, fp16 hi/lo) from the gradient gd of the (2H) x (2W) activation.
// Reference in float64: dt[ty, tx] = sum_{dy,dx} fy[dy] fx[dx] gd[ty + 1 - dy, tx + 1 - dx]  (the adjoint of fir_act3's FIR with pad 1);
// plane entries outside the (2H+1) x (2W+1) grid must be written as zeros (the dgrad GEMM reads whole planes).
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

template <int C, bool LO>
static int run(int N, int H, int W) {
  constexpr int JT = 16, KCOLS = 256 / (C / 4);
  const size_t gn = (size_t)N * 2 * H * 2 * W * C, plane_sz = (size_t)N * (H + 1) * (W + 1) * C;
  std::vector<__half> gd(gn), gd_lo(gn), planes(4 * plane_sz, (__half)NAN), planes_lo(4 * plane_sz, (__half)NAN);
  std::vector<double> g(gn);
  for (size_t i = 0; i < gn; ++i) {
    const float v = (float)frand();
    gd[i] = (__half)v;
    gd_lo[i] = LO ? (__half)(v - (float)gd[i]) : (__half)0.0f;
    g[i] = (double)(float)gd[i] + (LO ? (double)(float)gd_lo[i] : 0.0);
  }
  const float4 fy{0.125f, 0.375f, 0.375f, 0.125f}, fx{0.25f, 0.75f, 0.70f, 0.30f};
  emu_dim3 grid;
  grid.x = (W + 1 + KCOLS - 1) / KCOLS; grid.y = (H + 1 + JT - 1) / JT; grid.z = N;      // as launch_fir_bwd3
  emu_launch(grid, 256, 0, [&] { fir_bwd3_kernel<C, JT, LO, 3>(gd.data(), LO ? gd_lo.data() : nullptr, N, H, W, fy, fx, planes.data(), LO ? planes_lo.data() : nullptr); });
  const double fyd[4] = {fy.x, fy.y, fy.z, fy.w}, fxd[4] = {fx.x, fx.y, fx.z, fx.w};
  auto G = [&](int n, int oy, int ox, int ch) -> double {
    if (oy < 0 || ox < 0 || oy >= 2 * H || ox >= 2 * W) return 0.0;
    return g[(((size_t)n * 2 * H + oy) * 2 * W + ox) * C + ch];
  };
  double err = 0, m = 0;
  for (int r = 0; r < 2; ++r)
    for (int q = 0; q < 2; ++q)
      for (int n = 0; n < N; ++n)
        for (int a = 0; a <= H; ++a)
          for (int b = 0; b <= W; ++b)
            for (int ch = 0; ch < C; ++ch) {
              const int ty = 2 * a + r, tx = 2 * b + q;
              double want = 0;
              if (ty <= 2 * H && tx <= 2 * W)
                for (int dy = 0; dy < 4; ++dy)
                  for (int dx = 0; dx < 4; ++dx) want += fyd[dy] * fxd[dx] * G(n, ty + 1 - dy, tx + 1 - dx, ch);
              const size_t o = (size_t)(r * 2 + q) * plane_sz + (((size_t)n * (H + 1) + a) * (W + 1) + b) * C + ch;
              const double got = (double)(float)planes[o] + (LO ? (double)(float)planes_lo[o] : 0.0);
              err = std::max(err, std::fabs(got - want));
              m = std::max(m, std::fabs(want));
            }
  const bool ok = err <= (LO ? 3e-6 : 6e-4) * m;          // hi-only output: one fp16 rounding
  printf("%s fir_bwd3<C=%d, LO=%d> N=%d H=%d W=%d: max err %.2e (max |ref| %.2f)\n", ok ? "ok  " : "FAIL", C, (int)LO, N, H, W, err, m);
  return ok ? 0 : 1;
}

int main() {
  srand(19);
  int bad = 0;
  bad += run<32, true>(2, 16, 40);       // two column blocks and two row blocks (H + 1 = 17 > JT), ragged in both
  bad += run<64, false>(1, 8, 5);        // hi-only gradient, narrower than one column block
  bad += run<128, true>(1, 3, 9);
  return bad ? 1 : 0;
}
