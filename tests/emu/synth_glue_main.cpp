// Driver of the emulated shared-memory glue kernels of stylemc_b200/csrc/synth.cu (see cuda_emu.h): demodulation coefficients, the final
// style-gradient assembly, and the NCHW <-> NHWC tile transposes, each against a float64 restatement.
// kernels_extracted.inc is cut out of synth.cu by tests/test_kernels_emu.py.
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
static int report(const char* name, double err, double ref_max, double rel_tol) {
  const bool ok = err <= rel_tol * ref_max;
  printf("%s %-28s max err %.2e (max |ref| %.3g)\n", ok ? "ok  " : "FAIL", name, err, ref_max);
  return ok ? 0 : 1;
}
static emu_dim3 dim3(unsigned x, unsigned y = 1, unsigned z = 1) { emu_dim3 d; d.x = x; d.y = y; d.z = z; return d; }

static int test_demod_and_sgrad() {
  const int N = 3, cin = 72, cout = 45, s_stride = 100;     // cin, cout not multiples of 32; styles are rows of a wider tensor
  auto q = rnd((size_t)cout * cin, 0.5, 0.6), s = rnd((size_t)N * s_stride, 1.0, 1.0);
  std::vector<float> d((size_t)N * cout, NAN);
  // launched as smc_demod_coefs does: grid (min(ceil(cout / 8), 64), N), cin floats of shared memory
  emu_launch(dim3(std::min(ceil_div(cout, 8), 64), N), 256, cin * sizeof(float), [&] { demod_kernel(q.data(), s.data(), s_stride, d.data(), cin, cout); });
  double e1 = 0, m1 = 0;
  std::vector<double> dref((size_t)N * cout);
  for (int n = 0; n < N; ++n)
    for (int o = 0; o < cout; ++o) {
      double a = 0;
      for (int i = 0; i < cin; ++i) a += (double)q[(size_t)o * cin + i] * s[(size_t)n * s_stride + i] * s[(size_t)n * s_stride + i];
      dref[(size_t)n * cout + o] = 1.0 / std::sqrt(a + 1e-8);
      e1 = std::max(e1, std::fabs(d[(size_t)n * cout + o] - dref[(size_t)n * cout + o])); m1 = std::max(m1, dref[(size_t)n * cout + o]);
    }
  // dL/ds[i] += sum_n (T1[n,i] - s[n,i] * sum_o q[o,i] d[n,o]^2 R[n,o]) / gscale      (SURVEY.md 8a style-gradient algebra; R carries d * dL/dd)
  auto T1 = rnd((size_t)N * cin), R = rnd((size_t)N * cout);
  std::vector<float> grad = rnd(cin);
  const std::vector<float> grad0 = grad;
  const float gscale = 8.0f;
  // launched as smc_sgrad_finish does: per-sample kernel on grid (ceil(cin / 32), N), which overwrites T1 with ds, then the ordered sum over n
  const std::vector<float> T1in = T1;
  std::vector<float> each((size_t)N * 80, NAN);
  emu_launch(dim3(ceil_div(cin, 32), N), 256, (cout + 256) * sizeof(float),
             [&] { sgrad_sample_kernel(T1.data(), R.data(), q.data(), d.data(), s.data(), s_stride, &gscale, cin, cout, each.data(), 80); });
  emu_launch(ceil_div(cin, 128), 128, 0, [&] { sgrad_sum_kernel(T1.data(), &gscale, grad.data(), N, cin); });
  double e2 = 0, m2 = 0;
  for (int i = 0; i < cin; ++i) {
    double acc = 0;
    for (int n = 0; n < N; ++n) {
      double t = 0;
      for (int o = 0; o < cout; ++o) t += (double)q[(size_t)o * cin + i] * d[(size_t)n * cout + o] * d[(size_t)n * cout + o] * R[(size_t)n * cout + o];
      const double ds = T1in[(size_t)n * cin + i] - s[(size_t)n * s_stride + i] * t;
      e2 = std::max(e2, std::fabs(each[(size_t)n * 80 + i] - ds / gscale));      // the per-sample gradient (latent mapper)
      acc += ds;
    }
    const double want = grad0[i] + acc / gscale;
    e2 = std::max(e2, std::fabs(grad[i] - want)); m2 = std::max(m2, std::fabs(want));
  }
  return report("demod_kernel", e1, m1, 2e-6) + report("sgrad_sample/sum_kernel", e2, m2, 4e-6);
}

static int test_transposes() {
  const int N = 2, C = 40, HW = 75, CP = 48, s_stride = 64;   // ragged tiles in both directions, channel pitch > C
  auto x = rnd((size_t)N * C * HW), s = rnd((size_t)N * s_stride, 0.5, 1.0), noise = rnd(HW);
  std::vector<__half> hi((size_t)N * HW * CP, (__half)5.0f), lo((size_t)N * HW * CP, (__half)5.0f);
  emu_launch(dim3(ceil_div(HW, 32), ceil_div(C, 32), N), 256, 0,
             [&] { pack_nhwc_kernel(x.data(), (long long)C * HW, s.data(), s_stride, hi.data(), lo.data(), C, HW, CP); });
  double e1 = 0, pad = 0;
  for (int n = 0; n < N; ++n)
    for (int p = 0; p < HW; ++p)
      for (int c = 0; c < CP; ++c) {
        const size_t o = ((size_t)n * HW + p) * CP + c;
        if (c < C) e1 = std::max(e1, std::fabs((double)(float)hi[o] + (float)lo[o] - (double)x[((size_t)n * C + c) * HW + p] * s[(size_t)n * s_stride + c]));
        else pad = std::max(pad, std::fabs((float)hi[o] - 5.0) + std::fabs((float)lo[o] - 5.0));      // channels C .. CP-1 are left untouched
      }
  // b4 `const` input: x_stride_n = 0 broadcasts one [C, HW] block over the batch, no style multiply
  std::vector<__half> bh((size_t)N * HW * C, (__half)NAN);
  emu_launch(dim3(ceil_div(HW, 32), ceil_div(C, 32), N), 256, 0, [&] { pack_nhwc_kernel(x.data(), 0, nullptr, 0, bh.data(), nullptr, C, HW, C); });
  double e2 = 0;
  for (int n = 0; n < N; ++n)
    for (int p = 0; p < HW; ++p)
      for (int c = 0; c < C; ++c) e2 = std::max(e2, std::fabs((double)(float)bh[((size_t)n * HW + p) * C + c] - (double)(float)(__half)x[(size_t)c * HW + p]));
  // and back: NHWC (fp16 or fp32) -> NCHW fp32 with the noise plane added
  std::vector<float> y((size_t)N * C * HW, NAN), y2((size_t)N * C * HW, NAN), xf((size_t)N * HW * CP);
  for (size_t i = 0; i < xf.size(); ++i) xf[i] = (float)hi[i];
  emu_launch(dim3(ceil_div(HW, 32), ceil_div(C, 32), N), 256, 0, [&] { unpack_nchw_kernel<__half>(hi.data(), y.data(), noise.data(), C, HW, CP); });
  emu_launch(dim3(ceil_div(HW, 32), ceil_div(C, 32), N), 256, 0, [&] { unpack_nchw_kernel<float>(xf.data(), y2.data(), nullptr, C, HW, CP); });
  double e3 = 0, e4 = 0;
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c)
      for (int p = 0; p < HW; ++p) {
        const float h = (float)hi[((size_t)n * HW + p) * CP + c];
        e3 = std::max(e3, std::fabs((double)y[((size_t)n * C + c) * HW + p] - ((double)h + noise[p])));
        e4 = std::max(e4, std::fabs((double)y2[((size_t)n * C + c) * HW + p] - (double)h));
      }
  return report("pack_nhwc (styles, hi+lo)", e1, 2.0, 1e-6) + report("pack_nhwc pitch padding", pad, 1.0, 0.0) + report("pack_nhwc (broadcast)", e2, 1.0, 0.0) +
         report("unpack_nchw<half> + noise", e3, 3.0, 1e-7) + report("unpack_nchw<float>", e4, 1.0, 0.0);
}

int main() {
  srand(13);
  const int bad = test_demod_and_sgrad() + test_transposes();
  return bad ? 1 : 0;
}
