// Driver of the emulated ViT glue kernels of stylemc_b200/csrc/vit.cu (see cuda_emu.h): LayerNorm forward / backward, the embedding head and
// its transpose, QuickGELU forward / backward, split_rows and the directional CLIP loss, each against a float64 restatement.
// kernels_extracted.inc is cut out of vit.cu by tests/test_kernels_emu.py.
#include "cuda_emu.h"
#include "kernels_extracted.inc"
using namespace smc;

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }
static std::vector<float> rnd(size_t n, double scale = 1.0, double shift = 0.0) {
  std::vector<float> v(n);
  for (auto& x : v) x = (float)(scale * frand() + shift);
  return v;
}
static int report(const char* name, double err, double ref_max, double rel_tol) {
  const bool ok = err <= rel_tol * ref_max;       // NaN (an element never written) fails
  printf("%s %-28s max err %.2e (max |ref| %.3g)\n", ok ? "ok  " : "FAIL", name, err, ref_max);
  return ok ? 0 : 1;
}
static float joined(const std::vector<__half>& hi, const std::vector<__half>& lo, size_t i) { return (float)hi[i] + (float)lo[i]; }

static int test_layernorm() {
  // ln_post-style addressing: output row r reads input row r * stride + offset
  const int rows = 11, Wd = 200, stride = 3, offset = 1;
  auto x = rnd((size_t)(rows * stride + 2) * Wd, 2.0, 0.3), w = rnd(Wd, 0.2, 1.0), b = rnd(Wd, 0.2), dy = rnd((size_t)rows * Wd);
  std::vector<float> y32((size_t)rows * Wd, NAN), mean(rows, NAN), rstd(rows, NAN);
  std::vector<__half> yhi((size_t)rows * Wd, (__half)NAN), ylo((size_t)rows * Wd, (__half)NAN);
  emu_launch(2, 256, 0, [&] { layernorm_fwd_kernel(x.data(), stride, offset, w.data(), b.data(), y32.data(), yhi.data(), ylo.data(), mean.data(), rstd.data(), rows, Wd); });
  std::vector<float> dx = rnd(x.size());
  const std::vector<float> dx0 = dx;
  emu_launch(1, 256, 0, [&] { layernorm_bwd_kernel(dy.data(), x.data(), stride, offset, w.data(), mean.data(), rstd.data(), dx.data(), rows, Wd, 1); });
  std::vector<float> dxs((size_t)x.size(), 7.0f);
  emu_launch(3, 256, 0, [&] { layernorm_bwd_kernel(dy.data(), x.data(), stride, offset, w.data(), mean.data(), rstd.data(), dxs.data(), rows, Wd, 0); });
  double ey = 0, eh = 0, ed = 0, es = 0, my = 0, md = 0;
  for (int r = 0; r < rows; ++r) {
    const float* xr = x.data() + (size_t)(r * stride + offset) * Wd;
    double mu = 0, var = 0;
    for (int d = 0; d < Wd; ++d) mu += xr[d];
    mu /= Wd;
    for (int d = 0; d < Wd; ++d) var += (xr[d] - mu) * (xr[d] - mu);
    const double rs = 1.0 / std::sqrt(var / Wd + 1e-5);
    double a = 0, bs = 0;
    for (int d = 0; d < Wd; ++d) { const double g = (double)w[d] * dy[(size_t)r * Wd + d]; a += g; bs += g * (xr[d] - mu) * rs; }
    a /= Wd; bs /= Wd;
    for (int d = 0; d < Wd; ++d) {
      const double o = (xr[d] - mu) * rs * w[d] + b[d];
      const size_t i = (size_t)r * Wd + d, xi = (size_t)(r * stride + offset) * Wd + d;
      ey = std::max(ey, std::fabs(y32[i] - o)); eh = std::max(eh, std::fabs(joined(yhi, ylo, i) - o)); my = std::max(my, std::fabs(o));
      const double g = (double)w[d] * dy[i], want = rs * (g - a - (xr[d] - mu) * rs * bs);
      ed = std::max(ed, std::fabs(dx[xi] - (dx0[xi] + want))); es = std::max(es, std::fabs(dxs[xi] - want)); md = std::max(md, std::fabs(want));
    }
  }
  // rows the kernel must not touch (the input rows between the strided ones)
  double untouched = 0;
  for (int d = 0; d < Wd; ++d) untouched = std::max(untouched, (double)std::fabs(dxs[d] - 7.0f) + std::fabs(dx[d] - dx0[d]));
  return report("layernorm_fwd (fp32)", ey, my, 2e-6) + report("layernorm_fwd (hi+lo)", eh, my, 2e-6) + report("layernorm_bwd accumulate", ed, md, 4e-6) +
         report("layernorm_bwd store", es, md, 4e-6) + report("layernorm_bwd other rows", untouched, 1.0, 0.0);
}

static int test_head() {
  const int B = 3, Wd = 200, E = 70;
  auto ln = rnd((size_t)B * Wd), proj = rnd((size_t)Wd * E, 0.1), dE = rnd((size_t)B * E);
  std::vector<float> out((size_t)B * E, NAN), dln((size_t)B * Wd, NAN);
  emu_launch(B, 256, Wd * sizeof(float), [&] { head_proj_kernel(ln.data(), proj.data(), out.data(), Wd, E); });
  emu_launch(B, 256, E * sizeof(float), [&] { head_proj_bwd_kernel(dE.data(), proj.data(), dln.data(), Wd, E); });
  double e1 = 0, e2 = 0, m1 = 0, m2 = 0;
  for (int b = 0; b < B; ++b) {
    for (int j = 0; j < E; ++j) {
      double a = 0;
      for (int d = 0; d < Wd; ++d) a += (double)ln[(size_t)b * Wd + d] * proj[(size_t)d * E + j];
      e1 = std::max(e1, std::fabs(out[(size_t)b * E + j] - a)); m1 = std::max(m1, std::fabs(a));
    }
    for (int d = 0; d < Wd; ++d) {
      double a = 0;
      for (int j = 0; j < E; ++j) a += (double)dE[(size_t)b * E + j] * proj[(size_t)d * E + j];
      e2 = std::max(e2, std::fabs(dln[(size_t)b * Wd + d] - a)); m2 = std::max(m2, std::fabs(a));
    }
  }
  return report("head_proj", e1, m1, 4e-6) + report("head_proj_bwd", e2, m2, 4e-6);
}

static int test_elementwise() {
  const long long n = 3000;
  auto h = rnd(n, 4.0), dg = rnd(n);
  std::vector<__half> hi(n, (__half)NAN), lo(n, (__half)NAN), bhi(n, (__half)NAN), blo(n, (__half)NAN);
  emu_launch(2, 256, 0, [&] { quickgelu_fwd_kernel(h.data(), hi.data(), lo.data(), n); });
  emu_launch(2, 256, 0, [&] { quickgelu_bwd_kernel(dg.data(), h.data(), bhi.data(), blo.data(), n); });
  double e1 = 0, e2 = 0;
  for (long long i = 0; i < n; ++i) {
    const double x = h[i], sg = 1.0 / (1.0 + std::exp(-1.702 * x));
    e1 = std::max(e1, std::fabs(joined(hi, lo, i) - x * sg));
    e2 = std::max(e2, std::fabs(joined(bhi, blo, i) - dg[i] * (sg + 1.702 * x * sg * (1.0 - sg))));
  }
  // split_rows as the patch-embedding backward uses it: drop the class-token row of every group of t rows
  const int groups = 3, t = 6, Wd = 40, rows = groups * (t - 1);
  auto x = rnd((size_t)groups * t * Wd);
  std::vector<__half> shi((size_t)rows * Wd, (__half)NAN), slo((size_t)rows * Wd, (__half)NAN);
  emu_launch(1, 256, 0, [&] { split_rows_kernel(x.data(), shi.data(), slo.data(), rows, Wd, t - 1, t, 1); });
  double e3 = 0;
  for (int r = 0; r < rows; ++r)
    for (int d = 0; d < Wd; ++d)
      e3 = std::max(e3, std::fabs(joined(shi, slo, (size_t)r * Wd + d) - (double)x[(size_t)((r / (t - 1)) * t + 1 + r % (t - 1)) * Wd + d]));
  return report("quickgelu_fwd", e1, 4.0, 1e-6) + report("quickgelu_bwd", e2, 1.2, 1e-6) + report("split_rows", e3, 1.0, 1e-6);
}

static int test_clip_loss_mode(int normalize) {
  const int N = 5, E = 700;            // E not a multiple of the block size; sample 3 is degenerate (edited == original)
  auto es = rnd((size_t)N * E), et = rnd((size_t)N * E), text = rnd(E);
  for (int j = 0; j < E; ++j) et[(size_t)3 * E + j] = es[(size_t)3 * E + j];
  const float coef = 0.7f, inv_count = 1.0f / 9.0f, target = 64.0f;
  std::vector<float> d((size_t)N * E, NAN);
  float part = NAN, gs = NAN;
  emu_launch(1, 512, 0, [&] { clip_loss_kernel(es.data(), et.data(), text.data(), &part, d.data(), N, E, coef, inv_count, &gs, target, normalize); });
  std::vector<double> want((size_t)N * E, 0.0);
  double total = 0, tt = 0, dmax = 0;
  for (int j = 0; j < E; ++j) tt += (double)text[j] * text[j];
  const double nt = std::sqrt(tt);
  for (int n = 0; n < N; ++n) {
    // normalize: e = a/|a| - b/|b| (clip_loss_nada.py:162-168,213-216), gradient through the normalisation of a
    double na = 1, nb = 1;
    if (normalize) {
      double aa = 0, bb = 0;
      for (int j = 0; j < E; ++j) { aa += (double)et[(size_t)n * E + j] * et[(size_t)n * E + j]; bb += (double)es[(size_t)n * E + j] * es[(size_t)n * E + j]; }
      na = std::sqrt(aa); nb = std::sqrt(bb);
    }
    std::vector<double> e(E), ah(E);
    double ee = 0, ed = 0;
    for (int j = 0; j < E; ++j) {
      ah[j] = et[(size_t)n * E + j] / na;
      e[j] = ah[j] - es[(size_t)n * E + j] / nb;
      ee += e[j] * e[j]; ed += e[j] * text[j];
    }
    if (!(ee > 0)) continue;
    const double ne = std::sqrt(ee), c = ed / (ne * nt);
    total -= c;
    std::vector<double> g(E);
    double gdot = 0;
    for (int j = 0; j < E; ++j) { g[j] = -coef * inv_count * (text[j] / nt - c * e[j] / ne) / ne; gdot += g[j] * ah[j]; }
    for (int j = 0; j < E; ++j) {
      want[(size_t)n * E + j] = normalize ? (g[j] - gdot * ah[j]) / na : g[j];
      dmax = std::max(dmax, std::fabs(want[(size_t)n * E + j]));
    }
  }
  const double S = std::exp2(std::floor(std::log2(target / dmax)));
  double err = 0;
  for (size_t i = 0; i < want.size(); ++i) err = std::max(err, std::fabs(d[i] - S * want[i]));
  int bad = report(normalize ? "clip_loss (normalised): loss partial sum" : "clip_loss: loss partial sum", std::fabs(part - coef * inv_count * total),
                   std::fabs(coef * inv_count * total), 4e-6);
  bad += report("clip_loss: d_tgt * scale", err, S * dmax, 2e-5);
  bad += report("clip_loss: loss scale", std::fabs(gs - S), S, 0.0);
  bad += report("clip_loss: scaled max in range", (S * dmax >= target / 2 && S * dmax < target) ? 0.0 : 1.0, 1.0, 0.0);
  return bad;
}
static int test_clip_loss() { return test_clip_loss_mode(0) + test_clip_loss_mode(1); }

int main() {
  srand(11);
  const int bad = test_layernorm() + test_head() + test_elementwise() + test_clip_loss();
  return bad ? 1 : 0;
}
