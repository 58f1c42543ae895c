// Driver of the emulated token-assembly kernels of stylemc_b200/csrc/vit.cu (see cuda_emu.h): patchify (the im2col of the ViT patch
// embedding, kernel = stride = patch size), its transpose unpatchify, assemble_tokens (class token + positional embedding) and embed_text.
// Patch sizes other than ViT-B/32's 32 matter for the ViT-B/16 tower (clip_loss.py:12-13), which had no device run in round 1.
#include "cuda_emu.h"
#include "kernels_extracted.inc"
using namespace smc;

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }
static std::vector<float> rnd(size_t n) {
  std::vector<float> v(n);
  for (auto& x : v) x = (float)frand();
  return v;
}

static int run_patch(int B, int res, int ps) {
  const int grid = res / ps, kk = 3 * ps * ps, rows = B * grid * grid;
  auto img = rnd((size_t)B * 3 * res * res), gp = rnd((size_t)rows * kk);
  std::vector<__half> hi((size_t)rows * kk, (__half)NAN), lo((size_t)rows * kk, (__half)NAN);
  std::vector<float> gimg(img.size(), NAN);
  emu_launch(3, 256, 0, [&] { patchify_kernel(img.data(), hi.data(), lo.data(), B, res, ps); });
  emu_launch(2, 256, 0, [&] { unpatchify_kernel(gp.data(), gimg.data(), B, res, ps); });
  double e1 = 0, e2 = 0;
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < 3; ++c)
      for (int y = 0; y < res; ++y)
        for (int x = 0; x < res; ++x) {
          // F.conv2d(image, weight[vw, 3, ps, ps], stride = ps): patch (py, px) row-major, column = c * ps * ps + ky * ps + kx
          const size_t row = ((size_t)b * grid + y / ps) * grid + x / ps, col = ((size_t)c * ps + y % ps) * ps + x % ps;
          const size_t i = (((size_t)b * 3 + c) * res + y) * res + x;
          e1 = std::max(e1, std::fabs((double)(float)hi[row * kk + col] + (float)lo[row * kk + col] - (double)img[i]));
          e2 = std::max(e2, std::fabs((double)gimg[i] - (double)gp[row * kk + col]));
        }
  const bool ok = e1 <= 1e-6 && e2 == 0.0;
  printf("%s patchify / unpatchify B=%d res=%d patch=%d: split err %.2e, transpose err %.2e\n", ok ? "ok  " : "FAIL", B, res, ps, e1, e2);
  return ok ? 0 : 1;
}

static int run_tokens() {
  const int B = 2, T = 5, Wd = 24, vocab = 11;
  auto patch = rnd((size_t)B * (T - 1) * Wd), cls = rnd(Wd), pos = rnd((size_t)T * Wd), emb = rnd((size_t)vocab * Wd);
  std::vector<float> x0((size_t)B * T * Wd, NAN), xt((size_t)B * T * Wd, NAN);
  std::vector<long long> text((size_t)B * T);
  for (auto& t : text) t = rand() % vocab;
  emu_launch(1, 256, 0, [&] { assemble_tokens_kernel(patch.data(), cls.data(), pos.data(), x0.data(), B, T, Wd); });
  emu_launch(1, 256, 0, [&] { embed_text_kernel(text.data(), emb.data(), pos.data(), xt.data(), B, T, Wd); });
  double e1 = 0, e2 = 0;
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T; ++t)
      for (int d = 0; d < Wd; ++d) {
        const size_t i = ((size_t)b * T + t) * Wd + d;
        const float want = (t == 0 ? cls[d] : patch[((size_t)b * (T - 1) + t - 1) * Wd + d]) + pos[(size_t)t * Wd + d];
        e1 = std::max(e1, (double)std::fabs(x0[i] - want));
        e2 = std::max(e2, (double)std::fabs(xt[i] - (emb[(size_t)text[(size_t)b * T + t] * Wd + d] + pos[(size_t)t * Wd + d])));
      }
  const bool ok = e1 == 0.0 && e2 == 0.0;
  printf("%s assemble_tokens / embed_text: err %.2e / %.2e\n", ok ? "ok  " : "FAIL", e1, e2);
  return ok ? 0 : 1;
}

int main() {
  srand(41);
  const int bad = run_patch(2, 64, 32) + run_patch(1, 48, 16) + run_patch(2, 24, 8) + run_tokens();
  return bad ? 1 : 0;
}
