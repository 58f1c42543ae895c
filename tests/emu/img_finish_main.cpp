// Driver of the emulated img_finish4_kernel / img_finish_kernel of stylemc_b200/csrc/synth.cu (see cuda_emu.h): the tail of the fused ToRGB
// path from 256 px up (the conv1 epilogue accumulated the 1x1 modulated conv into img).  Reference (float64):
//   r = img + b_rgb[j];  pass_mask = |r| < clamp;  img = clamp(r) + upsample2d(img_prev)   (utils.py:45-49; 4x4 taps fk_up, pad [2,1,2,1]).
#include "cuda_emu.h"
#include "kernels_extracted.inc"
using namespace smc;

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

static int run(int N, int H, int W, bool with_prev, bool with_mask, float clamp, int parts = 1) {
  const size_t ne = (size_t)N * 3 * H * W;
  const int h2 = H / 2, w2 = W / 2;
  // parts > 1: one partial-sum image per N tile of the conv1 GEMM (smc_igemm_epilogue::rgb_snt), part_stride = ne + 4 floats apart
  const long long part_stride = parts > 1 ? (long long)ne + 4 : 0;
  std::vector<float> img(ne + (size_t)(parts - 1) * (size_t)part_stride), prev((size_t)N * 3 * h2 * w2), b(3);
  alignas(16) float fk[16];
  for (auto& v : img) v = (float)(1.5 * frand());
  for (auto& v : prev) v = (float)frand();
  for (auto& v : b) v = (float)(0.1 * frand());
  for (auto& v : fk) v = (float)(0.5 + 0.5 * frand());
  const std::vector<float> img0 = img;
  std::vector<unsigned char> mask(ne, 7);
  const bool vec = (W & 3) == 0;      // the choice smc_img_finish makes (the buffers here are 16-byte aligned)
  const long long items = vec ? (long long)N * 3 * H * (W >> 2) : (long long)ne;
  const int blocks = (int)std::max(1LL, std::min((items + 255) / 256, 148LL * 16));
  emu_launch(blocks, 256, 0, [&] {
    if (vec) img_finish4_kernel(img.data(), with_prev ? prev.data() : nullptr, b.data(), clamp, fk, N, H, W, with_mask ? mask.data() : nullptr, parts, part_stride);
    else img_finish_kernel(img.data(), with_prev ? prev.data() : nullptr, b.data(), clamp, fk, N, H, W, with_mask ? mask.data() : nullptr, parts, part_stride);
  });
  double err = 0;
  int bad_mask = 0, clamped = 0;
  for (int nj = 0; nj < N * 3; ++nj)
    for (int yy = 0; yy < H; ++yy)
      for (int xx = 0; xx < W; ++xx) {
        const size_t i = ((size_t)nj * H + yy) * W + xx;
        double r = b[nj % 3];
        for (int q = 0; q < parts; ++q) r += (double)img0[i + (size_t)q * (size_t)part_stride];
        const bool pass = clamp < 0 || std::fabs(r) < clamp;
        clamped += !pass;
        if (with_mask && mask[i] != (pass ? 1 : 0)) ++bad_mask;
        if (clamp >= 0) r = std::min(std::max(r, -(double)clamp), (double)clamp);
        if (with_prev)
          for (int fy = 0; fy < 4; ++fy)
            for (int fx = 0; fx < 4; ++fx) {
              const int ay = yy + fy - 2, ax = xx + fx - 2;
              if (ay < 0 || ax < 0 || (ay & 1) || (ax & 1) || ay / 2 >= h2 || ax / 2 >= w2) continue;
              r += (double)fk[fy * 4 + fx] * prev[((size_t)nj * h2 + ay / 2) * w2 + ax / 2];
            }
        err = std::max(err, std::fabs((double)img[i] - r));
      }
  const bool ok = err <= 2e-6 * parts && bad_mask == 0 && (clamp < 0 || clamp > 2 || clamped > 0);
  printf("%s %s N=%d %dx%d prev=%d mask=%d clamp=%g: max err %.2e, %d wrong mask bytes, %d clamped\n", ok ? "ok  " : "FAIL", vec ? "img_finish4" : "img_finish ", N, H, W,
         (int)with_prev, (int)with_mask, clamp, err, bad_mask, clamped);
  return ok ? 0 : 1;
}

int main() {
  srand(37);
  int bad = 0;
  bad += run(2, 12, 16, true, true, 1.0f);       // vector kernel: skip image, clamp mask saved for the backward pass
  bad += run(1, 8, 4, true, false, -1.0f);       // one float4 per row
  bad += run(1, 6, 8, false, true, 256.0f);      // no previous image
  bad += run(2, 6, 10, true, true, 1.0f);        // W % 4 != 0: scalar kernel
  bad += run(2, 12, 16, true, true, 2.0f, 4);    // four partial-sum images (512-channel conv1: four N tiles), vector kernel
  bad += run(1, 6, 10, true, true, 2.0f, 2);     // two partial-sum images, scalar kernel
  return bad ? 1 : 0;
}
