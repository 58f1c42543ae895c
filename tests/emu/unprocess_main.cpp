// Driver of the emulated `unprocess` kernels of stylemc_b200/csrc/vit.cu (see cuda_emu.h): resample_h / resample_v (forward) and
// resample_vT / resample_hT (backward), launched as smc_resample_fwd / smc_resample_bwd launch them (default pass order), on the tables
// that stylemc_b200/resample.py builds.  Reads one binary file written by tests/test_kernels_emu.py, writes y and gx; the test compares
// them with the oracle's unprocess (F.interpolate bicubic antialias) and its autograd.
#include "cuda_emu.h"
#include <fstream>
static inline void __trap() { fprintf(stderr, "FAIL __trap()\n"); abort(); }
#include "kernels_extracted.inc"
using namespace smc;

static emu_dim3 dim3(unsigned x, unsigned y = 1, unsigned z = 1) { emu_dim3 d; d.x = x; d.y = y; d.z = z; return d; }
// the launch of resample_rows_launch (csrc/vit.cu): span bound, shared-memory size, grid
template <int OBT, class F> static void rows_launch(long long rows, int in_w, int out_w, int taps, F&& body) {
  long long span = ((long long)(OBT - 1) * in_w + out_w - 1) / out_w + taps + 2;
  if (span > in_w) span = in_w;
  const size_t smem = ((size_t)32 * ((size_t)span | 1) + (size_t)32 * (OBT + 1)) * sizeof(float);
  emu_launch((int)(((rows + 31) / 32) * ((out_w + OBT - 1) / OBT)), 256, smem, [&] { body((int)span); });
}
static int same(const std::vector<float>& a, const std::vector<float>& b, const char* what) {
  for (size_t i = 0; i < a.size(); ++i)
    if (!(a[i] == b[i])) { printf("FAIL %s differs from the per-output kernel at %zu: %g vs %g\n", what, i, a[i], b[i]); return 1; }
  printf("ok   %s == per-output kernel (bitwise, %zu values)\n", what, a.size());
  return 0;
}

static int grid1d(long long items) {
  long long b = (items + 255) / 256;
  return (int)std::max(1LL, std::min(b, 148LL * 16));
}
template <class T> static std::vector<T> rd(std::ifstream& f, size_t n) {
  std::vector<T> v(n);
  f.read(reinterpret_cast<char*>(v.data()), n * sizeof(T));
  return v;
}

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  const auto hdr = rd<int>(f, 5);
  const int planes = hdr[0], in = hdr[1], out = hdr[2], taps = hdr[3], taps_t = hdr[4];
  const auto x = rd<float>(f, (size_t)planes * in * in);
  const auto start = rd<int>(f, out), count = rd<int>(f, out);
  const auto wgt = rd<float>(f, (size_t)out * taps);
  const auto oidx = rd<int>(f, (size_t)in * taps_t), count_t = rd<int>(f, in);
  const auto wgt_t = rd<float>(f, (size_t)in * taps_t);
  const auto g = rd<float>(f, (size_t)planes * out * out);
  const auto ms = rd<float>(f, 7);      // mean[3], std[3], unscale
  if (!f) return 3;
  // forward: horizontal pass with the denormalise + clamp, vertical pass with / 255 and the CLIP normalisation
  std::vector<float> tmp((size_t)planes * in * out, NAN), y((size_t)planes * out * out, NAN);
  const long long rows = (long long)planes * in;
  emu_launch(grid1d(rows * out), 256, 0, [&] {
    resample_h_kernel(x.data(), tmp.data(), start.data(), count.data(), wgt.data(), taps, rows, in, out, 1, 0, 1.f, 0.f, 0.f, 0.f, 1.f, 1.f, 1.f);
  });
  // the row-tile kernel that replaces it by default (lanes on rows, shared-memory staged): must reproduce it bit for bit
  std::vector<float> tmp_r((size_t)planes * in * out, NAN);
  rows_launch<32>(rows, in, out, taps, [&](int span) {
    resample_rows_kernel<32, 24>(x.data(), tmp_r.data(), start.data(), 1, count.data(), wgt.data(), taps, rows, in, out, span, 1, 0, nullptr, nullptr);
  });
  int bad = same(tmp_r, tmp, "resample_rows_kernel<32> (forward)");
  emu_launch(grid1d((long long)planes * out * out), 256, 0, [&] {
    resample_v_kernel(tmp.data(), y.data(), start.data(), count.data(), wgt.data(), taps, planes, in, out, out, 1.f / 255.f, ms[0], ms[1], ms[2], ms[3], ms[4],
                      ms[5], 1);
  });
  // backward: transposed vertical pass, then the transposed horizontal pass with the clamp mask, 127.5 and the loss scale
  std::vector<float> tmp2((size_t)planes * in * out, NAN), gx((size_t)planes * in * in, NAN);
  const float unscale = ms[6];
  emu_launch(grid1d((long long)planes * in * out), 256, 0, [&] {
    resample_vT_kernel(g.data(), tmp2.data(), oidx.data(), count_t.data(), wgt_t.data(), taps_t, planes, in, out, out, ms[3], ms[4], ms[5], nullptr, nullptr);
  });
  emu_launch(grid1d(rows * in), 256, 0, [&] {
    resample_hT_kernel(tmp2.data(), x.data(), gx.data(), oidx.data(), count_t.data(), wgt_t.data(), taps_t, rows, in, out, &unscale);
  });
  std::vector<float> gx_r((size_t)planes * in * in, NAN);
  rows_launch<128>(rows, out, in, taps_t, [&](int span) {
    resample_rows_kernel<128, 8>(tmp2.data(), gx_r.data(), oidx.data(), taps_t, count_t.data(), wgt_t.data(), taps_t, rows, out, in, span, 0, 1, x.data(), &unscale);
  });
  bad += same(gx_r, gx, "resample_rows_kernel<128> (backward)");
  if (bad) return 5;
  std::ofstream o(argv[2], std::ios::binary);
  o.write(reinterpret_cast<const char*>(y.data()), y.size() * sizeof(float));
  o.write(reinterpret_cast<const char*>(gx.data()), gx.size() * sizeof(float));
  return o ? 0 : 4;
}
