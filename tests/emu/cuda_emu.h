// Minimal CPU stand-in for the CUDA execution model, enough to run a shared-memory / warp-shuffle kernel of csrc/ unchanged on host
// threads: one std::thread per CUDA thread of a block, blocks one after another.  TEST INFRASTRUCTURE ONLY (tests/test_kernels_emu.py):
// it checks index arithmetic and data flow of kernels on a machine without a GPU; it says nothing about launch limits or speed.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
struct emu_block {
  explicit emu_block(int threads) : bar(threads), shfl(threads) {
    for (int w = 0; w < (threads + 31) / 32; ++w) warp_bar.emplace_back(new std::barrier<>(std::min(32, threads - 32 * w)));
  }
  std::barrier<> bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<float> shfl;
};
static thread_local emu_dim3 threadIdx, blockIdx;
static emu_dim3 blockDim, gridDim;
static thread_local float* emu_smem;
static thread_local emu_block* emu_ctx;

#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
typedef _Float16 __half;
static inline __half __float2half_rn(float v) { return (__half)v; }
static inline float __half2float(__half h) { return (float)h; }
static inline float rsqrtf(float v) { return 1.0f / std::sqrt(v); }
using std::min;
using std::max;
using std::isfinite;
#define __shared__ static          /* statically sized shared arrays: blocks run one after another, so one copy per process is one per block */
template <class T> static inline T __ldg(const T* p) { return *p; }

// vector types with CUDA's alignment, so that UBSan checks the alignment assumptions of 8- and 16-byte loads / stores
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(4) __half2 { __half x, y; };
struct alignas(4) uchar4 { unsigned char x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline __half2 __floats2half2_rn(float a, float b) { return __half2{(__half)a, (__half)b}; }
static inline float2 __half22float2(__half2 h) { return float2{(float)h.x, (float)h.y}; }
static inline float __int_as_float(int i) { float f; __builtin_memcpy(&f, &i, 4); return f; }
// packed fp32x2 arithmetic of common.cuh (inline PTX there): one rounding per component, like fma.rn.f32x2 / mul.rn.f32x2
static inline float2 ffma2(float2 a, float2 b, float2 c) { return float2{std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y)}; }
static inline float2 fmul2(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }

static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline float __uint_as_float(unsigned u) { float f; __builtin_memcpy(&f, &u, 4); return f; }
// 16-byte streaming load / store of common.cuh (inline PTX there); the cast keeps CUDA's 16-byte alignment requirement visible to UBSan
static inline uint4 ld_stream(const void* p) { return *static_cast<const uint4*>(p); }
static inline void st_stream(void* p, const uint4& v) { *static_cast<uint4*>(p) = v; }
static inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }

static inline void __syncthreads() { emu_ctx->bar.arrive_and_wait(); }
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  const int t = (int)threadIdx.x;
  emu_ctx->shfl[t] = v;
  emu_ctx->warp_bar[t >> 5]->arrive_and_wait();
  const float r = emu_ctx->shfl[t ^ lane_mask];
  emu_ctx->warp_bar[t >> 5]->arrive_and_wait();
  return r;
}

// kernel<<<grid, block, smem>>>(args...)  ->  emu_launch(grid, block, smem, [&] { kernel(args...); })   (grid: int or emu_dim3)
static void emu_launch(emu_dim3 grid, int block, size_t smem_bytes, const std::function<void()>& body) {
  gridDim = grid;
  blockDim.x = block;
  std::vector<float> smem((smem_bytes + 3) / 4);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::fill(smem.begin(), smem.end(), NAN);      // reads of never-written shared memory poison the result
        emu_block ctx(block);
        std::vector<std::thread> th;
        for (int t = 0; t < block; ++t)
          th.emplace_back([&, t] {
            threadIdx.x = t;
            blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
            emu_smem = smem.data();
            emu_ctx = &ctx;
            body();
          });
        for (auto& x : th) x.join();
      }
}
static void emu_launch(int grid, int block, size_t smem_bytes, const std::function<void()>& body) {
  emu_dim3 g;
  g.x = grid;
  emu_launch(g, block, smem_bytes, body);
}
