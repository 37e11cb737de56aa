// CPU execution of the plain CUDA kernels of univs_b200/csrc (TEST INFRASTRUCTURE, tests/test_kernels_cpu_emulation.py).
// The kernel SOURCES are compiled unchanged by g++ (tests/emu/build_emu.py rewrites only the <<<...>>> launch syntax):
// every CUDA thread of a block is a host thread, __syncthreads / warp shuffles / ballots are barriers over them, shared
// memory is a function-local static (blocks run one after the other), atomics take a lock.  What this checks is the
// kernels' index arithmetic, predication, reductions and output formats against the oracle -- not performance and not
// the hardware.  Kernels built on tcgen05 / TMA / mma.sync / cp.async are outside its reach.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __launch_bounds__(...)
#define __grid_constant__
#undef __shared__
#define __shared__ static

namespace emu {

struct Barrier {          // reusable barrier whose participants may leave (a CUDA thread that returned no longer takes part)
  std::mutex m;
  std::condition_variable cv;
  int expected = 0, arrived = 0;
  unsigned long long gen = 0;
  void wait() {
    std::unique_lock<std::mutex> l(m);
    const unsigned long long g = gen;
    if (++arrived >= expected) {
      arrived = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(l, [&] { return gen != g; });
    }
  }
  void drop() {
    std::unique_lock<std::mutex> l(m);
    --expected;
    if (expected > 0 && arrived >= expected) {
      arrived = 0;
      ++gen;
      cv.notify_all();
    }
  }
};
struct Warp {
  Barrier bar;
  uint32_t slot[32];
  uint32_t frag_a[32][4], frag_b[32][2];      // mma.sync operand fragments of the 32 lanes
};
struct Ctx {
  uint3 tid, bid;
  dim3 bdim, gdim;
  Barrier* block_bar;
  Barrier* cluster_bar;
  Warp* warp;
  int lane;
  int cta;          // rank of this CTA inside its cluster (0 without clusters): selects the emulated shared memory / TMEM
  int cluster_size;
};
extern thread_local Ctx ctx;
extern std::mutex atomic_lock;

void launch(dim3 grid, dim3 block, const std::function<void()>& body, int cluster = 1);
unsigned char* dyn_smem();                       // the CTA's dynamic shared memory
int emulated_sm_count();

inline uint32_t exchange(uint32_t v, int src_lane) {
  Warp* w = ctx.warp;
  w->slot[ctx.lane] = v;
  w->bar.wait();
  const uint32_t r = (src_lane >= 0 && src_lane < 32) ? w->slot[src_lane] : v;
  w->bar.wait();
  return r;
}
template <typename T>
inline T shfl(T v, int src_lane) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  uint32_t u;
  std::memcpy(&u, &v, 4);
  u = exchange(u, src_lane);
  std::memcpy(&v, &u, 4);
  return v;
}

// mma.sync.aligned.m16n8k8 (tf32) / m16n8k16 (f16), fp32 accumulate: every lane deposits its operand fragments, then
// computes its own four accumulator elements from the fragments of the lanes that hold the row / column it needs
// (fragment layouts of the PTX ISA: g = lane >> 2, t = lane & 3; A rows g, g+8; C columns 2t, 2t+1)
inline float half_of(uint32_t reg, int which) {
  const unsigned short h = (unsigned short)(which ? (reg >> 16) : (reg & 0xffffu));
  __half v;
  std::memcpy(&v, &h, 2);
  return __half2float(v);
}
inline float tf32_of(uint32_t reg) {
  reg &= 0xFFFFE000u;
  float f;
  std::memcpy(&f, &reg, 4);
  return f;
}
inline void mma_exchange(const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  Warp* w = ctx.warp;
  for (int i = 0; i < 4; ++i) w->frag_a[ctx.lane][i] = a[i];
  w->frag_b[ctx.lane][0] = b0;
  w->frag_b[ctx.lane][1] = b1;
  w->bar.wait();
}
inline void mma_m16n8k8_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  mma_exchange(a, b0, b1);
  Warp* w = ctx.warp;
  const int g = ctx.lane >> 2, t = ctx.lane & 3;
  for (int i = 0; i < 4; ++i) {
    const int row = g + 8 * (i >> 1), col = 2 * t + (i & 1);
    float acc = c[i];
    for (int k = 0; k < 8; ++k)
      acc += tf32_of(w->frag_a[(row & 7) * 4 + (k & 3)][(row >> 3) + 2 * (k >> 2)]) * tf32_of(w->frag_b[col * 4 + (k & 3)][k >> 2]);
    c[i] = acc;
  }
  w->bar.wait();
}
inline void mma_m16n8k16_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  mma_exchange(a, b0, b1);
  Warp* w = ctx.warp;
  const int g = ctx.lane >> 2, t = ctx.lane & 3;
  for (int i = 0; i < 4; ++i) {
    const int row = g + 8 * (i >> 1), col = 2 * t + (i & 1);
    float acc = c[i];
    for (int k = 0; k < 16; ++k) {
      const int kk = k & 7;
      acc += half_of(w->frag_a[(row & 7) * 4 + (kk >> 1)][(row >> 3) + 2 * (k >> 3)], kk & 1) *
             half_of(w->frag_b[col * 4 + (kk >> 1)][k >> 3], kk & 1);
    }
    c[i] = acc;
  }
  w->bar.wait();
}

}  // namespace emu

// threadIdx / blockIdx / blockDim / gridDim: objects whose .x / .y / .z read the calling host thread's coordinates (not
// macros: cudaLaunchConfig_t has members called gridDim and blockDim)
namespace emu {
struct Axis {
  int vec, axis;
  operator unsigned() const {
    const Ctx& c = ctx;
    const unsigned v[4][3] = {{c.tid.x, c.tid.y, c.tid.z}, {c.bid.x, c.bid.y, c.bid.z}, {c.bdim.x, c.bdim.y, c.bdim.z},
                              {c.gdim.x, c.gdim.y, c.gdim.z}};
    return v[vec][axis];
  }
};
struct Vec3 {
  Axis x, y, z;
};
}  // namespace emu
static const ::emu::Vec3 threadIdx{{0, 0}, {0, 1}, {0, 2}}, blockIdx{{1, 0}, {1, 1}, {1, 2}}, blockDim{{2, 0}, {2, 1}, {2, 2}},
    gridDim{{3, 0}, {3, 1}, {3, 2}};

inline void __syncthreads() { ::emu::ctx.block_bar->wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { ::emu::ctx.warp->bar.wait(); }
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return ::emu::shfl(v, ::emu::ctx.lane ^ lane_mask); }
template <typename T>
inline T __shfl_down_sync(unsigned, T v, int delta) {
  const int src = ::emu::ctx.lane + delta;
  return ::emu::shfl(v, src < 32 ? src : ::emu::ctx.lane);
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) { return ::emu::shfl(v, src & 31); }
inline unsigned __ballot_sync(unsigned, int pred) {
  ::emu::Warp* w = ::emu::ctx.warp;
  w->slot[::emu::ctx.lane] = pred ? 1u : 0u;
  w->bar.wait();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (w->slot[i] & 1u) << i;
  w->bar.wait();
  return r;
}
template <typename T>
inline T atomicAdd(T* p, T v) {
  std::lock_guard<std::mutex> l(::emu::atomic_lock);
  const T old = *p;
  *p = old + v;
  return old;
}
inline int atomicOr(int* p, int v) {
  std::lock_guard<std::mutex> l(::emu::atomic_lock);
  const int old = *p;
  *p = old | v;
  return old;
}
template <typename T>
inline T __ldg(const T* p) { return *p; }
template <typename T>
inline void __stcs(T* p, T v) { *p = v; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __expf(float x) { return std::exp(x); }
inline float __saturatef(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
template <typename T>
inline T min(T a, T b) { return a < b ? a : b; }
template <typename T>
inline T max(T a, T b) { return a > b ? a : b; }
inline long long min(long long a, int b) { return a < b ? a : b; }
inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }
// kernel attributes are meaningless here (nvcc offers this overload for __global__ functions)
template <typename R, typename... A>
inline cudaError_t cudaFuncSetAttribute(R (*)(A...), cudaFuncAttribute, int) { return cudaSuccess; }

// cluster launches: build_emu.py rewrites  cudaLaunchKernelEx(&cfg, kernel, args...)  into  ::emu::launch_ex(cfg, [=]{ kernel(args...); })
namespace emu {
inline cudaError_t launch_ex(const cudaLaunchConfig_t& cfg, const std::function<void()>& body) {
  int cluster = 1;
  for (unsigned i = 0; i < cfg.numAttrs; ++i)
    if (cfg.attrs[i].id == cudaLaunchAttributeClusterDimension) cluster = (int)cfg.attrs[i].val.clusterDim.x;
  launch(cfg.gridDim, cfg.blockDim, body, cluster);
  return cudaSuccess;
}
}  // namespace emu
template <typename R, typename... A>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, R (*)(A...), int, size_t) {
  *n = 2;
  return cudaSuccess;
}
template <typename R, typename... A>
inline cudaError_t cudaOccupancyMaxActiveClusters(int* n, R (*)(A...), const cudaLaunchConfig_t*) {
  *n = ::emu::emulated_sm_count() / 2;
  return cudaSuccess;
}
