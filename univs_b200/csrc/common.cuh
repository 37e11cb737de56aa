// Shared device helpers for the univs_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/univs_b200.h"

namespace univs {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define UNIVS_REQUIRE(cond, ...)                  \
  do {                                            \
    if (!(cond)) {                                \
      ::univs::set_error(__VA_ARGS__);            \
      return UNIVS_E_BADARG;                      \
    }                                             \
  } while (0)

// ---- TF32 helpers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2tf32(float x) {
#ifdef UNIVS_CPU_EMU      // tests/emu: round to nearest, ties away from zero, to 10 explicit significand bits
  const uint32_t u = __float_as_uint(x);
  return ((u & 0x7f800000u) == 0x7f800000u) ? u : ((u + 0x1000u) & 0xffffe000u);
#else
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
#endif
}
// big/small split: x ~= big + small, both exactly representable in TF32
__device__ __forceinline__ void split_tf32(float x, uint32_t& big, uint32_t& small) {
  big = f2tf32(x);
  small = f2tf32(x - __uint_as_float(big));
}

// D(16x8) += A(16x8, row) * B(8x8, col), TF32 inputs, FP32 accumulate.
// Fragment layouts (g = lane>>2, t = lane&3):
//   a0:(g,t) a1:(g+8,t) a2:(g,t+4) a3:(g+8,t+4);  b0:(k=t,n=g) b1:(k=t+4,n=g);
//   c0:(g,2t) c1:(g,2t+1) c2:(g+8,2t) c3:(g+8,2t+1)
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifdef UNIVS_CPU_EMU
  ::emu::mma_m16n8k8_tf32(c, a, b0, b1);
  return;
#endif
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 3xTF32: c += a*b with a = ab+as, b = bb+bs (small terms first)
__device__ __forceinline__ void mma_tf32x3(float (&c)[4], const uint32_t (&ab)[4], const uint32_t (&as)[4],
                                           uint32_t bb0, uint32_t bb1, uint32_t bs0, uint32_t bs1) {
  mma_tf32(c, as, bb0, bb1);
  mma_tf32(c, ab, bs0, bs1);
  mma_tf32(c, ab, bb0, bb1);
}

#ifdef UNIVS_CPU_EMU      // tests/emu: the copy happens at once (a legal schedule of the asynchronous one)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) { memcpy(smem, gmem, 16); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N>
__device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
#endif

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace univs
