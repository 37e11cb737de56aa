// Descriptor encodings, swizzle arithmetic and the fp16 hi|lo split of the tcgen05 kernels: plain C++ (no PTX), shared by
// the device build (tc05.cuh) and by the CPU emulation of the tensor-core kernels (tests/emu/tc05.cuh).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace univs {
namespace tc {

// Shared-memory matrix descriptor of a swizzled operand tile whose rows are ROW bytes (= one swizzle span: 128 ->
// SWIZZLE_128B, 64 -> SWIZZLE_64B, 32 -> SWIZZLE_32B), consecutive rows ROW bytes apart, 8-row groups 8*ROW bytes apart.
//   K-major operand : row = M/N index, the ROW bytes are consecutive K elements   ((8,n),(T,2)):((ROW/16,SBO),(1,T))
//   MN-major operand: row = K index,  the ROW bytes are consecutive M/N elements  ((T,ROW/16,1),(8,k)):((1,T,-),(ROW/16,SBO))
// Both read SBO (bits 32-45) as the byte distance between 8-row groups; LBO is unused for one swizzle span (set to 1);
// bit 46 = descriptor version 1 (Blackwell).  Which of the two the MMA assumes is the a_major / b_major bit of the
// instruction descriptor.  The tile base must be 1024-byte aligned; advancing the start address by 32 bytes inside a row
// selects the next 16-element K step of a K-major f16 operand.
template <int ROW>
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  static_assert(ROW == 128 || ROW == 64 || ROW == 32, "one swizzle span per row");
  constexpr uint64_t layout = ROW == 128 ? 2 : (ROW == 64 ? 4 : 6);
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((8 * ROW) >> 4) << 32) |
         ((uint64_t)1 << 46) | (layout << 61);
}
// Instruction descriptor, kind::f16 with fp16 operands and fp32 accumulation.
//   bits 4-5 D format (1 = f32), 7-9 / 10-12 A / B format (0 = f16), 15 / 16 A / B major (1 = MN-major),
//   17-22 N >> 3, 24-28 M >> 4
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int Nn, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (a_mn_major ? (1u << 15) : 0u) | (b_mn_major ? (1u << 16) : 0u) | ((uint32_t)(Nn >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// byte offset of the 16-byte chunk `chunk` of row `row` in a swizzled tile with ROW-byte rows (Swizzle<log2(ROW/16),4,3>
// on the byte address: the chunk index is XORed with the row bits that sit log2(ROW) .. above it)
template <int ROW>
__host__ __device__ constexpr uint32_t swz_off(int row, int chunk) {
  return ROW == 128 ? (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4))
                    : (ROW == 64 ? (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4))
                                 : (uint32_t)(row * 32 + ((chunk ^ ((row >> 2) & 1)) << 4)));
}

// fp32 -> fp16 hi + fp16 lo (hi + lo carries ~22 significand bits), two values at a time
__device__ __forceinline__ void split_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
}  // namespace tc
}  // namespace univs
