// Output formats shared by the fused row-wise kernels (layernorm / gelu / relu / split / groupnorm).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace univs {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// hi = x rounded to nearest TF32 (|lo| <= 2^-12 |x|, so the dropped lo*lo term is 2^-24 relative and unbiased)
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(f2tf32(x)); }

// Output formats of the row-wise kernels (`split` argument):
//   0          plain fp32 [rows, C]
//   Kc > 0     fp32 [rows, 2C] in K-chunks of Kc columns, chunk c = [hi_c | lo_c], hi = rna_tf32(x), lo = x - hi
//   -2         fp16 [rows, 2C] = [hi | lo], hi = fp16(x) (round-to-nearest, saturating), lo = x - hi   (einsum operands)
//   -3         fp16 [rows, 2C] = [hi | lo * 2^11]: the compact operand of the tcgen05 GEMM (gemm_tc.cu), which keeps the
//              correction terms in their own accumulator and therefore needs no hi * 2^-11 copy: 4 bytes per element
//   -Kc <= -4  fp16 [rows, 3C] in K-chunks of Kc columns, chunk c = [lo_c * 2^11 | hi_c * 2^-11 | hi_c]: the A operand of
//              the single-GEMM fp16x3 product  X W^T = [Xl' | Xh_s | Xh] [Wh_s | Wl' | Wh]^T  (correction terms first, so
//              the tensor core's truncating accumulator is still small while they are added; main term last).
// fp16 carries an 11-bit significand like TF32, so hi*hi products are exact in fp32 and two terms give 22 bits; the
// 2^11 scale keeps lo out of the fp16 subnormal range.
__device__ __forceinline__ __half sat_half(float x) { return __float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f)); }

__device__ __forceinline__ void store_maybe_split(float* __restrict__ out, size_t row, int C, int col, float4 v, int split) {
  if (split == 0) {
    *reinterpret_cast<float4*>(out + row * C + col) = v;
  } else if (split > 0) {
    float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    const int chunk = col / split;
    const size_t o = row * (2 * (size_t)C) + (size_t)chunk * (2 * split) + (col - chunk * split);
    *reinterpret_cast<float4*>(out + o) = h;
    *reinterpret_cast<float4*>(out + o + split) = l;
  } else {
    __half* o16 = reinterpret_cast<__half*>(out);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    __half h[4], l[4], hs[4];
    const float sc = (split == -2) ? 1.f : 2048.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      h[i] = sat_half(vv[i]);
      const float hf = __half2float(h[i]);
      l[i] = sat_half((vv[i] - hf) * sc);
      hs[i] = __float2half_rn(hf * (1.f / 2048.f));
    }
    if (split == -2 || split == -3) {
      const size_t o = row * (2 * (size_t)C) + col;
      *reinterpret_cast<__half2*>(o16 + o) = __halves2half2(h[0], h[1]);
      *reinterpret_cast<__half2*>(o16 + o + 2) = __halves2half2(h[2], h[3]);
      *reinterpret_cast<__half2*>(o16 + o + C) = __halves2half2(l[0], l[1]);
      *reinterpret_cast<__half2*>(o16 + o + C + 2) = __halves2half2(l[2], l[3]);
    } else {
      const int kc = -split;
      const int chunk = col / kc;
      const size_t o = row * (3 * (size_t)C) + (size_t)chunk * (3 * kc) + (col - chunk * kc);
      *reinterpret_cast<__half2*>(o16 + o) = __halves2half2(l[0], l[1]);
      *reinterpret_cast<__half2*>(o16 + o + 2) = __halves2half2(l[2], l[3]);
      *reinterpret_cast<__half2*>(o16 + o + kc) = __halves2half2(hs[0], hs[1]);
      *reinterpret_cast<__half2*>(o16 + o + kc + 2) = __halves2half2(hs[2], hs[3]);
      *reinterpret_cast<__half2*>(o16 + o + 2 * kc) = __halves2half2(h[0], h[1]);
      *reinterpret_cast<__half2*>(o16 + o + 2 * kc + 2) = __halves2half2(h[2], h[3]);
    }
  }
}

}  // namespace univs
