// Fused row-wise kernels around the library GEMMs: residual-add + LayerNorm, exact GELU, and the TF32 hi|lo operand
// split that lets TF32 tensor-core GEMMs reproduce fp32 products (3xTF32: X*W ~= Xh*Wh + Xl*Wh + Xh*Wl with
// hi = x rounded to nearest TF32 (exactly representable, so the tensor core's truncation is lossless) and lo = x - hi).
//
//   layernorm : s = x (+ residual);  [sum_out = s];  out = LN(s) * gamma + beta, written plain [rows,C] or split
//               [rows,2C] = [hi | lo]              (reference: nn.LayerNorm sites swin.py:246,292; msdeformattn.py:126-133;
//                                                    transformer_layers.py:42,113,176; decoder_norm ..._univs.py:500)
//   gelu      : out = 0.5*x*(1+erf(x/sqrt(2))) (nn.GELU default, swin.py:24-41), plain or split
//   split     : out = [hi | lo]
// One warp per row, 128-bit loads/stores, the row is held in registers (single read of x), two-pass mean/variance
// with warp-shuffle reductions.  All three are pure HBM streams.
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
    if (split == -2) {
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

template <int MAXV>  // float4 per lane; C <= MAXV*128
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ res_bias,
                 const float* __restrict__ gamma, const float* __restrict__ beta, long long rows, int C, float eps,
                 float* __restrict__ sum_out, float* __restrict__ out, int split) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = C >> 2;  // float4 per row
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      float4 a = *reinterpret_cast<const float4*>(x + row * C + idx * 4);
      if (res != nullptr) {
        const float4 b = *reinterpret_cast<const float4*>(res + row * C + idx * 4);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        if (res_bias != nullptr) {   // bias of the GEMM that produced `res`, deferred into this kernel
          const float4 rb = ldg_f4(res_bias + idx * 4);
          a.x += rb.x; a.y += rb.y; a.z += rb.z; a.w += rb.w;
        }
      }
      if (sum_out != nullptr) *reinterpret_cast<float4*>(sum_out + row * C + idx * 4) = a;
      v[i] = a;
      s += (a.x + a.y) + (a.z + a.w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      const float4 gm = ldg_f4(gamma + idx * 4), bt = ldg_f4(beta + idx * 4);
      float4 o;
      o.x = (v[i].x - mean) * rstd * gm.x + bt.x;
      o.y = (v[i].y - mean) * rstd * gm.y + bt.y;
      o.z = (v[i].z - mean) * rstd * gm.z + bt.z;
      o.w = (v[i].w - mean) * rstd * gm.w + bt.w;
      store_maybe_split(out, (size_t)row, C, idx * 4, o, split);
    }
  }
}

__global__ void __launch_bounds__(256)
gelu_split_kernel(const float* __restrict__ x, const float* __restrict__ bias, long long rows, int C,
                  float* __restrict__ out, int do_gelu, int split) {
  const int nv = C >> 2;
  const long long total = rows * nv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / nv;
    const int col = (int)(i - row * nv) * 4;
    float4 a = *reinterpret_cast<const float4*>(x + row * C + col);
    if (bias != nullptr) {   // bias of the producing GEMM, deferred into this kernel
      const float4 b = ldg_f4(bias + col);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (do_gelu == 2) {
      a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
    } else if (do_gelu == 1) {
      a.x = 0.5f * a.x * (1.f + erff(a.x * 0.70710678118654752440f));
      a.y = 0.5f * a.y * (1.f + erff(a.y * 0.70710678118654752440f));
      a.z = 0.5f * a.z * (1.f + erff(a.z * 0.70710678118654752440f));
      a.w = 0.5f * a.w * (1.f + erff(a.w * 0.70710678118654752440f));
    }
    store_maybe_split(out, (size_t)row, C, col, a, split);
  }
}

}  // namespace univs

using namespace univs;

extern "C" int univs_layernorm_f32(void* stream, const float* x, const float* residual, const float* residual_bias,
                                   const float* gamma, const float* beta, int64_t rows, int channels, float eps,
                                   float* sum_out, float* out, int split) {
  UNIVS_REQUIRE(rows >= 0 && channels > 0, "layernorm: bad sizes");
  if (rows == 0) return UNIVS_OK;
  UNIVS_REQUIRE(x && gamma && beta && out, "layernorm: null pointer");
  UNIVS_REQUIRE(residual_bias == nullptr || residual != nullptr, "layernorm: residual_bias needs a residual");
  UNIVS_REQUIRE(channels % 4 == 0 && channels <= 4096, "layernorm: channels must be a multiple of 4 and <= 4096 (got %d)", channels);
  UNIVS_REQUIRE(split == 0 || split == -2 || (split != -1 && split != -3 && (split > 0 ? split : -split) % 4 == 0 &&
                                               channels % (split > 0 ? split : -split) == 0),
                "layernorm: split chunk must divide channels");
  const unsigned grid = (unsigned)((rows + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
#define LN_LAUNCH(MV) layernorm_kernel<MV><<<grid, 256, 0, st>>>(x, residual, residual_bias, gamma, beta, rows, channels, eps, sum_out, out, split)
  if (channels <= 128) LN_LAUNCH(1);
  else if (channels <= 256) LN_LAUNCH(2);
  else if (channels <= 512) LN_LAUNCH(4);
  else if (channels <= 1024) LN_LAUNCH(8);
  else if (channels <= 2048) LN_LAUNCH(16);
  else LN_LAUNCH(32);
#undef LN_LAUNCH
  return check_launch("layernorm");
}

extern "C" int univs_gelu_f32(void* stream, const float* x, const float* bias, int64_t rows, int channels, float* out, int split) {
  UNIVS_REQUIRE(rows >= 0 && channels > 0 && channels % 4 == 0, "gelu: bad sizes (channels %% 4 == 0)");
  if (rows == 0) return UNIVS_OK;
  UNIVS_REQUIRE(x && out, "gelu: null pointer");
  long long blocks = (rows * (channels / 4) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gelu_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, bias, rows, channels, out, 1, split);
  return check_launch("gelu");
}

extern "C" int univs_relu_f32(void* stream, const float* x, const float* bias, int64_t rows, int channels, float* out, int split) {
  UNIVS_REQUIRE(rows >= 0 && channels > 0 && channels % 4 == 0, "relu: bad sizes (channels %% 4 == 0)");
  if (rows == 0) return UNIVS_OK;
  UNIVS_REQUIRE(x && out, "relu: null pointer");
  long long blocks = (rows * (channels / 4) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gelu_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, bias, rows, channels, out, 2, split);
  return check_launch("relu");
}

extern "C" int univs_split_tf32_f32(void* stream, const float* x, int64_t rows, int channels, int chunk, float* out) {
  UNIVS_REQUIRE(rows >= 0 && channels > 0 && channels % 4 == 0, "split_tf32: bad sizes (channels %% 4 == 0)");
  UNIVS_REQUIRE(chunk == -2 || (chunk != 0 && chunk != -1 && chunk != -3 && (chunk > 0 ? chunk : -chunk) % 4 == 0 &&
                               channels % (chunk > 0 ? chunk : -chunk) == 0),
                "split_tf32: |chunk| must divide channels and be a multiple of 4 (or -2 for the fp16 [hi|lo] format)");
  if (rows == 0) return UNIVS_OK;
  UNIVS_REQUIRE(x && out, "split_tf32: null pointer");
  long long blocks = (rows * (channels / 4) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gelu_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, nullptr, rows, channels, out, 0, chunk);
  return check_launch("split_tf32");
}
