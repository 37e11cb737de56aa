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
#include <stdlib.h>

#include "rowwise.cuh"

namespace univs {

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

// v2 of the bias / GELU / ReLU / operand-split stream for the fp16 operand formats (opt-in: UNIVS_ROWWISE_V2=1; results are
// bit-identical to gelu_split_kernel -- same erff, same round-to-nearest conversions in the same order).  The v1 kernel is
// issue-bound, not HBM-bound (ncu launch list: 0.55 ms for 2.26 GB at Swin-L stage 1 = 4.1 TB/s): a 64-bit division per
// float4, six 4-byte stores per float4 and scalar fp16 conversions.  Here a thread owns 8 consecutive columns: 32-bit
// index arithmetic (the host bounds rows * C / 8 below 2^31), two 128-bit loads, packed f16x2 conversions and one 128-bit
// store per output segment.
__device__ __forceinline__ uint32_t pack_sat_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }

template <int MODE>   // 0 = split only, 1 = GELU, 2 = ReLU
__global__ void __launch_bounds__(256)
gelu_split8_kernel(const float* __restrict__ x, const float* __restrict__ bias, int total8, int C8, int C, __half* __restrict__ out,
                   int split) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += gridDim.x * blockDim.x) {
    const int row = i / C8;
    const int col = (i - row * C8) * 8;
    const float4* src = reinterpret_cast<const float4*>(x + (size_t)row * C + col);
    const float4 a0 = src[0], a1 = src[1];
    float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    if (bias != nullptr) {
      const float4 b0 = ldg_f4(bias + col), b1 = ldg_f4(bias + col + 4);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (MODE == 2) v[e] = fmaxf(v[e], 0.f);
      if (MODE == 1) v[e] = 0.5f * v[e] * (1.f + erff(v[e] * 0.70710678118654752440f));
    }
    const float sc = (split == -2) ? 1.f : 2048.f;
    uint32_t hi[4], lo[4], hs[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[e] = pack_sat_h2(v[2 * e], v[2 * e + 1]);
      const float2 hf = unpack_h2(hi[e]);
      lo[e] = pack_sat_h2((v[2 * e] - hf.x) * sc, (v[2 * e + 1] - hf.y) * sc);
      const __half2 s2 = __floats2half2_rn(hf.x * (1.f / 2048.f), hf.y * (1.f / 2048.f));
      hs[e] = *reinterpret_cast<const uint32_t*>(&s2);
    }
    if (split == -2 || split == -3) {          // [hi | lo] / [hi | lo*2^11]
      __half* o = out + (size_t)row * (2 * (size_t)C) + col;
      *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(o + C) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    } else {                    // K-chunks [lo*2^11 | hi*2^-11 | hi]
      const int kc = -split;
      const int chunk = col / kc;
      __half* o = out + (size_t)row * (3 * (size_t)C) + (size_t)chunk * (3 * kc) + (col - chunk * kc);
      *reinterpret_cast<uint4*>(o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<uint4*>(o + kc) = make_uint4(hs[0], hs[1], hs[2], hs[3]);
      *reinterpret_cast<uint4*>(o + 2 * kc) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    }
  }
}

// UNIVS_ROWWISE_V2: bit 0 = 8-wide GELU / ReLU / split kernel, bit 1 = wide-store LayerNorm, bit 2 = streaming LayerNorm for
// the plain compact-operand case ("0" .. "7")
static int rowwise_v2_bits() {
  static int bits = -1;
  if (bits < 0) {
    const char* e = getenv("UNIVS_ROWWISE_V2");
    bits = (e != nullptr && e[0] >= '0' && e[0] <= '7') ? e[0] - '0' : 0;
  }
  return bits;
}
static bool rowwise_v2_enabled() { return (rowwise_v2_bits() & 1) != 0; }

// LayerNorm with 128-bit operand stores (opt-in: UNIVS_ROWWISE_V2 bit 1).  Loads, statistics and the normalisation are
// layernorm_kernel's, element for element and in the same order (same lane -> column mapping, same warp reductions), so the
// results are bit-identical; only the fp16 operand stores differ: layernorm_kernel issues six 4-byte stores per float4
// (store_maybe_split), here even lanes fetch their odd neighbour's packed halves with shuffles and write one 16-byte store
// per operand segment -- 4x fewer store instructions on a kernel that writes 6 bytes per element.
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_wide_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ res_bias,
                      const float* __restrict__ gamma, const float* __restrict__ beta, long long rows, int C, float eps,
                      float* __restrict__ sum_out, __half* __restrict__ out, int split) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = C >> 2;  // float4 per row (even: C % 8 == 0)
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
        if (res_bias != nullptr) {
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
  const float sc = (split == -2) ? 1.f : 2048.f;
  const int kc = -split;
  const bool two_blocks = split == -2 || split == -3;      // [hi | lo] containers: no hi * 2^-11 copy to compute or exchange
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    const bool has = idx < nv;                       // the same for both lanes of a pair (nv is even)
    uint32_t hi[2] = {0u, 0u}, lo[2] = {0u, 0u}, hs[2] = {0u, 0u};
    if (has) {
      const float4 gm = ldg_f4(gamma + idx * 4), bt = ldg_f4(beta + idx * 4);
      float o[4];
      o[0] = (v[i].x - mean) * rstd * gm.x + bt.x;
      o[1] = (v[i].y - mean) * rstd * gm.y + bt.y;
      o[2] = (v[i].z - mean) * rstd * gm.z + bt.z;
      o[3] = (v[i].w - mean) * rstd * gm.w + bt.w;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        hi[e] = pack_sat_h2(o[2 * e], o[2 * e + 1]);
        const float2 hf = unpack_h2(hi[e]);
        lo[e] = pack_sat_h2((o[2 * e] - hf.x) * sc, (o[2 * e + 1] - hf.y) * sc);
        if (!two_blocks) {
          const __half2 s2 = __floats2half2_rn(hf.x * (1.f / 2048.f), hf.y * (1.f / 2048.f));
          hs[e] = *reinterpret_cast<const uint32_t*>(&s2);
        }
      }
    }
    // every lane takes part in the exchange; even lanes store their own four halves followed by the odd neighbour's
    const uint32_t nh0 = __shfl_down_sync(0xffffffffu, hi[0], 1), nh1 = __shfl_down_sync(0xffffffffu, hi[1], 1);
    const uint32_t nl0 = __shfl_down_sync(0xffffffffu, lo[0], 1), nl1 = __shfl_down_sync(0xffffffffu, lo[1], 1);
    uint32_t ns0 = 0u, ns1 = 0u;
    if (!two_blocks) {                               // uniform branch (kernel argument)
      ns0 = __shfl_down_sync(0xffffffffu, hs[0], 1);
      ns1 = __shfl_down_sync(0xffffffffu, hs[1], 1);
    }
    if (has && (lane & 1) == 0) {
      const int col = idx * 4;                       // multiple of 8
      if (split == -2 || split == -3) {              // [hi | lo] / [hi | lo*2^11]
        __half* o16 = out + (size_t)row * (2 * (size_t)C) + col;
        *reinterpret_cast<uint4*>(o16) = make_uint4(hi[0], hi[1], nh0, nh1);
        *reinterpret_cast<uint4*>(o16 + C) = make_uint4(lo[0], lo[1], nl0, nl1);
      } else {                                       // K-chunks [lo*2^11 | hi*2^-11 | hi]
        const int chunk = col / kc;
        __half* o16 = out + (size_t)row * (3 * (size_t)C) + (size_t)chunk * (3 * kc) + (col - chunk * kc);
        *reinterpret_cast<uint4*>(o16) = make_uint4(lo[0], lo[1], nl0, nl1);
        *reinterpret_cast<uint4*>(o16 + kc) = make_uint4(hs[0], hs[1], ns0, ns1);
        *reinterpret_cast<uint4*>(o16 + 2 * kc) = make_uint4(hi[0], hi[1], nh0, nh1);
      }
    }
  }
}

// Streaming LayerNorm (opt-in: UNIVS_ROWWISE_V2 bit 2) for the case the Swin blocks of the default path launch 48 times per
// clip: no residual, no fp32 sum, compact operand output [hi | lo*2^11].  layernorm_wide_kernel gives every warp ONE row and
// lets the block retire (36800 blocks at Swin-L stage 1): a warp's life is load -> two dependent warp reductions -> stores, so
// half of the resident warps are not loading at any time and the launch runs at 3.7 TB/s (ncu launch list, all four stages).
// Here a warp walks rows with a grid stride and loads its NEXT row before it reduces the current one (two rows in flight per
// warp, no block turnover).  Per row the arithmetic is layernorm_wide_kernel's, element for element and in the same order:
// the results are bit-identical.
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_stream_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                        long long rows, int C, float eps, __half* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int nv = C >> 2;  // float4 per row (even: C % 8 == 0)
  const long long stride = (long long)gridDim.x * 8;
  long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  float4 v[MAXV], vn[MAXV];
  if (row < rows) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + i * 32;
      vn[i] = idx < nv ? *reinterpret_cast<const float4*>(x + row * C + idx * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (; row < rows; row += stride) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      v[i] = vn[i];
      if (lane + i * 32 < nv) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const long long nxt = row + stride;
    if (nxt < rows) {
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int idx = lane + i * 32;
        if (idx < nv) vn[i] = *reinterpret_cast<const float4*>(x + nxt * C + idx * 4);
      }
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (lane + i * 32 < nv) {
        const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + i * 32;
      const bool has = idx < nv;                       // the same for both lanes of a pair (nv is even)
      uint32_t hi[2] = {0u, 0u}, lo[2] = {0u, 0u};
      if (has) {
        const float4 gm = ldg_f4(gamma + idx * 4), bt = ldg_f4(beta + idx * 4);
        float o[4];
        o[0] = (v[i].x - mean) * rstd * gm.x + bt.x;
        o[1] = (v[i].y - mean) * rstd * gm.y + bt.y;
        o[2] = (v[i].z - mean) * rstd * gm.z + bt.z;
        o[3] = (v[i].w - mean) * rstd * gm.w + bt.w;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          hi[e] = pack_sat_h2(o[2 * e], o[2 * e + 1]);
          const float2 hf = unpack_h2(hi[e]);
          lo[e] = pack_sat_h2((o[2 * e] - hf.x) * 2048.f, (o[2 * e + 1] - hf.y) * 2048.f);
        }
      }
      const uint32_t nh0 = __shfl_down_sync(0xffffffffu, hi[0], 1), nh1 = __shfl_down_sync(0xffffffffu, hi[1], 1);
      const uint32_t nl0 = __shfl_down_sync(0xffffffffu, lo[0], 1), nl1 = __shfl_down_sync(0xffffffffu, lo[1], 1);
      if (has && (lane & 1) == 0) {
        __half* o16 = out + (size_t)row * (2 * (size_t)C) + idx * 4;
        *reinterpret_cast<uint4*>(o16) = make_uint4(hi[0], hi[1], nh0, nh1);
        *reinterpret_cast<uint4*>(o16 + C) = make_uint4(lo[0], lo[1], nl0, nl1);
      }
    }
  }
}

// returns true when the v2 kernel took the launch
static bool launch_split8(cudaStream_t st, const float* x, const float* bias, int64_t rows, int C, float* out, int mode, int split) {
  if (!rowwise_v2_enabled() || split >= 0 || C % 8 != 0) return false;
  if (split != -2 && split != -3 && ((-split) % 8 != 0)) return false;
  const int64_t total8 = rows * (C / 8);
  if (total8 >= (1ll << 31) - (1ll << 24)) return false;
  if (((uintptr_t)x | (uintptr_t)out) & 15) return false;
  int64_t blocks = (total8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  __half* o16 = reinterpret_cast<__half*>(out);
  if (mode == 1) gelu_split8_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(x, bias, (int)total8, C / 8, C, o16, split);
  else if (mode == 2) gelu_split8_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(x, bias, (int)total8, C / 8, C, o16, split);
  else gelu_split8_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(x, bias, (int)total8, C / 8, C, o16, split);
  return true;
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
  UNIVS_REQUIRE(split == 0 || split == -2 || split == -3 || (split != -1 && (split > 0 ? split : -split) % 4 == 0 &&
                                               channels % (split > 0 ? split : -split) == 0),
                "layernorm: split chunk must divide channels");
  const unsigned grid = (unsigned)((rows + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if ((rowwise_v2_bits() & 4) && split == -3 && residual == nullptr && sum_out == nullptr && channels % 8 == 0 && channels <= 1024 &&
      (((uintptr_t)x | (uintptr_t)out) & 15) == 0) {
    __half* o16 = reinterpret_cast<__half*>(out);
    // every block resident (occupancy x SMs), so that each warp walks several rows and its prefetch pays
#define LNS_LAUNCH(MV)                                                                                                  \
  do {                                                                                                                  \
    static int resident = 0;                                                                                            \
    if (!resident) {                                                                                                    \
      int per_sm = 0, dev = 0, sms = 0;                                                                                 \
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, layernorm_stream_kernel<MV>, 256, 0) != cudaSuccess || \
          cudaGetDevice(&dev) != cudaSuccess ||                                                                         \
          cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || per_sm <= 0 || sms <= 0) { \
        cudaGetLastError();                                                                                             \
        per_sm = 2;                                                                                                     \
        sms = 148;                                                                                                      \
      }                                                                                                                 \
      resident = per_sm * sms;                                                                                          \
    }                                                                                                                   \
    const unsigned sgrid = grid < (unsigned)resident ? grid : (unsigned)resident;                                       \
    layernorm_stream_kernel<MV><<<sgrid, 256, 0, st>>>(x, gamma, beta, rows, channels, eps, o16);                       \
  } while (0)
    if (channels <= 128) LNS_LAUNCH(1);
    else if (channels <= 256) LNS_LAUNCH(2);
    else if (channels <= 512) LNS_LAUNCH(4);
    else LNS_LAUNCH(8);
#undef LNS_LAUNCH
    return check_launch("layernorm_stream");
  }
  if ((rowwise_v2_bits() & 2) && split < 0 && channels % 8 == 0 && (split == -2 || split == -3 || (-split) % 8 == 0) &&
      (((uintptr_t)x | (uintptr_t)out | (uintptr_t)residual | (uintptr_t)sum_out) & 15) == 0) {
    __half* o16 = reinterpret_cast<__half*>(out);
#define LNW_LAUNCH(MV) layernorm_wide_kernel<MV><<<grid, 256, 0, st>>>(x, residual, residual_bias, gamma, beta, rows, channels, eps, sum_out, o16, split)
    if (channels <= 128) LNW_LAUNCH(1);
    else if (channels <= 256) LNW_LAUNCH(2);
    else if (channels <= 512) LNW_LAUNCH(4);
    else if (channels <= 1024) LNW_LAUNCH(8);
    else if (channels <= 2048) LNW_LAUNCH(16);
    else LNW_LAUNCH(32);
#undef LNW_LAUNCH
    return check_launch("layernorm_wide");
  }
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
  if (launch_split8((cudaStream_t)stream, x, bias, rows, channels, out, 1, split)) return check_launch("gelu");
  gelu_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, bias, rows, channels, out, 1, split);
  return check_launch("gelu");
}

extern "C" int univs_relu_f32(void* stream, const float* x, const float* bias, int64_t rows, int channels, float* out, int split) {
  UNIVS_REQUIRE(rows >= 0 && channels > 0 && channels % 4 == 0, "relu: bad sizes (channels %% 4 == 0)");
  if (rows == 0) return UNIVS_OK;
  UNIVS_REQUIRE(x && out, "relu: null pointer");
  long long blocks = (rows * (channels / 4) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (launch_split8((cudaStream_t)stream, x, bias, rows, channels, out, 2, split)) return check_launch("relu");
  gelu_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, bias, rows, channels, out, 2, split);
  return check_launch("relu");
}

extern "C" int univs_split_tf32_f32(void* stream, const float* x, int64_t rows, int channels, int chunk, float* out) {
  UNIVS_REQUIRE(rows >= 0 && channels > 0 && channels % 4 == 0, "split_tf32: bad sizes (channels %% 4 == 0)");
  UNIVS_REQUIRE(chunk == -2 || chunk == -3 || (chunk != 0 && chunk != -1 && (chunk > 0 ? chunk : -chunk) % 4 == 0 &&
                               channels % (chunk > 0 ? chunk : -chunk) == 0),
                "split_tf32: |chunk| must divide channels and be a multiple of 4 (or -2 for the fp16 [hi|lo] format)");
  if (rows == 0) return UNIVS_OK;
  UNIVS_REQUIRE(x && out, "split_tf32: null pointer");
  long long blocks = (rows * (channels / 4) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (launch_split8((cudaStream_t)stream, x, nullptr, rows, channels, out, 0, chunk)) return check_launch("split_tf32");
  gelu_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, nullptr, rows, channels, out, 0, chunk);
  return check_launch("split_tf32");
}
