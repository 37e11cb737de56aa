// Mask-logit einsum "btqc,btchw->btqhw" (+ transpose(1,2)) for sm_100a
// (reference: video_mask2former_transformer_decoder_univs.py:527-528; T=1 form 'bqc,bchw->bqhw' at
// mask2former_transformer_decoder.py:472).
//
// out[q, t, p] = sum_c E[t, q, c] * F[t, p, c]      E = mask_embed [T,Q,C], F = mask_features, channel-last [T,HW,C]
//
// HBM-bound (AI = 56 F/B at Q=200, C=256): per frame the kernel must stream F once (C*HW*4 B) and write the logits
// once (Q*HW*4 B).  Operands are rounded to nearest TF32 (unbiased, |err| <= 2^-12 relative), products accumulate
// in fp32.  This file holds the register-operand (mma.sync) version: CTA tile = all Q rows x 64 pixels, K streamed in
// 32-channel chunks through a cp.async double buffer.  (The TMA + tcgen05/TMEM version lives in
// mask_einsum_tc.cu once enabled.)
#include "common.cuh"

namespace univs {

constexpr int kEStride = 36;
constexpr int kPxTile = 64;
constexpr int kKChunk = 32;
constexpr int kEinThreads = 256;  // 8 warps, warp w owns query rows [32w, 32w+32)

template <bool X3>
__global__ void __launch_bounds__(kEinThreads)
mask_einsum_kernel(const float* __restrict__ E, const float* __restrict__ F, int T, int Q, int C, int HW,
                   float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  // stage layout: [2][ E: 256 x 36 | F: 64 x 36 ]
  constexpr int kStage = (256 + kPxTile) * kEStride;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int tiles_per_frame = (HW + kPxTile - 1) / kPxTile;
  const int frame = blockIdx.x / tiles_per_frame;
  const int p0 = (blockIdx.x - frame * tiles_per_frame) * kPxTile;
  const float* Ef = E + (size_t)frame * Q * C;
  const float* Ff = F + (size_t)frame * HW * C;
  const int qrows = (Q + 15) & ~15;

  auto stage = [&](int kc, int buf) {
    float* Es = sm + buf * kStage;
    float* Fs = Es + 256 * kEStride;
    const int c0 = kc * kKChunk;
    for (int idx = tid; idx < qrows * 8; idx += kEinThreads) {
      const int r = idx >> 3, ch = idx & 7;
      float* dst = Es + r * kEStride + ch * 4;
      if (r < Q) cp_async16(dst, Ef + (size_t)r * C + c0 + ch * 4);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int idx = tid; idx < kPxTile * 8; idx += kEinThreads) {
      const int r = idx >> 3, ch = idx & 7;
      float* dst = Fs + r * kEStride + ch * 4;
      if (p0 + r < HW) cp_async16(dst, Ff + (size_t)(p0 + r) * C + c0 + ch * 4);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };

  float acc[2][8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;

  const int nchunks = C / kKChunk;
  const bool act0 = warp * 32 < Q, act1 = warp * 32 + 16 < Q;
  stage(0, 0);
  for (int kc = 0; kc < nchunks; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < nchunks) {
      stage(kc + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* Es = sm + buf * kStage;
    const float* Fs = Es + 256 * kEStride;
    if (act0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[2][4], as[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int r = warp * 32 + mt * 16 + g;
          const float e0 = Es[r * kEStride + ks * 8 + t4], e1 = Es[(r + 8) * kEStride + ks * 8 + t4];
          const float e2 = Es[r * kEStride + ks * 8 + t4 + 4], e3 = Es[(r + 8) * kEStride + ks * 8 + t4 + 4];
          if (X3) {
            split_tf32(e0, a[mt][0], as[mt][0]); split_tf32(e1, a[mt][1], as[mt][1]);
            split_tf32(e2, a[mt][2], as[mt][2]); split_tf32(e3, a[mt][3], as[mt][3]);
          } else {
            a[mt][0] = f2tf32(e0); a[mt][1] = f2tf32(e1); a[mt][2] = f2tf32(e2); a[mt][3] = f2tf32(e3);
          }
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const float f0 = Fs[(nt * 8 + g) * kEStride + ks * 8 + t4], f1 = Fs[(nt * 8 + g) * kEStride + ks * 8 + t4 + 4];
          if (X3) {
            uint32_t b0, b1, s0, s1;
            split_tf32(f0, b0, s0);
            split_tf32(f1, b1, s1);
            mma_tf32x3(acc[0][nt], a[0], as[0], b0, b1, s0, s1);
            if (act1) mma_tf32x3(acc[1][nt], a[1], as[1], b0, b1, s0, s1);
          } else {
            const uint32_t b0 = f2tf32(f0), b1 = f2tf32(f1);
            mma_tf32(acc[0][nt], a[0], b0, b1);
            if (act1) mma_tf32(acc[1][nt], a[1], b0, b1);
          }
        }
      }
    }
    __syncthreads();
  }
  // store: out[q][frame][p]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int r0 = warp * 32 + mt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int p = p0 + nt * 8 + 2 * t4;
      if (p + 1 < HW) {
        if (r0 < Q) *reinterpret_cast<float2*>(out + ((size_t)r0 * T + frame) * HW + p) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
        if (r1 < Q) *reinterpret_cast<float2*>(out + ((size_t)r1 * T + frame) * HW + p) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
      } else if (p < HW) {
        if (r0 < Q) out[((size_t)r0 * T + frame) * HW + p] = acc[mt][nt][0];
        if (r1 < Q) out[((size_t)r1 * T + frame) * HW + p] = acc[mt][nt][2];
      }
    }
  }
}

}  // namespace univs

namespace univs {
int launch_mask_einsum_tc(cudaStream_t st, const float* E, const float* F, int T, int Q, int C, int HW, float* out);
int launch_mask_einsum_tc_f16(cudaStream_t st, const void* E16, const void* F16, int T, int Q, int C, int HW, float* out);
int launch_mask_einsum_mc_f16(cudaStream_t st, const void* E16, const void* F16, int T, int Q, int C, int HW, float* out);
}
using namespace univs;

static int check_einsum_args(const char* who, const float* e, const float* f, const float* out, int frames,
                             int queries, int channels, int pixels) {
  UNIVS_REQUIRE(frames >= 0 && queries >= 0 && pixels >= 0, "%s: negative size", who);
  UNIVS_REQUIRE(queries <= 256, "%s: at most 256 queries per call (got %d)", who, queries);
  UNIVS_REQUIRE(channels > 0 && channels % 32 == 0, "%s: channels must be a multiple of 32", who);
  if (frames == 0 || queries == 0 || pixels == 0) return 1;
  UNIVS_REQUIRE(e && f && out, "%s: null pointer", who);
  return 0;
}

extern "C" int univs_mask_einsum_f32(void* stream, const float* mask_embed, const float* mask_features_cl,
                                     int frames, int queries, int channels, int pixels, float* out) {
  int rc = check_einsum_args("mask_einsum", mask_embed, mask_features_cl, out, frames, queries, channels, pixels);
  if (rc) return rc < 0 ? rc : UNIVS_OK;
  UNIVS_REQUIRE(((uintptr_t)mask_embed & 15) == 0 && ((uintptr_t)mask_features_cl & 15) == 0,
                "mask_einsum: operands must be 16-byte aligned (TMA)");
  return launch_mask_einsum_tc((cudaStream_t)stream, mask_embed, mask_features_cl, frames, queries, channels, pixels, out);
}

extern "C" int univs_mask_einsum_f16x3(void* stream, const void* mask_embed16, const void* mask_features16, int frames,
                                       int queries, int channels, int pixels, float* out) {
  int rc = check_einsum_args("mask_einsum_f16x3", (const float*)mask_embed16, (const float*)mask_features16, out, frames,
                             queries, channels, pixels);
  if (rc) return rc < 0 ? rc : UNIVS_OK;
  UNIVS_REQUIRE(channels % 32 == 0, "mask_einsum_f16x3: channels must be a multiple of 32");
  UNIVS_REQUIRE(((uintptr_t)mask_embed16 & 15) == 0 && ((uintptr_t)mask_features16 & 15) == 0,
                "mask_einsum_f16x3: operands must be 16-byte aligned (TMA)");
  return launch_mask_einsum_tc_f16((cudaStream_t)stream, mask_embed16, mask_features16, frames, queries, channels, pixels, out);
}

extern "C" int univs_mask_einsum_f16x3_cluster(void* stream, const void* mask_embed16, const void* mask_features16, int frames,
                                               int queries, int channels, int pixels, float* out) {
  int rc = check_einsum_args("mask_einsum_f16x3_cluster", (const float*)mask_embed16, (const float*)mask_features16, out,
                             frames, queries, channels, pixels);
  if (rc) return rc < 0 ? rc : UNIVS_OK;
  UNIVS_REQUIRE(channels % 32 == 0, "mask_einsum_f16x3_cluster: channels must be a multiple of 32");
  UNIVS_REQUIRE(queries > 16, "mask_einsum_f16x3_cluster: the query split needs more than 16 queries");
  UNIVS_REQUIRE(((uintptr_t)mask_embed16 & 15) == 0 && ((uintptr_t)mask_features16 & 15) == 0,
                "mask_einsum_f16x3_cluster: operands must be 16-byte aligned (TMA)");
  return launch_mask_einsum_mc_f16((cudaStream_t)stream, mask_embed16, mask_features16, frames, queries, channels, pixels, out);
}

extern "C" int univs_mask_einsum_mma_f32(void* stream, const float* mask_embed, const float* mask_features_cl,
                                         int frames, int queries, int channels, int pixels, int precision, float* out) {
  int rc = check_einsum_args("mask_einsum_mma", mask_embed, mask_features_cl, out, frames, queries, channels, pixels);
  if (rc) return rc < 0 ? rc : UNIVS_OK;
  UNIVS_REQUIRE(pixels % 2 == 0, "mask_einsum_mma: pixels must be even");
  UNIVS_REQUIRE(precision == UNIVS_PREC_TF32X3 || precision == UNIVS_PREC_TF32, "mask_einsum_mma: bad precision");
  const size_t smem = 2 * (256 + kPxTile) * kEStride * sizeof(float);
  const long long grid = (long long)frames * ((pixels + kPxTile - 1) / kPxTile);
  UNIVS_REQUIRE(grid < (1ll << 31), "mask_einsum_mma: problem too large");
  cudaError_t e;
  if (precision == UNIVS_PREC_TF32X3) {
    e = cudaFuncSetAttribute(mask_einsum_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("mask_einsum_mma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNIVS_E_LAUNCH; }
    mask_einsum_kernel<true><<<(unsigned)grid, kEinThreads, smem, (cudaStream_t)stream>>>(mask_embed, mask_features_cl, frames, queries, channels, pixels, out);
  } else {
    e = cudaFuncSetAttribute(mask_einsum_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("mask_einsum_mma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNIVS_E_LAUNCH; }
    mask_einsum_kernel<false><<<(unsigned)grid, kEinThreads, smem, (cudaStream_t)stream>>>(mask_embed, mask_features_cl, frames, queries, channels, pixels, out);
  }
  return check_launch("mask_einsum_mma");
}
