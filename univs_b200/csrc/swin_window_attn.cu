// Swin (shifted-)window attention for sm_100a: QK^T (+relative-position bias, +shift mask) -> softmax -> PV,
// with the reference's pad / cyclic-roll / window-partition / window-reverse / un-roll / crop data movement
// (swin.py:247-289, 44-71) folded into the load and store addressing, so none of those six full-activation
// copies exists.  One CTA = one (window, head); each warp owns 16 query rows; scores live in registers
// (flash-style two key blocks for 12x12 windows), operands are staged once in padded shared memory
// (row stride 36 floats => conflict-free fragment loads).  Tensor-core products are 3xTF32 (fp32-equivalent)
// or single TF32, fp32 accumulate.
//
// Reference semantics restated: swin.py:131-171 (attention), :108-121 (relative_position_index, computed
// analytically here), :413-440 (shift mask = -100 between different regions of the PADDED, SHIFTED grid),
// :247-255 (zero pad AFTER norm1 => qkv of a pad token == qkv bias; pad tokens are real keys).
// The qkv input is the bias-free GEMM output; the kernel adds the bias while staging (one less pass over qkv).
#include "common.cuh"

namespace univs {

constexpr int kRowStride = 36;  // floats; 32 + 4 pad

template <int WS>
struct WinCfg {
  static constexpr int N = WS * WS;
  static constexpr int MT = (N + 15) / 16;          // 16-row query tiles == warps
  static constexpr int NT = (N + 7) / 8;            // 8-key tiles
  static constexpr int NBLK = (NT > 10) ? 2 : 1;    // key blocks (register budget)
  static constexpr int NT_BLK = NT / NBLK;
  static_assert(NT % NBLK == 0, "key tiles must split evenly");
  static constexpr int QROWS = MT * 16;
  static constexpr int KROWS = NT * 8;
  static constexpr int TABLE = (2 * WS - 1) * (2 * WS - 1);
  static constexpr int THREADS = MT * 32;
  static constexpr size_t SMEM =
      sizeof(float) * ((size_t)(QROWS + 2 * KROWS) * kRowStride + TABLE) + sizeof(int) * QROWS + QROWS;
};

template <int WS, bool X3>
__global__ void __launch_bounds__(WinCfg<WS>::THREADS, (WS >= 12 ? 2 : 4))
swin_window_attn_kernel(const float* __restrict__ qkv, const float* __restrict__ qkv_bias,
                        const float* __restrict__ table, int B, int H, int W, int C, int nH, int shift,
                        float scale, float* __restrict__ out) {
  using Cfg = WinCfg<WS>;
  constexpr int N = Cfg::N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Qs = reinterpret_cast<float*>(smem_raw);
  float* Ks = Qs + Cfg::QROWS * kRowStride;
  float* Vs = Ks + Cfg::KROWS * kRowStride;
  float* sBias = Vs + Cfg::KROWS * kRowStride;
  int* sSrc = reinterpret_cast<int*>(sBias + Cfg::TABLE);
  unsigned char* sLab = reinterpret_cast<unsigned char*>(sSrc + Cfg::QROWS);

  const int Hp = (H + WS - 1) / WS * WS, Wp = (W + WS - 1) / WS * WS;
  const int nWw = Wp / WS, nWh = Hp / WS;
  const int head = blockIdx.y;
  int win = blockIdx.x;
  const int wx = win % nWw;
  win /= nWw;
  const int wy = win % nWh;
  const int b = win / nWh;
  const int tid = threadIdx.x;

  // ---- token table: source pixel of every window slot (or -1 for a pad token) + shift-mask region label
  for (int i = tid; i < Cfg::QROWS; i += Cfg::THREADS) {
    int src = -1;
    int lab = 0;
    if (i < N) {
      const int iy = i / WS, ix = i - iy * WS;
      const int hp = wy * WS + iy, wp = wx * WS + ix;  // coordinates in the rolled, padded grid
      int hs = hp + shift, wsrc = wp + shift;          // roll(-shift): rolled[h] = x[(h + shift) % Hp]
      if (hs >= Hp) hs -= Hp;
      if (wsrc >= Wp) wsrc -= Wp;
      if (hs < H && wsrc < W) src = (b * H + hs) * W + wsrc;
      if (shift > 0) {
        const int lh = hp < Hp - WS ? 0 : (hp < Hp - shift ? 1 : 2);
        const int lw = wp < Wp - WS ? 0 : (wp < Wp - shift ? 1 : 2);
        lab = lh * 3 + lw;
      }
    }
    sSrc[i] = src;
    sLab[i] = (unsigned char)lab;
  }
  for (int i = tid; i < Cfg::TABLE; i += Cfg::THREADS) sBias[i] = __ldg(table + (size_t)i * nH + head);
  __syncthreads();

  // ---- stage Q (pre-scaled), K, V of this head: 8 lanes x float4 per token row
  {
    const int lane8 = tid & 7;
    const int c = head * 32 + lane8 * 4;
    const float4 bq = ldg_f4(qkv_bias + c), bk = ldg_f4(qkv_bias + C + c), bv = ldg_f4(qkv_bias + 2 * C + c);
    for (int i = tid >> 3; i < Cfg::QROWS; i += Cfg::THREADS / 8) {
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f), k = q, v = q;
      if (i < N) {
        const int src = sSrc[i];
        if (src >= 0) {
          const float* p = qkv + (size_t)src * (3 * C) + c;
          q = ldg_f4(p);
          k = ldg_f4(p + C);
          v = ldg_f4(p + 2 * C);
          q.x += bq.x; q.y += bq.y; q.z += bq.z; q.w += bq.w;   // qkv arrives without its bias
          k.x += bk.x; k.y += bk.y; k.z += bk.z; k.w += bk.w;
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
        } else {
          q = bq; k = bk; v = bv;
        }
        q.x *= scale; q.y *= scale; q.z *= scale; q.w *= scale;
      }
      *reinterpret_cast<float4*>(Qs + i * kRowStride + lane8 * 4) = q;
      if (i < Cfg::KROWS) {
        *reinterpret_cast<float4*>(Ks + i * kRowStride + lane8 * 4) = k;
        *reinterpret_cast<float4*>(Vs + i * kRowStride + lane8 * 4) = v;
      }
    }
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16 + g, r1 = r0 + 8;

  // Q fragments (A operand), 4 k-steps of 8
  uint32_t qb[4][4], qs[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const float a0 = Qs[r0 * kRowStride + ks * 8 + t], a1 = Qs[r1 * kRowStride + ks * 8 + t];
    const float a2 = Qs[r0 * kRowStride + ks * 8 + t + 4], a3 = Qs[r1 * kRowStride + ks * 8 + t + 4];
    if (X3) {
      split_tf32(a0, qb[ks][0], qs[ks][0]);
      split_tf32(a1, qb[ks][1], qs[ks][1]);
      split_tf32(a2, qb[ks][2], qs[ks][2]);
      split_tf32(a3, qb[ks][3], qs[ks][3]);
    } else {
      qb[ks][0] = f2tf32(a0); qb[ks][1] = f2tf32(a1); qb[ks][2] = f2tf32(a2); qb[ks][3] = f2tf32(a3);
    }
  }
  const int r0y = r0 / WS, r0x = r0 - r0y * WS, r1y = r1 / WS, r1x = r1 - r1y * WS;
  const int lab0 = sLab[min(r0, Cfg::QROWS - 1)], lab1 = sLab[min(r1, Cfg::QROWS - 1)];

  float o[4][4];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[nb][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

#pragma unroll
  for (int kb = 0; kb < Cfg::NBLK; ++kb) {
    float s[Cfg::NT_BLK][4];
#pragma unroll
    for (int nt = 0; nt < Cfg::NT_BLK; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const int krow = (kb * Cfg::NT_BLK + nt) * 8 + g;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float k0 = Ks[krow * kRowStride + ks * 8 + t], k1 = Ks[krow * kRowStride + ks * 8 + t + 4];
        if (X3) {
          uint32_t b0, b1, s0, s1;
          split_tf32(k0, b0, s0);
          split_tf32(k1, b1, s1);
          mma_tf32x3(s[nt], qb[ks], qs[ks], b0, b1, s0, s1);
        } else {
          mma_tf32(s[nt], qb[ks], f2tf32(k0), f2tf32(k1));
        }
      }
    }
    // bias + shift mask + key padding, running max
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < Cfg::NT_BLK; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = (kb * Cfg::NT_BLK + nt) * 8 + 2 * t + e;
        if (j < N) {
          const int jy = j / WS, jx = j - jy * WS;
          float v0 = s[nt][e], v1 = s[nt][2 + e];
          if (r0 < N) v0 += sBias[(r0y - jy + WS - 1) * (2 * WS - 1) + (r0x - jx + WS - 1)];
          if (r1 < N) v1 += sBias[(r1y - jy + WS - 1) * (2 * WS - 1) + (r1x - jx + WS - 1)];
          if (shift > 0) {
            const int lj = sLab[j];
            if (lj != lab0) v0 += -100.f;
            if (lj != lab1) v1 += -100.f;
          }
          s[nt][e] = v0;
          s[nt][2 + e] = v1;
        } else {
          s[nt][e] = -INFINITY;
          s[nt][2 + e] = -INFINITY;
        }
        bm0 = fmaxf(bm0, s[nt][e]);
        bm1 = fmaxf(bm1, s[nt][2 + e]);
      }
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);
    const float sc0 = expf(m0 - nm0), sc1 = expf(m1 - nm1);  // exp(-inf) = 0 on the first block
    m0 = nm0;
    m1 = nm1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < Cfg::NT_BLK; ++nt) {
      s[nt][0] = expf(s[nt][0] - m0);
      s[nt][1] = expf(s[nt][1] - m0);
      s[nt][2] = expf(s[nt][2] - m1);
      s[nt][3] = expf(s[nt][3] - m1);
      ps0 += s[nt][0] + s[nt][1];
      ps1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * sc0 + ps0;
    l1 = l1 * sc1 + ps1;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      o[nb][0] *= sc0; o[nb][1] *= sc0; o[nb][2] *= sc1; o[nb][3] *= sc1;
    }
    // O += P V.  The accumulator fragment of S is reused as the A fragment of P with the k-permutation
    // (k=t <-> key 2t, k=t+4 <-> key 2t+1); V rows are fetched with the same permutation.
#pragma unroll
    for (int nt = 0; nt < Cfg::NT_BLK; ++nt) {
      uint32_t pb[4], ps[4];
      if (X3) {
        split_tf32(s[nt][0], pb[0], ps[0]);
        split_tf32(s[nt][2], pb[1], ps[1]);
        split_tf32(s[nt][1], pb[2], ps[2]);
        split_tf32(s[nt][3], pb[3], ps[3]);
      } else {
        pb[0] = f2tf32(s[nt][0]); pb[1] = f2tf32(s[nt][2]); pb[2] = f2tf32(s[nt][1]); pb[3] = f2tf32(s[nt][3]);
      }
      const int kbase = (kb * Cfg::NT_BLK + nt) * 8;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        const float v0 = Vs[(kbase + 2 * t) * kRowStride + nb * 8 + g];
        const float v1 = Vs[(kbase + 2 * t + 1) * kRowStride + nb * 8 + g];
        if (X3) {
          uint32_t b0, b1, s0, s1;
          split_tf32(v0, b0, s0);
          split_tf32(v1, b1, s1);
          mma_tf32x3(o[nb], pb, ps, b0, b1, s0, s1);
        } else {
          mma_tf32(o[nb], pb, f2tf32(v0), f2tf32(v1));
        }
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;

  // ---- store through the inverse addressing (window_reverse + roll(+shift) + crop)
  const int src0 = r0 < N ? sSrc[r0] : -1, src1 = r1 < N ? sSrc[r1] : -1;
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) {
    const int c = head * 32 + nb * 8 + 2 * t;
    if (src0 >= 0) *reinterpret_cast<float2*>(out + (size_t)src0 * C + c) = make_float2(o[nb][0] * i0, o[nb][1] * i0);
    if (src1 >= 0) *reinterpret_cast<float2*>(out + (size_t)src1 * C + c) = make_float2(o[nb][2] * i1, o[nb][3] * i1);
  }
}

template <int WS>
static int launch_window(cudaStream_t st, const float* qkv, const float* bias, const float* table, int B, int H,
                         int W, int C, int nH, int shift, int precision, float* out) {
  using Cfg = WinCfg<WS>;
  const int Hp = (H + WS - 1) / WS * WS, Wp = (W + WS - 1) / WS * WS;
  const long long wins = (long long)B * (Hp / WS) * (Wp / WS);
  UNIVS_REQUIRE(wins < (1ll << 31) && nH <= 65535, "swin_window_attention: grid too large");
  dim3 grid((unsigned)wins, (unsigned)nH);
  const float scale = 0.17677669529663687f;  // 32^-0.5
  cudaError_t e;
  if (precision == UNIVS_PREC_TF32X3) {
    e = cudaFuncSetAttribute(swin_window_attn_kernel<WS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Cfg::SMEM);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNIVS_E_LAUNCH; }
    swin_window_attn_kernel<WS, true><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(qkv, bias, table, B, H, W, C, nH,
                                                                             shift, scale, out);
  } else {
    e = cudaFuncSetAttribute(swin_window_attn_kernel<WS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Cfg::SMEM);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNIVS_E_LAUNCH; }
    swin_window_attn_kernel<WS, false><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(qkv, bias, table, B, H, W, C, nH,
                                                                              shift, scale, out);
  }
  return check_launch("swin_window_attention");
}

}  // namespace univs

using namespace univs;

extern "C" int univs_swin_window_attention_f32(void* stream, const float* qkv, const float* qkv_bias,
                                               const float* rel_bias_table, int batch, int height, int width,
                                               int channels, int num_heads, int window, int shift, int precision,
                                               float* out) {
  UNIVS_REQUIRE(qkv && qkv_bias && rel_bias_table && out, "swin_window_attention: null pointer");
  UNIVS_REQUIRE(batch >= 0 && height > 0 && width > 0, "swin_window_attention: bad sizes");
  UNIVS_REQUIRE(num_heads > 0 && channels == num_heads * 32,
                "swin_window_attention: head_dim must be 32 (channels=%d heads=%d)", channels, num_heads);
  UNIVS_REQUIRE(shift >= 0 && shift < window, "swin_window_attention: shift must be in [0, window)");
  UNIVS_REQUIRE(precision == UNIVS_PREC_TF32X3 || precision == UNIVS_PREC_TF32, "swin_window_attention: bad precision");
  if (batch == 0) return UNIVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  switch (window) {
    case 4: return launch_window<4>(st, qkv, qkv_bias, rel_bias_table, batch, height, width, channels, num_heads, shift, precision, out);
    case 7: return launch_window<7>(st, qkv, qkv_bias, rel_bias_table, batch, height, width, channels, num_heads, shift, precision, out);
    case 12: return launch_window<12>(st, qkv, qkv_bias, rel_bias_table, batch, height, width, channels, num_heads, shift, precision, out);
    default:
      set_error("swin_window_attention: window %d not instantiated (4, 7, 12)", window);
      return UNIVS_E_BADARG;
  }
}
