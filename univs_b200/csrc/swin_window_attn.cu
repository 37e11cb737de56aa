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
#include <cuda_fp16.h>

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

// ---------------------------------------------------------------------------------------------------------------
// Strict-precision variant: fp16 hi|lo split operands (fp16 has TF32's 11-bit significand; hi*hi is exact in fp32 and
// the two correction terms restore ~22 bits), m16n8k16 MMAs at twice the TF32 rate and half the instruction count.
// Q, K, V are split ONCE while staging (the TF32 kernel above re-splits K/V fragments in every warp); V is staged
// transposed so that the PV B-fragments are contiguous half2 loads; the S accumulator fragments map 1:1 onto the
// m16n8k16 A fragments of P (two adjacent key tiles = one 16-key step).  Optional epilogue: emit the output directly
// in the fp16x3 GEMM operand layout [lo*2^11 | hi*2^-11 | hi] consumed by the projection GEMM.
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_h(float x, __half& h, __half& l) {
  h = __float2half_rn(x);
  l = __float2half_rn(x - __half2float(h));
}
__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
  __half2 v = __halves2half2(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int WS>
struct WinCfg16 {
  static constexpr int N = WS * WS;
  static constexpr int MT = (N + 15) / 16;
  static constexpr int NTP = ((N + 15) / 16) * 2;        // 8-key tiles, even (16-key MMA steps)
  static constexpr int KB = (NTP % 6 == 0) ? 6 : NTP;    // key tiles per register block
  static constexpr int NBLK = NTP / KB;
  static constexpr int QROWS = MT * 16;
  static constexpr int KROWS = NTP * 8;
  static constexpr int QK_STRIDE = 40;                    // halfs per row (32 + 8): conflict-free fragment loads
  static constexpr int VT_STRIDE = KROWS + 8;             // halfs per dim row of the transposed V
  static constexpr int TABLE = (2 * WS - 1) * (2 * WS - 1);
  static constexpr int THREADS = MT * 32;
  static constexpr size_t SMEM = sizeof(__half) * (2 * (size_t)QROWS * QK_STRIDE + 2 * (size_t)KROWS * QK_STRIDE +
                                                   2 * 32 * (size_t)VT_STRIDE) +
                                 sizeof(float) * TABLE + sizeof(int) * QROWS + QROWS + 16;
};

template <int WS>
__global__ void __launch_bounds__(WinCfg16<WS>::THREADS, (WS >= 12 ? 2 : 4))
swin_window_attn_f16x3_kernel(const float* __restrict__ qkv, const float* __restrict__ qkv_bias,
                              const float* __restrict__ table, int B, int H, int W, int C, int nH, int shift,
                              float scale, float* __restrict__ out, __half* __restrict__ out16) {
  using Cfg = WinCfg16<WS>;
  constexpr int N = Cfg::N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half* Qh = reinterpret_cast<__half*>(smem_raw);
  __half* Ql = Qh + Cfg::QROWS * Cfg::QK_STRIDE;
  __half* Kh = Ql + Cfg::QROWS * Cfg::QK_STRIDE;
  __half* Kl = Kh + Cfg::KROWS * Cfg::QK_STRIDE;
  __half* Vth = Kl + Cfg::KROWS * Cfg::QK_STRIDE;
  __half* Vtl = Vth + 32 * Cfg::VT_STRIDE;
  float* sBias = reinterpret_cast<float*>(Vtl + 32 * Cfg::VT_STRIDE);
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

  for (int i = tid; i < Cfg::QROWS; i += Cfg::THREADS) {
    int src = -1, lab = 0;
    if (i < N) {
      const int iy = i / WS, ix = i - iy * WS;
      const int hp = wy * WS + iy, wp = wx * WS + ix;
      int hs = hp + shift, wsrc = wp + shift;
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

  // ---- stage + split Q (pre-scaled), K, V^T of this head
  {
    const int lane8 = tid & 7;
    const int c = head * 32 + lane8 * 4;
    const float4 bq = ldg_f4(qkv_bias + c), bk = ldg_f4(qkv_bias + C + c), bv = ldg_f4(qkv_bias + 2 * C + c);
    constexpr int ROWS = (Cfg::QROWS > Cfg::KROWS) ? Cfg::QROWS : Cfg::KROWS;
    for (int i = tid >> 3; i < ROWS; i += Cfg::THREADS / 8) {
      float q[4] = {0.f, 0.f, 0.f, 0.f}, k[4] = {0.f, 0.f, 0.f, 0.f}, v[4] = {0.f, 0.f, 0.f, 0.f};
      if (i < N) {
        const int src = sSrc[i];
        float4 q4 = bq, k4 = bk, v4 = bv;
        if (src >= 0) {
          const float* p = qkv + (size_t)src * (3 * C) + c;
          const float4 a = ldg_f4(p), bb = ldg_f4(p + C), cc = ldg_f4(p + 2 * C);
          q4.x += a.x; q4.y += a.y; q4.z += a.z; q4.w += a.w;
          k4.x += bb.x; k4.y += bb.y; k4.z += bb.z; k4.w += bb.w;
          v4.x += cc.x; v4.y += cc.y; v4.z += cc.z; v4.w += cc.w;
        }
        q[0] = q4.x * scale; q[1] = q4.y * scale; q[2] = q4.z * scale; q[3] = q4.w * scale;
        k[0] = k4.x; k[1] = k4.y; k[2] = k4.z; k[3] = k4.w;
        v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
      }
      __half h[4], l[4];
      if (i < Cfg::QROWS) {
#pragma unroll
        for (int e = 0; e < 4; ++e) split_h(q[e], h[e], l[e]);
        *reinterpret_cast<uint2*>(Qh + i * Cfg::QK_STRIDE + lane8 * 4) = make_uint2(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]));
        *reinterpret_cast<uint2*>(Ql + i * Cfg::QK_STRIDE + lane8 * 4) = make_uint2(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]));
      }
      if (i < Cfg::KROWS) {
#pragma unroll
        for (int e = 0; e < 4; ++e) split_h(k[e], h[e], l[e]);
        *reinterpret_cast<uint2*>(Kh + i * Cfg::QK_STRIDE + lane8 * 4) = make_uint2(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]));
        *reinterpret_cast<uint2*>(Kl + i * Cfg::QK_STRIDE + lane8 * 4) = make_uint2(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          split_h(v[e], h[e], l[e]);
          Vth[(lane8 * 4 + e) * Cfg::VT_STRIDE + i] = h[e];
          Vtl[(lane8 * 4 + e) * Cfg::VT_STRIDE + i] = l[e];
        }
      }
    }
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16 + g, r1 = r0 + 8;

  // Q fragments: 2 k16-steps x {hi, lo}
  uint32_t qh[2][4], ql[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const int c0 = ks * 16 + 2 * t;
    qh[ks][0] = *reinterpret_cast<const uint32_t*>(Qh + r0 * Cfg::QK_STRIDE + c0);
    qh[ks][1] = *reinterpret_cast<const uint32_t*>(Qh + r1 * Cfg::QK_STRIDE + c0);
    qh[ks][2] = *reinterpret_cast<const uint32_t*>(Qh + r0 * Cfg::QK_STRIDE + c0 + 8);
    qh[ks][3] = *reinterpret_cast<const uint32_t*>(Qh + r1 * Cfg::QK_STRIDE + c0 + 8);
    ql[ks][0] = *reinterpret_cast<const uint32_t*>(Ql + r0 * Cfg::QK_STRIDE + c0);
    ql[ks][1] = *reinterpret_cast<const uint32_t*>(Ql + r1 * Cfg::QK_STRIDE + c0);
    ql[ks][2] = *reinterpret_cast<const uint32_t*>(Ql + r0 * Cfg::QK_STRIDE + c0 + 8);
    ql[ks][3] = *reinterpret_cast<const uint32_t*>(Ql + r1 * Cfg::QK_STRIDE + c0 + 8);
  }
  const int r0y = r0 / WS, r0x = r0 - r0y * WS, r1y = r1 / WS, r1x = r1 - r1y * WS;
  const int lab0 = sLab[min(r0, Cfg::QROWS - 1)], lab1 = sLab[min(r1, Cfg::QROWS - 1)];

  float o[4][4];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

#pragma unroll
  for (int kb = 0; kb < Cfg::NBLK; ++kb) {
    float s[Cfg::KB][4];
#pragma unroll
    for (int nt = 0; nt < Cfg::KB; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const int krow = (kb * Cfg::KB + nt) * 8 + g;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int c0 = ks * 16 + 2 * t;
        const uint32_t kh0 = *reinterpret_cast<const uint32_t*>(Kh + krow * Cfg::QK_STRIDE + c0);
        const uint32_t kh1 = *reinterpret_cast<const uint32_t*>(Kh + krow * Cfg::QK_STRIDE + c0 + 8);
        const uint32_t kl0 = *reinterpret_cast<const uint32_t*>(Kl + krow * Cfg::QK_STRIDE + c0);
        const uint32_t kl1 = *reinterpret_cast<const uint32_t*>(Kl + krow * Cfg::QK_STRIDE + c0 + 8);
        mma_f16(s[nt], ql[ks], kh0, kh1);
        mma_f16(s[nt], qh[ks], kl0, kl1);
        mma_f16(s[nt], qh[ks], kh0, kh1);
      }
    }
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < Cfg::KB; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = (kb * Cfg::KB + nt) * 8 + 2 * t + e;
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
    const float sc0 = expf(m0 - nm0), sc1 = expf(m1 - nm1);
    m0 = nm0;
    m1 = nm1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < Cfg::KB; ++nt) {
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
    // O += P V : one 16-key MMA step = two adjacent key tiles of S
#pragma unroll
    for (int kk = 0; kk < Cfg::KB / 2; ++kk) {
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int nt = 2 * kk + half;
        __half h0, h1, h2, h3, x0, x1, x2, x3;
        split_h(s[nt][0], h0, x0);
        split_h(s[nt][1], h1, x1);
        split_h(s[nt][2], h2, x2);
        split_h(s[nt][3], h3, x3);
        ph[2 * half] = pack_h2(h0, h1);       // a0 / a2: row g
        ph[2 * half + 1] = pack_h2(h2, h3);   // a1 / a3: row g+8
        pl[2 * half] = pack_h2(x0, x1);
        pl[2 * half + 1] = pack_h2(x2, x3);
      }
      const int key0 = (kb * Cfg::KB + 2 * kk) * 8 + 2 * t;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        const int d = nb * 8 + g;
        const uint32_t vh0 = *reinterpret_cast<const uint32_t*>(Vth + d * Cfg::VT_STRIDE + key0);
        const uint32_t vh1 = *reinterpret_cast<const uint32_t*>(Vth + d * Cfg::VT_STRIDE + key0 + 8);
        const uint32_t vl0 = *reinterpret_cast<const uint32_t*>(Vtl + d * Cfg::VT_STRIDE + key0);
        const uint32_t vl1 = *reinterpret_cast<const uint32_t*>(Vtl + d * Cfg::VT_STRIDE + key0 + 8);
        mma_f16(o[nb], pl, vh0, vh1);
        mma_f16(o[nb], ph, vl0, vl1);
        mma_f16(o[nb], ph, vh0, vh1);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;

  const int src0 = r0 < N ? sSrc[r0] : -1, src1 = r1 < N ? sSrc[r1] : -1;
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) {
    const int c = head * 32 + nb * 8 + 2 * t;
    const float v00 = o[nb][0] * i0, v01 = o[nb][1] * i0, v10 = o[nb][2] * i1, v11 = o[nb][3] * i1;
    if (out16 == nullptr) {
      if (src0 >= 0) *reinterpret_cast<float2*>(out + (size_t)src0 * C + c) = make_float2(v00, v01);
      if (src1 >= 0) *reinterpret_cast<float2*>(out + (size_t)src1 * C + c) = make_float2(v10, v11);
    } else {
      // fp16x3 GEMM operand layout (single K-chunk, C <= 1536): [lo*2^11 (C) | hi*2^-11 (C) | hi (C)]
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int src = r ? src1 : src0;
        if (src < 0) continue;
        const float a = r ? v10 : v00, bb = r ? v11 : v01;
        const __half ha = __float2half_rn(a), hb = __float2half_rn(bb);
        const float fa = __half2float(ha), fb = __half2float(hb);
        __half* row = out16 + (size_t)src * (3 * (size_t)C);
        *reinterpret_cast<uint32_t*>(row + c) = pack_h2(__float2half_rn((a - fa) * 2048.f), __float2half_rn((bb - fb) * 2048.f));
        *reinterpret_cast<uint32_t*>(row + C + c) = pack_h2(__float2half_rn(fa * (1.f / 2048.f)), __float2half_rn(fb * (1.f / 2048.f)));
        *reinterpret_cast<uint32_t*>(row + 2 * C + c) = pack_h2(ha, hb);
      }
    }
  }
}

template <int WS>
static int launch_window(cudaStream_t st, const float* qkv, const float* bias, const float* table, int B, int H,
                         int W, int C, int nH, int shift, int precision, float* out, __half* out16) {
  using Cfg = WinCfg<WS>;
  const int Hp = (H + WS - 1) / WS * WS, Wp = (W + WS - 1) / WS * WS;
  const long long wins = (long long)B * (Hp / WS) * (Wp / WS);
  UNIVS_REQUIRE(wins < (1ll << 31) && nH <= 65535, "swin_window_attention: grid too large");
  dim3 grid((unsigned)wins, (unsigned)nH);
  const float scale = 0.17677669529663687f;  // 32^-0.5
  cudaError_t e;
  if (precision == UNIVS_PREC_TF32X3) {
    using Cfg16 = WinCfg16<WS>;
    e = cudaFuncSetAttribute(swin_window_attn_f16x3_kernel<WS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Cfg16::SMEM);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNIVS_E_LAUNCH; }
    swin_window_attn_f16x3_kernel<WS><<<grid, Cfg16::THREADS, Cfg16::SMEM, st>>>(qkv, bias, table, B, H, W, C, nH,
                                                                                 shift, scale, out, out16);
  } else {
    UNIVS_REQUIRE(out16 == nullptr, "swin_window_attention: the fp16x3 operand output needs UNIVS_PREC_TF32X3");
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

static int window_dispatch(void* stream, const float* qkv, const float* qkv_bias, const float* rel_bias_table, int batch,
                           int height, int width, int channels, int num_heads, int window, int shift, int precision,
                           float* out, __half* out16);

extern "C" int univs_swin_window_attention_f32(void* stream, const float* qkv, const float* qkv_bias,
                                               const float* rel_bias_table, int batch, int height, int width,
                                               int channels, int num_heads, int window, int shift, int precision,
                                               float* out) {
  UNIVS_REQUIRE(out, "swin_window_attention: null pointer");
  return window_dispatch(stream, qkv, qkv_bias, rel_bias_table, batch, height, width, channels, num_heads, window, shift,
                         precision, out, nullptr);
}

extern "C" int univs_swin_window_attention_f16x3out(void* stream, const float* qkv, const float* qkv_bias,
                                                    const float* rel_bias_table, int batch, int height, int width,
                                                    int channels, int num_heads, int window, int shift, void* out16) {
  UNIVS_REQUIRE(out16, "swin_window_attention: null pointer");
  UNIVS_REQUIRE(channels <= 1536, "swin_window_attention_f16x3out: single K-chunk layout needs channels <= 1536");
  return window_dispatch(stream, qkv, qkv_bias, rel_bias_table, batch, height, width, channels, num_heads, window, shift,
                         UNIVS_PREC_TF32X3, nullptr, reinterpret_cast<__half*>(out16));
}

static int window_dispatch(void* stream, const float* qkv, const float* qkv_bias, const float* rel_bias_table, int batch,
                           int height, int width, int channels, int num_heads, int window, int shift, int precision,
                           float* out, __half* out16) {
  UNIVS_REQUIRE(qkv && qkv_bias && rel_bias_table, "swin_window_attention: null pointer");
  UNIVS_REQUIRE(batch >= 0 && height > 0 && width > 0, "swin_window_attention: bad sizes");
  UNIVS_REQUIRE(num_heads > 0 && channels == num_heads * 32,
                "swin_window_attention: head_dim must be 32 (channels=%d heads=%d)", channels, num_heads);
  UNIVS_REQUIRE(shift >= 0 && shift < window, "swin_window_attention: shift must be in [0, window)");
  UNIVS_REQUIRE(precision == UNIVS_PREC_TF32X3 || precision == UNIVS_PREC_TF32, "swin_window_attention: bad precision");
  if (batch == 0) return UNIVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  switch (window) {
    case 4: return launch_window<4>(st, qkv, qkv_bias, rel_bias_table, batch, height, width, channels, num_heads, shift, precision, out, out16);
    case 7: return launch_window<7>(st, qkv, qkv_bias, rel_bias_table, batch, height, width, channels, num_heads, shift, precision, out, out16);
    case 12: return launch_window<12>(st, qkv, qkv_bias, rel_bias_table, batch, height, width, channels, num_heads, shift, precision, out, out16);
    default:
      set_error("swin_window_attention: window %d not instantiated (4, 7, 12)", window);
      return UNIVS_E_BADARG;
  }
}
