// Decoder attention cores for sm_100a (head_dim 32, hidden = heads*32):
//   * univs_mha_forward_f32  -- flash-style masked attention used for the per-frame masked cross-attention
//     (transformer_layers.py:95-115 called at ..._univs.py:399-405) and the Q*T spatio-temporal self-attention
//     (transformer_layers.py:34-44 at :408-416).  The boolean mask is consumed bit-packed and shared by all heads
//     (the reference materialises a [T*8, Q, S] bool + a float copy, ..._univs.py:564); the "fully blocked row is
//     unblocked" rule (:390) is a per-row flag.  Keys are split across CTAs (split-K) so that even Q=200 fills
//     148 SMs; partial (m, l, o) are merged by a second tiny kernel.
//   * univs_proca_forward_f32 -- ProCA (..._univs.py:456-496): q-len 1, kv = [own token ; L memory tokens];
//     one warp per (prompt, frame, head), 128-bit K/V streaming, warp-shuffle reductions; the T-invariant
//     prompt memory is read with a frame stride of 0 instead of being repeated T times (:819-820).
//   * univs_attn_mask_bits_f32 -- (..._univs.py:555-566) bilinear-downsample + sigmoid<0.5 threshold of the
//     mask logits straight to bits + the row flag.
#include <cuda_fp16.h>

#include "common.cuh"

namespace univs {

constexpr int kKeyBlk = 64;
constexpr int kStride = 36;
constexpr int kRowsPerCta = 128;  // 8 warps x 16 query rows
constexpr int kMhaThreads = 256;

__device__ __forceinline__ float safe_exp_diff(float a, float b) {  // exp(a-b) with (-inf)-(-inf) -> 0
  return (a == -INFINITY) ? 0.f : expf(a - b);
}

template <bool X3>
__global__ void __launch_bounds__(kMhaThreads)
mha_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
               const uint32_t* __restrict__ mask_bits, const int32_t* __restrict__ row_open, int mask_batch, int B,
               int Lq, int Lk, int C, int heads, int nsplit, int blocks_per_split, float* __restrict__ out,
               float* __restrict__ part_o, float* __restrict__ part_ml) {
  __shared__ __align__(16) float Ks[2][kKeyBlk * kStride];
  __shared__ __align__(16) float Vs[2][kKeyBlk * kStride];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y, b = bh / heads, h = bh - b * heads;
  const int split = blockIdx.z;
  const int row_base = blockIdx.x * kRowsPerCta + warp * 16;
  const int r0 = row_base + g, r1 = r0 + 8;
  const int nkb = (Lk + kKeyBlk - 1) / kKeyBlk;
  const int kb_begin = split * blocks_per_split;
  const int kb_end = min(nkb, kb_begin + blocks_per_split);
  const float scale = 0.17677669529663687f;

  const float* kbase = k + (size_t)b * Lk * C + h * 32;
  const float* vbase = v + (size_t)b * Lk * C + h * 32;

  auto stage_block = [&](int kb, int buf) {
    // 64 rows x 8 16-byte chunks, for K and V: 1024 chunks / 256 threads
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * kMhaThreads;  // 0..1023
      const int mat = idx >> 9;                // 0: K, 1: V
      const int rr = (idx & 511) >> 3, ch = idx & 7;
      const int key = kb * kKeyBlk + rr;
      float* dst = (mat ? Vs[buf] : Ks[buf]) + rr * kStride + ch * 4;
      if (key < Lk) {
        cp_async16(dst, (mat ? vbase : kbase) + (size_t)key * C + ch * 4);
      } else {
        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    cp_async_commit();
  };

  // Q fragments straight from global (one-time, 16 scalars per thread)
  uint32_t qb[4][4], qs[4][4];
  {
    const float* q0 = q + ((size_t)b * Lq + min(r0, Lq - 1)) * C + h * 32;
    const float* q1 = q + ((size_t)b * Lq + min(r1, Lq - 1)) * C + h * 32;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const float a0 = (r0 < Lq ? __ldg(q0 + ks * 8 + t) : 0.f) * scale;
      const float a1 = (r1 < Lq ? __ldg(q1 + ks * 8 + t) : 0.f) * scale;
      const float a2 = (r0 < Lq ? __ldg(q0 + ks * 8 + t + 4) : 0.f) * scale;
      const float a3 = (r1 < Lq ? __ldg(q1 + ks * 8 + t + 4) : 0.f) * scale;
      if (X3) {
        split_tf32(a0, qb[ks][0], qs[ks][0]);
        split_tf32(a1, qb[ks][1], qs[ks][1]);
        split_tf32(a2, qb[ks][2], qs[ks][2]);
        split_tf32(a3, qb[ks][3], qs[ks][3]);
      } else {
        qb[ks][0] = f2tf32(a0); qb[ks][1] = f2tf32(a1); qb[ks][2] = f2tf32(a2); qb[ks][3] = f2tf32(a3);
      }
    }
  }
  // mask rows
  const int words = (Lk + 31) >> 5;
  const int mb = (mask_batch == 1) ? 0 : b;
  const uint32_t* mrow0 = nullptr;
  const uint32_t* mrow1 = nullptr;
  if (mask_bits != nullptr) {
    bool use0 = r0 < Lq, use1 = r1 < Lq;
    if (row_open != nullptr) {
      if (use0) use0 = __ldg(row_open + (size_t)mb * Lq + r0) != 0;
      if (use1) use1 = __ldg(row_open + (size_t)mb * Lq + r1) != 0;
    }
    if (use0) mrow0 = mask_bits + ((size_t)mb * Lq + r0) * words;
    if (use1) mrow1 = mask_bits + ((size_t)mb * Lq + r1) * words;
  }

  float o[4][4];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  if (kb_begin < kb_end) stage_block(kb_begin, 0);
  for (int kb = kb_begin; kb < kb_end; ++kb) {
    const int buf = (kb - kb_begin) & 1;
    if (kb + 1 < kb_end) {
      stage_block(kb + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* Kb = Ks[buf];
    const float* Vb = Vs[buf];

    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const float k0 = Kb[(nt * 8 + g) * kStride + ks * 8 + t], k1 = Kb[(nt * 8 + g) * kStride + ks * 8 + t + 4];
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
    // mask (bit set = blocked) + key padding
    uint32_t w00 = 0, w01 = 0, w10 = 0, w11 = 0;
    const int wbase = kb * 2;
    if (mrow0) {
      w00 = __ldg(mrow0 + wbase);
      if (wbase + 1 < words) w01 = __ldg(mrow0 + wbase + 1);
    }
    if (mrow1) {
      w10 = __ldg(mrow1 + wbase);
      if (wbase + 1 < words) w11 = __ldg(mrow1 + wbase + 1);
    }
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int jj = nt * 8 + 2 * t + e;  // 0..63 in block
        const int key = kb * kKeyBlk + jj;
        const uint32_t bit0 = ((jj < 32 ? w00 : w01) >> (jj & 31)) & 1u;
        const uint32_t bit1 = ((jj < 32 ? w10 : w11) >> (jj & 31)) & 1u;
        if (key >= Lk || bit0) s[nt][e] = -INFINITY;
        if (key >= Lk || bit1) s[nt][2 + e] = -INFINITY;
        bm0 = fmaxf(bm0, s[nt][e]);
        bm1 = fmaxf(bm1, s[nt][2 + e]);
      }
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);
    const float e0 = (nm0 == -INFINITY) ? 0.f : nm0, e1 = (nm1 == -INFINITY) ? 0.f : nm1;
    const float sc0 = safe_exp_diff(m0, e0), sc1 = safe_exp_diff(m1, e1);
    m0 = nm0;
    m1 = nm1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = expf(s[nt][0] - e0);
      s[nt][1] = expf(s[nt][1] - e0);
      s[nt][2] = expf(s[nt][2] - e1);
      s[nt][3] = expf(s[nt][3] - e1);
      ps0 += s[nt][0] + s[nt][1];
      ps1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * sc0 + ps0;
    l1 = l1 * sc1 + ps1;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      o[nb][0] *= sc0; o[nb][1] *= sc0; o[nb][2] *= sc1; o[nb][3] *= sc1;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      uint32_t pb[4], ps[4];
      if (X3) {
        split_tf32(s[nt][0], pb[0], ps[0]);
        split_tf32(s[nt][2], pb[1], ps[1]);
        split_tf32(s[nt][1], pb[2], ps[2]);
        split_tf32(s[nt][3], pb[3], ps[3]);
      } else {
        pb[0] = f2tf32(s[nt][0]); pb[1] = f2tf32(s[nt][2]); pb[2] = f2tf32(s[nt][1]); pb[3] = f2tf32(s[nt][3]);
      }
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        const float v0 = Vb[(nt * 8 + 2 * t) * kStride + nb * 8 + g];
        const float v1 = Vb[(nt * 8 + 2 * t + 1) * kStride + nb * 8 + g];
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
    __syncthreads();  // buffer `buf` is refilled two iterations later
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

  if (nsplit == 1) {
    const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      const int c = h * 32 + nb * 8 + 2 * t;
      if (r0 < Lq) *reinterpret_cast<float2*>(out + ((size_t)b * Lq + r0) * C + c) = make_float2(o[nb][0] * i0, o[nb][1] * i0);
      if (r1 < Lq) *reinterpret_cast<float2*>(out + ((size_t)b * Lq + r1) * C + c) = make_float2(o[nb][2] * i1, o[nb][3] * i1);
    }
  } else {
    // partial layout: part_o [split][B*heads][Lq][32], part_ml [split][B*heads][Lq][2]
    const size_t rows = (size_t)B * heads * Lq;
    float* po = part_o + ((size_t)split * rows + (size_t)bh * Lq) * 32;
    float* pm = part_ml + ((size_t)split * rows + (size_t)bh * Lq) * 2;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      const int c = nb * 8 + 2 * t;
      if (r0 < Lq) *reinterpret_cast<float2*>(po + (size_t)r0 * 32 + c) = make_float2(o[nb][0], o[nb][1]);
      if (r1 < Lq) *reinterpret_cast<float2*>(po + (size_t)r1 * 32 + c) = make_float2(o[nb][2], o[nb][3]);
    }
    if (t == 0) {
      if (r0 < Lq) *reinterpret_cast<float2*>(pm + (size_t)r0 * 2) = make_float2(m0, l0);
      if (r1 < Lq) *reinterpret_cast<float2*>(pm + (size_t)r1 * 2) = make_float2(m1, l1);
    }
  }
}

// ---- strict-precision variant: fp16 hi|lo split operands, m16n8k16 MMAs (see swin_window_attn.cu) ----------------
__device__ __forceinline__ void mha_mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mha_split_h(float x, __half& h, __half& l) {
  h = __float2half_rn(x);
  l = __float2half_rn(x - __half2float(h));
}
__device__ __forceinline__ uint32_t mha_pack(__half a, __half b) {
  __half2 v = __halves2half2(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

constexpr int kKS16 = 40;            // halfs per K row (32 + 8)
constexpr int kVS16 = kKeyBlk + 8;   // halfs per V^T row (64 keys + 8)

__global__ void __launch_bounds__(kMhaThreads)
mha_fwd_f16x3_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                     const uint32_t* __restrict__ mask_bits, const int32_t* __restrict__ row_open, int mask_batch, int B,
                     int Lq, int Lk, int C, int heads, int nsplit, int blocks_per_split, float* __restrict__ out,
                     float* __restrict__ part_o, float* __restrict__ part_ml) {
  __shared__ __align__(16) __half Kh[kKeyBlk * kKS16], Kl[kKeyBlk * kKS16];
  __shared__ __align__(16) __half Vth[32 * kVS16], Vtl[32 * kVS16];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y, b = bh / heads, h = bh - b * heads;
  const int split = blockIdx.z;
  const int r0 = blockIdx.x * kRowsPerCta + warp * 16 + g, r1 = r0 + 8;
  const int nkb = (Lk + kKeyBlk - 1) / kKeyBlk;
  const int kb_begin = split * blocks_per_split;
  const int kb_end = min(nkb, kb_begin + blocks_per_split);
  const float scale = 0.17677669529663687f;
  const float* kbase = k + (size_t)b * Lk * C + h * 32;
  const float* vbase = v + (size_t)b * Lk * C + h * 32;

  // register prefetch of one 64-key block: thread -> (row = tid/8 + 32*it, 4 dims = (tid%8)*4), K and V
  float4 pk[2], pv[2];
  auto fetch = [&](int kb) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int key = kb * kKeyBlk + (tid >> 3) + 32 * it;
      if (key < Lk) {
        pk[it] = ldg_f4(kbase + (size_t)key * C + (tid & 7) * 4);
        pv[it] = ldg_f4(vbase + (size_t)key * C + (tid & 7) * 4);
      } else {
        pk[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        pv[it] = pk[it];
      }
    }
  };
  auto stash = [&]() {   // split to fp16 hi|lo; K row-major, V transposed
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int rr = (tid >> 3) + 32 * it, c4 = (tid & 7) * 4;
      const float kk[4] = {pk[it].x, pk[it].y, pk[it].z, pk[it].w};
      const float vv[4] = {pv[it].x, pv[it].y, pv[it].z, pv[it].w};
      __half hh[4], ll[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) mha_split_h(kk[e], hh[e], ll[e]);
      *reinterpret_cast<uint2*>(Kh + rr * kKS16 + c4) = make_uint2(mha_pack(hh[0], hh[1]), mha_pack(hh[2], hh[3]));
      *reinterpret_cast<uint2*>(Kl + rr * kKS16 + c4) = make_uint2(mha_pack(ll[0], ll[1]), mha_pack(ll[2], ll[3]));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        mha_split_h(vv[e], hh[e], ll[e]);
        Vth[(c4 + e) * kVS16 + rr] = hh[e];
        Vtl[(c4 + e) * kVS16 + rr] = ll[e];
      }
    }
  };

  // Q fragments (scaled, split) straight from global
  uint32_t qh[2][4], ql[2][4];
  {
    const float* q0 = q + ((size_t)b * Lq + min(r0, Lq - 1)) * C + h * 32;
    const float* q1 = q + ((size_t)b * Lq + min(r1, Lq - 1)) * C + h * 32;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        const int c0 = ks * 16 + 2 * t + 8 * part;
        const float2 x0 = r0 < Lq ? __ldg(reinterpret_cast<const float2*>(q0 + c0)) : make_float2(0.f, 0.f);
        const float2 x1 = r1 < Lq ? __ldg(reinterpret_cast<const float2*>(q1 + c0)) : make_float2(0.f, 0.f);
        __half a, bb, la, lb;
        mha_split_h(x0.x * scale, a, la);
        mha_split_h(x0.y * scale, bb, lb);
        qh[ks][2 * part] = mha_pack(a, bb);
        ql[ks][2 * part] = mha_pack(la, lb);
        mha_split_h(x1.x * scale, a, la);
        mha_split_h(x1.y * scale, bb, lb);
        qh[ks][2 * part + 1] = mha_pack(a, bb);
        ql[ks][2 * part + 1] = mha_pack(la, lb);
      }
    }
  }
  const int words = (Lk + 31) >> 5;
  const int mb = (mask_batch == 1) ? 0 : b;
  const uint32_t* mrow0 = nullptr;
  const uint32_t* mrow1 = nullptr;
  if (mask_bits != nullptr) {
    bool use0 = r0 < Lq, use1 = r1 < Lq;
    if (row_open != nullptr) {
      if (use0) use0 = __ldg(row_open + (size_t)mb * Lq + r0) != 0;
      if (use1) use1 = __ldg(row_open + (size_t)mb * Lq + r1) != 0;
    }
    if (use0) mrow0 = mask_bits + ((size_t)mb * Lq + r0) * words;
    if (use1) mrow1 = mask_bits + ((size_t)mb * Lq + r1) * words;
  }

  float o[4][4];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  if (kb_begin < kb_end) fetch(kb_begin);
  for (int kb = kb_begin; kb < kb_end; ++kb) {
    __syncthreads();            // previous block's fragments are no longer read
    stash();
    __syncthreads();
    if (kb + 1 < kb_end) fetch(kb + 1);   // global loads of the next block fly during the MMAs below

    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const int krow = nt * 8 + g;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int c0 = ks * 16 + 2 * t;
        const uint32_t kh0 = *reinterpret_cast<const uint32_t*>(Kh + krow * kKS16 + c0);
        const uint32_t kh1 = *reinterpret_cast<const uint32_t*>(Kh + krow * kKS16 + c0 + 8);
        const uint32_t kl0 = *reinterpret_cast<const uint32_t*>(Kl + krow * kKS16 + c0);
        const uint32_t kl1 = *reinterpret_cast<const uint32_t*>(Kl + krow * kKS16 + c0 + 8);
        mha_mma_f16(s[nt], ql[ks], kh0, kh1);
        mha_mma_f16(s[nt], qh[ks], kl0, kl1);
        mha_mma_f16(s[nt], qh[ks], kh0, kh1);
      }
    }
    uint32_t w00 = 0, w01 = 0, w10 = 0, w11 = 0;
    const int wbase = kb * 2;
    if (mrow0) {
      w00 = __ldg(mrow0 + wbase);
      if (wbase + 1 < words) w01 = __ldg(mrow0 + wbase + 1);
    }
    if (mrow1) {
      w10 = __ldg(mrow1 + wbase);
      if (wbase + 1 < words) w11 = __ldg(mrow1 + wbase + 1);
    }
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int jj = nt * 8 + 2 * t + e;
        const int key = kb * kKeyBlk + jj;
        const uint32_t bit0 = ((jj < 32 ? w00 : w01) >> (jj & 31)) & 1u;
        const uint32_t bit1 = ((jj < 32 ? w10 : w11) >> (jj & 31)) & 1u;
        if (key >= Lk || bit0) s[nt][e] = -INFINITY;
        if (key >= Lk || bit1) s[nt][2 + e] = -INFINITY;
        bm0 = fmaxf(bm0, s[nt][e]);
        bm1 = fmaxf(bm1, s[nt][2 + e]);
      }
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);
    const float e0 = (nm0 == -INFINITY) ? 0.f : nm0, e1 = (nm1 == -INFINITY) ? 0.f : nm1;
    const float sc0 = safe_exp_diff(m0, e0), sc1 = safe_exp_diff(m1, e1);
    m0 = nm0;
    m1 = nm1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = expf(s[nt][0] - e0);
      s[nt][1] = expf(s[nt][1] - e0);
      s[nt][2] = expf(s[nt][2] - e1);
      s[nt][3] = expf(s[nt][3] - e1);
      ps0 += s[nt][0] + s[nt][1];
      ps1 += s[nt][2] + s[nt][3];
    }
    l0 = l0 * sc0 + ps0;
    l1 = l1 * sc1 + ps1;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      o[nb][0] *= sc0; o[nb][1] *= sc0; o[nb][2] *= sc1; o[nb][3] *= sc1;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int nt = 2 * kk + half;
        __half h0, h1, h2, h3, x0, x1, x2, x3;
        mha_split_h(s[nt][0], h0, x0);
        mha_split_h(s[nt][1], h1, x1);
        mha_split_h(s[nt][2], h2, x2);
        mha_split_h(s[nt][3], h3, x3);
        ph[2 * half] = mha_pack(h0, h1);
        ph[2 * half + 1] = mha_pack(h2, h3);
        pl[2 * half] = mha_pack(x0, x1);
        pl[2 * half + 1] = mha_pack(x2, x3);
      }
      const int key0 = kk * 16 + 2 * t;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        const int d = nb * 8 + g;
        const uint32_t vh0 = *reinterpret_cast<const uint32_t*>(Vth + d * kVS16 + key0);
        const uint32_t vh1 = *reinterpret_cast<const uint32_t*>(Vth + d * kVS16 + key0 + 8);
        const uint32_t vl0 = *reinterpret_cast<const uint32_t*>(Vtl + d * kVS16 + key0);
        const uint32_t vl1 = *reinterpret_cast<const uint32_t*>(Vtl + d * kVS16 + key0 + 8);
        mha_mma_f16(o[nb], pl, vh0, vh1);
        mha_mma_f16(o[nb], ph, vl0, vl1);
        mha_mma_f16(o[nb], ph, vh0, vh1);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

  if (nsplit == 1) {
    const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      const int c = h * 32 + nb * 8 + 2 * t;
      if (r0 < Lq) *reinterpret_cast<float2*>(out + ((size_t)b * Lq + r0) * C + c) = make_float2(o[nb][0] * i0, o[nb][1] * i0);
      if (r1 < Lq) *reinterpret_cast<float2*>(out + ((size_t)b * Lq + r1) * C + c) = make_float2(o[nb][2] * i1, o[nb][3] * i1);
    }
  } else {
    const size_t rows = (size_t)B * heads * Lq;
    float* po = part_o + ((size_t)split * rows + (size_t)bh * Lq) * 32;
    float* pm = part_ml + ((size_t)split * rows + (size_t)bh * Lq) * 2;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      const int c = nb * 8 + 2 * t;
      if (r0 < Lq) *reinterpret_cast<float2*>(po + (size_t)r0 * 32 + c) = make_float2(o[nb][0], o[nb][1]);
      if (r1 < Lq) *reinterpret_cast<float2*>(po + (size_t)r1 * 32 + c) = make_float2(o[nb][2], o[nb][3]);
    }
    if (t == 0) {
      if (r0 < Lq) *reinterpret_cast<float2*>(pm + (size_t)r0 * 2) = make_float2(m0, l0);
      if (r1 < Lq) *reinterpret_cast<float2*>(pm + (size_t)r1 * 2) = make_float2(m1, l1);
    }
  }
}

// merge split-K partials: one warp per (b, h, q) row, lane = channel
__global__ void __launch_bounds__(256)
mha_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml, int B, int heads, int Lq,
                   int C, int nsplit, float* __restrict__ out) {
  const size_t rows = (size_t)B * heads * Lq;
  const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, __ldg(part_ml + ((size_t)s * rows + row) * 2));
  float acc = 0.f, l = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float ms = __ldg(part_ml + ((size_t)s * rows + row) * 2);
    const float ls = __ldg(part_ml + ((size_t)s * rows + row) * 2 + 1);
    const float w = (ms == -INFINITY) ? 0.f : expf(ms - M);
    acc = fmaf(w, __ldg(part_o + ((size_t)s * rows + row) * 32 + lane), acc);
    l = fmaf(w, ls, l);
  }
  const int qi = (int)(row % Lq);
  const size_t bh = row / Lq;
  const int h = (int)(bh % heads);
  const size_t b = bh / heads;
  out[(b * Lq + qi) * C + h * 32 + lane] = acc / l;
}

// ---- ProCA ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
proca_kernel(const float* __restrict__ q, const float* __restrict__ k_self, const float* __restrict__ v_self,
             const float* __restrict__ k_mem, const float* __restrict__ v_mem, int P, int T, int Tm, int L, int C,
             int heads, float* __restrict__ out) {
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= (long long)P * T * heads) return;
  const int lane = threadIdx.x & 31, grp = lane >> 3, l8 = lane & 7;
  const int h = (int)(wid % heads);
  const long long pt = wid / heads;
  const int tt = (int)(pt % T);
  const int p = (int)(pt / T);
  const float scale = 0.17677669529663687f;
  const size_t tok = ((size_t)p * T + tt) * C + h * 32 + l8 * 4;
  float4 qv = ldg_f4(q + tok);
  qv.x *= scale; qv.y *= scale; qv.z *= scale; qv.w *= scale;
  const size_t mem_base = ((size_t)p * Tm + (Tm == 1 ? 0 : tt)) * (size_t)L * C + h * 32 + l8 * 4;
  float m = -INFINITY, l = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j0 = 0; j0 < 1 + L; j0 += 4) {
    const int j = j0 + grp;
    float sc = -INFINITY;
    float4 vv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < 1 + L) {
      const float4 kv = (j == 0) ? ldg_f4(k_self + tok) : ldg_f4(k_mem + mem_base + (size_t)(j - 1) * C);
      vv = (j == 0) ? ldg_f4(v_self + tok) : ldg_f4(v_mem + mem_base + (size_t)(j - 1) * C);
      sc = qv.x * kv.x + qv.y * kv.y + qv.z * kv.z + qv.w * kv.w;
    }
    // 8-lane dot-product reduction (all lanes participate; inactive groups carry -inf)
    float d = (j < 1 + L) ? sc : 0.f;
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    if (j < 1 + L) {
      const float nm = fmaxf(m, d);
      const float sco = safe_exp_diff(m, nm);
      const float pw = expf(d - nm);
      l = l * sco + pw;
      acc.x = acc.x * sco + pw * vv.x;
      acc.y = acc.y * sco + pw * vv.y;
      acc.z = acc.z * sco + pw * vv.z;
      acc.w = acc.w * sco + pw * vv.w;
      m = nm;
    }
  }
  // merge the four key groups (lanes l8, l8+8, l8+16, l8+24)
#pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, off);
    const float ol = __shfl_xor_sync(0xffffffffu, l, off);
    const float ox = __shfl_xor_sync(0xffffffffu, acc.x, off), oy = __shfl_xor_sync(0xffffffffu, acc.y, off);
    const float oz = __shfl_xor_sync(0xffffffffu, acc.z, off), ow = __shfl_xor_sync(0xffffffffu, acc.w, off);
    const float nm = fmaxf(m, om);
    const float a = safe_exp_diff(m, nm), bsc = safe_exp_diff(om, nm);
    l = l * a + ol * bsc;
    acc.x = acc.x * a + ox * bsc;
    acc.y = acc.y * a + oy * bsc;
    acc.z = acc.z * a + oz * bsc;
    acc.w = acc.w * a + ow * bsc;
    m = nm;
  }
  if (grp == 0) {
    const float inv = 1.f / l;
    *reinterpret_cast<float4*>(out + tok) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
}

// ---- attention-mask bits -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
attn_mask_bits_kernel(const float* __restrict__ logits, int Q, int T, int H, int W, int h, int w, int ry, int rx,
                      uint32_t* __restrict__ bits, int32_t* __restrict__ row_open) {
  const int S = h * w, words = (S + 31) >> 5;
  const long long total_words = (long long)T * Q * words;
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= total_words) return;
  const int lane = threadIdx.x & 31;
  const int word = (int)(wid % words);
  const long long tq = wid / words;  // t*Q + q
  const int qi = (int)(tq % Q), ti = (int)(tq / Q);
  const int key = word * 32 + lane;
  bool blocked = true;
  if (key < S) {
    const int ky = key / w, kx = key - ky * w;
    const int y0 = ky * ry + (ry >> 1) - 1, x0 = kx * rx + (rx >> 1) - 1;
    const float* p = logits + (((size_t)qi * T + ti) * H + y0) * W + x0;
    const float a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + W), d = __ldg(p + W + 1);
    // upsample_bilinear2d with both lambdas == 0.5 (even integer ratios, align_corners=False)
    const float val = 0.5f * (0.5f * a + 0.5f * b) + 0.5f * (0.5f * c + 0.5f * d);
    const float sg = 1.f / (1.f + expf(-val));
    blocked = sg < 0.5f;
  }
  const uint32_t ballot = __ballot_sync(0xffffffffu, blocked);
  if (lane == 0) {
    bits[wid] = ballot;
    const int valid = min(32, S - word * 32);
    const uint32_t vmask = valid == 32 ? 0xffffffffu : ((1u << valid) - 1u);
    if ((ballot & vmask) != vmask) atomicOr(row_open + tq, 1);
  }
}

__global__ void round_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __uint_as_float(f2tf32(in[i]));
}

static void mha_plan(int B, int heads, int Lq, int Lk, int& qtiles, int& nsplit, int& bps) {
  qtiles = (Lq + kRowsPerCta - 1) / kRowsPerCta;
  const int nkb = (Lk + kKeyBlk - 1) / kKeyBlk;
  const long long base = (long long)qtiles * B * heads;
  int want = (int)((2 * 148 + base - 1) / base);
  if (want < 1) want = 1;
  if (want > nkb) want = nkb;
  if (want > 64) want = 64;
  bps = (nkb + want - 1) / want;
  nsplit = (nkb + bps - 1) / bps;
  if (nsplit < 1) nsplit = 1;
}

// split-K merge for the tensor-core kernel of mha_tc.cu (same partial format)
int launch_mha_combine(cudaStream_t st, const float* part_o, const float* part_ml, int B, int heads, int Lq, int C, int nsplit,
                       float* out) {
  const size_t rows = (size_t)B * heads * Lq;
  mha_combine_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(part_o, part_ml, B, heads, Lq, C, nsplit, out);
  return check_launch("mha_combine");
}

}  // namespace univs

using namespace univs;

extern "C" int64_t univs_mha_workspace_bytes(int batch, int len_q, int len_k, int channels) {
  if (batch <= 0 || len_q <= 0 || len_k <= 0 || channels <= 0) return 0;
  int qt, ns, bps;
  mha_plan(batch, channels / 32, len_q, len_k, qt, ns, bps);
  if (ns == 1) return 16;
  return (int64_t)ns * batch * (channels / 32) * len_q * (32 + 2) * (int64_t)sizeof(float);
}

extern "C" int univs_mha_forward_f32(void* stream, const float* q, const float* k, const float* v,
                                     const uint32_t* mask_bits, const int32_t* row_open, int mask_batch, int batch,
                                     int len_q, int len_k, int channels, int precision, void* workspace, float* out) {
  UNIVS_REQUIRE(q && k && v && out, "mha_forward: null pointer");
  UNIVS_REQUIRE(batch >= 0 && len_q >= 0 && len_k > 0, "mha_forward: bad sizes (len_k must be > 0)");
  UNIVS_REQUIRE(channels > 0 && channels % 32 == 0, "mha_forward: channels must be heads*32");
  UNIVS_REQUIRE(mask_bits == nullptr || mask_batch == 1 || mask_batch == batch, "mha_forward: mask_batch must be 1 or batch");
  UNIVS_REQUIRE(precision == UNIVS_PREC_TF32X3 || precision == UNIVS_PREC_TF32, "mha_forward: bad precision");
  if (batch == 0 || len_q == 0) return UNIVS_OK;
  const int heads = channels / 32;
  int qt, ns, bps;
  mha_plan(batch, heads, len_q, len_k, qt, ns, bps);
  UNIVS_REQUIRE(ns == 1 || workspace != nullptr, "mha_forward: workspace required (split-K = %d)", ns);
  UNIVS_REQUIRE((long long)batch * heads <= 65535, "mha_forward: batch*heads too large");
  float* part_o = reinterpret_cast<float*>(workspace);
  float* part_ml = part_o ? part_o + (size_t)ns * batch * heads * len_q * 32 : nullptr;
  dim3 grid(qt, batch * heads, ns);
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == UNIVS_PREC_TF32X3)   // strict: fp16 hi|lo split kernel (fp32-equivalent products)
    mha_fwd_f16x3_kernel<<<grid, kMhaThreads, 0, st>>>(q, k, v, mask_bits, row_open, mask_batch, batch, len_q, len_k,
                                                       channels, heads, ns, bps, out, part_o, part_ml);
  else
    mha_fwd_kernel<false><<<grid, kMhaThreads, 0, st>>>(q, k, v, mask_bits, row_open, mask_batch, batch, len_q, len_k,
                                                        channels, heads, ns, bps, out, part_o, part_ml);
  int rc = check_launch("mha_forward");
  if (rc || ns == 1) return rc;
  const size_t rows = (size_t)batch * heads * len_q;
  mha_combine_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(part_o, part_ml, batch, heads, len_q, channels, ns, out);
  return check_launch("mha_combine");
}

extern "C" int univs_proca_forward_f32(void* stream, const float* q, const float* k_self, const float* v_self,
                                       const float* k_mem, const float* v_mem, int prompts, int frames,
                                       int mem_frames, int mem_len, int channels, float* out) {
  UNIVS_REQUIRE(q && k_self && v_self && out, "proca_forward: null pointer");
  UNIVS_REQUIRE(mem_len == 0 || (k_mem && v_mem), "proca_forward: null memory pointer");
  UNIVS_REQUIRE(prompts >= 0 && frames >= 0 && mem_len >= 0, "proca_forward: bad sizes");
  UNIVS_REQUIRE(mem_frames == 1 || mem_frames == frames, "proca_forward: mem_frames must be 1 or frames");
  UNIVS_REQUIRE(channels > 0 && channels % 32 == 0, "proca_forward: channels must be heads*32");
  if (prompts == 0 || frames == 0) return UNIVS_OK;
  const int heads = channels / 32;
  const long long warps = (long long)prompts * frames * heads;
  proca_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(q, k_self, v_self, k_mem, v_mem, prompts,
                                                                              frames, mem_frames, mem_len, channels,
                                                                              heads, out);
  return check_launch("proca_forward");
}

extern "C" int univs_attn_mask_bits_f32(void* stream, const float* logits, int queries, int frames, int height,
                                        int width, int tgt_h, int tgt_w, uint32_t* bits, int32_t* row_open) {
  UNIVS_REQUIRE(logits && bits && row_open, "attn_mask_bits: null pointer");
  UNIVS_REQUIRE(queries >= 0 && frames >= 0 && height > 0 && width > 0 && tgt_h > 0 && tgt_w > 0, "attn_mask_bits: bad sizes");
  UNIVS_REQUIRE(height % tgt_h == 0 && width % tgt_w == 0, "attn_mask_bits: target size must divide the logit size");
  const int ry = height / tgt_h, rx = width / tgt_w;
  UNIVS_REQUIRE(ry % 2 == 0 && rx % 2 == 0, "attn_mask_bits: resize ratios must be even (got %d, %d)", ry, rx);
  if (queries == 0 || frames == 0) return UNIVS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(row_open, 0, sizeof(int32_t) * (size_t)queries * frames, st);
  if (e != cudaSuccess) { set_error("attn_mask_bits: memset failed: %s", cudaGetErrorString(e)); return UNIVS_E_LAUNCH; }
  const int words = (tgt_h * tgt_w + 31) / 32;
  const long long total = (long long)frames * queries * words;
  attn_mask_bits_kernel<<<(unsigned)((total + 7) / 8), 256, 0, st>>>(logits, queries, frames, height, width, tgt_h,
                                                                      tgt_w, ry, rx, bits, row_open);
  return check_launch("attn_mask_bits");
}

extern "C" int univs_round_tf32_f32(void* stream, const float* in, float* out, int64_t n) {
  UNIVS_REQUIRE(in && out && n >= 0, "round_tf32: bad arguments");
  if (n == 0) return UNIVS_OK;
  long long gl = (n + 255) / 256; if (gl > 148 * 32) gl = 148 * 32; const int grid = (int)gl;
  round_tf32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, n);
  return check_launch("round_tf32");
}
