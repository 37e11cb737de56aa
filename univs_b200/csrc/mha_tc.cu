// Masked multi-head cross-attention of the decoder on the 5th-gen tensor cores (strict fp16 hi|lo policy), for the
// shape that dominates the decoder: few queries (Lq <= 256: the 200 learnable queries + prompts of one frame) against a
// long per-frame memory (S = 920 / 3680 / 14720 keys at the north-star size), head_dim = 32.
// Reference: nn.MultiheadAttention core inside CrossAttentionLayer (mask2former_video/.../transformer_layers.py:95-115
// through ..._univs.py:386-400): softmax(q k^T / sqrt(32) + mask) v per head, mask = per-frame boolean [Q, S] shared by
// the heads (here: bits, bit set = key blocked; a row whose `row_open` flag is 0 ignores its mask, ..._univs.py:390).
// Same C-ABI contract and split-K partial format as univs_mha_forward_f32 (mha.cu), whose combine kernel merges the
// partials; that mma.sync kernel stays the path for Lq > 256 (the Q*T self-attention) and the cross-check of this one.
//
// One CTA = (batch b, head h, key split z); it walks its key blocks of 128 keys.  Work unit n = (key block, query row
// tile): 128 query rows x 128 keys.  384 threads, warp-specialised like swin_window_attn_tc.cu:
//   warps 9-11 loaders  : Q once (scaled by 32^-0.5, rows >= Lq zero), then K / V blocks (contiguous 128-byte head slices,
//                         22 x 128-bit loads in flight per thread), split fp32 -> fp16 hi + lo, K-major SWIZZLE_64B tiles,
//                         3-stage ring (one stage = one key block, used by both row tiles)
//   warp 8      MMA     : S = Ql Kh^T + Qh Kl^T + Qh Kh^T (M=128, N=128, K=32), O_blk = Pl Vh + Ph Vl + Ph Vh (M=128, N=32,
//                         K=128, V through the MN-major B descriptor)
//   warps 0-7   softmax : warp = (TMEM lane quarter, key half); a thread owns 64 scores of one query row of the unit,
//                         applies mask bits / key bound, exchanges the block max with its partner warp (shared memory +
//                         64-thread named barrier), writes P relative to the NEW running max (so P <= 1) and, one unit
//                         later, folds the block's O (read back with tcgen05.ld) into its running output in registers:
//                         o = o * exp(m_old - m_new) + O_blk.  No correction pass over TMEM is needed.
//   warp 8 also allocates TMEM (256 columns: S [0,128), O_blk [128,160)).
// Pipelining: softmax warps run softmax(n+1) -> merge(n), the MMA lane S(n+1) -> PV(n); consecutive units alternate
// between the two row tiles, so the running state touched by softmax(n+1) and merge(n) is disjoint.
// HBM-bound by design: K and V are read once per (frame, head) = 2 * S * 128 B; everything else is on chip.
#include <stdlib.h>

#include "tc05.cuh"

namespace univs {
namespace mhatc {

using namespace tc;

// 12 warps: warps are allocated in groups of four, so the 14 warps of the first version were budgeted like 16 (128 registers
// per thread) and the softmax warps' running state (two row tiles x {m, l, alpha, o[16]} beside 64 scores) spilled ~20 values
// per unit pair -- with 197 KB of shared memory there is no L1 left, so every spill is an L2 round trip.  384 threads get 168.
constexpr int kThreads = 384;
constexpr int kMmaWarp = 8;
constexpr int kAllocWarp = 8;
constexpr int kLoaderWarp0 = 9;
constexpr int kLoaderThreads = 96;
constexpr int kLoaderRows = kLoaderThreads / 8;                         // rows per pass: 12
constexpr int kKvPasses = (128 + kLoaderRows - 1) / kLoaderRows;        // 11 (the last one covers 8 rows)
constexpr int kBlk = 128;                        // keys per block
constexpr int kStages = 3;
constexpr int kMaxLq = 256;

constexpr int kRow = 64;
constexpr int kTile = 128 * kRow;                // 8192: one 128-row operand tile
constexpr int kOffQh = 0, kOffQl = 2 * kTile;    // two row tiles each
constexpr int kOffKV = 4 * kTile;                // stages: Kh | Kl | Vh | Vl
constexpr int kStageBytes = 4 * kTile;           // 32768
constexpr int kOffPh = kOffKV + kStages * kStageBytes;
constexpr int kOffPl = kOffPh + 4 * kTile;       // P: 4 atoms of 32 keys x 128 rows
constexpr int kOffXch = kOffPl + 4 * kTile;      // max[2 parity][2 half][128] floats + final l [2 tile][128]
constexpr int kOffBars = kOffXch + (2 * 2 * 128 + 2 * 128) * 4;
constexpr int kNumBars = 2 * kStages + 7;
constexpr int kSmemBytes = kOffBars + kNumBars * 8 + 16;
static_assert(kOffBars % 8 == 0 && kSmemBytes <= 227 * 1024, "shared memory budget");

constexpr uint32_t kColS = 0, kColO = 128;

enum Bar { KV_FULL = 0, KV_EMPTY = kStages, Q_FULL = 2 * kStages, S_FULL, S_FREE, P_FULL, P_FREE, O_FULL, O_FREE };

struct Args {
  const float* q;
  const float* k;
  const float* v;
  const uint32_t* mask_bits;
  const int32_t* row_open;
  int mask_batch, B, Lq, Lk, C, heads, nsplit, bps;
  float* out;
  float* part_o;
  float* part_ml;
};

__device__ __forceinline__ void store_row(unsigned char* tile_h, unsigned char* tile_l, int row, int lane8, float4 x) {
  uint32_t h0, h1, l0, l1;
  const uint32_t off = swz_off<kRow>(row, lane8 >> 1) + (uint32_t)(lane8 & 1) * 8u;
  split_h2(x.x, x.y, h0, l0);
  split_h2(x.z, x.w, h1, l1);
  sts_v2(smem_u32(tile_h) + off, h0, h1);
  sts_v2(smem_u32(tile_l) + off, l0, l1);
}

// transposed V (diagnostic variant: K-major B operand instead of the MN-major one): row = dim, 32-key atoms of 32 rows x 64 B
__device__ __forceinline__ void store_row_vt(unsigned char* tile_h, unsigned char* tile_l, int key, int lane8, float4 x) {
  uint32_t h0, h1, l0, l1;
  split_h2(x.x, x.y, h0, l0);
  split_h2(x.z, x.w, h1, l1);
  const uint32_t hv[2] = {h0, h1}, lv[2] = {l0, l1};
  __half* vh = reinterpret_cast<__half*>(tile_h);
  __half* vl = reinterpret_cast<__half*>(tile_l);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int d = lane8 * 4 + e;
    const uint32_t o = (uint32_t)(key >> 5) * 2048u + swz_off<kRow>(d, (key & 31) >> 3) + (uint32_t)(key & 7) * 2u;
    const uint32_t hw = hv[e >> 1], lw = lv[e >> 1];
    vh[o >> 1] = __ushort_as_half((unsigned short)((e & 1) ? (hw >> 16) : (hw & 0xffffu)));
    vl[o >> 1] = __ushort_as_half((unsigned short)((e & 1) ? (lw >> 16) : (lw & 0xffffu)));
  }
}

template <bool VT>
__device__ void loader_loop(unsigned char* smem, uint64_t* bars, const Args& a, int b, int h, int kb_begin, int kb_end) {
  const int lt = threadIdx.x - kLoaderWarp0 * 32;
  const int lane8 = lt & 7, slot = lt >> 3;       // kLoaderRows rows per pass
  const int C = a.C;
  const float scale = 0.17677669529663687f;      // 32^-0.5
  {
    const float* qbase = a.q + (size_t)b * a.Lq * C + h * 32 + lane8 * 4;
#pragma unroll 4
    for (int p = 0; p < (kMaxLq + kLoaderRows - 1) / kLoaderRows; ++p) {
      const int r = p * kLoaderRows + slot;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < a.Lq) {
        x = ldg_f4(qbase + (size_t)r * C);
        x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale;
      }
      if (r < kMaxLq) store_row(smem + kOffQh, smem + kOffQl, r, lane8, x);     // rows 128-255 continue into the second tile
    }
    fence_proxy_async_smem();
    mbar_arrive(&bars[Q_FULL]);
  }
  const float* kbase = a.k + (size_t)b * a.Lk * C + h * 32 + lane8 * 4;
  const float* vbase = a.v + (size_t)b * a.Lk * C + h * 32 + lane8 * 4;
  int it = 0;
  for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
    const int s = it % kStages, use = it / kStages;
    if (use > 0) mbar_wait(&bars[KV_EMPTY + s], (uint32_t)((use - 1) & 1), KV_EMPTY + s);
    unsigned char* st = smem + kOffKV + (size_t)s * kStageBytes;
    float4 kk[kKvPasses], vv[kKvPasses];
#pragma unroll
    for (int p = 0; p < kKvPasses; ++p) {
      const int row = p * kLoaderRows + slot;
      const int key = kb * kBlk + row;
      kk[p] = vv[p] = make_float4(0.f, 0.f, 0.f, 0.f);       // keys >= Lk: masked in the softmax, V must still be finite
      if (row < kBlk && key < a.Lk) {
        kk[p] = ldg_f4(kbase + (size_t)key * C);
        vv[p] = ldg_f4(vbase + (size_t)key * C);
      }
    }
#pragma unroll
    for (int p = 0; p < kKvPasses; ++p) {
      const int row = p * kLoaderRows + slot;
      if (row < kBlk) {
        store_row(st, st + kTile, row, lane8, kk[p]);
        if (VT) store_row_vt(st + 2 * kTile, st + 3 * kTile, row, lane8, vv[p]);
        else store_row(st + 2 * kTile, st + 3 * kTile, row, lane8, vv[p]);
      }
    }
    fence_proxy_async_smem();
    mbar_arrive(&bars[KV_FULL + s]);
  }
}

template <bool VT>
__device__ void mma_loop(unsigned char* smem, uint64_t* bars, uint32_t tmem_base, int nblocks) {
  const uint32_t base = smem_u32(smem);
  const int count = 2 * nblocks;                  // units: (block, row tile)
  constexpr uint32_t idesc_s = make_idesc_f16(128, kBlk, false, false);
  constexpr uint32_t idesc_pv = make_idesc_f16(128, 32, false, !VT);
  mbar_wait(&bars[Q_FULL], 0, Q_FULL);
  for (int n = -1; n < count; ++n) {
    if (n + 1 < count) {          // S of unit n+1
      const int m = n + 1, blk = m >> 1, tile = m & 1, s = blk % kStages;
      if (tile == 0) mbar_wait(&bars[KV_FULL + s], (uint32_t)((blk / kStages) & 1), KV_FULL + s);
      if (m > 0) mbar_wait(&bars[S_FREE], (uint32_t)((m - 1) & 1), S_FREE);
      fence_after();
      if (elect_one()) {
        const uint32_t st = base + kOffKV + (uint32_t)s * kStageBytes;
        const uint64_t ah = make_desc<kRow>(base + kOffQh + (uint32_t)tile * kTile), al = make_desc<kRow>(base + kOffQl + (uint32_t)tile * kTile);
        const uint64_t bh = make_desc<kRow>(st), bl = make_desc<kRow>(st + kTile);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          umma_f16(tmem_base + kColS, al + (uint64_t)(2 * k), bh + (uint64_t)(2 * k), idesc_s, k ? 1u : 0u);
          umma_f16(tmem_base + kColS, ah + (uint64_t)(2 * k), bl + (uint64_t)(2 * k), idesc_s, 1u);
          umma_f16(tmem_base + kColS, ah + (uint64_t)(2 * k), bh + (uint64_t)(2 * k), idesc_s, 1u);
        }
        umma_commit(&bars[S_FULL]);
      }
      __syncwarp();
    }
    if (n < 0) continue;
    {                             // O_blk of unit n
      const int blk = n >> 1, tile = n & 1, s = blk % kStages;
      mbar_wait(&bars[P_FULL], (uint32_t)(n & 1), P_FULL);
      if (n > 0) mbar_wait(&bars[O_FREE], (uint32_t)((n - 1) & 1), O_FREE);
      fence_after();
      if (elect_one()) {
        const uint32_t st = base + kOffKV + (uint32_t)s * kStageBytes;
#pragma unroll
        for (int ks = 0; ks < kBlk / 16; ++ks) {
          const uint32_t aoff = (uint32_t)(ks >> 1) * kTile + (uint32_t)(ks & 1) * 32u;
          const uint32_t boff = VT ? (uint32_t)(ks >> 1) * 2048u + (uint32_t)(ks & 1) * 32u : (uint32_t)ks * 16u * kRow;
          const uint64_t ah = make_desc<kRow>(base + kOffPh + aoff), al = make_desc<kRow>(base + kOffPl + aoff);
          const uint64_t bh = make_desc<kRow>(st + 2 * kTile + boff), bl = make_desc<kRow>(st + 3 * kTile + boff);
          umma_f16(tmem_base + kColO, al, bh, idesc_pv, ks ? 1u : 0u);
          umma_f16(tmem_base + kColO, ah, bl, idesc_pv, 1u);
          umma_f16(tmem_base + kColO, ah, bh, idesc_pv, 1u);
        }
        umma_commit(&bars[O_FULL]);
        umma_commit(&bars[P_FREE]);
        if (tile == 1) umma_commit(&bars[KV_EMPTY + s]);     // both row tiles have consumed this key block
      }
      __syncwarp();
    }
  }
}

// running softmax state of one query row (one row tile) in the registers of its two owner threads
struct Run {
  float m, l, alpha;     // running max, this thread's partial denominator (its key half), rescale of the pending merge
  float o[16];           // dims [16*half, 16*half + 16)
};

struct Thr {
  int quarter, half, lane;
  const uint32_t* mrow[2];   // mask row of this thread's query per tile (nullptr: no mask)
};

template <int TILE>
__device__ __forceinline__ void softmax_unit(unsigned char* smem, uint64_t* bars, uint32_t tmem_base, const Thr& t, Run& run,
                                             int n, int kb, int Lk, int words) {
  constexpr float kLog2e = 1.4426950408889634f;
  const int trow = t.quarter * 32 + t.lane;          // row inside the tile = TMEM lane
  uint32_t sr[64];
  mbar_wait(&bars[S_FULL], (uint32_t)(n & 1), S_FULL);
  fence_after();
  {
    const uint32_t taddr = tmem_base + ((uint32_t)(t.quarter * 32) << 16) + kColS + (uint32_t)t.half * 64u;
    uint32_t* r0 = sr;
    uint32_t* r1 = sr + 32;
    UNIVS_TMEM_LD_X32(taddr, r0);
    UNIVS_TMEM_LD_X32(taddr + 32u, r1);
    tmem_wait_ld();
  }
  fence_before();
  __syncwarp();
  if (t.lane == 0) mbar_arrive(&bars[S_FREE]);

  // mask bits (bit set = blocked) and the key bound of the last block
  const int key0 = kb * kBlk + t.half * 64;
  uint32_t w0 = 0u, w1 = 0u;
  if (t.mrow[TILE] != nullptr) {
    const int wi = key0 >> 5;
    if (wi < words) w0 = __ldg(t.mrow[TILE] + wi);
    if (wi + 1 < words) w1 = __ldg(t.mrow[TILE] + wi + 1);
  }
  const int valid = Lk - key0;                       // columns [0, valid) are real keys
  if (valid < 64) {
    if (valid <= 0) { w0 = 0xffffffffu; w1 = 0xffffffffu; }
    else if (valid < 32) { w0 |= ~((1u << valid) - 1u); w1 = 0xffffffffu; }
    else if (valid == 32) { w1 = 0xffffffffu; }
    else { w1 |= ~((1u << (valid - 32)) - 1u); }
  }
  float sc[64];
  float bm = -INFINITY;
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    const uint32_t bit = ((j < 32 ? w0 : w1) >> (j & 31)) & 1u;
    sc[j] = bit ? -INFINITY : __uint_as_float(sr[j]);
    bm = fmaxf(bm, sc[j]);
  }
  float* xmax = reinterpret_cast<float*>(smem + kOffXch) + (n & 1) * 256;
  xmax[t.half * 128 + trow] = bm;
  named_bar_sync(1 + t.quarter, 64);
  bm = fmaxf(bm, xmax[(t.half ^ 1) * 128 + trow]);
  const float m_new = fmaxf(run.m, bm);
  const float e = (m_new == -INFINITY) ? 0.f : m_new;                 // nothing visible yet: exp(-inf - 0) = 0 everywhere
  run.alpha = (run.m == -INFINITY) ? 0.f : ex2_approx((run.m - e) * kLog2e);
  run.m = m_new;
  const float mneg = -e * kLog2e;

  if (n > 0) mbar_wait(&bars[P_FREE], (uint32_t)((n - 1) & 1), P_FREE);
  const uint32_t ph = smem_u32(smem + kOffPh), pl = smem_u32(smem + kOffPl);
  const uint32_t rowoff = (uint32_t)trow * kRow;
  const uint32_t sw = (uint32_t)(trow >> 1) & 3u;
  float psum = 0.f;
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const float p0 = ex2_approx(fmaf(sc[cc * 8 + 2 * x], kLog2e, mneg));
      const float p1 = ex2_approx(fmaf(sc[cc * 8 + 2 * x + 1], kLog2e, mneg));
      psum += p0 + p1;
      split_h2(p0, p1, hi[x], lo[x]);
    }
    const uint32_t g = (uint32_t)(t.half * 8 + cc);                   // 16-byte chunk (8 keys) along the key axis
    const uint32_t off = (g >> 2) * kTile + rowoff + (((g & 3u) ^ sw) << 4);
    sts_v4(ph + off, hi[0], hi[1], hi[2], hi[3]);
    sts_v4(pl + off, lo[0], lo[1], lo[2], lo[3]);
  }
  run.l = run.l * run.alpha + psum;
  fence_proxy_async_smem();
  mbar_arrive(&bars[P_FULL]);
}

__device__ __forceinline__ void merge_unit(uint64_t* bars, uint32_t tmem_base, const Thr& t, Run& run, int n) {
  mbar_wait(&bars[O_FULL], (uint32_t)(n & 1), O_FULL);
  fence_after();
  uint32_t r[16];
  const uint32_t taddr = tmem_base + ((uint32_t)(t.quarter * 32) << 16) + kColO + (uint32_t)t.half * 16u;
  UNIVS_TMEM_LD_X16(taddr, r);
  tmem_wait_ld();
  fence_before();
  __syncwarp();
  if (t.lane == 0) mbar_arrive(&bars[O_FREE]);
#pragma unroll
  for (int x = 0; x < 16; ++x) run.o[x] = fmaf(run.o[x], run.alpha, __uint_as_float(r[x]));
}

__device__ void softmax_loop(unsigned char* smem, uint64_t* bars, uint32_t tmem_base, const Args& a, int b, int h, int split,
                             int kb_begin, int kb_end) {
  Thr t;
  const int warp = threadIdx.x >> 5;
  t.quarter = warp & 3;
  t.half = warp >> 2;
  t.lane = threadIdx.x & 31;
  const int words = (a.Lk + 31) >> 5;
  const int mb = (a.mask_batch == 1) ? 0 : b;
#pragma unroll
  for (int tile = 0; tile < 2; ++tile) {
    const int row = tile * 128 + t.quarter * 32 + t.lane;
    t.mrow[tile] = nullptr;
    if (a.mask_bits != nullptr && row < a.Lq) {
      bool use = true;
      if (a.row_open != nullptr) use = __ldg(a.row_open + (size_t)mb * a.Lq + row) != 0;   // closed row: attends everywhere
      if (use) t.mrow[tile] = a.mask_bits + ((size_t)mb * a.Lq + row) * words;
    }
  }
  Run run0, run1;
  run0.m = run1.m = -INFINITY;
  run0.l = run1.l = 0.f;
  run0.alpha = run1.alpha = 0.f;
#pragma unroll
  for (int x = 0; x < 16; ++x) run0.o[x] = run1.o[x] = 0.f;

  const int nblocks = kb_end - kb_begin;
  if (nblocks > 0) softmax_unit<0>(smem, bars, tmem_base, t, run0, 0, kb_begin, a.Lk, words);
  for (int i = 0; i < nblocks; ++i) {
    softmax_unit<1>(smem, bars, tmem_base, t, run1, 2 * i + 1, kb_begin + i, a.Lk, words);
    merge_unit(bars, tmem_base, t, run0, 2 * i);
    if (i + 1 < nblocks) softmax_unit<0>(smem, bars, tmem_base, t, run0, 2 * i + 2, kb_begin + i + 1, a.Lk, words);
    merge_unit(bars, tmem_base, t, run1, 2 * i + 1);
  }

  // the two key halves of a row hold partial denominators relative to the same running max
  float* xl = reinterpret_cast<float*>(smem + kOffXch) + 512;
  const int trow = t.quarter * 32 + t.lane;
  if (t.half == 1) {
    xl[trow] = run0.l;
    xl[128 + trow] = run1.l;
  }
  named_bar_sync(1 + t.quarter, 64);
  float l0 = run0.l, l1 = run1.l;
  if (t.half == 0) {
    l0 += xl[trow];
    l1 += xl[128 + trow];
  }
  named_bar_sync(1 + t.quarter, 64);
  if (t.half == 0) {
    xl[trow] = l0;
    xl[128 + trow] = l1;
  }
  named_bar_sync(1 + t.quarter, 64);
  l0 = xl[trow];
  l1 = xl[128 + trow];

  const size_t rows = (size_t)a.B * a.heads * a.Lq;
  const size_t bh = (size_t)b * a.heads + h;
#pragma unroll
  for (int tile = 0; tile < 2; ++tile) {
    const int row = tile * 128 + trow;
    if (row >= a.Lq) continue;
    const Run& run = tile ? run1 : run0;
    const float l = tile ? l1 : l0;
    if (a.nsplit == 1) {
      const float inv = 1.f / l;
      float4* dst = reinterpret_cast<float4*>(a.out + ((size_t)b * a.Lq + row) * a.C + h * 32 + t.half * 16);
#pragma unroll
      for (int x = 0; x < 4; ++x)
        dst[x] = make_float4(run.o[4 * x] * inv, run.o[4 * x + 1] * inv, run.o[4 * x + 2] * inv, run.o[4 * x + 3] * inv);
    } else {
      float4* dst = reinterpret_cast<float4*>(a.part_o + ((size_t)split * rows + bh * a.Lq + row) * 32 + t.half * 16);
#pragma unroll
      for (int x = 0; x < 4; ++x) dst[x] = make_float4(run.o[4 * x], run.o[4 * x + 1], run.o[4 * x + 2], run.o[4 * x + 3]);
      if (t.half == 0)
        *reinterpret_cast<float2*>(a.part_ml + ((size_t)split * rows + bh * a.Lq + row) * 2) = make_float2(run.m, l);
    }
  }
}

template <bool VT>
__global__ void __launch_bounds__(kThreads, 1) mha_tc_kernel(const Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x, b = bh / a.heads, h = bh - b * a.heads;
  const int split = blockIdx.y;
  const int nkb = (a.Lk + kBlk - 1) / kBlk;
  const int kb_begin = split * a.bps;
  const int kb_end = min(nkb, kb_begin + a.bps);

  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars[KV_FULL + s], kLoaderThreads);
      mbar_init(&bars[KV_EMPTY + s], 1);
    }
    mbar_init(&bars[Q_FULL], kLoaderThreads);
    mbar_init(&bars[S_FULL], 1);
    mbar_init(&bars[S_FREE], 8);
    mbar_init(&bars[P_FULL], 256);
    mbar_init(&bars[P_FREE], 1);
    mbar_init(&bars[O_FULL], 1);
    mbar_init(&bars[O_FREE], 8);
    mbar_init_fence();
  }
  if (warp == kAllocWarp) {
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kLoaderWarp0) {
    loader_loop<VT>(smem, bars, a, b, h, kb_begin, kb_end);
  } else if (warp == kMmaWarp) {
    mma_loop<VT>(smem, bars, tmem_base, kb_end - kb_begin);
  } else if (warp < 8) {
    softmax_loop(smem, bars, tmem_base, a, b, h, split, kb_begin, kb_end);
  }
  fence_before();
  __syncthreads();
  if (warp == kAllocWarp) {
    fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static void plan(int B, int heads, int Lk, int& nsplit, int& bps) {
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        num_sms <= 0) {
      cudaGetLastError();
      num_sms = 148;
    }
  }
  const int nkb = (Lk + kBlk - 1) / kBlk;
  const long long pairs = (long long)B * heads;
  int want = (int)(num_sms / pairs);      // one wave: at most one CTA per SM
  static int forced = -1;                 // tuning aid: UNIVS_MHA_TC_SPLITS overrides the number of key splits
  if (forced < 0) {
    const char* e = getenv("UNIVS_MHA_TC_SPLITS");
    forced = e ? atoi(e) : 0;
  }
  if (forced > 0) want = forced;
  if (want < 1) want = 1;
  if (want > nkb) want = nkb;
  bps = (nkb + want - 1) / want;
  nsplit = (nkb + bps - 1) / bps;
}

}  // namespace mhatc

// merge of split-K partials, defined in mha.cu
int launch_mha_combine(cudaStream_t st, const float* part_o, const float* part_ml, int B, int heads, int Lq, int C, int nsplit,
                       float* out);

}  // namespace univs

using namespace univs;

extern "C" int64_t univs_mha_tc_workspace_bytes(int batch, int len_q, int len_k, int channels) {
  if (batch <= 0 || len_q <= 0 || len_k <= 0 || channels <= 0) return 0;
  int ns, bps;
  mhatc::plan(batch, channels / 32, len_k, ns, bps);
  if (ns == 1) return 16;
  return (int64_t)ns * batch * (channels / 32) * len_q * (32 + 2) * (int64_t)sizeof(float);
}

template <bool VT>
static int launch_mha_tc(cudaStream_t st, const mhatc::Args& a) {
  cudaError_t e = cudaFuncSetAttribute(mhatc::mha_tc_kernel<VT>, cudaFuncAttributeMaxDynamicSharedMemorySize, mhatc::kSmemBytes);
  if (e != cudaSuccess) {
    set_error("mha_tc_forward: cudaFuncSetAttribute(%d): %s", mhatc::kSmemBytes, cudaGetErrorString(e));
    return UNIVS_E_LAUNCH;
  }
  mhatc::mha_tc_kernel<VT><<<dim3((unsigned)(a.B * a.heads), (unsigned)a.nsplit), mhatc::kThreads, mhatc::kSmemBytes, st>>>(a);
  return check_launch("mha_tc_forward");
}

extern "C" int univs_mha_tc_forward_f32(void* stream, const float* q, const float* k, const float* v,
                                        const uint32_t* mask_bits, const int32_t* row_open, int mask_batch, int batch,
                                        int len_q, int len_k, int channels, int flags, void* workspace, float* out) {
  UNIVS_REQUIRE(q && k && v && out, "mha_tc_forward: null pointer");
  UNIVS_REQUIRE(batch >= 0 && len_q >= 0 && len_k > 0, "mha_tc_forward: bad sizes (len_k must be > 0)");
  UNIVS_REQUIRE(len_q <= mhatc::kMaxLq, "mha_tc_forward: at most %d queries per batch element (got %d); use univs_mha_forward_f32",
                mhatc::kMaxLq, len_q);
  UNIVS_REQUIRE(channels > 0 && channels % 32 == 0, "mha_tc_forward: channels must be heads*32");
  UNIVS_REQUIRE(mask_bits == nullptr || mask_batch == 1 || mask_batch == batch, "mha_tc_forward: mask_batch must be 1 or batch");
  if (batch == 0 || len_q == 0) return UNIVS_OK;
  const int heads = channels / 32;
  UNIVS_REQUIRE((long long)batch * heads < (1ll << 31), "mha_tc_forward: batch*heads too large");
  mhatc::Args a;
  a.q = q; a.k = k; a.v = v; a.mask_bits = mask_bits; a.row_open = row_open; a.mask_batch = mask_batch;
  a.B = batch; a.Lq = len_q; a.Lk = len_k; a.C = channels; a.heads = heads;
  mhatc::plan(batch, heads, len_k, a.nsplit, a.bps);
  UNIVS_REQUIRE(a.nsplit == 1 || workspace != nullptr, "mha_tc_forward: workspace required (split-K = %d)", a.nsplit);
  UNIVS_REQUIRE(a.nsplit <= 65535, "mha_tc_forward: too many key splits");
  a.out = out;
  a.part_o = reinterpret_cast<float*>(workspace);
  a.part_ml = a.part_o ? a.part_o + (size_t)a.nsplit * batch * heads * len_q * 32 : nullptr;
  UNIVS_REQUIRE((flags & ~1) == 0, "mha_tc_forward: unknown flags %d", flags);
  cudaStream_t st = (cudaStream_t)stream;
  const int rc = (flags & 1) ? launch_mha_tc<true>(st, a) : launch_mha_tc<false>(st, a);
  if (rc || a.nsplit == 1) return rc;
  return launch_mha_combine(st, a.part_o, a.part_ml, batch, heads, len_q, channels, a.nsplit, out);
}
