// Swin (shifted-)window attention for 12x12 windows on the 5th-gen tensor cores, second version (flags bit 0 of
// univs_swin_window_attention_tc).  Same mathematics, operand formats and pipeline as swin_window_attn_tc.cu (read its
// header first); what changed follows the ncu source views of that kernel on the B200 (profiles/r2_wintc_stalls.txt):
//
//   * THE TAIL.  144 query rows = one M=128 row tile + 16 tail rows.  Version 1 gave the tail to ONE warp with its own
//     code (80 score columns per lane through a key -> table LUT, zero-filling of the half rows it does not own, 32
//     shuffles in the epilogue: ~2600 instructions per unit); that warp never waited for anything while the eight main
//     softmax warps spent 17 % of their time waiting for scores the MMA lane could not issue behind it.  Here the tail is
//     simply a SECOND ROW TILE that starts 16 rows further down the (now plain, 144-row) Q buffer: rows 16-143, so the
//     tail rows 128-143 land on TMEM lanes 112-127 and what precedes them is a recomputation nobody reads.  Two tail
//     warps of lane quarter 3 (warps 11 and 15) own those rows exactly like the two warps of a main lane quarter own
//     theirs -- 72 keys each, compile-time bias offsets, max / sum exchange through shared memory, 16 output dims each --
//     with the SAME instructions: one softmax / epilogue code path for all ten warps (a first attempt with two
//     specialised tail kernels was slower than version 1: 25 % of the issue slots went to instruction-cache misses).
//     The tail's P tile is 16 rows per 32-key atom (1 KB) addressed through a tile base 112 rows before it.
//   * relative-position table rows are 44 floats apart instead of 23: 44 = 12 (mod 32), so the 32 consecutive query rows
//     of a warp read 32 different banks (the 23-float rows gave a 2-way conflict on every one of the 72 loads per score row).
//   * the (frame, window) of the next unit is advanced incrementally (three integer divisions per unit and thread were
//     7 % of the softmax warps' issue slots).
//   * NO L1.  The CTA takes ~215 KB of shared memory, so local memory lives in L2: a spilled register costs an L2 round
//     trip.  The first build of this file spent 40 % of the loaders' time on reloads of hoisted per-pass window
//     coordinates (96 registers of loads in flight) and 11 % of the softmax warps' on a spilled loop bound.  Hence: the
//     coordinates are recomputed per unit, the loop bound and the unit stride come from kernel parameters, a softmax
//     thread derives its place from its thread index (RowCtx), the loop holds ONE inlined softmax body and ONE epilogue
//     body, and the production variant (GENERIC = false: operand output only, no score dump) carries no optional
//     pointers.  4 reloads per unit are left (ptxas -v: 24 bytes).
//   * the production variant writes the COMPACT operand [hi | lo * 2^11] the own GEMM reads: 4 instead of 6 bytes per
//     element and no hi * 2^-11 block to compute.
// B200, Swin-L stage 1 of the north-star clip (12960 units): 0.621 ms (version 1) -> 0.438 ms; 24 launches of the step
// 6.3 -> 4.5 ms.
//
// Warps: 0-7 softmax of row tile 0, 11 / 15 softmax of the tail tile, 8 MMA + TMEM allocation, 9 10 12 13 14 loaders
// (160 threads, 20 tokens per pass).
// TMEM columns: S tile0 [0,144)  S tail tile [144,288)  O tile0 [288,320)  O tail [320,352).
#include "tc05.cuh"

namespace univs {
namespace wintc2 {

using namespace tc;

constexpr int kWS = 12;
constexpr int kN = 144;
constexpr int kThreads = 512;
constexpr int kTailWarp0 = 11, kTailWarp1 = 15;   // lane quarter 3
constexpr int kTailRow0 = 16;                  // first Q row of the tail tile: rows 16-143, the tail rows on lanes 112-127
constexpr int kMmaWarp = 8;
constexpr int kAllocWarp = 8;
constexpr int kLoaderThreads = 160;
constexpr int kLoaderSlots = kLoaderThreads / 8;             // tokens per pass: 20
constexpr int kLoaderPasses = (kN + kLoaderSlots - 1) / kLoaderSlots;   // 8 (the last one covers 4 tokens)
constexpr int kTabStride = 44;                 // floats between relative-position table rows in shared memory

// ---- shared memory map (bytes); every operand tile base is a multiple of 1024 --------------------------------------
constexpr int kRow = 64;                       // operand row: 32 halfs = one SWIZZLE_64B span
constexpr int kTile = kN * kRow;               // 9216: q, k or v of a unit, hi or lo
constexpr int kOffQh = 0, kOffQl = kTile, kOffKh = 2 * kTile, kOffKl = 3 * kTile, kOffVh = 4 * kTile, kOffVl = 5 * kTile;
constexpr int kStageBytes = 6 * kTile;         // 55296
constexpr int kP1Atom = 16 * kRow;             // tail tile: 16 stored rows per 32-key atom = rows 112-127 of a tile that starts
constexpr int kP1Lead = (128 - 16) * kRow;     // kP1Lead bytes earlier (the MMA reads all 128 rows: the rest lands in lanes nobody reads)
constexpr int kP0Atom = 128 * kRow;            // 8192
constexpr int kP1Bytes = 5 * kP1Atom;          // 5120
constexpr int kP0Bytes = 5 * kP0Atom;          // 40960
constexpr int kOffP1h = 2 * kStageBytes, kOffP1l = kOffP1h + kP1Bytes;
constexpr int kOffP0h = kOffP1l + kP1Bytes, kOffP0l = kOffP0h + kP0Bytes;
constexpr int kOffBias = kOffP0l + kP0Bytes;   // 23 rows x 44 floats
constexpr int kXchBytes = 2 * 2 * 2 * 128 * 4;  // max[2 parity][2 half][128 tile rows] + sum[2][2][128] floats
constexpr int kOffXch = kOffBias + 1024 * 4;   // one block per row tile
constexpr int kOffBars = kOffXch + 2 * kXchBytes;
constexpr int kNumBars = 16;
constexpr int kSmemBytes = kOffBars + kNumBars * 8 + 16;
static_assert(kStageBytes % 1024 == 0 && kOffP1h % 1024 == 0 && kOffP1l % 1024 == 0 && kOffP0h % 1024 == 0 &&
                  kOffP0l % 1024 == 0 && kOffBars % 8 == 0,
              "tile alignment");
static_assert(22 * kTabStride + 23 <= 1024, "relative-position table");
static_assert(kOffP1h >= kP1Lead && (kOffP1h - kP1Lead) % 1024 == 0 && (kOffP1l - kP1Lead) % 1024 == 0, "tail P tile base");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

constexpr uint32_t kColS0 = 0, kColS1 = 144, kColO0 = 288, kColO1 = 320;

enum Bar { QKV_FULL = 0, QKV_EMPTY = 2, S_FULL = 4, S_FREE = 6, P_FULL = 8, P_FREE = 10, O_FULL = 12, O_FREE = 14 };

struct Geo {
  int B, H, W, C, nH, shift, Hp, Wp, nWh, nWw;
  int sb, swy, swx;    // (frame, window row, window column) digits of the window stride grid / nH between a CTA's units
};
// (frame, window) of this CTA's units: unit n of CTA c is window c / nH + n * (grid / nH) of head c % nH (the grid is a
// multiple of the head count); the window index is advanced digit by digit instead of being divided again
struct Unit {
  int b, wy, wx;
};
__device__ __forceinline__ Unit split_window(int w, const Geo& g) {
  Unit u;
  u.wx = w % g.nWw; w /= g.nWw; u.wy = w % g.nWh; u.b = w / g.nWh;
  return u;
}
__device__ __forceinline__ Unit first_unit(const Geo& g) { return split_window((int)(blockIdx.x / (unsigned)g.nH), g); }
__device__ __forceinline__ void next_unit(Unit& u, const Geo& g) {
  u.wx += g.swx;
  int c = u.wx >= g.nWw;
  u.wx -= c ? g.nWw : 0;
  u.wy += g.swy + c;
  c = u.wy >= g.nWh;
  u.wy -= c ? g.nWh : 0;
  u.b += g.sb + c;
}
// token index of window slot i (row-major inside the window) in the unpadded grid, or -1 for a pad token
__device__ __forceinline__ int source_token(const Geo& g, const Unit& un, int i) {
  const int iy = i / kWS, ix = i - iy * kWS;
  int hs = un.wy * kWS + iy + g.shift, ws = un.wx * kWS + ix + g.shift;   // roll(-shift): rolled[h] = x[(h + shift) % Hp]
  if (hs >= g.Hp) hs -= g.Hp;
  if (ws >= g.Wp) ws -= g.Wp;
  return (hs < g.H && ws < g.W) ? (un.b * g.H + hs) * g.W + ws : -1;
}

// ---- loader ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_token(unsigned char* st, int i, int lane8, float4 q, float4 k, float4 v) {
  uint32_t h0, h1, l0, l1;
  const uint32_t off = swz_off<kRow>(i, lane8 >> 1) + (uint32_t)(lane8 & 1) * 8u;
  split_h2(q.x, q.y, h0, l0);
  split_h2(q.z, q.w, h1, l1);
  sts_v2(smem_u32(st + kOffQh) + off, h0, h1);
  sts_v2(smem_u32(st + kOffQl) + off, l0, l1);
  split_h2(k.x, k.y, h0, l0);
  split_h2(k.z, k.w, h1, l1);
  sts_v2(smem_u32(st + kOffKh) + off, h0, h1);
  sts_v2(smem_u32(st + kOffKl) + off, l0, l1);
  split_h2(v.x, v.y, h0, l0);
  split_h2(v.z, v.w, h1, l1);
  sts_v2(smem_u32(st + kOffVh) + off, h0, h1);
  sts_v2(smem_u32(st + kOffVl) + off, l0, l1);
}

__device__ void loader_loop(unsigned char* smem, uint64_t* bars, const float* __restrict__ qkv,
                            const float* __restrict__ qkv_bias, const float* __restrict__ table, const Geo g,
                            int count, float scale) {
  const int lw = (int)(threadIdx.x >> 5);           // loader warps 9 10 12 13 14 -> 0..4
  const int lt = (lw - 9 - (lw > 11 ? 1 : 0)) * 32 + (int)(threadIdx.x & 31);
  const int lane8 = lt & 7, slot = lt >> 3;
  const int C = g.C;
  // one head per CTA (gridDim.x % nH == 0): its relative-position table once, rows kTabStride floats apart
  const int head = (int)(blockIdx.x % (unsigned)g.nH);
  {
    float* sb = reinterpret_cast<float*>(smem + kOffBias);
    for (int i = lt; i < 23 * 23; i += kLoaderThreads) sb[(i / 23) * kTabStride + i % 23] = __ldg(table + (size_t)i * g.nH + head);
  }
  const int c = head * 32 + lane8 * 4;
  Unit un = first_unit(g);
  int slot_v = slot;
  for (int it = 0; it < count; ++it, next_unit(un, g)) {
    const int s = it & 1;
    unsigned char* st = smem + (size_t)s * kStageBytes;
    // the per-pass window coordinates of this thread's tokens are loop invariants the compiler would keep (and, with 96
    // registers of loads in flight, spill: no L1 is left beside the shared memory, so a spill costs an L2 round trip --
    // the first build spent 40 % of the loaders' time there); recomputing them per unit is a handful of integer operations
    asm volatile("" : "+r"(slot_v));
    const int slot = slot_v;
    float4 q[kLoaderPasses], k[kLoaderPasses], v[kLoaderPasses];
#pragma unroll
    for (int p = 0; p < kLoaderPasses; ++p) {        // the loads do not depend on the stage being free: issue them first
      const int i = p * kLoaderSlots + slot;
      const int src = i < kN ? source_token(g, un, i) : -1;
      q[p] = k[p] = v[p] = make_float4(0.f, 0.f, 0.f, 0.f);   // pad token: qkv == bias (swin.py:247-255)
      if (src >= 0) {
        const float* ptr = qkv + (size_t)src * (3 * C) + c;
        q[p] = ldg_f4(ptr);
        k[p] = ldg_f4(ptr + C);
        v[p] = ldg_f4(ptr + 2 * C);
      }
    }
    if (it >= 2) mbar_wait(&bars[QKV_EMPTY + s], (uint32_t)(((it >> 1) - 1) & 1), QKV_EMPTY + s);
    // the head's qkv bias: L1-resident after the first unit (kept out of the registers the loads in flight need)
    const float4 bq = ldg_f4(qkv_bias + c), bk = ldg_f4(qkv_bias + C + c), bv = ldg_f4(qkv_bias + 2 * C + c);
#pragma unroll
    for (int p = 0; p < kLoaderPasses; ++p) {
      const int i = p * kLoaderSlots + slot;
      if (i < kN) {
        float4 qq = q[p], kk = k[p], vv = v[p];
        qq.x = (qq.x + bq.x) * scale; qq.y = (qq.y + bq.y) * scale; qq.z = (qq.z + bq.z) * scale; qq.w = (qq.w + bq.w) * scale;
        kk.x += bk.x; kk.y += bk.y; kk.z += bk.z; kk.w += bk.w;
        vv.x += bv.x; vv.y += bv.y; vv.z += bv.z; vv.w += bv.w;
        store_token(st, i, lane8, qq, kk, vv);
      }
    }
    fence_proxy_async_smem();
    mbar_arrive(&bars[QKV_FULL + s]);
  }
}

// ---- MMA issuer --------------------------------------------------------------------------------------------------------
// three MMAs per 16-dim k-step: lo*hi + hi*lo + hi*hi (correction terms first)
__device__ __forceinline__ void issue_qk(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t idesc) {
  const uint64_t ah = make_desc<kRow>(a_hi), al = make_desc<kRow>(a_lo);
  const uint64_t bh = make_desc<kRow>(b_hi), bl = make_desc<kRow>(b_lo);
#pragma unroll
  for (int k = 0; k < 2; ++k) {   // 16 dims = 32 bytes inside the swizzle row: +2 in the (addr >> 4) field
    umma_f16(d, al + (uint64_t)(2 * k), bh + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
    umma_f16(d, ah + (uint64_t)(2 * k), bl + (uint64_t)(2 * k), idesc, 1u);
    umma_f16(d, ah + (uint64_t)(2 * k), bh + (uint64_t)(2 * k), idesc, 1u);
  }
}
__device__ __forceinline__ void issue_scores(uint32_t st, int tile, uint32_t tmem_base) {
  if (tile == 0) {
    issue_qk(tmem_base + kColS0, st + kOffQh, st + kOffQl, st + kOffKh, st + kOffKl, make_idesc_f16(128, kN, false, false));
  } else {
    // the same product with the A tile starting at row 16: Q rows 16-143, the tail rows 128-143 on TMEM lanes 112-127
    issue_qk(tmem_base + kColS1, st + kOffQh + kTailRow0 * kRow, st + kOffQl + kTailRow0 * kRow, st + kOffKh, st + kOffKl,
             make_idesc_f16(128, kN, false, false));
  }
}
__device__ __forceinline__ void issue_pv(uint32_t smem_base, uint32_t st, int tile, uint32_t tmem_base) {
  constexpr uint32_t idesc = make_idesc_f16(128, 32, false, true);   // B = V [key][dim]: MN-major
  const uint32_t ph = smem_base + (tile ? kOffP1h - kP1Lead : kOffP0h), pl = smem_base + (tile ? kOffP1l - kP1Lead : kOffP0l);
  const uint32_t atom = tile ? kP1Atom : kP0Atom;
  const uint32_t d = tmem_base + (tile ? kColO1 : kColO0);
#pragma unroll
  for (int s = 0; s < 9; ++s) {   // 16 keys per step
    const uint32_t aoff = (uint32_t)(s >> 1) * atom + (uint32_t)(s & 1) * 32u;
    const uint32_t boff = (uint32_t)s * 16u * kRow;
    const uint64_t ah = make_desc<kRow>(ph + aoff), al = make_desc<kRow>(pl + aoff);
    const uint64_t bh = make_desc<kRow>(st + kOffVh + boff), bl = make_desc<kRow>(st + kOffVl + boff);
    umma_f16(d, al, bh, idesc, s ? 1u : 0u);
    umma_f16(d, ah, bl, idesc, 1u);
    umma_f16(d, ah, bh, idesc, 1u);
  }
}

__device__ void mma_loop(unsigned char* smem, uint64_t* bars, uint32_t tmem_base, int count) {
  const uint32_t smem_base = smem_u32(smem);
  for (int n = -1; n < count; ++n) {
    // scores of unit n+1 (needs its operands, and the softmax of unit n to have read S out of TMEM)
    if (n + 1 < count) {
      const int m = n + 1, s = m & 1;
      mbar_wait(&bars[QKV_FULL + s], (uint32_t)((m >> 1) & 1), QKV_FULL + s);
      for (int tile = 0; tile < 2; ++tile) {
        if (m > 0) mbar_wait(&bars[S_FREE + tile], (uint32_t)((m - 1) & 1), S_FREE + tile);
        fence_after();
        if (elect_one()) {
          issue_scores(smem_base + (uint32_t)s * kStageBytes, tile, tmem_base);
          umma_commit(&bars[S_FULL + tile]);
        }
        __syncwarp();
      }
    }
    if (n < 0) continue;
    // O of unit n (needs P from the softmax, and the epilogue of unit n-1 to have read O out of TMEM)
    const int s = n & 1;
    for (int tile = 0; tile < 2; ++tile) {
      mbar_wait(&bars[P_FULL + tile], (uint32_t)(n & 1), P_FULL + tile);
      if (n > 0) mbar_wait(&bars[O_FREE + tile], (uint32_t)((n - 1) & 1), O_FREE + tile);
      fence_after();
      if (elect_one()) {
        issue_pv(smem_base, smem_base + (uint32_t)s * kStageBytes, tile, tmem_base);
        umma_commit(&bars[O_FULL + tile]);
        umma_commit(&bars[P_FREE + tile]);
        if (tile == 1) umma_commit(&bars[QKV_EMPTY + s]);   // last reader of this stage's operand tiles
      }
      __syncwarp();
    }
  }
}

// ---- softmax + epilogue ------------------------------------------------------------------------------------------------
// One thread = 72 consecutive score columns of one query row; one code path for both row tiles:
//   tile 0 (warps 0-7)   : tile row = 32*quarter + lane = window row, keys [72*half, 72*half + 72)
//   tile 1 (warps 11, 15): tile row = 96 + lane = window row - 16; lanes 16-31 own the tail rows 128-143, lanes 0-15 hold
//                          a second copy of rows 112-127: they run along (the TMEM loads are warp-wide) and store nothing
// The partner of a thread is the same row in the warp of the other key half.
// Everything a softmax thread needs to know about its place follows from its thread index t (warps 0-7: tile 0,
// warps 11 / 15: tail tile), so no per-thread context has to stay in registers across the unrolled bodies:
struct RowCtx {
  int t;
  __device__ __forceinline__ int lane() const { return t & 31; }
  __device__ __forceinline__ int quarter() const { return (t >> 5) & 3; }
  __device__ __forceinline__ int half() const { return (t >> 7) & 1; }       // key half: warps 4-7 and 15
  __device__ __forceinline__ int tb() const { return t >> 8; }               // row tile (barrier index)
  __device__ __forceinline__ int trow() const { return t & 127; }            // row inside the row tile (TMEM lane, P tile row)
  __device__ __forceinline__ int row() const { return (t & 127) + (t >> 8) * kTailRow0; }   // token slot in the window
  __device__ __forceinline__ bool live() const { return t < 256 || (t & 16); }              // this lane's results are stored
};

// GENERIC == false is the production variant: fp16x3 operand output only, no score dump (fewer live registers: with
// ~225 KB of shared memory per CTA almost no L1 is left, so every spill is an L2 round trip)
template <bool GENERIC>
__device__ __forceinline__ float softmax_unit(unsigned char* smem, uint64_t* bars, uint32_t tmem_base, const RowCtx& rc,
                                              const Geo& g, const Unit& un, int n, int u, float* __restrict__ dbg) {
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr int NCOL = 72;
  const int TB = rc.tb();
  uint32_t sr[NCOL];
  mbar_wait(&bars[S_FULL + TB], (uint32_t)(n & 1), S_FULL + TB);
  fence_after();
  {
    // warp-uniform address: lane quarter, row tile and column half are per-warp values
    const uint32_t taddr = tmem_base + ((uint32_t)(rc.quarter() * 32) << 16) + (TB ? kColS1 : kColS0) + (uint32_t)rc.half() * 72u;
    uint32_t* r0 = sr;
    uint32_t* r1 = sr + 32;
    uint32_t* r2 = sr + 64;
    UNIVS_TMEM_LD_X32(taddr, r0);
    UNIVS_TMEM_LD_X32(taddr + 32u, r1);
    UNIVS_TMEM_LD_X8(taddr + 64u, r2);
    tmem_wait_ld();
  }
  fence_before();
  __syncwarp();
  if (rc.lane() == 0) mbar_arrive(&bars[S_FREE + TB]);   // the MMA lane may overwrite S with the next unit

  // relative-position bias (swin.py:108-121: index = (qy-ky+11)*23 + (qx-kx+11)) and the shift mask; 72 keys = 6 key
  // rows, so the per-column part of the table offset is a compile-time constant
  const int qy = rc.row() / kWS, qx = rc.row() - qy * kWS;
  const int key0 = rc.half() * 72;
  const float* bp = reinterpret_cast<const float*>(smem + kOffBias) + (qy + 11 - rc.half() * 6) * kTabStride + (qx + 11);
  float sc[NCOL];
#pragma unroll
  for (int j = 0; j < NCOL; ++j) sc[j] = __uint_as_float(sr[j]) + bp[-((j / kWS) * kTabStride + (j % kWS))];
  // shift mask (swin.py:413-440): -100 between tokens of different regions; only windows in the last window row /
  // column of the padded grid contain more than one region: rows (cols) >= 12 - shift belong to the wrapped part
  const bool mh = g.shift > 0 && un.wy == g.nWh - 1, mw = g.shift > 0 && un.wx == g.nWw - 1;
  if (mh || mw) {
    const int thr = kWS - g.shift;
    uint32_t dh = 0, dw = 0;   // bit y: key row y (key col x) lies in another region than this query
    if (mh) dh = (qy >= thr) ? ((1u << thr) - 1u) : (0xfffu & ~((1u << thr) - 1u));
    if (mw) dw = (qx >= thr) ? ((1u << thr) - 1u) : (0xfffu & ~((1u << thr) - 1u));
    dh >>= rc.half() * 6;
#pragma unroll
    for (int j = 0; j < NCOL; ++j)
      if (((dh >> (j / kWS)) | (dw >> (j % kWS))) & 1u) sc[j] += -100.f;
  }
  if (GENERIC && dbg != nullptr && rc.live()) {
    float* drow = dbg + ((size_t)u * kN + rc.row()) * kN + key0;
#pragma unroll
    for (int j = 0; j < NCOL; ++j) drow[j] = sc[j];
  }
  float mx = sc[0];
#pragma unroll
  for (int j = 1; j < NCOL; ++j) mx = fmaxf(mx, sc[j]);
  float* xch = reinterpret_cast<float*>(smem + kOffXch + TB * kXchBytes);
  {
    float* xmax = xch + (n & 1) * 256;
    xmax[rc.half() * 128 + rc.trow()] = mx;
    named_bar_sync(TB ? 5 : 1 + rc.quarter(), 64);
    mx = fmaxf(mx, xmax[(rc.half() ^ 1) * 128 + rc.trow()]);
  }
  const float mneg = -mx * kLog2e;

  // P = exp(s - max) as fp16 hi + lo into the K-major SWIZZLE_64B A tile of the PV MMA (32-key atoms)
  if (n > 0) mbar_wait(&bars[P_FREE + TB], (uint32_t)((n - 1) & 1), P_FREE + TB);
  const uint32_t ph = smem_u32(smem) + (uint32_t)(TB ? kOffP1h - kP1Lead : kOffP0h);
  const uint32_t pl = ph + (uint32_t)(TB ? kP1Bytes : kP0Bytes);
  const uint32_t atom = TB ? kP1Atom : kP0Atom;
  const uint32_t rowoff = (uint32_t)rc.trow() * kRow;
  const uint32_t sw = (uint32_t)(rc.trow() >> 1) & 3u;
  const uint32_t chunk0 = (uint32_t)key0 >> 3;                               // first 16-byte chunk (8 keys) of this thread
  float sum = 0.f;
#pragma unroll
  for (int cc = 0; cc < NCOL / 8; ++cc) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float p0 = ex2_approx(fmaf(sc[cc * 8 + 2 * e], kLog2e, mneg));
      const float p1 = ex2_approx(fmaf(sc[cc * 8 + 2 * e + 1], kLog2e, mneg));
      sum += p0 + p1;
      split_h2(p0, p1, hi[e], lo[e]);
    }
    const uint32_t gch = chunk0 + (uint32_t)cc;
    const uint32_t off = (gch >> 2) * atom + rowoff + (((gch & 3u) ^ sw) << 4);
    if (rc.live()) {
      sts_v4(ph + off, hi[0], hi[1], hi[2], hi[3]);
      sts_v4(pl + off, lo[0], lo[1], lo[2], lo[3]);
    }
  }
  xch[512 + (n & 1) * 256 + rc.half() * 128 + rc.trow()] = sum;
  fence_proxy_async_smem();
  mbar_arrive(&bars[P_FULL + TB]);
  return sum;
}

template <bool GENERIC>
__device__ __forceinline__ void epilogue_unit(unsigned char* smem, uint64_t* bars, uint32_t tmem_base, const RowCtx& rc,
                                              const Geo& g, int src, int head, int n, float sum, float* __restrict__ out,
                                              __half* __restrict__ out16, bool compact) {
  const int TB = rc.tb();
  mbar_wait(&bars[O_FULL + TB], (uint32_t)(n & 1), O_FULL + TB);
  fence_after();
  uint32_t r[16];
  const uint32_t taddr = tmem_base + ((uint32_t)(rc.quarter() * 32) << 16) + (TB ? kColO1 : kColO0) + (uint32_t)rc.half() * 16u;
  UNIVS_TMEM_LD_X16(taddr, r);
  tmem_wait_ld();
  const float total =
      sum + reinterpret_cast<const float*>(smem + kOffXch + TB * kXchBytes)[512 + (n & 1) * 256 + (rc.half() ^ 1) * 128 + rc.trow()];
  fence_before();
  __syncwarp();
  if (rc.lane() == 0) mbar_arrive(&bars[O_FREE + TB]);

  if (src < 0) return;              // pad row (dropped by window_reverse + roll(+shift) + crop) or a lane without a row
  const float inv = 1.f / total;
  const int c = head * 32 + rc.half() * 16;
  if (GENERIC && out != nullptr) {
    float4* dst = reinterpret_cast<float4*>(out + (size_t)src * g.C + c);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      dst[e] = make_float4(__uint_as_float(r[4 * e]) * inv, __uint_as_float(r[4 * e + 1]) * inv, __uint_as_float(r[4 * e + 2]) * inv,
                           __uint_as_float(r[4 * e + 3]) * inv);
  }
  if (!GENERIC || out16 != nullptr) {
    // GEMM operand of the projection: value = hi + lo' * 2^-11.  compact: [hi (C) | lo' (C)] (what gemm_tc.cu reads);
    // otherwise the round-1 K-chunk container [lo' (C) | hi*2^-11 (C) | hi (C)] of the library-GEMM path (C <= 1536)
    uint32_t lo2[8], hi2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float a = __uint_as_float(r[2 * e]) * inv, b = __uint_as_float(r[2 * e + 1]) * inv;
      const __half2 h = __floats2half2_rn(a, b);
      const float2 f = __half22float2(h);
      const __half2 l = __floats2half2_rn((a - f.x) * 2048.f, (b - f.y) * 2048.f);
      hi2[e] = *reinterpret_cast<const uint32_t*>(&h);
      lo2[e] = *reinterpret_cast<const uint32_t*>(&l);
    }
    __half* rowp = out16 + (size_t)src * ((compact ? 2 : 3) * (size_t)g.C) + c;
    uint4* dl = reinterpret_cast<uint4*>(compact ? rowp + g.C : rowp);
    uint4* dh = reinterpret_cast<uint4*>(compact ? rowp : rowp + 2 * g.C);
    dl[0] = make_uint4(lo2[0], lo2[1], lo2[2], lo2[3]);
    dl[1] = make_uint4(lo2[4], lo2[5], lo2[6], lo2[7]);
    dh[0] = make_uint4(hi2[0], hi2[1], hi2[2], hi2[3]);
    dh[1] = make_uint4(hi2[4], hi2[5], hi2[6], hi2[7]);
    if (!compact) {
      uint32_t hs2[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi2[e]));
        const __half2 s = __floats2half2_rn(f.x * (1.f / 2048.f), f.y * (1.f / 2048.f));
        hs2[e] = *reinterpret_cast<const uint32_t*>(&s);
      }
      uint4* d1 = reinterpret_cast<uint4*>(rowp + g.C);
      d1[0] = make_uint4(hs2[0], hs2[1], hs2[2], hs2[3]);
      d1[1] = make_uint4(hs2[4], hs2[5], hs2[6], hs2[7]);
    }
  }
}

template <bool GENERIC>
__device__ void softmax_loop(unsigned char* smem, uint64_t* bars, uint32_t tmem_base, const RowCtx rc, const Geo g, int units,
                             float* __restrict__ out, __half* __restrict__ out16, float* __restrict__ dbg, bool compact) {
  // the loop bound is re-derived from kernel parameters every iteration (a unit count kept across the two inlined
  // bodies was spilled to local memory and re-read at the loop head: 11 % of the softmax warps' stalls)
  const int head = (int)(blockIdx.x % (unsigned)g.nH);
  const int G = (int)gridDim.x;
  int u = (int)blockIdx.x;
  Unit un = first_unit(g);
  float sum_prev = 0.f;
  int src_prev = -1;
  // iteration n: softmax of unit n (if there is one), then the epilogue of unit n-1 -- one inlined copy of each body
  for (int n = 0;; ++n, u += G) {
    float sum = 0.f;
    int src = -1;
    const bool more = u < units;
    if (more) {
      sum = softmax_unit<GENERIC>(smem, bars, tmem_base, rc, g, un, n, u, dbg);
      src = rc.live() ? source_token(g, un, rc.row()) : -1;
      next_unit(un, g);
    }
    if (n > 0) epilogue_unit<GENERIC>(smem, bars, tmem_base, rc, g, src_prev, head, n - 1, sum_prev, out, out16, compact);
    if (!more) break;
    src_prev = src;
    sum_prev = sum;
  }
}

template <bool GENERIC>
__global__ void __launch_bounds__(kThreads, 1)
swin_window_attn_tc12v2_kernel(const float* __restrict__ qkv, const float* __restrict__ qkv_bias,
                               const float* __restrict__ table, const Geo g, long long units, float scale,
                               float* __restrict__ out, __half* __restrict__ out16, float* __restrict__ dbg, int compact) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int count = (int)((units - blockIdx.x + gridDim.x - 1) / gridDim.x);   // units of this CTA (grid <= units)

  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars[QKV_FULL + s], kLoaderThreads);
      mbar_init(&bars[QKV_EMPTY + s], 1);
    }
    mbar_init(&bars[S_FULL + 0], 1);
    mbar_init(&bars[S_FULL + 1], 1);
    mbar_init(&bars[S_FREE + 0], 8);      // one arrive per softmax warp of tile 0
    mbar_init(&bars[S_FREE + 1], 2);      // and per tail-tile warp
    mbar_init(&bars[P_FULL + 0], 256);    // every softmax thread arrives after its own P stores + proxy fence
    mbar_init(&bars[P_FULL + 1], 64);
    mbar_init(&bars[P_FREE + 0], 1);
    mbar_init(&bars[P_FREE + 1], 1);
    mbar_init(&bars[O_FULL + 0], 1);
    mbar_init(&bars[O_FULL + 1], 1);
    mbar_init(&bars[O_FREE + 0], 8);
    mbar_init(&bars[O_FREE + 1], 2);
    mbar_init_fence();
  }
  if (warp == kAllocWarp) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kMmaWarp) {
    mma_loop(smem, bars, tmem_base, count);
  } else if (warp < 8 || warp == kTailWarp0 || warp == kTailWarp1) {
    RowCtx rc;
    rc.t = (int)threadIdx.x;
    softmax_loop<GENERIC>(smem, bars, tmem_base, rc, g, (int)units, out, out16, dbg, compact != 0);
  } else {
    loader_loop(smem, bars, qkv, qkv_bias, table, g, count, scale);
  }
  fence_before();
  __syncthreads();
  if (warp == kAllocWarp) {
    fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch(cudaStream_t st, const float* qkv, const float* bias, const float* table, int batch, int height, int width,
           int channels, int num_heads, int shift, float* out, __half* out16, float* dbg, int compact) {
  Geo g;
  g.B = batch; g.H = height; g.W = width; g.C = channels; g.nH = num_heads; g.shift = shift;
  g.Hp = (height + kWS - 1) / kWS * kWS;
  g.Wp = (width + kWS - 1) / kWS * kWS;
  g.nWh = g.Hp / kWS;
  g.nWw = g.Wp / kWS;
  const long long units = (long long)g.B * g.nWh * g.nWw * g.nH;
  UNIVS_REQUIRE(units < (1ll << 31), "swin_window_attention_tc: too many (window, head) units");
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const bool generic = out != nullptr || dbg != nullptr;
  auto kernel = generic ? swin_window_attn_tc12v2_kernel<true> : swin_window_attn_tc12v2_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) {
    set_error("swin_window_attention_tc: cudaFuncSetAttribute(%d): %s", kSmemBytes, cudaGetErrorString(e));
    return UNIVS_E_LAUNCH;
  }
  // one head per CTA: the grid is a multiple of the head count (units = windows * nH is one too), unit u has head u % nH
  long long grid_ll = g.nH <= num_sms ? (long long)(num_sms / g.nH) * g.nH : g.nH;
  if (grid_ll > units) grid_ll = units;
  {
    int d = (int)(grid_ll / g.nH);
    g.swx = d % g.nWw; d /= g.nWw; g.swy = d % g.nWh; g.sb = d / g.nWh;
  }
  const float scale = 0.17677669529663687f;   // 32^-0.5 (swin.py:96)
  if (generic)
    swin_window_attn_tc12v2_kernel<true><<<(int)grid_ll, kThreads, kSmemBytes, st>>>(qkv, bias, table, g, units, scale, out, out16, dbg, compact);
  else
    swin_window_attn_tc12v2_kernel<false><<<(int)grid_ll, kThreads, kSmemBytes, st>>>(qkv, bias, table, g, units, scale, out, out16, dbg, compact);
  return check_launch("swin_window_attention_tc");
}

}  // namespace wintc2
}  // namespace univs
