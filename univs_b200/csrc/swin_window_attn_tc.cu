// Swin (shifted-)window attention for 12x12 windows on the 5th-gen tensor cores (strict fp16 hi|lo policy):
// S = (q*scale) K^T and O = P V are tcgen05.mma (kind::f16, fp32 accumulation in TMEM), the softmax runs on rows read
// back with tcgen05.ld (TMEM lane = query row, so row max / sum are thread-local), P returns to shared memory as the
// A operand of the second MMA.  Reference semantics: swin.py:131-171 (attention), :108-121 (relative-position index),
// :413-440 (shift mask), :247-289 (pad / roll / partition / reverse -- folded into addressing, as in
// swin_window_attn.cu, whose mma.sync kernels remain the path for 7x7 / 4x4 windows and the cross-check of this one).
//
// Work unit = (frame, window, head): 144 tokens x 32 dims of q, k, v.  Persistent CTAs (one per SM, 512 threads),
// warp-specialised, units round-robin over CTAs:
//   warps 10-15  loaders : gather q/k/v rows through the roll/pad addressing (128-bit loads, all 18 of a thread's unit share
//                          in flight at once: one memory latency per unit), add the qkv bias, scale q, split fp32 -> fp16
//                          hi + lo once, store K-major SWIZZLE_64B operand tiles (64-byte rows = 32 dims); 2-stage ring.
//                          The grid is a multiple of the head count, so a CTA only ever sees ONE head: its
//                          relative-position table and qkv bias are fetched once per CTA, not per unit (ncu, round 2: the
//                          per-unit table gather was 17 % of the loader's time, the three dependent load batches most of
//                          the rest, and the softmax warps spent 43 % of theirs waiting for scores)
//   warp 9       MMA     : one elected lane issues  S = Ql Kh^T + Qh Kl^T + Qh Kh^T  and  O = Pl Vh + Ph Vl + Ph Vh;
//                          V is consumed as stored ([key][dim] rows) through the MN-major B descriptor
//   warps 0-7    softmax of row tile 0 (query rows 0-127; M=128, N=144): warp = (TMEM lane quarter, column half); a thread
//                          owns 72 scores of one row; the halves exchange max / sum through shared memory
//   warp 8       softmax of the tail tile (query rows 128-143).  tcgen05.ld addresses are warp-uniform, so the two lane
//                          halves cannot read different columns; instead the MMA puts different KEYS under the same
//                          columns: the 16 tail rows of Q sit between two blocks of 16 zero rows ("Z Q Z"), the A tile
//                          starting at Z (rows = [0 | Q]) multiplies keys 64-143 (N=80, overwrite), the A tile starting
//                          at Q (rows = [Q | 0]) multiplies keys 0-63 (N=64, accumulate): TMEM lanes 0-15 end up with
//                          keys 0-63 of the 16 rows, lanes 16-31 with keys 64-143 of the same rows, both in columns
//                          0-79.  Each lane writes its part of its P row and zeros elsewhere, the PV MMA leaves two
//                          partial sums per row and the epilogue adds them with one shuffle.
//   warp 9 also allocates TMEM.
// Softmax warps run  softmax(n+1) -> epilogue(n), the MMA lane  S(n+1) -> PV(n), so the tensor pipe works on the next
// unit's scores while the softmax of the current one is in flight, and no softmax warp waits for a PV it just enabled.
// TMEM columns: S tile0 [0,144)  S tail [160,240)  O tile0 [320,352)  O tail [352,384).
//
// HBM-bound by design: per unit 55.3 KB of qkv in, 18.4 KB (fp32) or 27.6 KB (fp16x3 operand) out; tensor work per unit
// is 45 MMAs (~1.5k tensor cycles), the issue-slot budget is dominated by the softmax (~9 instructions per score).
#include "tc05.cuh"

namespace univs {
namespace wintc {

using namespace tc;

constexpr int kWS = 12;
constexpr int kN = 144;
constexpr int kThreads = 512;
constexpr int kMmaWarp = 9;
constexpr int kAllocWarp = 9;
constexpr int kLoaderWarp0 = 10;
constexpr int kLoaderThreads = 192;
constexpr int kItemsPerLoader = kN * 8 / kLoaderThreads;   // (token, 4-dim group) items per loader thread and unit: 6
constexpr int kTable = 23 * 23;
constexpr int kTailKeys0 = 64;                 // keys of the tail rows handled by TMEM lanes 0-15 (lanes 16-31: the other 80)

// ---- shared memory map (bytes); every operand tile base is a multiple of 1024 --------------------------------------
constexpr int kRow = 64;                       // operand row: 32 halfs = one SWIZZLE_64B span
constexpr int kQTailOff = 128 * kRow;          // "Z Q Z": 16 zero rows, the 16 tail rows, 16 zero rows
constexpr int kQBytes = kQTailOff + 48 * kRow; // 11264 (an MMA reads 128 rows from its start: what follows is garbage in
                                               // TMEM lanes nobody reads)
constexpr int kKBytes = kN * kRow;             // 9216
constexpr int kVBytes = kN * kRow;             // [key][dim] rows, MN-major B operand
constexpr int kOffQh = 0, kOffQl = kQBytes, kOffKh = 2 * kQBytes, kOffKl = kOffKh + kKBytes;
constexpr int kOffVh = kOffKl + kKBytes, kOffVl = kOffVh + kVBytes;
constexpr int kStageBytes = kOffVl + kVBytes;  // 59392
constexpr int kP1Atom = 32 * kRow;             // tail tile: 32 valid rows per 32-key atom
constexpr int kP0Atom = 128 * kRow;            // 8192
constexpr int kP1Bytes = 5 * kP1Atom;          // 10240 (the MMA reads up to 6 KB past it: lands in the P0 buffers)
constexpr int kP0Bytes = 5 * kP0Atom;          // 40960
constexpr int kOffP1h = 2 * kStageBytes, kOffP1l = kOffP1h + kP1Bytes;
constexpr int kOffP0h = kOffP1l + kP1Bytes, kOffP0l = kOffP0h + kP0Bytes;
constexpr int kOffBias = kOffP0l + kP0Bytes;   // 544 floats (+ 544 unused)
constexpr int kBiasStride = 544;
constexpr int kOffXch = kOffBias + 2 * kBiasStride * 4;   // max[2 parity][2 half][128] + sum[2][2][128] floats
constexpr int kOffKidx = kOffXch + 2 * 2 * 2 * 128 * 4;   // int[144]: key -> ky * 23 + kx
constexpr int kOffBars = kOffKidx + kN * 4;
constexpr int kNumBars = 16;
constexpr int kSmemBytes = kOffBars + kNumBars * 8 + 16;
static_assert(kQBytes % 1024 == 0 && kStageBytes % 1024 == 0 && kOffP1h % 1024 == 0 && kOffP0h % 1024 == 0 &&
                  kOffP0l % 1024 == 0 && kOffBars % 8 == 0,
              "tile alignment");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

constexpr uint32_t kColS0 = 0, kColS1 = 160, kColO0 = 320, kColO1 = 352;

enum Bar { QKV_FULL = 0, QKV_EMPTY = 2, S_FULL = 4, S_FREE = 6, P_FULL = 8, P_FREE = 10, O_FULL = 12, O_FREE = 14 };

struct Geo {
  int B, H, W, C, nH, shift, Hp, Wp, nWh, nWw;
};
struct Unit {
  int b, wy, wx, head;
};
__device__ __forceinline__ Unit decode(long long u, const Geo& g) {
  Unit r;
  r.head = (int)(u % g.nH);
  long long w = u / g.nH;
  r.wx = (int)(w % g.nWw);
  w /= g.nWw;
  r.wy = (int)(w % g.nWh);
  r.b = (int)(w / g.nWh);
  return r;
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
  const uint32_t sub = (uint32_t)(lane8 & 1) * 8u;
  const uint32_t off = swz_off<kRow>(i, lane8 >> 1) + sub;
  // query rows 0-127: tile 0; rows 128-143: rows 16-31 of the "Z Q Z" block behind it
  const uint32_t qoff = i < 128 ? off : (uint32_t)kQTailOff + swz_off<kRow>(16 + (i - 128), lane8 >> 1) + sub;
  split_h2(q.x, q.y, h0, l0);
  split_h2(q.z, q.w, h1, l1);
  sts_v2(smem_u32(st + kOffQh) + qoff, h0, h1);
  sts_v2(smem_u32(st + kOffQl) + qoff, l0, l1);
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
                            long long units, float scale) {
  const int lt = threadIdx.x - kLoaderWarp0 * 32;
  const int lane8 = lt & 7, slot = lt >> 3;
  const int C = g.C;
  // the zero rows of the "Z Q Z" blocks (both stages, hi and lo) are written once; token stores never touch them
  for (int i = lt; i < 2 * 2 * 2 * 16 * 4; i += kLoaderThreads) {     // stage x {hi,lo} x {first,last Z} x 16 rows x 4 chunks
    const int chunk = i & 3, row = (i >> 2) & 15, z = (i >> 6) & 1, hl = (i >> 7) & 1, s = i >> 8;
    sts_v4(smem_u32(smem + (size_t)s * kStageBytes + (hl ? kOffQl : kOffQh) + kQTailOff) + (uint32_t)(z * 32 + row) * kRow +
               (uint32_t)chunk * 16u,
           0u, 0u, 0u, 0u);
  }
  // one head per CTA (gridDim.x % nH == 0): relative-position table, key LUT and qkv bias once
  const int head = (int)(blockIdx.x % (unsigned)g.nH);
  {
    float* sb = reinterpret_cast<float*>(smem + kOffBias);
    for (int i = lt; i < kTable; i += kLoaderThreads) sb[i] = __ldg(table + (size_t)i * g.nH + head);
    int* kidx = reinterpret_cast<int*>(smem + kOffKidx);
    for (int k = lt; k < kN; k += kLoaderThreads) kidx[k] = (k / kWS) * 23 + (k % kWS);
  }
  const int c = head * 32 + lane8 * 4;
  int it = 0;
  for (long long u = blockIdx.x; u < units; u += gridDim.x, ++it) {
    const int s = it & 1;
    const Unit un = decode(u, g);
    unsigned char* st = smem + (size_t)s * kStageBytes;
    float4 q[kItemsPerLoader], k[kItemsPerLoader], v[kItemsPerLoader];
#pragma unroll
    for (int p = 0; p < kItemsPerLoader; ++p) {      // the loads do not depend on the stage being free: issue them first
      const int i = p * (kLoaderThreads / 8) + slot;
      const int src = source_token(g, un, i);
      q[p] = k[p] = v[p] = make_float4(0.f, 0.f, 0.f, 0.f);   // pad token: qkv == bias (swin.py:247-255)
      if (src >= 0) {
        const float* ptr = qkv + (size_t)src * (3 * C) + c;
        q[p] = ldg_f4(ptr);
        k[p] = ldg_f4(ptr + C);
        v[p] = ldg_f4(ptr + 2 * C);
      }
    }
    if (it >= 2) mbar_wait(&bars[QKV_EMPTY + s], (uint32_t)(((it >> 1) - 1) & 1), QKV_EMPTY + s);
    // the head's qkv bias: L1-resident after the first unit (kept out of the registers the 18 loads in flight need)
    const float4 bq = ldg_f4(qkv_bias + c), bk = ldg_f4(qkv_bias + C + c), bv = ldg_f4(qkv_bias + 2 * C + c);
#pragma unroll
    for (int p = 0; p < kItemsPerLoader; ++p) {
      const int i = p * (kLoaderThreads / 8) + slot;
      float4 qq = q[p], kk = k[p], vv = v[p];
      qq.x = (qq.x + bq.x) * scale; qq.y = (qq.y + bq.y) * scale; qq.z = (qq.z + bq.z) * scale; qq.w = (qq.w + bq.w) * scale;
      kk.x += bk.x; kk.y += bk.y; kk.z += bk.z; kk.w += bk.w;
      vv.x += bv.x; vv.y += bv.y; vv.z += bv.z; vv.w += bv.w;
      store_token(st, i, lane8, qq, kk, vv);
    }
    fence_proxy_async_smem();
    mbar_arrive(&bars[QKV_FULL + s]);
  }
}

// ---- MMA issuer --------------------------------------------------------------------------------------------------------
// three MMAs per 16-dim k-step: lo*hi + hi*lo + hi*hi (correction terms first)
__device__ __forceinline__ void issue_qk(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t idesc,
                                         bool accumulate) {
  const uint64_t ah = make_desc<kRow>(a_hi), al = make_desc<kRow>(a_lo);
  const uint64_t bh = make_desc<kRow>(b_hi), bl = make_desc<kRow>(b_lo);
#pragma unroll
  for (int k = 0; k < 2; ++k) {   // 16 dims = 32 bytes inside the swizzle row: +2 in the (addr >> 4) field
    umma_f16(d, al + (uint64_t)(2 * k), bh + (uint64_t)(2 * k), idesc, (accumulate || k) ? 1u : 0u);
    umma_f16(d, ah + (uint64_t)(2 * k), bl + (uint64_t)(2 * k), idesc, 1u);
    umma_f16(d, ah + (uint64_t)(2 * k), bh + (uint64_t)(2 * k), idesc, 1u);
  }
}
__device__ __forceinline__ void issue_scores(uint32_t st, int tile, uint32_t tmem_base) {
  if (tile == 0) {
    issue_qk(tmem_base + kColS0, st + kOffQh, st + kOffQl, st + kOffKh, st + kOffKl, make_idesc_f16(128, kN, false, false), false);
  } else {
    // A = [0 | Q_tail] x keys 64..143 (overwrites columns 0-79), then A = [Q_tail | 0] x keys 0..63 (accumulates into 0-63)
    constexpr uint32_t k1 = (uint32_t)kTailKeys0 * kRow;
    issue_qk(tmem_base + kColS1, st + kOffQh + kQTailOff, st + kOffQl + kQTailOff, st + kOffKh + k1, st + kOffKl + k1,
             make_idesc_f16(128, kN - kTailKeys0, false, false), false);
    issue_qk(tmem_base + kColS1, st + kOffQh + kQTailOff + 16 * kRow, st + kOffQl + kQTailOff + 16 * kRow, st + kOffKh,
             st + kOffKl, make_idesc_f16(128, kTailKeys0, false, false), true);
  }
}
__device__ __forceinline__ void issue_pv(uint32_t smem_base, uint32_t st, int tile, uint32_t tmem_base) {
  constexpr uint32_t idesc = make_idesc_f16(128, 32, false, true);   // B = V [key][dim]: MN-major
  const uint32_t ph = smem_base + (tile ? kOffP1h : kOffP0h), pl = smem_base + (tile ? kOffP1l : kOffP0l);
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
// One thread = NCOL consecutive score columns of one query row.
//   TAIL == false: row = 32*quarter + lane of tile 0, keys [72*half, 72*half + 72); partner = same row in warp (quarter, half^1)
//   TAIL == true : row = 128 + (lane & 15), half = lane >> 4: keys [0,64) (half 0; columns 64-79 are not its keys) or
//                  [64,144) (half 1); partner = lane ^ 16
struct RowCtx {
  int quarter, half, lane;
  int row;       // token slot in the window, 0..143
  int prow;      // row inside the P tile
};

template <bool TAIL>
__device__ __forceinline__ float softmax_unit(unsigned char* smem, uint64_t* bars, uint32_t tmem_base, const RowCtx& rc,
                                              const Geo& g, const Unit& un, int n, long long u, float* __restrict__ dbg) {
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr int NCOL = TAIL ? (kN - kTailKeys0) : 72;
  constexpr int TB = TAIL ? 1 : 0;
  uint32_t sr[NCOL];
  mbar_wait(&bars[S_FULL + TB], (uint32_t)(n & 1), S_FULL + TB);
  fence_after();
  {
    // warp-uniform address: lane quarter and (tile 0) column half are per-warp values
    const uint32_t taddr = TAIL ? tmem_base + kColS1
                                : tmem_base + ((uint32_t)(rc.quarter * 32) << 16) + kColS0 + (uint32_t)rc.half * 72u;
    uint32_t* r0 = sr;
    uint32_t* r1 = sr + 32;
    uint32_t* r2 = sr + 64;
    UNIVS_TMEM_LD_X32(taddr, r0);
    UNIVS_TMEM_LD_X32(taddr + 32u, r1);
    if (TAIL) {
      UNIVS_TMEM_LD_X16(taddr + 64u, r2);
    } else {
      UNIVS_TMEM_LD_X8(taddr + 64u, r2);
    }
    tmem_wait_ld();
  }
  fence_before();
  __syncwarp();
  if (rc.lane == 0) mbar_arrive(&bars[S_FREE + TB]);   // the MMA lane may overwrite S with the next unit

  // relative-position bias (swin.py:108-121: index = (qy-ky+11)*23 + (qx-kx+11)) and the shift mask
  const int qy = rc.row / kWS, qx = rc.row - qy * kWS;
  const float* sbias = reinterpret_cast<const float*>(smem + kOffBias) + (qy + 11) * 23 + (qx + 11);
  const int* kidx = reinterpret_cast<const int*>(smem + kOffKidx);
  const int key0 = TAIL ? rc.half * kTailKeys0 : rc.half * 72;
  float sc[NCOL];
  if (!TAIL) {
    const float* bp = sbias - rc.half * 6 * 23;     // 72 keys = 6 key rows: the per-column part is a compile-time constant
#pragma unroll
    for (int j = 0; j < NCOL; ++j) sc[j] = __uint_as_float(sr[j]) + bp[-((j / kWS) * 23 + (j % kWS))];
  } else {
    // key -> table offset through the LUT (64 keys are not a whole number of key rows); in blocks of 16 columns with a
    // scheduling barrier in between, so that the 160 dependent loads are not all hoisted (register pressure)
#pragma unroll
    for (int jb = 0; jb < NCOL; jb += 16) {
#pragma unroll
      for (int j = jb; j < jb + 16; ++j) sc[j] = __uint_as_float(sr[j]) + sbias[-kidx[key0 + j]];
      asm volatile("" ::: "memory");
    }
  }
  // shift mask (swin.py:413-440): -100 between tokens of different regions; only windows in the last window row /
  // column of the padded grid contain more than one region: rows (cols) >= 12 - shift belong to the wrapped part
  const bool mh = g.shift > 0 && un.wy == g.nWh - 1, mw = g.shift > 0 && un.wx == g.nWw - 1;
  if (mh || mw) {
    const int thr = kWS - g.shift;
    uint32_t dh = 0, dw = 0;   // bit y: key row y (key col x) lies in another region than this query
    if (mh) dh = (qy >= thr) ? ((1u << thr) - 1u) : (0xfffu & ~((1u << thr) - 1u));
    if (mw) dw = (qx >= thr) ? ((1u << thr) - 1u) : (0xfffu & ~((1u << thr) - 1u));
    if (!TAIL) {
      dh >>= rc.half * 6;
#pragma unroll
      for (int j = 0; j < NCOL; ++j)
        if (((dh >> (j / kWS)) | (dw >> (j % kWS))) & 1u) sc[j] += -100.f;
    } else {
#pragma unroll
      for (int j = 0; j < NCOL; ++j) {
        const int c = kidx[key0 + j];
        const int ky = c / 23, kx = c - ky * 23;
        if (((dh >> ky) | (dw >> kx)) & 1u) sc[j] += -100.f;
      }
    }
  }
  if (TAIL) {      // lanes 0-15 own 64 keys only: the other 16 columns hold nothing of theirs
#pragma unroll
    for (int j = kTailKeys0; j < NCOL; ++j)
      if (rc.half == 0) sc[j] = -INFINITY;
  }
  if (dbg != nullptr) {
    float* drow = dbg + ((size_t)u * kN + rc.row) * kN + key0;
#pragma unroll
    for (int j = 0; j < NCOL; ++j)
      if (!TAIL || rc.half == 1 || j < kTailKeys0) drow[j] = sc[j];
  }
  float mx = sc[0];
#pragma unroll
  for (int j = 1; j < NCOL; ++j) mx = fmaxf(mx, sc[j]);
  float* xch = reinterpret_cast<float*>(smem + kOffXch);
  if (!TAIL) {
    float* xmax = xch + (n & 1) * 256;
    xmax[rc.half * 128 + rc.row] = mx;
    named_bar_sync(1 + rc.quarter, 64);
    mx = fmaxf(mx, xmax[(rc.half ^ 1) * 128 + rc.row]);
  } else {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
  }
  const float mneg = -mx * kLog2e;

  // P = exp(s - max) as fp16 hi + lo into the K-major SWIZZLE_64B A tile of the PV MMA (32-key atoms)
  if (n > 0) mbar_wait(&bars[P_FREE + TB], (uint32_t)((n - 1) & 1), P_FREE + TB);
  const uint32_t ph = smem_u32(smem + (TAIL ? kOffP1h : kOffP0h)), pl = smem_u32(smem + (TAIL ? kOffP1l : kOffP0l));
  constexpr uint32_t atom = TAIL ? kP1Atom : kP0Atom;
  const uint32_t rowoff = (uint32_t)rc.prow * kRow;
  const uint32_t sw = (uint32_t)(rc.prow >> 1) & 3u;
  const uint32_t chunk0 = (uint32_t)key0 >> 3;                               // first 16-byte chunk (8 keys) of this thread
  float sum = 0.f;
#pragma unroll
  for (int cc = 0; cc < NCOL / 8; ++cc) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float p0 = ex2_approx(fmaf(sc[cc * 8 + 2 * e], kLog2e, mneg));        // exp2(-inf) = 0 in the unowned columns
      const float p1 = ex2_approx(fmaf(sc[cc * 8 + 2 * e + 1], kLog2e, mneg));
      sum += p0 + p1;
      split_h2(p0, p1, hi[e], lo[e]);
    }
    const uint32_t gch = chunk0 + (uint32_t)cc;
    const uint32_t off = (gch >> 2) * atom + rowoff + (((gch & 3u) ^ sw) << 4);
    sts_v4(ph + off, hi[0], hi[1], hi[2], hi[3]);
    sts_v4(pl + off, lo[0], lo[1], lo[2], lo[3]);
  }
  if (TAIL) {
    // the rest of this lane's P row is zero (its keys belong to the partner lane's copy of the row):
    // half 0 wrote chunks 0-9 (8, 9 as zeros) -> zero 10-17;  half 1 wrote chunks 8-17 -> zero 0-7
#pragma unroll
    for (int cz = 0; cz < 8; ++cz) {
      const uint32_t gz = (rc.half ? 0u : 10u) + (uint32_t)cz;
      const uint32_t offz = (gz >> 2) * atom + rowoff + (((gz & 3u) ^ sw) << 4);
      sts_v4(ph + offz, 0u, 0u, 0u, 0u);
      sts_v4(pl + offz, 0u, 0u, 0u, 0u);
    }
  } else {
    xch[512 + (n & 1) * 256 + rc.half * 128 + rc.row] = sum;
  }
  fence_proxy_async_smem();
  mbar_arrive(&bars[P_FULL + TB]);
  return sum;
}

template <bool TAIL>
__device__ __forceinline__ void epilogue_unit(unsigned char* smem, uint64_t* bars, uint32_t tmem_base, const RowCtx& rc,
                                              const Geo& g, const Unit& un, int n, float sum, float* __restrict__ out,
                                              __half* __restrict__ out16) {
  constexpr int TB = TAIL ? 1 : 0;
  mbar_wait(&bars[O_FULL + TB], (uint32_t)(n & 1), O_FULL + TB);
  fence_after();
  float o[16];
  float total;
  if (!TAIL) {
    uint32_t r[16];
    const uint32_t taddr = tmem_base + ((uint32_t)(rc.quarter * 32) << 16) + kColO0 + (uint32_t)rc.half * 16u;
    UNIVS_TMEM_LD_X16(taddr, r);
    tmem_wait_ld();
#pragma unroll
    for (int e = 0; e < 16; ++e) o[e] = __uint_as_float(r[e]);
    total = sum + reinterpret_cast<const float*>(smem + kOffXch)[512 + (n & 1) * 256 + (rc.half ^ 1) * 128 + rc.row];
  } else {
    uint32_t r[32];
    const uint32_t taddr = tmem_base + kColO1;
    UNIVS_TMEM_LD_X32(taddr, r);
    tmem_wait_ld();
#pragma unroll
    for (int e = 0; e < 16; ++e) {   // lanes l and l^16 hold the two key-range partial sums of the same row
      const float a = __uint_as_float(r[e]), b = __uint_as_float(r[16 + e]);
      const float a2 = a + __shfl_xor_sync(0xffffffffu, a, 16), b2 = b + __shfl_xor_sync(0xffffffffu, b, 16);
      o[e] = rc.half ? b2 : a2;      // lanes 0-15 store dims 0-15, lanes 16-31 dims 16-31
    }
    total = sum + __shfl_xor_sync(0xffffffffu, sum, 16);
  }
  fence_before();
  __syncwarp();
  if (rc.lane == 0) mbar_arrive(&bars[O_FREE + TB]);

  const int src = source_token(g, un, rc.row);     // window_reverse + roll(+shift) + crop: pad rows are dropped
  if (src < 0) return;
  const float inv = 1.f / total;
  const int c = un.head * 32 + rc.half * 16;
  if (out != nullptr) {
    float4* dst = reinterpret_cast<float4*>(out + (size_t)src * g.C + c);
#pragma unroll
    for (int e = 0; e < 4; ++e) dst[e] = make_float4(o[4 * e] * inv, o[4 * e + 1] * inv, o[4 * e + 2] * inv, o[4 * e + 3] * inv);
  }
  if (out16 != nullptr) {
    // fp16x3 GEMM operand (single K-chunk, C <= 1536): [lo*2^11 (C) | hi*2^-11 (C) | hi (C)]
    uint32_t lo2[8], hs2[8], hi2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float a = o[2 * e] * inv, b = o[2 * e + 1] * inv;
      const __half2 h = __floats2half2_rn(a, b);
      const float2 f = __half22float2(h);
      const __half2 l = __floats2half2_rn((a - f.x) * 2048.f, (b - f.y) * 2048.f);
      const __half2 s = __floats2half2_rn(f.x * (1.f / 2048.f), f.y * (1.f / 2048.f));
      hi2[e] = *reinterpret_cast<const uint32_t*>(&h);
      lo2[e] = *reinterpret_cast<const uint32_t*>(&l);
      hs2[e] = *reinterpret_cast<const uint32_t*>(&s);
    }
    __half* rowp = out16 + (size_t)src * (3 * (size_t)g.C) + c;
    uint4* d0 = reinterpret_cast<uint4*>(rowp);
    uint4* d1 = reinterpret_cast<uint4*>(rowp + g.C);
    uint4* d2 = reinterpret_cast<uint4*>(rowp + 2 * g.C);
    d0[0] = make_uint4(lo2[0], lo2[1], lo2[2], lo2[3]);
    d0[1] = make_uint4(lo2[4], lo2[5], lo2[6], lo2[7]);
    d1[0] = make_uint4(hs2[0], hs2[1], hs2[2], hs2[3]);
    d1[1] = make_uint4(hs2[4], hs2[5], hs2[6], hs2[7]);
    d2[0] = make_uint4(hi2[0], hi2[1], hi2[2], hi2[3]);
    d2[1] = make_uint4(hi2[4], hi2[5], hi2[6], hi2[7]);
  }
}

template <bool TAIL>
__device__ void softmax_loop(unsigned char* smem, uint64_t* bars, uint32_t tmem_base, const RowCtx rc, const Geo g, int count,
                             float* __restrict__ out, __half* __restrict__ out16, float* __restrict__ dbg) {
  if (count <= 0) return;
  long long u = blockIdx.x;
  Unit cur = decode(u, g);
  float sum_cur = softmax_unit<TAIL>(smem, bars, tmem_base, rc, g, cur, 0, u, dbg);
  for (int n = 0; n < count; ++n) {
    Unit nxt = cur;
    float sum_nxt = 0.f;
    if (n + 1 < count) {
      const long long u2 = u + gridDim.x;
      nxt = decode(u2, g);
      sum_nxt = softmax_unit<TAIL>(smem, bars, tmem_base, rc, g, nxt, n + 1, u2, dbg);
    }
    epilogue_unit<TAIL>(smem, bars, tmem_base, rc, g, cur, n, sum_cur, out, out16);
    cur = nxt;
    sum_cur = sum_nxt;
    u += gridDim.x;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
swin_window_attn_tc12_kernel(const float* __restrict__ qkv, const float* __restrict__ qkv_bias,
                             const float* __restrict__ table, const Geo g, long long units, float scale,
                             float* __restrict__ out, __half* __restrict__ out16, float* __restrict__ dbg) {
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
    mbar_init(&bars[S_FREE + 1], 1);
    mbar_init(&bars[P_FULL + 0], 256);    // every softmax thread arrives after its own P stores + proxy fence
    mbar_init(&bars[P_FULL + 1], 32);
    mbar_init(&bars[P_FREE + 0], 1);
    mbar_init(&bars[P_FREE + 1], 1);
    mbar_init(&bars[O_FULL + 0], 1);
    mbar_init(&bars[O_FULL + 1], 1);
    mbar_init(&bars[O_FREE + 0], 8);
    mbar_init(&bars[O_FREE + 1], 1);
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

  if (warp >= kLoaderWarp0) {
    loader_loop(smem, bars, qkv, qkv_bias, table, g, units, scale);
  } else if (warp == kMmaWarp) {
    mma_loop(smem, bars, tmem_base, count);
  } else if (warp < 8) {
    RowCtx rc;
    rc.quarter = warp & 3;
    rc.half = warp >> 2;
    rc.lane = lane;
    rc.row = rc.quarter * 32 + lane;
    rc.prow = rc.row;
    softmax_loop<false>(smem, bars, tmem_base, rc, g, count, out, out16, dbg);
  } else if (warp == 8) {
    RowCtx rc;
    rc.quarter = 0;
    rc.half = lane >> 4;
    rc.lane = lane;
    rc.row = 128 + (lane & 15);
    rc.prow = lane;
    softmax_loop<true>(smem, bars, tmem_base, rc, g, count, out, out16, dbg);
  }
  fence_before();
  __syncthreads();
  if (warp == kAllocWarp) {
    fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int launch(cudaStream_t st, const float* qkv, const float* bias, const float* table, const Geo& g, float* out,
                  __half* out16, float* dbg) {
  const long long units = (long long)g.B * g.nWh * g.nWw * g.nH;
  UNIVS_REQUIRE(units < (1ll << 31), "swin_window_attention_tc: too many (window, head) units");
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  cudaError_t e = cudaFuncSetAttribute(swin_window_attn_tc12_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) {
    set_error("swin_window_attention_tc: cudaFuncSetAttribute(%d): %s", kSmemBytes, cudaGetErrorString(e));
    return UNIVS_E_LAUNCH;
  }
  // one head per CTA: the grid is a multiple of the head count (units = windows * nH is one too), unit u has head u % nH
  long long grid_ll = g.nH <= num_sms ? (long long)(num_sms / g.nH) * g.nH : g.nH;
  if (grid_ll > units) grid_ll = units;
  const int grid = (int)grid_ll;
  const float scale = 0.17677669529663687f;   // 32^-0.5 (swin.py:96)
  swin_window_attn_tc12_kernel<<<grid, kThreads, kSmemBytes, st>>>(qkv, bias, table, g, units, scale, out, out16, dbg);
  return check_launch("swin_window_attention_tc");
}

}  // namespace wintc
}  // namespace univs

namespace univs {
namespace wintc2 {   // swin_window_attn_tc2.cu: the second version of the kernel (flags bit 0)
int launch(cudaStream_t st, const float* qkv, const float* bias, const float* table, int batch, int height, int width,
           int channels, int num_heads, int shift, float* out, __half* out16, float* dbg, int compact);
}
}  // namespace univs

using namespace univs;

extern "C" int univs_swin_window_attention_tc(void* stream, const float* qkv, const float* qkv_bias,
                                              const float* rel_bias_table, int batch, int height, int width, int channels,
                                              int num_heads, int window, int shift, int flags, float* out, void* out16,
                                              float* debug_scores) {
  UNIVS_REQUIRE(qkv && qkv_bias && rel_bias_table && (out || out16), "swin_window_attention_tc: null pointer");
  UNIVS_REQUIRE(window == wintc::kWS, "swin_window_attention_tc: only 12x12 windows (got %d); use univs_swin_window_attention_f32", window);
  UNIVS_REQUIRE(batch >= 0 && height > 0 && width > 0, "swin_window_attention_tc: bad sizes");
  UNIVS_REQUIRE(num_heads > 0 && channels == num_heads * 32, "swin_window_attention_tc: head_dim must be 32 (channels=%d heads=%d)",
                channels, num_heads);
  UNIVS_REQUIRE(shift >= 0 && shift < window, "swin_window_attention_tc: shift must be in [0, window)");
  UNIVS_REQUIRE(out16 == nullptr || channels <= 1536 || (flags & 2), "swin_window_attention_tc: the operand output needs channels <= 1536");
  UNIVS_REQUIRE((flags & ~3) == 0 && flags != 2,
                "swin_window_attention_tc: unknown flags %d (bit 0: second kernel version; bit 1, with bit 0: compact operand)", flags);
  if (batch == 0) return UNIVS_OK;
  if (flags & 1)
    return wintc2::launch((cudaStream_t)stream, qkv, qkv_bias, rel_bias_table, batch, height, width, channels, num_heads, shift, out,
                          reinterpret_cast<__half*>(out16), debug_scores, (flags >> 1) & 1);
  wintc::Geo g;
  g.B = batch; g.H = height; g.W = width; g.C = channels; g.nH = num_heads; g.shift = shift;
  g.Hp = (height + window - 1) / window * window;
  g.Wp = (width + window - 1) / window * window;
  g.nWh = g.Hp / window;
  g.nWw = g.Wp / window;
  return wintc::launch((cudaStream_t)stream, qkv, qkv_bias, rel_bias_table, g, out, reinterpret_cast<__half*>(out16), debug_scores);
}
