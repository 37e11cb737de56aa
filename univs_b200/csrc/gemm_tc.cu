// Dense layers of the path on the 5th-gen tensor cores with fused epilogues: the fp32-equivalent
// fp16 hi|lo product of the strict policy as ONE kernel -- TMA -> shared memory -> tcgen05.mma -> TMEM -> tcgen05.ld ->
// bias / GELU / ReLU / residual / operand emission -> coalesced global stores.
// Replaces, on the reference side, nn.Linear + the activation / residual that follows it:
//   swin.py:35-41 (Mlp fc1 -> GELU -> fc2), :138-169 (qkv / proj), :292-293 (residual), ms_deform_attn.py:98-120
//   (value / offset / output projections), transformer_layers.py (decoder linears, FFN).
//
//   Y[t, n] = act( alpha * sum_k X[t, k] * W[n, k] + bias[n] ) + addend[t, n]
//
// Operands are fp16 pairs: x = hi + lo' * 2^-11 with hi = fp16(x), lo' = fp16((x - hi) * 2^11) (the scale keeps lo out of
// fp16's subnormal range); the weight tensor is pre-scaled by a per-tensor power of two (alpha undoes it).  hi*hi products
// are exact in fp32, two correction products restore ~22 bits:
//   main = Wh Xh^T          corr = Wh Xl'^T + Wl' Xh^T          Y = alpha * (main + corr * 2^-11)
// (issued as two MMAs per k-step: Wh * [Xh ; Xl']^T with N = 256 into the adjacent {main, corr} columns, then Wl' * Xh^T)
// with main and corr in SEPARATE TMEM accumulators: the tensor core accumulates with truncation (tools/accum_probe.py), and
// the correction terms must not be added to an accumulator 2^11 times their size.  Three kind::f16 MMAs per 16-wide k-step
// read four operand tiles (the single-GEMM formulation of round 1 read six: [lo'|hi_s|hi] x [hi_s|lo'|hi]).
//
// Mapping onto UMMA (D = A * B^T, both operands K-major, 128-byte rows = 64 halfs, SWIZZLE_128B):
//   A (M side, 128 rows) = weight tile: 128 output channels   -> TMEM lane   = output channel n
//   B (N side, 128 rows) = activation tile: 128 tokens        -> TMEM column = token
// so that a 32x32b TMEM load gives the 32 lanes of a warp 32 CONSECUTIVE channels of one token: every global store of the
// epilogue is one full 128-byte line of the row-major [tokens, N] output (or 64 bytes of the fp16 operand output), and
// bias[n] is a per-thread scalar.
//
// Persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer (3-stage ring of {Wh, Wl', Xh, Xl'} = 64 KB),
// warp 1 = MMA issuer (one elected lane), warp 2 = TMEM allocator, warps 4-19 = epilogue (four per TMEM lane quarter, 32
// token columns each).  Two TMEM stages of {main, corr} (4 x 128 columns): the epilogue of tile i -- which for the GELU
// variant is as long as the main loop -- overlaps the MMAs of tile i+1.  Tiles are ordered channel-tile fastest, so
// concurrently running CTAs share activation tiles in L2; weights stay L2-resident.
//
// The operand tensors are addressed through four 2-D tensor maps (rows x K window with an arbitrary row pitch), which makes
// the kernel independent of the operand container: the round-1 [lo' | hi_s | hi] K-chunk format (hi_s simply is not read)
// and the compact [hi | lo'] format this kernel's own operand epilogue writes.  K beyond 1536 is sliced on the host
// (one launch per slice, `addend` = the partial result): bounds the main accumulation chain to 96 MMA steps.
#include "rowwise.cuh"
#include "tc05.cuh"

namespace univs {
namespace gemmtc {

using namespace tc;

constexpr int kBM = 128;                 // channels per tile (UMMA M)
constexpr int kBT = 128;                 // tokens per tile (UMMA N)
constexpr int kBK = 64;                  // halfs per operand row in shared memory (one SWIZZLE_128B span)
constexpr int kTileBytes = 128 * 128;    // one operand tile: 128 rows x 128 bytes
// one CTA per tile: stage = Wh | Wl | Xh | Xl' tiles (64 KB), 3 stages.  CTA pair (cta_group::2): stage = Wh | Wl tiles of this
// CTA's 128 channels + this CTA's 64-token halves of Xh and Xl' (48 KB), 4 stages.  Both rings are 192 KB.
template <bool PAIR>
struct Ring {
  static constexpr int kStages = PAIR ? 4 : 3;
  static constexpr int kStageBytes = PAIR ? 3 * kTileBytes : 4 * kTileBytes;
  static constexpr int kOffBars = kStages * kStageBytes;
  static constexpr int kNumBars = 2 * kStages + 4;
  static constexpr int kSmemBytes = kOffBars + kNumBars * 8 + 16;
};
constexpr int kEpiWarps = 16;            // four per TMEM lane quarter, 32 token columns each
constexpr int kThreads = 128 + 32 * kEpiWarps;   // 4 control warps + the epilogue warps
constexpr int kMaxTaps = 9;
constexpr int kCh = 16;                  // token columns per epilogue chunk (registers: 2 x 16 accumulator values + 16 results)
constexpr int kSmemBytes = Ring<true>::kSmemBytes > Ring<false>::kSmemBytes ? Ring<true>::kSmemBytes : Ring<false>::kSmemBytes;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
constexpr uint32_t kColsPerStage = 2 * kBT;   // main [0,128) corr [128,256)

struct Args {
  int M, N, K;                 // tokens, output channels, reduction length
  float alpha;
  const float* bias;           // [N] or null
  const float* addend;         // fp32 [M, ldadd] or null (may alias out)
  long long ldadd;
  float* out;                  // fp32 [M, ldo] or null
  long long ldo;
  __half* out16;               // fp16 operand output or null: hi at column n, lo*2^11 at column lo_off16 + n
  long long ld16, lo_off16;
  int act;                     // 0 none, 1 GELU (erf), 2 ReLU
  // shifted-row taps (kxk convolution as ONE accumulation): K = taps * kpt k-blocks; k-block kb multiplies the weight columns
  // [64 kb, 64 kb + 64) with the activation rows shifted by tap_off[kb / kpt], columns [64 (kb % kpt), ..).  taps == 1: a plain GEMM.
  int taps, kpt;
  int tap_off[kMaxTaps];
};

// erf(z) in one branch-free form:  erf(|z|) = 1 - exp(-|z| * q(|z|)),  q = degree-7 polynomial fitted to -ln(erfc(t)) / t on
// [0, 4] (weighted for the absolute error of erf; beyond 4 erf rounds to 1).  Max absolute error 1.3e-7 (2 ulp of 1, the
// same class as erff) at about a third of erff's instruction count -- erff evaluates two polynomial branches and merges them
// with selects, ~35 instructions per element in an epilogue whose issue slots bound the kernel (ncu, round 2).  The result
// for negative z is built as 1 + erf = exp(...) directly, so GELU keeps its relative accuracy in the negative tail.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = x * 0.70710678118654752440f;
  const float t = fminf(fabsf(z), 4.0f);
  float q = 3.1440224120160565e-05f;
  q = fmaf(q, t, -0.00030880147824063897f);
  q = fmaf(q, t, 0.001032395986840129f);
  q = fmaf(q, t, 0.0005369476275518537f);
  q = fmaf(q, t, -0.01958397589623928f);
  q = fmaf(q, t, 0.10291962325572968f);
  q = fmaf(q, t, 0.636597752571106f);
  q = fmaf(q, t, 1.128380298614502f);
  const float e = ex2_approx(q * t * -1.4426950408889634f);      // exp(-t q(t)) = erfc(t)
  // 1 + erf(z):  z >= 0 -> 2 - e ;  z < 0 -> e
  const float one_plus_erf = z >= 0.f ? 2.f - e : e;
  return 0.5f * x * one_plus_erf;
}

template <int ACT>
__device__ __forceinline__ float activate(float v) {
  if (ACT == 1) return gelu_erf(v);
  if (ACT == 2) return fmaxf(v, 0.f);
  return v;
}

__device__ __forceinline__ uint32_t pack_sat_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}

// One 32-token chunk of the epilogue for this thread's channel n: v[j] = act(alpha * (main + corr * 2^-11) + bias) (+ addend),
// written as fp32 (one 128-byte line per token and warp) and / or as the fp16 [hi | lo'] operand.  The variant is fixed at
// compile time: the first version took `act` and the nullable pointers at run time, and its per-element branches and
// argument reloads made the epilogue several times longer than the main loop (ncu: instruction-fetch and dependency stalls).
// FULL = all 32 tokens of the chunk valid (every chunk but those of the last token tile): the only predicate left is the
// per-lane channel validity `n_ok`, one branch around each load / store loop.  (Per-element predicates -- the first form
// of the partial-channel-tile path -- cost 5x: N = 192 and 576 have a partial last channel tile in EVERY token tile.)
template <int ACT, bool OUT32, bool OUT16, bool ADD, bool FULL>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&rm)[kCh], const uint32_t (&rc)[kCh], const Args& a, float bias, int n,
                                               bool n_ok, int t0, int nt, int lane) {
  float v[kCh];
#pragma unroll
  for (int j = 0; j < kCh; ++j) {
    const float s = fmaf(__uint_as_float(rc[j]), 1.f / 2048.f, __uint_as_float(rm[j]));
    v[j] = activate<ACT>(fmaf(s, a.alpha, bias));
  }
  if (ADD && n_ok) {
    const float* ap = a.addend + (size_t)t0 * a.ldadd + n;
#pragma unroll
    for (int j = 0; j < kCh; ++j) {
      if (FULL || j < nt) v[j] += __ldg(ap);
      ap += a.ldadd;
    }
  }
  if (OUT32 && n_ok) {
    float* op = a.out + (size_t)t0 * a.ldo + n;
#pragma unroll
    for (int j = 0; j < kCh; ++j) {
      if (FULL || j < nt) *op = v[j];
      op += a.ldo;
    }
  }
  if (OUT16) {
    // lane pairs exchange so that a lane holds TWO adjacent channels of one token (even lane: token j, odd lane: token
    // j + 1) and stores them as one half2: half as many store instructions as 2-byte stores per element
    const bool odd = lane & 1;
    const int ne = n & ~1;                                   // first channel of the pair
    __half* hp = a.out16 + (size_t)(t0 + (odd ? 1 : 0)) * a.ld16 + ne;
    // 4-byte stores need an even channel count, lo offset and row pitch (true for every layer of the path); otherwise the
    // two channels are stored one by one.  With an even N a pair is valid or invalid as a whole.
    const bool even = ((a.N | a.lo_off16 | a.ld16) & 1) == 0;
    const bool pair_ok = ne + 1 < a.N;
#pragma unroll
    for (int j = 0; j < kCh; j += 2) {
      const float mine = odd ? v[j + 1] : v[j];              // my channel, my token
      const float other = __shfl_xor_sync(0xffffffffu, odd ? v[j] : v[j + 1], 1);   // partner's channel, my token
      const float c0 = odd ? other : mine, c1 = odd ? mine : other;   // channels ne, ne + 1
      const __half2 h = __floats2half2_rn(fminf(fmaxf(c0, -65504.f), 65504.f), fminf(fmaxf(c1, -65504.f), 65504.f));
      const float2 hf = __half22float2(h);
      const uint32_t lo = pack_sat_half2((c0 - hf.x) * 2048.f, (c1 - hf.y) * 2048.f);
      const int tok = j + (odd ? 1 : 0);
      if (even) {
        if (pair_ok && (FULL || tok < nt)) {
          *reinterpret_cast<__half2*>(hp) = h;
          *reinterpret_cast<uint32_t*>(hp + a.lo_off16) = lo;
        }
      } else if (FULL || tok < nt) {
        if (ne < a.N) {
          hp[0] = __low2half(h);
          hp[a.lo_off16] = __ushort_as_half((unsigned short)(lo & 0xffffu));
        }
        if (ne + 1 < a.N) {
          hp[1] = __high2half(h);
          hp[a.lo_off16 + 1] = __ushort_as_half((unsigned short)(lo >> 16));
        }
      }
      hp += 2 * a.ld16;
    }
  }
}

// PAIR (round 2): two CTAs of a cluster work as ONE tensor-core unit (tcgen05.mma.cta_group::2, M = 256) on a tile of 256
// channels x 128 tokens.  Why: the one-CTA kernel runs the K >= 768 layers at 1.15-1.25 PFLOP/s whatever their shape.  A
// cluster variant that only MULTICAST the shared activation tile (48 instead of 64 KB per k-block out of L2, same shared-memory
// traffic) gained nothing (profiles/r2_gemm_mc.txt), so L2 is not the limit; shared-memory bandwidth is: per k-block a CTA's
// MMAs read 80 KB of operands while TMA writes 64 KB, ~187 B/clk at the tensor peak against the ~150 B/clk the SM delivers.
// In a pair each CTA holds its 128 channels of Wh / Wl and HALF of the token tile (the hardware feeds both tensor cores from
// both shared memories), so per k-block it reads 72 KB and TMA writes 48 KB, and the ring is 4 stages deep.  Protocol (as in
// CUTLASS' 2-SM kernels): both CTAs allocate TMEM and load with TMA, crediting the LEADER's `full` barrier; the leader (rank 0)
// issues the MMAs and multicasts its commits to both CTAs' `empty` / `tfull` barriers; each CTA runs the epilogue of its 128
// channels from its own TMEM and releases the accumulator stage on the leader's `tempty` barrier.  With an odd number of
// channel tiles the second CTA of the last pair works on an all-zero (out-of-range) weight tile and stores nothing.
template <int ACT, bool OUT32, bool OUT16, bool ADD, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
gemm_f16x3_tc_kernel(const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl,
                     const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                     const __grid_constant__ Args a) {   // __grid_constant__: tap_off[] is indexed at run time -- from the constant bank, not a local copy
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int kStages = Ring<PAIR>::kStages, kStageBytes = Ring<PAIR>::kStageBytes, kNumBars = Ring<PAIR>::kNumBars;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Ring<PAIR>::kOffBars);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ct = (a.N + kBM - 1) / kBM;
  const int n_tt = (a.M + kBT - 1) / kBT;
  const int kblocks = (a.K + kBK - 1) / kBK;
  // work items: tiles (one CTA) or pairs of channel tiles of one token tile (cluster of two), channel index fastest
  const int rank = PAIR ? (int)cluster_rank() : 0;
  const int n_cu = PAIR ? (n_ct + 1) / 2 : n_ct;           // channel units per token tile
  const int num_items = n_cu * n_tt;
  const int item0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], PAIR ? 2 * kEpiWarps : kEpiWarps);   // one arrive per epilogue warp (pair: of both CTAs, on the leader's)
    }
    mbar_init_fence();
  } else if (warp == 2) {
    if (PAIR) tmem_alloc_2sm(tmem_slot, 512);
    else tmem_alloc(tmem_slot, 512);
  }
  fence_before();
  __syncthreads();
  fence_after();
  if (PAIR) cluster_sync_all();         // the peer's barriers exist before anything is signalled on them
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = item0; item < num_items; item += item_step) {
        const int tt = item / n_cu, cu = item - tt * n_cu;
        const int ct = PAIR ? 2 * cu + rank : cu;          // pair, odd n_ct: ct == n_ct reads an out-of-range (zero-filled) weight tile
        for (int kb = 0, tap = 0, kin = 0; kb < kblocks; ++kb) {      // k-block kb = k-block kin of tap `tap`
          const int xk = kin * kBK;                          // activation column of this k-block
          const int xr = tt * kBT + a.tap_off[tap];          // first activation row
          if (++kin == a.kpt) { kin = 0; ++tap; }
          mbar_wait(&empty[stage], phase ^ 1, 10 + stage);
          const uint32_t s0 = smem_u32(smem + (size_t)stage * kStageBytes);
          // a box always counts in full (out-of-range parts are zero-filled)
          if (PAIR) {   // both CTAs' boxes are credited to the leader's barrier (the leader issues the MMAs of the pair)
            if (rank == 0) mbar_expect_tx(&full[stage], (uint32_t)(2 * kStageBytes));
            tma_load_3d_2sm(s0, &map_wh, &full[stage], kb * kBK, ct * kBM, 0);
            tma_load_3d_2sm(s0 + kTileBytes, &map_wl, &full[stage], kb * kBK, ct * kBM, 0);
            // my half of the token tile (map_x*: 64-row boxes): B rows (N/2) * rank .. of every MMA of the pair
            tma_load_3d_2sm(s0 + 2 * kTileBytes, &map_xh, &full[stage], xk, xr + rank * (kBT / 2), 0);
            tma_load_3d_2sm(s0 + 2 * kTileBytes + kTileBytes / 2, &map_xl, &full[stage], xk, xr + rank * (kBT / 2), 0);
          } else {
            mbar_expect_tx(&full[stage], (uint32_t)kStageBytes);
            tma_load_3d(s0, &map_wh, &full[stage], kb * kBK, ct * kBM, 0);
            tma_load_3d(s0 + kTileBytes, &map_wl, &full[stage], kb * kBK, ct * kBM, 0);
            tma_load_3d(s0 + 2 * kTileBytes, &map_xh, &full[stage], xk, xr, 0);
            tma_load_3d(s0 + 3 * kTileBytes, &map_xl, &full[stage], xk, xr, 0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (pair: the leader CTA only) =====
    // Two MMA instructions per k-step carry the three products: the Xh and Xl' tiles are adjacent in shared memory (256 rows of
    // one K-major operand), and {main, corr} are adjacent in TMEM, so ONE N = 256 MMA computes Wh * [Xh ; Xl']^T -- the main
    // term and the first correction term -- reading the Wh tile once; the N = 128 MMA adds Wl' * Xh^T to corr.  Same products
    // and accumulation order as three N = 128 MMAs, with 20 KB instead of 24 KB of shared-memory operand reads per k-step
    // (the one-CTA 128 x 128 tile is bound by shared-memory bandwidth, not by the tensor pipe).
    constexpr uint32_t idesc = make_idesc_f16(kBM, kBT, false, false);
    constexpr uint32_t idesc2 = make_idesc_f16(kBM, 2 * kBT, false, false);
    constexpr uint32_t idesc_pair = make_idesc_f16(2 * kBM, kBT, false, false);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = item0; item < num_items; item += item_step) {
      mbar_wait(&tempty[acc], acc_phase ^ 1, 20 + acc);
      fence_after();
      const uint32_t d_main = tmem_base + (uint32_t)acc * kColsPerStage;
      const uint32_t d_corr = d_main + (uint32_t)kBT;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full[stage], phase, 30 + stage);
        fence_after();
        if (elect_one()) {
          const uint32_t s0 = smem_u32(smem + (size_t)stage * kStageBytes);
          const uint64_t wh = make_desc<128>(s0), wl = make_desc<128>(s0 + kTileBytes);
          const uint64_t xh = make_desc<128>(s0 + 2 * kTileBytes);     // rows 0-127: Xh tile, rows 128-255: the Xl' tile behind it
          if (PAIR) {
            // M = 256 over the pair, N = 128 tokens (64 rows of B from each CTA): the same three products per k-step in the
            // same order as the one-CTA kernel -- main (+)= Wh Xh^T; corr (+)= Wh Xl'^T; corr += Wl' Xh^T
            const uint64_t xl = make_desc<128>(s0 + 2 * kTileBytes + kTileBytes / 2);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint32_t first = (kb | k) != 0 ? 1u : 0u;
              umma_f16_2sm(d_main, wh + (uint64_t)(2 * k), xh + (uint64_t)(2 * k), idesc_pair, first);
              umma_f16_2sm(d_corr, wh + (uint64_t)(2 * k), xl + (uint64_t)(2 * k), idesc_pair, first);
              umma_f16_2sm(d_corr, wl + (uint64_t)(2 * k), xh + (uint64_t)(2 * k), idesc_pair, 1u);
            }
            umma_commit_multicast_2sm(&empty[stage], (uint16_t)0x3);   // frees the stage in both CTAs when these MMAs retire
            if (kb == kblocks - 1) umma_commit_multicast_2sm(&tfull[acc], (uint16_t)0x3);
          } else {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {   // one k-step = 16 halfs = 32 bytes inside the swizzle row: +2 in the (addr >> 4) field
              const uint32_t first = (kb | k) != 0 ? 1u : 0u;
              umma_f16(d_main, wh + (uint64_t)(2 * k), xh + (uint64_t)(2 * k), idesc2, first);  // [main | corr] (+)= Wh * [Xh ; Xl']^T
              umma_f16(d_corr, wl + (uint64_t)(2 * k), xh + (uint64_t)(2 * k), idesc, 1u);      // corr += Wl' * Xh^T
            }
            umma_commit(&empty[stage]);                       // frees the stage when these MMAs retire
            if (kb == kblocks - 1) umma_commit(&tfull[acc]);  // both accumulators complete -> epilogue
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global.  16 warps: four per TMEM lane quarter, 32 token columns each, in two
    // chunks of 16.  (Eight warps with 64 columns each left the epilogue latency-bound: two warps per scheduler walking
    // dependent convert / shuffle / store chains; for K = 192 layers the epilogue, not the main loop, set the tile time.) =====
    const int ew = warp - 4;
    const int wq = ew & 3;       // TMEM lane quarter this warp may access (== warp % 4)
    const int part = ew >> 2;    // token columns [32 * part, 32 * part + 32)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = item0; item < num_items; item += item_step) {
      const int tt = item / n_cu, cu = item - tt * n_cu;
      const int ct = PAIR ? 2 * cu + rank : cu;             // pair, odd n_ct: ct == n_ct -> n >= N, no stores
      const int n = ct * kBM + wq * 32 + lane;              // this thread's output channel
      const bool n_ok = n < a.N;
      const float bias = (a.bias != nullptr && n_ok) ? __ldg(a.bias + n) : 0.f;
      mbar_wait(&tfull[acc], acc_phase, 40 + acc);
      fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)acc * kColsPerStage + (uint32_t)(part * 32);
      uint32_t rm[2][kCh], rc[2][kCh];
      UNIVS_TMEM_LD_X16(taddr, rm[0]);
      UNIVS_TMEM_LD_X16(taddr + (uint32_t)kBT, rc[0]);
      UNIVS_TMEM_LD_X16(taddr + (uint32_t)kCh, rm[1]);
      UNIVS_TMEM_LD_X16(taddr + (uint32_t)(kBT + kCh), rc[1]);
      tmem_wait_ld();
      fence_before();             // this warp's share of the accumulator stage is in registers: release it to the MMA lane
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_remote(&tempty[acc], 0u);     // the leader's MMA lane waits for both CTAs' epilogues
        else mbar_arrive(&tempty[acc]);
      }
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        const int t0 = tt * kBT + part * 32 + cb * kCh;     // first token of this chunk
        const int nt = a.M - t0 < kCh ? a.M - t0 : kCh;     // valid tokens (<= 0: none)
        if (nt == kCh) {
          epilogue_chunk<ACT, OUT32, OUT16, ADD, true>(rm[cb], rc[cb], a, bias, n, n_ok, t0, kCh, lane);
        } else if (nt > 0) {
          epilogue_chunk<ACT, OUT32, OUT16, ADD, false>(rm[cb], rc[cb], a, bias, n, n_ok, t0, nt, lane);
        }
        __syncwarp();           // the shuffles / TMEM loads that follow are warp-collective
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();         // no CTA leaves (or frees TMEM) while the pair's MMAs / barrier signals may still touch it
  if (warp == 2) {
    fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encoder() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp16 window [rows][K] of a row-major tensor with `pitch` halfs per row; box [64 halfs][128 rows], SWIZZLE_128B; rows and
// columns beyond the window read as zeros
static int make_map(CUtensorMap* m, const __half* base, long long rows, int K, long long pitch, int box_rows = 128) {
  EncodeTiledFn enc = encoder();
  if (!enc) { set_error("gemm_f16x3_tc: cuTensorMapEncodeTiled entry point unavailable"); return UNIVS_E_LAUNCH; }
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 1};
  cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)pitch * 2 * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gemm_f16x3_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return UNIVS_E_LAUNCH; }
  return 0;
}

constexpr int kPairDefault = 2;     // 0: off unless UNIVS_GEMM_PAIR=1; 2: heuristic (see the launcher)
constexpr int kPairMinK = 768;

// cluster launch: persistent pairs of CTAs, as many as can be co-resident (a GPC with an odd number of free SMs leaves one out)
template <typename Kern>
static int launch_pair(Kern kern, int num_sms, long long pairs, cudaStream_t st, const CUtensorMap& mwh, const CUtensorMap& mwl,
                     const CUtensorMap& mxh, const CUtensorMap& mxl, const Args& a) {
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int resident = 0;       // the same for every epilogue variant: one CTA per SM, same shared memory
  if (!resident) {
    cfg.gridDim = dim3((unsigned)(2 * (num_sms / 2)));
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess || n < 1) {
      set_error("gemm_f16x3_tc: cudaOccupancyMaxActiveClusters: %s (%d clusters)", cudaGetErrorString(e), n);
      (void)cudaGetLastError();
      return UNIVS_E_LAUNCH;
    }
    resident = n < num_sms / 2 ? n : num_sms / 2;
  }
  const long long clusters = resident < pairs ? resident : pairs;
  cfg.gridDim = dim3((unsigned)(2 * clusters));
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mwh, mwl, mxh, mxl, a);
  if (e != cudaSuccess) { set_error("gemm_f16x3_tc: cluster launch: %s", cudaGetErrorString(e)); return UNIVS_E_LAUNCH; }
  return check_launch("gemm_f16x3_tc");
}

}  // namespace gemmtc
}  // namespace univs

using namespace univs;

extern "C" int univs_gemm_f16x3_tc(void* stream, const void* x16, int64_t ldx, int64_t x_hi_off, int64_t x_lo_off,
                                   const void* w16, int64_t ldw, int64_t w_hi_off, int64_t w_lo_off, int64_t tokens,
                                   int channels, int k, float alpha, const float* bias, const float* addend, int64_t ldadd,
                                   float* out, int64_t ldo, void* out16, int64_t ld16, int64_t out16_lo_off, int activation) {
  const int64_t zero = 0;
  return univs_gemm_f16x3_tc_taps(stream, x16, ldx, x_hi_off, x_lo_off, tokens, w16, ldw, w_hi_off, w_lo_off, tokens, channels, k, 1,
                                  &zero, alpha, bias, addend, ldadd, out, ldo, out16, ld16, out16_lo_off, activation);
}

extern "C" int univs_gemm_f16x3_tc_taps(void* stream, const void* x16, int64_t ldx, int64_t x_hi_off, int64_t x_lo_off,
                                        int64_t x_rows, const void* w16, int64_t ldw, int64_t w_hi_off, int64_t w_lo_off,
                                        int64_t tokens, int channels, int k_tap, int taps, const int64_t* tap_row_offsets,
                                        float alpha, const float* bias, const float* addend, int64_t ldadd, float* out,
                                        int64_t ldo, void* out16, int64_t ld16, int64_t out16_lo_off, int activation) {
  using namespace gemmtc;
  UNIVS_REQUIRE(x16 && w16 && (out || out16), "gemm_f16x3_tc: null pointer");
  UNIVS_REQUIRE(tokens >= 0 && tokens < (1ll << 31) - 256 && channels > 0 && k_tap > 0, "gemm_f16x3_tc: bad sizes");
  UNIVS_REQUIRE(taps >= 1 && taps <= kMaxTaps && tap_row_offsets != nullptr, "gemm_f16x3_tc: taps must be in [1, %d]", kMaxTaps);
  UNIVS_REQUIRE(taps == 1 || k_tap % kBK == 0, "gemm_f16x3_tc: with taps the per-tap K must be a multiple of %d", kBK);
  UNIVS_REQUIRE(x_rows >= 0 && x_rows < (1ll << 31) - 256, "gemm_f16x3_tc: bad activation row count");
  for (int t = 0; t < taps; ++t)
    UNIVS_REQUIRE(tap_row_offsets[t] >= 0 && tap_row_offsets[t] < (1ll << 31) - 256, "gemm_f16x3_tc: bad tap row offset");
  const int k = k_tap * taps;                 // accumulation length = weight columns
  UNIVS_REQUIRE(k % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0 && x_hi_off % 8 == 0 && x_lo_off % 8 == 0 && w_hi_off % 8 == 0 &&
                    w_lo_off % 8 == 0,
                "gemm_f16x3_tc: K, row pitches and block offsets must be multiples of 8 halfs (16-byte TMA alignment)");
  UNIVS_REQUIRE(((uintptr_t)x16 | (uintptr_t)w16) % 16 == 0, "gemm_f16x3_tc: operands must be 16-byte aligned");
  UNIVS_REQUIRE(x_hi_off + k_tap <= ldx && x_lo_off + k_tap <= ldx && w_hi_off + k <= ldw && w_lo_off + k <= ldw,
                "gemm_f16x3_tc: operand blocks exceed the row pitch");
  UNIVS_REQUIRE(activation >= 0 && activation <= 2, "gemm_f16x3_tc: activation must be 0 (none), 1 (GELU) or 2 (ReLU)");
  UNIVS_REQUIRE(out == nullptr || ldo >= channels, "gemm_f16x3_tc: ldo < channels");
  UNIVS_REQUIRE(addend == nullptr || ldadd >= channels, "gemm_f16x3_tc: ldadd < channels");
  UNIVS_REQUIRE(out16 == nullptr || (ld16 >= channels && out16_lo_off >= 0 && out16_lo_off + channels <= ld16),
                "gemm_f16x3_tc: operand output blocks exceed its row pitch");
  if (tokens == 0) return UNIVS_OK;
  const __half* x = reinterpret_cast<const __half*>(x16);
  const __half* w = reinterpret_cast<const __half*>(w16);
  CUtensorMap mwh, mwl, mxh, mxl;
  int rc;
  if ((rc = make_map(&mwh, w + w_hi_off, channels, k, ldw))) return rc;
  if ((rc = make_map(&mwl, w + w_lo_off, channels, k, ldw))) return rc;
  // CTA-pair variant (see the kernel): UNIVS_GEMM_PAIR = 0 never, 1 whenever there are two channel tiles, unset = where it
  // pays (deep K, enough pairs to fill the SMs).  Read per call: the tests toggle it.
  const char* pair_env = getenv("UNIVS_GEMM_PAIR");
  const int pair_mode = pair_env ? (atoi(pair_env) ? 1 : 0) : kPairDefault;
  const int n_ct_h = (channels + kBM - 1) / kBM;
  const long long n_tt_h = (tokens + kBT - 1) / kBT;
  // measured (tools/gemm_tc_pair.py, profiles/r2_gemm_pair.txt): the pair wins where the main loop dominates (K >= 768: stage 3 / 4
  // of Swin-L, 5-15 %) and loses on shallow K and on an odd number of channel tiles (N = 384: a quarter of the pair slots idle)
  const bool mc = pair_mode == 1 ? n_ct_h >= 2
                                 : (pair_mode == 2 && k >= kPairMinK && n_ct_h >= 2 && (n_ct_h % 2 == 0 || n_ct_h >= 9) &&
                                    (long long)((n_ct_h + 1) / 2) * n_tt_h >= 74);
  // activation rows beyond x_rows (and columns beyond the per-tap K) read as zeros
  if ((rc = make_map(&mxh, x + x_hi_off, x_rows, k_tap, ldx, mc ? kBT / 2 : kBT))) return rc;
  if ((rc = make_map(&mxl, x + x_lo_off, x_rows, k_tap, ldx, mc ? kBT / 2 : kBT))) return rc;
  Args a;
  a.M = (int)tokens; a.N = channels; a.K = k; a.alpha = alpha; a.bias = bias; a.addend = addend; a.ldadd = ldadd;
  a.out = out; a.ldo = ldo; a.out16 = reinterpret_cast<__half*>(out16); a.ld16 = ld16; a.lo_off16 = out16_lo_off;
  a.act = activation;
  a.taps = taps;
  a.kpt = (k_tap + kBK - 1) / kBK;
  for (int t = 0; t < kMaxTaps; ++t) a.tap_off[t] = t < taps ? (int)tap_row_offsets[t] : 0;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const long long tiles = (long long)n_ct_h * n_tt_h;
  const int grid = (int)(tiles < num_sms ? tiles : num_sms);
  const long long pairs = (long long)((n_ct_h + 1) / 2) * n_tt_h;
  cudaStream_t st = (cudaStream_t)stream;
  const bool o32 = out != nullptr, o16 = out16 != nullptr, add = addend != nullptr;
  // the epilogue variants the path uses are compiled; anything else is a caller error
#define UNIVS_GEMM_CASE(ACT, O32, O16, ADDF)                                                                                   \
  if (activation == ACT && o32 == O32 && o16 == O16 && add == ADDF) {                                                          \
    static bool attr_set = false;                                                                                              \
    if (!attr_set) {                                                                                                           \
      cudaError_t e = cudaFuncSetAttribute(gemm_f16x3_tc_kernel<ACT, O32, O16, ADDF, false>,                                   \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);                           \
      if (e == cudaSuccess)                                                                                                    \
        e = cudaFuncSetAttribute(gemm_f16x3_tc_kernel<ACT, O32, O16, ADDF, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                 kSmemBytes);                                                                                  \
      if (e != cudaSuccess) { set_error("gemm_f16x3_tc: cudaFuncSetAttribute(%d): %s", kSmemBytes, cudaGetErrorString(e)); return UNIVS_E_LAUNCH; } \
      attr_set = true;                                                                                                         \
    }                                                                                                                          \
    if (mc) return launch_pair(gemm_f16x3_tc_kernel<ACT, O32, O16, ADDF, true>, num_sms, pairs, st, mwh, mwl, mxh, mxl, a);        \
    gemm_f16x3_tc_kernel<ACT, O32, O16, ADDF, false><<<grid, kThreads, kSmemBytes, st>>>(mwh, mwl, mxh, mxl, a);                  \
    return check_launch("gemm_f16x3_tc");                                                                                      \
  }
  UNIVS_GEMM_CASE(0, true, false, false)      // plain dense layer
  UNIVS_GEMM_CASE(0, true, false, true)       // + residual / previous K-slice
  UNIVS_GEMM_CASE(1, false, true, false)      // GELU -> next operand (Swin MLP fc1)
  UNIVS_GEMM_CASE(2, false, true, false)      // ReLU -> next operand (FFN linear1)
  UNIVS_GEMM_CASE(0, false, true, false)      // operand only
  UNIVS_GEMM_CASE(0, true, true, false)       // fp32 + operand
  UNIVS_GEMM_CASE(0, true, true, true)
  UNIVS_GEMM_CASE(1, true, true, false)
  UNIVS_GEMM_CASE(1, true, true, true)
  UNIVS_GEMM_CASE(2, true, true, false)
  UNIVS_GEMM_CASE(2, true, true, true)
  UNIVS_GEMM_CASE(1, true, false, false)
  UNIVS_GEMM_CASE(2, true, false, false)
#undef UNIVS_GEMM_CASE
  set_error("gemm_f16x3_tc: epilogue variant (activation %d, f32 %d, operand %d, addend %d) is not instantiated", activation,
            (int)o32, (int)o16, (int)add);
  return UNIVS_E_BADARG;
}
