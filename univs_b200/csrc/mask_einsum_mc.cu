// Mask-logit einsum, cluster variant of mask_einsum_tc.cu<F16X3> (opt-in UNIVS_EINSUM_MC; written after the round-1 GPU
// budget was spent: never run on hardware).  Same contraction and operand format:
//   out[q, t, p] = sum_c E[t, q, c] * F[t, p, c],   E [T,Q,2C] / F [T,HW,2C] fp16 [hi | lo],   out [Q,T,HW] fp32
//
// What it changes.  The validated kernel re-streams the whole E tile (npad x 2C halfs = 213 KB at Q = 200) from L2 for
// EVERY 128-pixel tile: 791 MB of L2->SM traffic per launch next to 269 MB of F (ncu, profiles/), which is what holds it
// at 0.73 of the HBM roofline with the tensor pipe at 48 %.  Here two CTAs form a cluster and split the QUERIES:
//   * CTA r keeps ITS half of E (112 / 96 queries at Q = 200, all 2C columns, 112 KB) resident in shared memory for a
//     whole frame -- E is read once per (cluster, frame) instead of once per tile;
//   * both CTAs need the same F tile: each issues the TMA load of one half of a stage (CTA 0 the hi chunk, CTA 1 the lo
//     chunk) with .multicast::cluster, so a stage reaches both SMs with one L2 read;
//   * each CTA runs the 1-SM MMA of the validated kernel (M = 128 pixels, N = its queries, three kind::f16 MMAs per
//     k-step: lo*hi + hi*lo + hi*hi) into its own TMEM and stores its own query planes.
// A stage may be refilled only when BOTH CTAs have consumed it (the refill writes into both): the MMA warp of each CTA
// commits with .multicast::cluster to the `empty` barrier of both CTAs (arrival count 2).
// Tiles are handed out as contiguous ranges per cluster (31-32 tiles at the north-star shape), so a cluster crosses at
// most one or two frame boundaries and reloads E only there.
// Warp roles as in mask_einsum_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11 epilogue.
#include <cuda.h>

#include "tc05.cuh"

namespace univs {
namespace einmc {

using namespace tc;

constexpr int kTileM = 128;
constexpr int kThreads = 384;
constexpr int kStages = 6;
constexpr int kRow = 64;                          // operand row: 32 halfs = one SWIZZLE_64B span
constexpr int kChunk = 32;                        // channels per stage
constexpr int kFHalf = kTileM * kRow;             // 8192: one F chunk (hi or lo)
constexpr int kStageBytes = 2 * kFHalf;           // hi | lo
constexpr int kMaxNHalf = 128;
constexpr int kAccStride = 128;                   // TMEM columns per accumulator stage

// barrier slots
enum Bar { FULL = 0, EMPTY = kStages, E_FULL = 2 * kStages, E_EMPTY, T_FULL, T_EMPTY = T_FULL + 2, NUM_BARS = T_EMPTY + 2 };

struct Split {          // query range of one CTA of the pair
  int n0, n, valid;     // first query, MMA N (multiple of 16), queries that exist (<= n)
};
__host__ __device__ inline Split query_split(int Q, int rank) {
  const int npad = (Q + 15) & ~15;
  const int first = ((npad / 16 + 1) / 2) * 16;
  Split s;
  s.n0 = rank == 0 ? 0 : first;
  s.n = rank == 0 ? first : npad - first;
  const int end = s.n0 + s.n < Q ? s.n0 + s.n : Q;
  s.valid = end - s.n0 > 0 ? end - s.n0 : 0;
  return s;
}
__host__ __device__ inline void tile_range(int num_tiles, int clusters, int c, int& begin, int& end) {
  begin = (int)((long long)num_tiles * c / clusters);
  end = (int)((long long)num_tiles * (c + 1) / clusters);
}

__global__ void __launch_bounds__(kThreads, 1)
mask_einsum_mc_kernel(const __grid_constant__ CUtensorMap map_f, const __grid_constant__ CUtensorMap map_e0,
                      const __grid_constant__ CUtensorMap map_e1, int T, int Q, int C, int HW, float* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t rank = cluster_rank();
  const Split sp = query_split(Q, (int)rank);
  const int kchunks = C / kChunk;
  const int e_chunk_bytes = sp.n * kRow;                       // one E chunk (hi or lo) of this CTA: n rows x 64 B
  // layout: [E resident: kchunks x {hi, lo} x kMaxNHalf rows] [F ring] [barriers]   (identical offsets in both CTAs)
  const uint32_t e_base = smem_u32(smem);
  const int e_slot = kMaxNHalf * kRow;                         // 8192: slot per E chunk, independent of n
  const uint32_t f_base = e_base + (uint32_t)(2 * kchunks * e_slot);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)2 * kchunks * e_slot + (size_t)kStages * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_frame = (HW + kTileM - 1) / kTileM;
  int tile_begin, tile_end;
  tile_range(T * tiles_per_frame, (int)(gridDim.x >> 1), (int)(blockIdx.x >> 1), tile_begin, tile_end);

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars[FULL + i], 1);       // this CTA's expect_tx arrive; bytes come from both CTAs' multicasts
      mbar_init(&bars[EMPTY + i], 2);      // one commit from the MMA warp of each CTA
    }
    mbar_init(&bars[E_FULL], 1);
    mbar_init(&bars[E_EMPTY], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[T_FULL + i], 1);
      mbar_init(&bars[T_EMPTY + i], 8);    // one arrive per epilogue warp
    }
    mbar_init_fence();
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, 256);
  }
  fence_before();
  __syncthreads();
  cluster_sync_all();                      // the peer's barriers exist before anything is multicast into this CTA
  fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      const CUtensorMap* map_e = rank == 0 ? &map_e0 : &map_e1;
      int stage = 0, frame_loaded = -1, e_loads = 0;
      uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int t = tile / tiles_per_frame;
        const int p0 = (tile - t * tiles_per_frame) * kTileM;
        if (t != frame_loaded) {           // new frame: this CTA's half of E becomes resident
          if (e_loads > 0) mbar_wait(&bars[E_EMPTY], (uint32_t)((e_loads - 1) & 1), E_EMPTY);
          mbar_expect_tx(&bars[E_FULL], (uint32_t)(2 * kchunks * e_chunk_bytes));
          for (int kc = 0; kc < kchunks; ++kc) {
            tma_load_3d(e_base + (uint32_t)((2 * kc) * e_slot), map_e, &bars[E_FULL], kc * kChunk, sp.n0, t);          // hi
            tma_load_3d(e_base + (uint32_t)((2 * kc + 1) * e_slot), map_e, &bars[E_FULL], C + kc * kChunk, sp.n0, t);  // lo
          }
          frame_loaded = t;
          ++e_loads;
        }
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&bars[EMPTY + stage], phase ^ 1, EMPTY + stage);
          mbar_expect_tx(&bars[FULL + stage], (uint32_t)kStageBytes);
          // CTA 0 fetches the hi chunk, CTA 1 the lo chunk; each box goes to both CTAs
          const uint32_t dst = f_base + (uint32_t)(stage * kStageBytes) + rank * (uint32_t)kFHalf;
          tma_load_3d_multicast(dst, &map_f, &bars[FULL + stage], (int)rank * C + kc * kChunk, p0, t, (uint16_t)0x3);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc_f16(kTileM, sp.n, false, false);
    int stage = 0, acc = 0, frame_ready = -1, e_uses = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int t = tile / tiles_per_frame;
      if (t != frame_ready) {
        mbar_wait(&bars[E_FULL], (uint32_t)(e_uses & 1), E_FULL);
        frame_ready = t;
        ++e_uses;
      }
      const bool last_of_frame = tile + 1 == tile_end || (tile + 1) / tiles_per_frame != t;
      mbar_wait(&bars[T_EMPTY + acc], acc_phase ^ 1, T_EMPTY + acc);
      fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kAccStride);
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&bars[FULL + stage], phase, FULL + stage);
        fence_after();
        if (elect_one()) {
          const uint32_t f_hi = f_base + (uint32_t)(stage * kStageBytes), f_lo = f_hi + (uint32_t)kFHalf;
          const uint32_t e_hi = e_base + (uint32_t)((2 * kc) * e_slot), e_lo = e_hi + (uint32_t)e_slot;
          const uint64_t a_hi = make_desc<kRow>(f_hi), a_lo = make_desc<kRow>(f_lo);
          const uint64_t b_hi = make_desc<kRow>(e_hi), b_lo = make_desc<kRow>(e_lo);
#pragma unroll
          for (int k = 0; k < 2; ++k) {     // 16 channels = 32 bytes inside the swizzle row: +2 in the (addr >> 4) field
            umma_f16(tmem_d, a_lo + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);   // Fl*Eh
            umma_f16(tmem_d, a_hi + (uint64_t)(2 * k), b_lo + (uint64_t)(2 * k), idesc, 1u);                        // Fh*El
            umma_f16(tmem_d, a_hi + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, 1u);                        // Fh*Eh
          }
          umma_commit_multicast(&bars[EMPTY + stage], (uint16_t)0x3);    // both CTAs' producers refill this stage
          if (kc == kchunks - 1) {
            umma_commit(&bars[T_FULL + acc]);                            // accumulator complete -> epilogue
            if (last_of_frame) umma_commit(&bars[E_EMPTY]);              // E may be replaced by the next frame's
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global; two warps per TMEM lane quarter, alternating 32-column chunks =====
    const int ew = warp - 4;
    const int wq = ew & 3, half = ew >> 2;
    const size_t plane = (size_t)T * HW;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int t = tile / tiles_per_frame;
      const int p = (tile - t * tiles_per_frame) * kTileM + wq * 32 + lane;
      mbar_wait(&bars[T_FULL + acc], acc_phase, T_FULL + acc);
      fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * kAccStride);
      float* orow = out + (size_t)sp.n0 * plane + (size_t)t * HW + p;
      const bool row_ok = p < HW;
      for (int c0 = half * 32; c0 < sp.valid; c0 += 64) {
        uint32_t r[32];
        UNIVS_TMEM_LD_X32(taddr + (uint32_t)c0, r);
        tmem_wait_ld();
        if (row_ok) {
          float* dst = orow + (size_t)c0 * plane;
          if (c0 + 32 <= sp.valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              __stcs(dst, __uint_as_float(r[j]));
              dst += plane;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (c0 + j < sp.valid) __stcs(dst + (size_t)j * plane, __uint_as_float(r[j]));
            }
          }
        }
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[T_EMPTY + acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  fence_before();
  __syncthreads();
  cluster_sync_all();          // no CTA leaves while its peer may still multicast into it or signal its barriers
  if (warp == 2) {
    fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
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

// fp16 tensor [d2][d1][d0] (d0 contiguous), box [32 halfs][box_rows][1], SWIZZLE_64B; rows beyond d1 read as zeros
static int make_map(CUtensorMap* m, const void* base, int d0, int d1, int d2, int box_rows) {
  EncodeTiledFn enc = encoder();
  if (!enc) { set_error("mask_einsum_mc: cuTensorMapEncodeTiled entry point unavailable"); return UNIVS_E_LAUNCH; }
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)d0 * 2, (cuuint64_t)d0 * (cuuint64_t)d1 * 2};
  cuuint32_t box[3] = {(cuuint32_t)kChunk, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("mask_einsum_mc: cuTensorMapEncodeTiled failed (%d)", (int)r); return UNIVS_E_LAUNCH; }
  return 0;
}

}  // namespace einmc

size_t mask_einsum_mc_smem_bytes(int C) {
  using namespace einmc;
  return (size_t)2 * (C / kChunk) * kMaxNHalf * kRow + (size_t)kStages * kStageBytes + NUM_BARS * 8 + 16;
}

int launch_mask_einsum_mc_f16(cudaStream_t st, const void* E16, const void* F16, int T, int Q, int C, int HW, float* out) {
  using namespace einmc;
  const Split s0 = query_split(Q, 0), s1 = query_split(Q, 1);
  if (s1.n < 16 || s0.n > kMaxNHalf) { set_error("mask_einsum_mc: needs 16 < Q <= 256"); return UNIVS_E_BADARG; }
  const size_t smem = mask_einsum_mc_smem_bytes(C);
  if (smem > 227 * 1024) { set_error("mask_einsum_mc: %d channels do not fit (E resident + F ring = %zu bytes)", C, smem); return UNIVS_E_BADARG; }
  CUtensorMap map_f, map_e0, map_e1;
  int rc = make_map(&map_f, F16, 2 * C, HW, T, kTileM);
  if (rc) return rc;
  if ((rc = make_map(&map_e0, E16, 2 * C, Q, T, s0.n))) return rc;
  if ((rc = make_map(&map_e1, E16, 2 * C, Q, T, s1.n))) return rc;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  cudaError_t e = cudaFuncSetAttribute(mask_einsum_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("mask_einsum_mc: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e)); return UNIVS_E_LAUNCH; }
  const int tiles = T * ((HW + kTileM - 1) / kTileM);
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent clusters: as many as can be co-resident (a GPC with an odd number of free SMs leaves one SM out), so that
  // no cluster waits for a second wave
  static int resident = 0;
  static size_t resident_smem = 0;
  if (!resident || resident_smem != smem) {
    cfg.gridDim = dim3((unsigned)(2 * (num_sms / 2)));
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, mask_einsum_mc_kernel, &cfg);
    if (e != cudaSuccess || n < 1) {
      set_error("mask_einsum_mc: cudaOccupancyMaxActiveClusters: %s (%d clusters)", cudaGetErrorString(e), n);
      (void)cudaGetLastError();
      return UNIVS_E_LAUNCH;
    }
    resident = n < num_sms / 2 ? n : num_sms / 2;
    resident_smem = smem;
  }
  const int clusters = resident < tiles ? resident : tiles;
  cfg.gridDim = dim3((unsigned)(2 * clusters));
  e = cudaLaunchKernelEx(&cfg, mask_einsum_mc_kernel, map_f, map_e0, map_e1, T, Q, C, HW, out);
  if (e != cudaSuccess) { set_error("mask_einsum_mc: launch: %s", cudaGetErrorString(e)); return UNIVS_E_LAUNCH; }
  return check_launch("mask_einsum_mc");
}

}  // namespace univs
