// Mask-logit einsum on the 5th-gen tensor cores: TMA -> shared memory -> tcgen05.mma -> TMEM -> tcgen05.ld -> coalesced
// global stores.  (reference op: ..._univs.py:527-528 "btqc,btchw->btqhw" + transpose)
//
//   out[q, t, p] = sum_c E[t, q, c] * F[t, p, c]      E [T,Q,C], F channel-last [T,HW,C], out [Q,T,HW] fp32
//
// Mapping onto UMMA (D = A * B^T, both operands K-major):
//   A (M side) = F tile: 128 pixels x one K-chunk   -> TMEM lane  = pixel
//   B (N side) = E tile: Npad queries x one K-chunk -> TMEM column = query        (Npad = ceil16(Q) <= 256)
// so that one TMEM column (= one query) is 128 consecutive pixels = 512 contiguous bytes of `out`: the epilogue's
// 32x32b TMEM loads give every lane one pixel and the stores of a warp are one full 128-byte line per query.
//
// Two instantiations:
//   <false> fp32 operands consumed as TF32 (kind::tf32, truncation: callers pre-round with univs_round_tf32_f32),
//           128-byte operand rows (32 floats), SWIZZLE_128B, 4-stage ring;
//   <true>  fp16 [hi | lo] operands (univs_split_tf32_f32 chunk = UNIVS_SPLIT_F16U), three kind::f16 MMAs per k-step
//           (lo*hi + hi*lo + hi*hi): fp32-equivalent products at the byte count of the fp32 tensors; 64-byte operand
//           rows (32 halfs), SWIZZLE_64B, 4-stage ring of {F_hi, E_hi, F_lo, E_lo}.
// Accumulation is fp32 in TMEM in both.
//
// HBM-bound by design (AI = 56 F/B): per 128-pixel tile the kernel streams 128 KB of F once and writes Q*512 B of logits
// once; E (Q*C*4 B per frame) is re-read from L2 per tile.  Persistent CTAs (one per SM), warp-specialised:
// warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM allocator, warps 4-11 = epilogue
// (two per TMEM lane quarter, alternating 32-column chunks, branch-free pointer-bump stores -- the 4-warp predicated
// epilogue of the first version was the bottleneck: 0.177 ms -> 0.109 ms at the north-star shape).  Two TMEM accumulator
// stages let the epilogue of tile i overlap the MMAs of tile i+1.
#include <cuda.h>

#include "common.cuh"

namespace univs {

constexpr int kTcTileM = 128;
constexpr int kTcStagesTf32 = 4;
constexpr int kTcStagesF16 = 4;   // 64-byte operand rows (SWIZZLE_64B): four 42 KB stages instead of two 85 KB ones
constexpr int kTcThreads = 384;   // 4 control warps + 8 epilogue warps
constexpr int kTcMaxN = 256;
constexpr int kTcABytes = kTcTileM * 128;  // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// K-major swizzled operand tile: rows of ROW bytes (128 -> SWIZZLE_128B, layout code 2; 64 -> SWIZZLE_64B, code 4),
// 8-row groups 8*ROW bytes apart (cute canonical K-major layouts ((8,n),2):((ROW/16,SBO),1), LBO = 1)
template <int ROW>
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((8 * ROW) >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)(ROW == 128 ? 2 : 4) << 61);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

#define TMEM_LD_32x32b_X32(taddr, r)                                                                                   \
  asm volatile(                                                                                                        \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19," \
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                       \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),    \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),         \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),        \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                      \
      : "r"(taddr))

// F16X3 == false: fp32 operands consumed as TF32 (one MMA per k-step, K-chunk = 32 floats).
// F16X3 == true : fp16 operands [hi | lo] (lo unscaled), three MMAs per k-step  lo*hi + hi*lo + hi*hi  (fp32-equivalent
//                 products, correction terms first), K-chunk = 64 halfs; `C` is the logical channel count, the tensors
//                 are [.., 2C] wide with hi at columns [0,C) and lo at [C,2C).
template <bool F16X3>
__global__ void __launch_bounds__(kTcThreads, 1)
mask_einsum_tc_kernel(const __grid_constant__ CUtensorMap map_f, const __grid_constant__ CUtensorMap map_e, int T, int Q,
                      int C, int HW, int npad, float* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int kTcStages = F16X3 ? kTcStagesF16 : kTcStagesTf32;
  constexpr int kRowBytes = F16X3 ? 64 : 128;       // operand row = one swizzle span
  constexpr int kChunkElems = 32;                   // 32 halfs (64 B) or 32 floats (128 B) per stage
  constexpr int kKSteps = F16X3 ? 2 : 4;            // MMA k-steps (32 B each) per stage
  constexpr int kABytes = kTcTileM * kRowBytes;     // 8 KB / 16 KB
  const int b_bytes = npad * kRowBytes;             // multiple of 1024 (npad % 16 == 0)
  const int stage_bytes = (F16X3 ? 2 : 1) * (kABytes + b_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kTcStages * stage_bytes);
  uint64_t* full = bars;                   // [kTcStages]
  uint64_t* empty = bars + kTcStages;      // [kTcStages]
  uint64_t* tfull = bars + 2 * kTcStages;  // [2]
  uint64_t* tempty = tfull + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_frame = (HW + kTcTileM - 1) / kTcTileM;
  const int num_tiles = T * tiles_per_frame;
  const int kchunks = C / kChunkElems;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kTcStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int t = tile / tiles_per_frame;
        const int p0 = (tile - t * tiles_per_frame) * kTcTileM;
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* sA = smem + (size_t)stage * stage_bytes;
          mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);
          tma_load_3d(sA, &map_f, &full[stage], kc * kChunkElems, p0, t);
          tma_load_3d(sA + kABytes, &map_e, &full[stage], kc * kChunkElems, 0, t);
          if (F16X3) {   // lo halves live C columns further right
            tma_load_3d(sA + kABytes + b_bytes, &map_f, &full[stage], C + kc * kChunkElems, p0, t);
            tma_load_3d(sA + 2 * kABytes + b_bytes, &map_e, &full[stage], C + kc * kChunkElems, 0, t);
          }
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // instruction descriptor: D=f32 (bits 4-5 = 1), A=B=tf32 (2 at bits 7-9 / 10-12), K-major both, N>>3 @17, M>>4 @24
    // (kind::f16: A=B=f16 -> format code 0)
    const uint32_t fmt = F16X3 ? 0u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(kTcTileM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kTcMaxN);
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = make_desc<kRowBytes>(a_addr);
          const uint64_t bdesc = make_desc<kRowBytes>(a_addr + kABytes);
#pragma unroll
          for (int k = 0; k < kKSteps; ++k) {
            // one k-step = 32 bytes (8 tf32 / 16 f16) inside the swizzle row: +2 in the (addr >> 4) field
            if (F16X3) {
              const uint64_t adesc_lo = make_desc<kRowBytes>(a_addr + kABytes + b_bytes);
              const uint64_t bdesc_lo = make_desc<kRowBytes>(a_addr + 2 * kABytes + b_bytes);
              umma_f16(tmem_d, adesc_lo + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);  // Fl*Eh
              umma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc_lo + (uint64_t)(2 * k), idesc, 1u);                        // Fh*El
              umma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);                           // Fh*Eh
            } else {
              umma_tf32(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);                       // frees the smem stage when these MMAs retire
          if (kc == kchunks - 1) umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == kTcStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global.  8 warps: two per TMEM lane quarter, each taking every other
    // 32-column chunk; branch-free stores (one pointer bump per query column) for all full chunks =====
    const int ew = warp - 4;
    const int wq = ew & 3;      // TMEM lane quarter this warp may access (== warp % 4)
    const int half = ew >> 2;   // 0: chunks 0,2,4,..   1: chunks 1,3,5,..
    const size_t plane = (size_t)T * HW;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int t = tile / tiles_per_frame;
      const int p0 = (tile - t * tiles_per_frame) * kTcTileM;
      const int p = p0 + wq * 32 + lane;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * kTcMaxN);
      float* orow = out + (size_t)t * HW + p;
      const bool row_ok = p < HW;
      for (int c0 = half * 32; c0 < npad; c0 += 64) {
        uint32_t r[32];
        TMEM_LD_32x32b_X32(taddr + (uint32_t)c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row_ok) {
          float* dst = orow + (size_t)c0 * plane;
          if (c0 + 32 <= Q) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              __stcs(dst, __uint_as_float(r[j]));
              dst += plane;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (c0 + j < Q) __stcs(dst + (size_t)j * plane, __uint_as_float(r[j]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---- host side: tensor maps through the driver entry point (no link-time libcuda dependency) ------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 3-D tensor [d2][d1][d0] (d0 contiguous) of fp32 (esize 4) or fp16 (esize 2), box [128 bytes][box_rows][1], SWIZZLE_128B
static int make_map(CUtensorMap* m, const void* base, int d0, int d1, int d2, int box_rows, int esize, int row_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("mask_einsum_tc: cuTensorMapEncodeTiled entry point unavailable"); return UNIVS_E_LAUNCH; }
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)d0 * esize, (cuuint64_t)d0 * (cuuint64_t)d1 * esize};
  cuuint32_t box[3] = {(cuuint32_t)(row_bytes / esize), (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                   const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("mask_einsum_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return UNIVS_E_LAUNCH; }
  return 0;
}

template <bool F16X3>
static int launch_tc(cudaStream_t st, const void* E, const void* F, int T, int Q, int C, int HW, float* out) {
  const int npad = (Q + 15) & ~15;
  const int esize = F16X3 ? 2 : 4;
  const int width = F16X3 ? 2 * C : C;      // fp16 operands carry [hi | lo]
  CUtensorMap map_f, map_e;
  const int row_bytes = F16X3 ? 64 : 128;
  int rc = make_map(&map_f, F, width, HW, T, kTcTileM, esize, row_bytes);
  if (rc) return rc;
  rc = make_map(&map_e, E, width, Q, T, npad, esize, row_bytes);
  if (rc) return rc;
  const int stages = F16X3 ? kTcStagesF16 : kTcStagesTf32;
  const size_t smem = (size_t)stages * (F16X3 ? 2 : 1) * ((size_t)kTcTileM * row_bytes + (size_t)npad * row_bytes) + 256;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  cudaError_t e = cudaFuncSetAttribute(mask_einsum_tc_kernel<F16X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("mask_einsum_tc: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e)); return UNIVS_E_LAUNCH; }
  const int tiles = T * ((HW + kTcTileM - 1) / kTcTileM);
  const int grid = tiles < num_sms ? tiles : num_sms;
  mask_einsum_tc_kernel<F16X3><<<grid, kTcThreads, smem, st>>>(map_f, map_e, T, Q, C, HW, npad, out);
  return check_launch("mask_einsum_tc");
}

int launch_mask_einsum_tc(cudaStream_t st, const float* E, const float* F, int T, int Q, int C, int HW, float* out) {
  return launch_tc<false>(st, E, F, T, Q, C, HW, out);
}
int launch_mask_einsum_tc_f16(cudaStream_t st, const void* E16, const void* F16, int T, int Q, int C, int HW, float* out) {
  return launch_tc<true>(st, E16, F16, T, Q, C, HW, out);
}

}  // namespace univs
