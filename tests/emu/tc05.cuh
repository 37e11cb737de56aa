// CPU implementation of univs_b200/csrc/tc05.cuh (the PTX wrappers of the tcgen05 kernels) for tests/emu: the kernels'
// sources compile against THIS header instead (tests/emu/build_emu.py puts tests/emu first on the include path); the
// descriptor encodings and swizzle arithmetic stay the production ones (tc05_math.cuh).
//
// What is modelled (PTX ISA / CuTe mma_sm100_desc.hpp semantics, the same assumptions the kernels are written against):
//   * mbarrier: arrival count + transaction bytes + phase bit; try_wait.parity spins (yielding) and aborts after 120 s;
//   * tcgen05.mma kind::f16, cta_group::1: M = 128, N from the instruction descriptor, K = 16; operands are read from
//     (emulated) shared memory through the matrix descriptors -- start address, stride byte offset between 8-row groups,
//     swizzle mode applied to the ABSOLUTE shared-memory address (XOR of address bits 4.. with bits 7..), K-major or
//     MN-major as the instruction descriptor says -- multiplied in fp32 and accumulated into TMEM (lane = M row,
//     column = N index); issued work goes to an in-order tensor pipe that executes it LATER on its own thread, and
//     tcgen05.commit arrives when the pipe reaches it (operands overwritten too early give wrong results);
//   * TMEM: 128 lanes x 512 fp32 columns per CTA, NaN-filled at launch (reading what no MMA wrote is visible);
//   * tcgen05.ld 32x32b: thread i of a warp reads lane (warp % 4) * 32 + i; the address' lane field must name that
//     quarter (the hardware restriction is checked);
//   * bar.sync id, n: a barrier per id over n threads; elect.sync: lane 0.
// Shared memory is a NaN-filled 232 KB buffer per CTA; `smem_u32` is the offset into it.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include <chrono>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <thread>

#include "common.cuh"
#include "tc05_math.cuh"

namespace emu {
unsigned char* dyn_smem();                       // base of the CTA's dynamic shared memory (1024-byte aligned)
unsigned char* dyn_smem_of(int cta);             // the same of another CTA of the cluster
CUresult encode_tiled(CUtensorMap* m, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave,
                      CUtensorMapSwizzle swizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct TensorMap {                               // what the emulated cuTensorMapEncodeTiled stores in the opaque CUtensorMap
  unsigned long long magic;
  const unsigned char* base;
  unsigned long long dims[3], strides[2];        // elements / bytes
  unsigned box[3];
  int esize, swizzle_bits, rank;
};
static_assert(sizeof(TensorMap) <= sizeof(CUtensorMap), "fits the opaque descriptor");
float* tmem();                                   // [128][512]
struct MBar {
  int count = 0, pending = 0;
  long long tx = 0;
  unsigned phase = 0;
  unsigned long long completions = 0;
};
MBar& mbar_of(const void* p);                    // state of the mbarrier stored at this shared-memory word
std::mutex& mbar_lock();
Barrier& named_barrier(int id, int threads);
[[noreturn]] void fail(const char* what);
float* tmem_of(int cta);                       // TMEM of CTA `cta` of the running cluster (cta_group::2 MMAs write both)
void pipe_push(std::function<void()> fn);        // enqueue work on this CTA's (asynchronous, in-order) tensor pipe
void note_wait(int thread, int id, int parity);  // diagnostics: what every thread of the CTA is waiting for (-1: nothing)
void dump_waits();
}  // namespace emu

namespace univs {
namespace tc {

// UNIVS_EMU_CHAOS=<max microseconds>: every barrier operation, MMA issue, TMEM load and TMA copy is preceded by a random
// delay of the calling thread, so that roles overtake each other in ways the natural scheduling rarely produces
inline void chaos() {
  static const int max_us = [] { const char* e = std::getenv("UNIVS_EMU_CHAOS"); return e ? std::atoi(e) : 0; }();
  if (max_us <= 0) return;
  static thread_local unsigned state = 0x9E3779B9u ^ (unsigned)(size_t)&state;
  state = state * 1664525u + 1013904223u;
  const unsigned r = state >> 8;
  if ((r & 3u) == 0) std::this_thread::sleep_for(std::chrono::microseconds((r >> 2) % (unsigned)max_us));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)(reinterpret_cast<const unsigned char*>(p) - ::emu::dyn_smem());
}
inline void mbar_complete(::emu::MBar& b) {      // caller holds the lock
  if (b.pending == 0 && b.tx == 0) {
    b.phase ^= 1u;
    b.pending = b.count;
    ++b.completions;
  }
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  std::lock_guard<std::mutex> l(::emu::mbar_lock());
  ::emu::MBar& b = ::emu::mbar_of(bar);
  b = ::emu::MBar();
  b.count = b.pending = (int)count;
}
__device__ __forceinline__ void mbar_init_fence() {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  chaos();
  std::lock_guard<std::mutex> l(::emu::mbar_lock());
  ::emu::MBar& b = ::emu::mbar_of(bar);
  if (b.pending <= 0) ::emu::fail("mbarrier: more arrivals than its count in one phase");
  b.tx += bytes;
  --b.pending;
  mbar_complete(b);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  chaos();
  std::lock_guard<std::mutex> l(::emu::mbar_lock());
  ::emu::MBar& b = ::emu::mbar_of(bar);
  if (b.pending <= 0) ::emu::fail("mbarrier: more arrivals than its count in one phase");
  --b.pending;
  mbar_complete(b);
}
inline void mbar_complete_tx(uint64_t* bar, long long bytes) {
  std::lock_guard<std::mutex> l(::emu::mbar_lock());
  ::emu::MBar& b = ::emu::mbar_of(bar);
  b.tx -= bytes;
  mbar_complete(b);
}
// try_wait.parity succeeds when the barrier's phase bit differs from `parity`.  The lanes of a converged warp execute the
// instruction together and therefore all see the completion the first of them sees; host threads do not run in lockstep
// (a lane may be descheduled for milliseconds while lane 0 issues work that flips the same barrier again), so the wait is
// made STICKY per thread: it also succeeds if the barrier completed at least once since this thread's last successful wait
// on it -- the completion it would have seen in lockstep.  (Whether a wait can really be overtaken by two completions is
// the business of tests/test_tc_protocol_sim.py / test_einsum_mc_protocol.py, at the granularity of roles.)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int id = 0) {
  static thread_local std::map<const void*, unsigned long long> seen;      // completions at this thread's last pass
  chaos();
  const auto t0 = std::chrono::steady_clock::now();
  ::emu::note_wait((int)threadIdx.x, id, (int)parity);
  for (unsigned spin = 0;; ++spin) {
    {
      std::lock_guard<std::mutex> l(::emu::mbar_lock());
      const ::emu::MBar& b = ::emu::mbar_of(bar);
      unsigned long long& last = seen[bar];
      if (b.phase != parity || b.completions > last) {
        last = b.completions;
        ::emu::note_wait((int)threadIdx.x, -1, 0);
        return;
      }
    }
    if (spin < 64) {
      std::this_thread::yield();
    } else {
      std::this_thread::sleep_for(std::chrono::microseconds(50));
      if ((spin & 1023u) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) {
        std::fprintf(stderr, "emu: mbarrier wait timed out (id %d, parity %u, thread %u)\n", id, parity, (unsigned)threadIdx.x);
        ::emu::dump_waits();
        ::emu::fail("mbarrier wait timed out: protocol deadlock");
      }
    }
  }
}
__device__ __forceinline__ void fence_before() {}
__device__ __forceinline__ void fence_after() {}
__device__ __forceinline__ void fence_proxy_async_smem() {}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {      // arrives when everything issued before it has executed
  ::emu::pipe_push([bar] { mbar_arrive(bar); });
}

// ---- thread-block clusters and TMA ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() { return (uint32_t)::emu::ctx.cta; }
__device__ __forceinline__ void cluster_sync_all() { ::emu::ctx.cluster_bar->wait(); }
inline uint64_t* peer_bar(uint64_t* bar, int cta) {          // the barrier at the same shared-memory offset in CTA `cta`
  return reinterpret_cast<uint64_t*>(::emu::dyn_smem_of(cta) + (reinterpret_cast<unsigned char*>(bar) - ::emu::dyn_smem()));
}
// cp.async.bulk.tensor.3d: box [box0 elements][box1 rows][1] at (c0, c1, c2); elements outside the tensor read as zero; the
// box lands row by row (box0 * esize bytes per row) with the map's swizzle applied to the absolute shared-memory address;
// the full box size is credited to the barrier.  Executed synchronously (a legal schedule of the asynchronous copy).
inline void tma_copy(int cta, uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int bar_cta = -1) {
  chaos();
  const ::emu::TensorMap& m = *reinterpret_cast<const ::emu::TensorMap*>(map);
  if (m.magic != 0x554e495653ull || m.rank != 3) ::emu::fail("TMA: not a tensor map of the emulated encoder");
  const unsigned row_bytes = m.box[0] * (unsigned)m.esize;
  if (dst % 1024u) ::emu::fail("TMA: swizzled destination must be 1024-byte aligned");
  if (row_bytes != (16u << m.swizzle_bits)) ::emu::fail("TMA: box row must be one swizzle span");
  unsigned char* smem = ::emu::dyn_smem_of(cta);
  for (unsigned r = 0; r < m.box[1]; ++r)
    for (unsigned e = 0; e < m.box[0]; ++e) {
      const long long x = (long long)c0 + e, y = (long long)c1 + r, z = c2;
      unsigned char v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (x >= 0 && y >= 0 && z >= 0 && (unsigned long long)x < m.dims[0] && (unsigned long long)y < m.dims[1] && (unsigned long long)z < m.dims[2])
        std::memcpy(v, m.base + (size_t)z * m.strides[1] + (size_t)y * m.strides[0] + (size_t)x * m.esize, m.esize);
      const uint32_t addr = dst + r * row_bytes + e * (unsigned)m.esize;
      const uint32_t sw = addr ^ (((addr >> 7) & ((1u << m.swizzle_bits) - 1u)) << 4);
      if (sw + m.esize > 232448u) ::emu::fail("TMA: box beyond shared memory");
      std::memcpy(smem + sw, v, m.esize);
    }
  mbar_complete_tx(peer_bar(bar, bar_cta < 0 ? cta : bar_cta), (long long)row_bytes * m.box[1]);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  tma_copy(::emu::ctx.cta, dst, map, bar, c0, c1, c2);
}
__device__ __forceinline__ void tma_load_3d_multicast(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                      int c2, uint16_t mask) {
  for (int c = 0; c < ::emu::ctx.cluster_size; ++c)
    if (mask & (1u << c)) tma_copy(c, dst, map, bar, c0, c1, c2);
}
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
  for (int c = 0; c < ::emu::ctx.cluster_size; ++c)
    if (mask & (1u << c)) {
      uint64_t* target = peer_bar(bar, c);
      ::emu::pipe_push([target] { mbar_arrive(target); });
    }
}

// ---- operand access through the shared-memory matrix descriptor ------------------------------------------------------
struct Desc {
  uint32_t start, lbo, sbo;
  int swizzle_bits;                              // 0: none, 1: 32B, 2: 64B, 3: 128B
};
inline Desc decode_desc(uint64_t d) {
  Desc r;
  r.start = (uint32_t)(d & 0x3FFF) << 4;
  r.lbo = (uint32_t)((d >> 16) & 0x3FFF) << 4;
  r.sbo = (uint32_t)((d >> 32) & 0x3FFF) << 4;
  const int layout = (int)((d >> 61) & 7);
  r.swizzle_bits = layout == 2 ? 3 : (layout == 4 ? 2 : (layout == 6 ? 1 : 0));
  if (((d >> 46) & 1) != 1) ::emu::fail("matrix descriptor without the Blackwell version bit");
  if (layout != 0 && layout != 2 && layout != 4 && layout != 6) ::emu::fail("matrix descriptor: unknown layout type");
  return r;
}
inline uint32_t swizzled(uint32_t addr, int bits) {      // Swizzle<bits,4,3> on the absolute shared-memory address
  return addr ^ (((addr >> 7) & ((1u << bits) - 1u)) << 4);
}
// element (r = M/N index, k = K index inside one MMA) of an operand whose elements are `esize` bytes (2: fp16, 4: tf32)
inline float operand_elem(const Desc& d, bool mn_major, int r, int k, int esize, const unsigned char* smem_base = nullptr) {
  const int span = 16 << d.swizzle_bits;          // bytes of one swizzle row (32 / 64 / 128)
  const int per16 = 16 / esize;                   // elements per 16-byte unit (the "T" of the canonical layouts)
  uint32_t off;
  if (!mn_major) {                                // K-major: rows of `span` bytes, 8-row groups SBO apart, K contiguous
    off = (uint32_t)(r >> 3) * d.sbo + (uint32_t)(r & 7) * span + (uint32_t)k * esize;
  } else {                                        // MN-major: K rows of `span` bytes (span/esize MN elements), 8-row groups SBO apart
    const int per_row = span / esize;
    off = (uint32_t)(r / per_row) * d.lbo + (uint32_t)(r % per_row) * esize + (uint32_t)(k >> 3) * d.sbo + (uint32_t)(k & 7) * span;
  }
  (void)per16;
  const uint32_t addr = swizzled(d.start + off, d.swizzle_bits);
  if (addr + esize > 232448u) ::emu::fail("tcgen05.mma operand read beyond shared memory");
  const unsigned char* p = (smem_base ? smem_base : ::emu::dyn_smem()) + addr;
  if (esize == 2) return __half2float(*reinterpret_cast<const __half*>(p));
  uint32_t u;
  std::memcpy(&u, p, 4);
  u &= 0xFFFFE000u;                               // kind::tf32 consumes the upper 19 bits of each fp32 operand
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline void umma_emulated(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum, int fmt, int esize, int K) {
  chaos();
  const int M = (int)((idesc >> 24) & 0x1F) << 4, N = (int)((idesc >> 17) & 0x3F) << 3;
  const bool a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
  if (M != 128 || N < 16 || N > 256 || (N & 15)) ::emu::fail("tcgen05.mma: unsupported shape (M = 128, 16 <= N <= 256, N % 16 == 0)");
  if (((idesc >> 4) & 3) != 1 || (int)((idesc >> 7) & 7) != fmt || (int)((idesc >> 10) & 7) != fmt)
    ::emu::fail("tcgen05.mma: instruction descriptor formats do not match the instruction kind (f32 accumulate)");
  const Desc a = decode_desc(adesc), b = decode_desc(bdesc);
  const int lane0 = (int)(tmem_d >> 16), col0 = (int)(tmem_d & 0xFFFF);
  if (lane0 != 0 || col0 + N > 512) ::emu::fail("tcgen05.mma: accumulator outside TMEM");
  float* T = ::emu::tmem();
  static thread_local float A[128][16], B[256][16];
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) A[m][k] = operand_elem(a, a_mn, m, k, esize);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) B[n][k] = operand_elem(b, b_mn, n, k, esize);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = accum ? T[m * 512 + col0 + n] : 0.f;
      for (int k = 0; k < K; ++k) acc += A[m][k] * B[n][k];
      T[m * 512 + col0 + n] = acc;
    }
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  ::emu::pipe_push([=] { umma_emulated(tmem_d, adesc, bdesc, idesc, accum, /*fmt f16*/ 0, 2, 16); });
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  ::emu::pipe_push([=] { umma_emulated(tmem_d, adesc, bdesc, idesc, accum, /*fmt tf32*/ 2, 4, 8); });
}
// ---- CTA pairs (cta_group::2) ---------------------------------------------------------------------------------------------
// One MMA of M = 256 issued by the leader (CTA 0): CTA c contributes rows 128c..128c+127 of A and rows (N/2)c..(N/2)c+N/2-1 of
// B from ITS shared memory (same descriptors = same offsets in both) and receives rows 128c.. of D in ITS TMEM.
inline void umma_emulated_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  chaos();
  const int M = (int)((idesc >> 24) & 0x1F) << 4, N = (int)((idesc >> 17) & 0x3F) << 3;
  const bool a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
  if (M != 256 || N < 32 || N > 256 || (N & 31)) ::emu::fail("tcgen05.mma.cta_group::2: unsupported shape (M = 256, 32 <= N <= 256, N % 32 == 0)");
  if (((idesc >> 4) & 3) != 1 || ((idesc >> 7) & 7) != 0 || ((idesc >> 10) & 7) != 0)
    ::emu::fail("tcgen05.mma.cta_group::2: instruction descriptor is not f16 x f16 -> f32");
  if (::emu::ctx.cluster_size != 2) ::emu::fail("tcgen05.mma.cta_group::2 outside a cluster of two CTAs");
  const Desc a = decode_desc(adesc), b = decode_desc(bdesc);
  const int lane0 = (int)(tmem_d >> 16), col0 = (int)(tmem_d & 0xFFFF);
  if (lane0 != 0 || col0 + N > 512) ::emu::fail("tcgen05.mma: accumulator outside TMEM");
  static thread_local float A[128][16], B[256][16];
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 16; ++k) B[n][k] = operand_elem(b, b_mn, n % (N / 2), k, 2, ::emu::dyn_smem_of(n / (N / 2)));
  for (int c = 0; c < 2; ++c) {
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < 16; ++k) A[m][k] = operand_elem(a, a_mn, m, k, 2, ::emu::dyn_smem_of(c));
    float* T = ::emu::tmem_of(c);
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        float acc = accum ? T[m * 512 + col0 + n] : 0.f;
        for (int k = 0; k < 16; ++k) acc += A[m][k] * B[n][k];
        T[m * 512 + col0 + n] = acc;
      }
  }
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (::emu::ctx.cta != 0) ::emu::fail("tcgen05.mma.cta_group::2 must be issued by the leader CTA");
  ::emu::pipe_push([=] { umma_emulated_2sm(tmem_d, adesc, bdesc, idesc, accum); });
}
__device__ __forceinline__ void umma_commit_multicast_2sm(uint64_t* bar, uint16_t mask) { umma_commit_multicast(bar, mask); }
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  tma_copy(::emu::ctx.cta, dst, map, bar, c0, c1, c2, /*barrier in CTA*/ 0);
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) { mbar_arrive(peer_bar(bar, (int)cta)); }

__device__ __forceinline__ bool elect_one() { return ::emu::ctx.lane == 0; }
__device__ __forceinline__ void named_bar_sync(int id, int threads) { ::emu::named_barrier(id, threads).wait(); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  if (cols < 32 || cols > 512 || (cols & (cols - 1))) ::emu::fail("tcgen05.alloc: columns must be a power of two in [32, 512]");
  if (::emu::ctx.lane == 0) *slot = 0;
}
__device__ __forceinline__ void tmem_dealloc(uint32_t, uint32_t) {}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* slot, uint32_t cols) { tmem_alloc(slot, cols); }
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t, uint32_t) {}
__device__ __forceinline__ void tmem_wait_ld() {}

inline void tmem_ld(uint32_t taddr, uint32_t* r, int n) {
  chaos();
  const int lane_field = (int)(taddr >> 16), col = (int)(taddr & 0xFFFF);
  const int warp = (int)(threadIdx.x >> 5);
  if (lane_field != (warp & 3) * 32) ::emu::fail("tcgen05.ld 32x32b: a warp may only access TMEM lanes 32 * (warp % 4) ..+31");
  if (col + n > 512) ::emu::fail("tcgen05.ld beyond TMEM");
  const float* T = ::emu::tmem() + (size_t)(lane_field + ::emu::ctx.lane) * 512 + col;
  std::memcpy(r, T, sizeof(float) * n);
}
#define UNIVS_TMEM_LD_X4(taddr, r) ::univs::tc::tmem_ld(taddr, r, 4)
#define UNIVS_TMEM_LD_X8(taddr, r) ::univs::tc::tmem_ld(taddr, r, 8)
#define UNIVS_TMEM_LD_X16(taddr, r) ::univs::tc::tmem_ld(taddr, r, 16)
#define UNIVS_TMEM_LD_X32(taddr, r) ::univs::tc::tmem_ld(taddr, r, 32)

__device__ __forceinline__ float ex2_approx(float x) { return std::exp2(x); }
__device__ __forceinline__ void sts_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  const uint32_t v[4] = {a, b, c, d};
  if (saddr & 15) ::emu::fail("st.shared.v4 misaligned");
  std::memcpy(::emu::dyn_smem() + saddr, v, 16);
}
__device__ __forceinline__ void sts_v2(uint32_t saddr, uint32_t a, uint32_t b) {
  const uint32_t v[2] = {a, b};
  if (saddr & 7) ::emu::fail("st.shared.v2 misaligned");
  std::memcpy(::emu::dyn_smem() + saddr, v, 8);
}

}  // namespace tc
}  // namespace univs
