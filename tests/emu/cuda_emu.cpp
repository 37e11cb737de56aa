// Runtime of the CPU emulation (see cuda_emu.h): block / warp scheduling and the handful of CUDA runtime entry points
// the launch wrappers call.
#include "cuda_emu.h"

#include <cstdio>
#include <cstdlib>
#include <deque>
#include <map>
#include <memory>
#include <string>

#include "tc05.cuh"

namespace emu {

thread_local Ctx ctx;
std::mutex atomic_lock;

// ---- per-CTA state of the tensor-core emulation (the CTAs of one cluster run concurrently, clusters one after the other)
static constexpr size_t kSmemBytes = 232448;                       // 227 KB
static constexpr int kMaxCluster = 2;
struct CtaState {
  alignas(1024) unsigned char smem[kSmemBytes + 1024];
  float tmem[128 * 512];
  std::map<const void*, MBar> mbars;
  std::map<int, std::unique_ptr<Barrier>> named;
};
static CtaState g_cta[kMaxCluster];

// the tensor pipe of a CTA: tcgen05.mma / tcgen05.commit are ASYNCHRONOUS -- issued work is queued and executed in order by a
// worker thread (with random delays under UNIVS_EMU_CHAOS), so an operand tile that is overwritten before "its" MMA has
// actually run, or an accumulator read before the commit retired, shows up as a wrong result
struct Pipe {
  std::mutex m;
  std::condition_variable cv;
  std::deque<std::function<void()>> q;
  bool stop = false;
  std::thread worker;
};
static Pipe g_pipe[kMaxCluster];
void pipe_push(std::function<void()> fn) {
  Pipe& p = g_pipe[ctx.cta];
  {
    std::lock_guard<std::mutex> l(p.m);
    p.q.push_back(std::move(fn));
  }
  p.cv.notify_one();
}
static void pipe_start(int c) {
  Pipe& p = g_pipe[c];
  p.stop = false;
  p.worker = std::thread([c] {
    ctx.cta = c;
    ctx.cluster_size = kMaxCluster;
    ctx.lane = 0;
    Pipe& p = g_pipe[c];
    for (;;) {
      std::function<void()> fn;
      {
        std::unique_lock<std::mutex> l(p.m);
        p.cv.wait(l, [&] { return p.stop || !p.q.empty(); });
        if (p.q.empty()) return;
        fn = std::move(p.q.front());
        p.q.pop_front();
      }
      fn();
    }
  });
}
static void pipe_finish(int c) {
  Pipe& p = g_pipe[c];
  {
    std::lock_guard<std::mutex> l(p.m);
    p.stop = true;
  }
  p.cv.notify_one();
  p.worker.join();
}
static std::mutex g_mbar_lock, g_named_lock;
static int g_wait_id[kMaxCluster][1024], g_wait_parity[kMaxCluster][1024];

unsigned char* dyn_smem() { return g_cta[ctx.cta].smem; }
unsigned char* dyn_smem_of(int cta) { return g_cta[cta].smem; }
float* tmem() { return g_cta[ctx.cta].tmem; }
float* tmem_of(int cta) { return g_cta[cta].tmem; }
MBar& mbar_of(const void* p) {                                     // the barrier lives in the CTA whose shared memory holds it
  for (int c = 0; c < kMaxCluster; ++c) {
    const unsigned char* b = g_cta[c].smem;
    if (p >= (const void*)b && p < (const void*)(b + sizeof(g_cta[c].smem))) return g_cta[c].mbars[p];
  }
  fail("mbarrier outside shared memory");
}
std::mutex& mbar_lock() { return g_mbar_lock; }
Barrier& named_barrier(int id, int threads) {
  std::lock_guard<std::mutex> l(g_named_lock);
  auto& b = g_cta[ctx.cta].named[id];
  if (!b) {
    b.reset(new Barrier());
    b->expected = threads;
  }
  if (b->expected != threads) fail("bar.sync: the same barrier id used with two thread counts");
  return *b;
}
void note_wait(int thread, int id, int parity) {
  if (thread >= 0 && thread < 1024) {
    g_wait_id[ctx.cta][thread] = id;
    g_wait_parity[ctx.cta][thread] = parity;
  }
}
void dump_waits() {
  static std::mutex once;
  static bool done = false;
  std::lock_guard<std::mutex> l(once);
  if (done) return;
  done = true;
  std::string out;
  for (int c = 0; c < kMaxCluster; ++c) {
    std::map<std::pair<int, int>, std::string> groups;
    for (int t = 0; t < 1024; ++t)
      if (g_wait_id[c][t] >= 0) groups[{g_wait_id[c][t], g_wait_parity[c][t]}] += " " + std::to_string(t);
    if (groups.empty() && g_cta[c].mbars.empty()) continue;
    out += "emu: CTA " + std::to_string(c) + ": threads waiting on mbarriers (id/parity: threads):";
    for (auto& kv : groups) out += "\n  " + std::to_string(kv.first.first) + "/" + std::to_string(kv.first.second) + ":" + kv.second;
    out += "\nemu: CTA " + std::to_string(c) + ": mbarrier states (index: count pending tx phase):";
    std::lock_guard<std::mutex> l2(g_mbar_lock);
    int i = 0;
    for (auto& kv : g_cta[c].mbars)
      out += " [" + std::to_string(i++) + "] " + std::to_string(kv.second.count) + "/" + std::to_string(kv.second.pending) + "/" +
             std::to_string(kv.second.tx) + "/" + std::to_string(kv.second.phase);
    out += "\n";
  }
  std::fprintf(stderr, "%s", out.c_str());
  std::fflush(stderr);
}
void fail(const char* what) {
  std::fprintf(stderr, "emu: %s\n", what);
  std::fflush(stderr);
  std::abort();
}
CUresult encode_tiled(CUtensorMap* m, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* dims,
                      const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il,
                      CUtensorMapSwizzle swizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  if (rank != 3 || il != CU_TENSOR_MAP_INTERLEAVE_NONE || estr[0] != 1 || estr[1] != 1 || estr[2] != 1) return CUDA_ERROR_INVALID_VALUE;
  TensorMap t;
  std::memset(&t, 0, sizeof(t));
  t.magic = 0x554e495653ull;
  t.base = static_cast<const unsigned char*>(base);
  t.esize = dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT16 ? 2 : (dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 0);
  t.swizzle_bits = swizzle == CU_TENSOR_MAP_SWIZZLE_128B ? 3 : (swizzle == CU_TENSOR_MAP_SWIZZLE_64B ? 2 : (swizzle == CU_TENSOR_MAP_SWIZZLE_32B ? 1 : 0));
  if (t.esize == 0 || (reinterpret_cast<uintptr_t>(base) & 15)) return CUDA_ERROR_INVALID_VALUE;
  for (int i = 0; i < 3; ++i) {
    t.dims[i] = dims[i];
    t.box[i] = box[i];
    if (box[i] == 0 || box[i] > 256) return CUDA_ERROR_INVALID_VALUE;
  }
  for (int i = 0; i < 2; ++i) {
    t.strides[i] = strides[i];
    if (strides[i] % 16) return CUDA_ERROR_INVALID_VALUE;          // global strides must be multiples of 16 bytes
  }
  t.rank = 3;
  std::memset(m, 0, sizeof(*m));
  std::memcpy(m, &t, sizeof(t));
  return CUDA_SUCCESS;
}
int emulated_sm_count() {
  const char* e = std::getenv("UNIVS_EMU_SMS");      // few "SMs": persistent kernels walk many work units per CTA
  return (e != nullptr && std::atoi(e) > 0) ? std::atoi(e) : 148;
}
static void reset_cta_state(int c) {
  std::memset(g_cta[c].smem, 0xFF, sizeof(g_cta[c].smem));         // fp16 / fp32 NaN patterns: unwritten data is visible
  std::memset(g_cta[c].tmem, 0xFF, sizeof(g_cta[c].tmem));
  g_cta[c].mbars.clear();
  g_cta[c].named.clear();
  for (int t = 0; t < 1024; ++t) g_wait_id[c][t] = -1;
}

void launch(dim3 grid, dim3 block, const std::function<void()>& body, int cluster) {
  const int nthreads = (int)(block.x * block.y * block.z);
  const int nwarps = (nthreads + 31) / 32;
  if (cluster < 1 || cluster > kMaxCluster || grid.x % cluster) fail("launch: unsupported cluster shape");
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx0 = 0; bx0 < grid.x; bx0 += cluster) {
        Barrier cluster_bar;
        cluster_bar.expected = nthreads * cluster;
        std::vector<Barrier> block_bars(cluster);
        std::vector<std::vector<Warp>> warps(cluster);
        for (int c = 0; c < cluster; ++c) {
          reset_cta_state(c);
          block_bars[c].expected = nthreads;
          warps[c] = std::vector<Warp>(nwarps);
          for (int w = 0; w < nwarps; ++w) {
            warps[c][w].bar.expected = std::min(32, nthreads - 32 * w);
            std::memset(warps[c][w].slot, 0, sizeof(warps[c][w].slot));
          }
        }
        for (int c = 0; c < cluster; ++c) pipe_start(c);
        std::vector<std::thread> threads;
        threads.reserve((size_t)nthreads * cluster);
        for (int c = 0; c < cluster; ++c)
          for (int t = 0; t < nthreads; ++t) {
            threads.emplace_back([&, c, t] {
              Ctx& x = ctx;
              x.tid = make_uint3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
              x.bid = make_uint3(bx0 + c, by, bz);
              x.bdim = block;
              x.gdim = grid;
              x.block_bar = &block_bars[c];
              x.cluster_bar = &cluster_bar;
              x.warp = &warps[c][t >> 5];
              x.lane = t & 31;
              x.cta = c;
              x.cluster_size = cluster;
              body();
              x.warp->bar.drop();      // a returned thread no longer takes part in barriers / collectives
              x.block_bar->drop();
              x.cluster_bar->drop();
            });
          }
        for (auto& th : threads) th.join();
        for (int c = 0; c < cluster; ++c) pipe_finish(c);
      }
}

}  // namespace emu

extern "C" {
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
  std::memset(p, v, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  std::memcpy(d, s, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  std::memcpy(d, s, n);
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) {
  *d = 0;
  return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) {
  *v = emu::emulated_sm_count();
  return cudaSuccess;
}
cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* status) {
  if (std::strcmp(symbol, "cuTensorMapEncodeTiled") == 0) {
    *fn = reinterpret_cast<void*>(&emu::encode_tiled);
    if (status) *status = cudaDriverEntryPointSuccess;
    return cudaSuccess;
  }
  if (status) *status = cudaDriverEntryPointSymbolNotFound;
  return cudaErrorSymbolNotFound;
}
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  std::memset(a, 0, sizeof(*a));
  a->type = cudaMemoryTypeUnregistered;      // every pointer is a host pointer here
  a->hostPointer = const_cast<void*>(p);
  return cudaSuccess;
}
}
