// Runtime of the CPU emulation (see cuda_emu.h): block / warp scheduling and the handful of CUDA runtime entry points
// the launch wrappers call.
#include "cuda_emu.h"

#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>

#include "tc05.cuh"

namespace emu {

thread_local Ctx ctx;
std::mutex atomic_lock;

// ---- per-CTA state of the tensor-core emulation (blocks run one after the other) ---------------------------------------
static constexpr size_t kSmemBytes = 232448;                       // 227 KB
alignas(1024) static unsigned char g_smem[kSmemBytes + 1024];
static float g_tmem[128 * 512];
static std::map<const void*, MBar> g_mbars;
static std::mutex g_mbar_lock, g_named_lock;
static std::map<int, std::unique_ptr<Barrier>> g_named;

unsigned char* dyn_smem() { return g_smem; }
float* tmem() { return g_tmem; }
MBar& mbar_of(const void* p) { return g_mbars[p]; }
std::mutex& mbar_lock() { return g_mbar_lock; }
Barrier& named_barrier(int id, int threads) {
  std::lock_guard<std::mutex> l(g_named_lock);
  auto& b = g_named[id];
  if (!b) {
    b.reset(new Barrier());
    b->expected = threads;
  }
  if (b->expected != threads) fail("bar.sync: the same barrier id used with two thread counts");
  return *b;
}
static int g_wait_id[1024], g_wait_parity[1024];
void note_wait(int thread, int id, int parity) {
  if (thread >= 0 && thread < 1024) {
    g_wait_id[thread] = id;
    g_wait_parity[thread] = parity;
  }
}
void dump_waits() {
  static std::mutex once;
  static bool done = false;
  std::lock_guard<std::mutex> l(once);
  if (done) return;
  done = true;
  std::string out = "emu: threads waiting on mbarriers (id/parity: threads):";
  std::map<std::pair<int, int>, std::string> groups;
  for (int t = 0; t < 1024; ++t)
    if (g_wait_id[t] >= 0) groups[{g_wait_id[t], g_wait_parity[t]}] += " " + std::to_string(t);
  for (auto& kv : groups) out += "\n  " + std::to_string(kv.first.first) + "/" + std::to_string(kv.first.second) + ":" + kv.second;
  out += "\nemu: mbarrier states (index: count pending tx phase):";
  {
    std::lock_guard<std::mutex> l2(g_mbar_lock);
    int i = 0;
    for (auto& kv : g_mbars)
      out += " [" + std::to_string(i++) + "] " + std::to_string(kv.second.count) + "/" + std::to_string(kv.second.pending) + "/" +
             std::to_string(kv.second.tx) + "/" + std::to_string(kv.second.phase);
  }
  std::fprintf(stderr, "%s\n", out.c_str());
  std::fflush(stderr);
}
void fail(const char* what) {
  std::fprintf(stderr, "emu: %s\n", what);
  std::fflush(stderr);
  std::abort();
}
static void reset_cta_state() {
  std::memset(g_smem, 0xFF, sizeof(g_smem));                       // fp16 / fp32 NaN patterns: unwritten data is visible
  std::memset(g_tmem, 0xFF, sizeof(g_tmem));
  g_mbars.clear();
  g_named.clear();
  for (int t = 0; t < 1024; ++t) g_wait_id[t] = -1;
}

void launch(dim3 grid, dim3 block, const std::function<void()>& body) {
  const int nthreads = (int)(block.x * block.y * block.z);
  const int nwarps = (nthreads + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        reset_cta_state();
        Barrier block_bar;
        block_bar.expected = nthreads;
        std::vector<Warp> warps(nwarps);
        for (int w = 0; w < nwarps; ++w) {
          warps[w].bar.expected = std::min(32, nthreads - 32 * w);
          std::memset(warps[w].slot, 0, sizeof(warps[w].slot));
        }
        std::vector<std::thread> threads;
        threads.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t) {
          threads.emplace_back([&, t] {
            Ctx& c = ctx;
            c.tid = make_uint3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            c.bid = make_uint3(bx, by, bz);
            c.bdim = block;
            c.gdim = grid;
            c.block_bar = &block_bar;
            c.warp = &warps[t >> 5];
            c.lane = t & 31;
            body();
            c.warp->bar.drop();      // a returned thread no longer takes part in barriers / collectives
            block_bar.drop();
          });
        }
        for (auto& th : threads) th.join();
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
  const char* e = std::getenv("UNIVS_EMU_SMS");      // few "SMs": persistent kernels walk many work units per CTA
  *v = (e != nullptr && std::atoi(e) > 0) ? std::atoi(e) : 148;
  return cudaSuccess;
}
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  std::memset(a, 0, sizeof(*a));
  a->type = cudaMemoryTypeUnregistered;      // every pointer is a host pointer here
  a->hostPointer = const_cast<void*>(p);
  return cudaSuccess;
}
}
