// Runtime of the CPU emulation (see cuda_emu.h): block / warp scheduling and the handful of CUDA runtime entry points
// the launch wrappers call.
#include "cuda_emu.h"

namespace emu {

thread_local Ctx ctx;
std::mutex atomic_lock;

void launch(dim3 grid, dim3 block, const std::function<void()>& body) {
  const int nthreads = (int)(block.x * block.y * block.z);
  const int nwarps = (nthreads + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
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
  *v = 148;
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
