// emu_runtime.hpp -- DEVELOPMENT TOOLING ONLY (see aleo_b200/csrc/platform.cuh).
// Lets g++ compile the CUDA kernel sources and run them with one OS thread per CUDA thread so
// that indexing / synchronisation logic can be debugged in a container without a GPU.
// Never linked into libaleo_b200.so, never imported by the aleo_b200 package, never timed.
#pragma once
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define DEV inline
#define DEV_NOINLINE inline
#define HOSTDEV inline
#define KERNEL static
#define CONSTFN inline constexpr
#define DEVCONST const
#define SHARED static
#define __restrict__
#define __launch_bounds__(...)

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(uint32_t a, uint32_t b) { return uint2{a, b}; }

namespace emu {
inline thread_local dim3 t_threadIdx, t_blockIdx;
inline dim3 g_blockDim, g_gridDim;
inline unsigned char* g_smem = nullptr;
inline std::barrier<>* g_barrier = nullptr;
inline thread_local uint32_t t_cf = 0;  // PTX carry flag CC.CF

template <class F>
void launch(dim3 grid, dim3 block, size_t smem, bool needs_sync, F body) {
  g_blockDim = block;
  g_gridDim = grid;
  const unsigned nthreads = block.x * block.y * block.z;
  const unsigned nblocks = grid.x * grid.y * grid.z;
  if (nblocks == 0 || nthreads == 0) return;
  if (!needs_sync) {  // blocks in parallel over host cores, CUDA threads of a block in sequence
    unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::atomic<unsigned> next{0};
    std::vector<std::thread> pool;
    for (unsigned h = 0; h < hw; h++)
      pool.emplace_back([&] {
        for (;;) {
          unsigned b = next.fetch_add(1);
          if (b >= nblocks) break;
          t_blockIdx = dim3(b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y));
          for (unsigned t = 0; t < nthreads; t++) {
            t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            body();
          }
        }
      });
    for (auto& th : pool) th.join();
    return;
  }
  std::vector<unsigned char> smem_buf(smem + 64);
  g_smem = smem_buf.data();
  std::barrier<> bar(nthreads);
  g_barrier = &bar;
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nthreads; t++)
    pool.emplace_back([&, t] {
      t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
      for (unsigned b = 0; b < nblocks; b++) {
        t_blockIdx = dim3(b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y));
        body();
        bar.arrive_and_wait();  // block boundary: shared memory is reused by the next block
      }
    });
  for (auto& th : pool) th.join();
  g_barrier = nullptr;
  g_smem = nullptr;
}
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::g_smem)
#define SYNC_THREADS() emu::g_barrier->arrive_and_wait()
#define SYNC_WARP() ((void)0)
#define LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), true, [=]() { kern(__VA_ARGS__); })
#define LAUNCH_NOSYNC(kern, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), false, [=]() { kern(__VA_ARGS__); })

inline uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline void fence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void store_release_sys_u32(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline uint32_t load_acquire_sys_u32(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void spin_pause() {}
inline uint32_t atomic_add_shared_u32(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }

// ---- the handful of runtime calls the host drivers use ----------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { std::free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
enum { cudaStreamNonBlocking = 1 };
typedef void* cudaEvent_t;
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
template <class T> static inline cudaError_t cudaFuncSetAttribute(T, int, int) { return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
