// platform.cuh -- one source, two compilers.
//
// Product build: nvcc, sm_100a only.  Every kernel in this directory is hand-written CUDA; there
// is no CPU fallback in the product library (capi.cu fails loudly when no B200 is present).
//
// ALEO_EMU build: the dev container has no GPU, so the *same kernel source* can also be compiled
// by g++ into a lock-step-free thread emulator (tests/emu/) to debug indexing logic before GPU
// time is spent.  The emulator is development tooling: it is never linked into libaleo_b200.so,
// never imported by the aleo_b200 package and never timed.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>

#ifndef ALEO_EMU
// ------------------------------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
#define DEV __device__ __forceinline__
#define DEV_NOINLINE static __device__ __noinline__
#define HOSTDEV __host__ __device__ __forceinline__
#define KERNEL __global__
#define CONSTFN __host__ __device__ __forceinline__ constexpr
#define DEVCONST __constant__
#define SHARED __shared__
#define DYN_SMEM(type, name)                                   \
  extern __shared__ __align__(16) unsigned char _dyn_smem[];   \
  type* name = reinterpret_cast<type*>(_dyn_smem)
#define LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define LAUNCH_NOSYNC LAUNCH
#define SYNC_THREADS() __syncthreads()
#define SYNC_WARP() __syncwarp()
DEV uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
DEV uint32_t atomic_add_shared_u32(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
// cross-GPU flags (peer memory over NVLink): release store / acquire load at system scope
DEV void fence_system() { __threadfence_system(); }
DEV void store_release_sys_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
DEV uint32_t load_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
DEV void spin_pause() { __nanosleep(64); }
#else
// ------------------------------------------------------------------------------------------ EMU
#include "../../tests/emu/emu_runtime.hpp"
#endif

typedef uint32_t u32;
typedef uint64_t u64;

// Part constants of the calling thread's current device, read once per device from the runtime (grids are sized in
// multiples of the SM count; the sort picks its pass structure from the L2 size).  The emulator reports a B200.
struct DevProps {
  u32 sms;
  size_t l2_bytes;
};
#ifndef ALEO_EMU
static inline DevProps dev_props() {
  static DevProps cache[64];
  static bool have[64];  // benign race: every thread computes the same values
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return DevProps{148u, (size_t)126 << 20};
  if (!have[d]) {
    int sms = 0, l2 = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || sms <= 0) sms = 148;
    if (cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, d) != cudaSuccess || l2 <= 0) l2 = 126 << 20;
    cache[d] = DevProps{(u32)sms, (size_t)l2};
    have[d] = true;
  }
  return cache[d];
}
#else
static inline DevProps dev_props() { return DevProps{148u, (size_t)126 << 20}; }
#endif
// bucket heads / tables that should stay L2 resident next to the streaming traffic: 3/8 of the L2 (47 MB on B200)
static inline size_t l2_resident_budget() { return dev_props().l2_bytes / 8 * 3; }
