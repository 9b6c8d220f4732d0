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
#else
// ------------------------------------------------------------------------------------------ EMU
#include "../../tests/emu/emu_runtime.hpp"
#endif

typedef uint32_t u32;
typedef uint64_t u64;
