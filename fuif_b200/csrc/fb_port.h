// Portability shim for kernels that are also executed by the CPU-side execution-model emulator of the test tier
// (tests/emu/cuemu.h, compiled only by the tests with -DFB_EMULATE).  Under nvcc these macros are plain CUDA.
#pragma once
#include <stdint.h>

#if defined(FB_EMULATE)
#include "cuemu.h"
#define FB_HD inline
#define FB_DEV inline
#define FB_KERNEL(maxthreads) inline void
#define FB_GRID_CONSTANT
#define FB_DYN_SMEM(name) unsigned char *name = cuemu::S().dyn_smem
inline void fb_bar_sync(int id, int count) { cuemu::bar_sync(id, count); }
inline void fb_grid_sync() { cuemu::grid_sync(); }
#else
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#define FB_HD __host__ __device__ __forceinline__
#define FB_DEV __device__ __forceinline__
#define FB_KERNEL(maxthreads) __global__ void __launch_bounds__(maxthreads)
#define FB_GRID_CONSTANT __grid_constant__
#define FB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
// named barrier: `count` threads (a multiple of 32) of the block meet on hardware barrier `id` (1..15)
__device__ __forceinline__ void fb_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void fb_grid_sync() { cooperative_groups::this_grid().sync(); }
#endif
