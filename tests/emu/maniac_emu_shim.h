// TEST INFRASTRUCTURE ONLY.
//
// Lets tests/emu/emu_maniac.cpp compile the product's MANIAC decode kernel SOURCE (fuif_b200/csrc/fb_maniac.cu, device part)
// for the CPU execution-model emulator (cuemu.h): one fibre per CUDA thread, warp collectives as 32-fibre rendezvous,
// "shared-memory addresses" = byte offsets into the block's emulated shared memory, every spin-wait poll yields to the
// other fibres.  What this checks is the kernel's LOGIC (stream tickets, row wavefront, walker / prologue / decoder
// protocol, tree walks with forks, leaf cache, integer coder) against the oracle in the CPU-only test tier.  It does not
// model SIMT lockstep: the decoder's redundant all-lanes pixel loop runs on lane 0 only here (see FB_UNIFORM_LOOP_LANE0).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "cuemu.h"
#include "../../include/fuif_b200.h"

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n)
#define FB_DYN_SMEM_DECL(name) unsigned char *name = cuemu::S().dyn_smem

#pragma push_macro("threadIdx")
#pragma push_macro("blockIdx")
#pragma push_macro("blockDim")
#pragma push_macro("gridDim")
#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim
namespace cuemu {
// ---- warp collectives: the 32 fibres of a warp meet twice (publish, consume) -------------------------------------
struct WarpX { Barrier bar; unsigned long long v[32]; };
inline WarpX &warpx() {
    static std::vector<WarpX> w;
    State &s = S();
    const size_t nwarps = (size_t)s.gridDim.x * ((s.blockDim.x + 31) / 32);
    if (w.size() < nwarps) w.resize(nwarps);
    return w[(size_t)s.blockIdx.x * ((s.blockDim.x + 31) / 32) + s.threadIdx.x / 32];
}
inline void warp_barrier() { barrier_wait(warpx().bar, 32, 0); }
template <class T> inline T shfl(T v, int src) {
    WarpX &w = warpx();
    unsigned long long x = 0;
    memcpy(&x, &v, sizeof(T));
    w.v[S().threadIdx.x & 31] = x;
    warp_barrier();
    const unsigned long long r = w.v[src & 31];
    warp_barrier();
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
inline unsigned ballot(int pred) {
    WarpX &w = warpx();
    w.v[S().threadIdx.x & 31] = pred ? 1 : 0;
    warp_barrier();
    unsigned m = 0;
    for (int i = 0; i < 32; i++) if (w.v[i]) m |= 1u << i;
    warp_barrier();
    return m;
}
inline unsigned reduce_max(unsigned v) {
    WarpX &w = warpx();
    w.v[S().threadIdx.x & 31] = v;
    warp_barrier();
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m = std::max(m, (unsigned)w.v[i]);
    warp_barrier();
    return m;
}
}  // namespace cuemu
#pragma pop_macro("gridDim")
#pragma pop_macro("blockDim")
#pragma pop_macro("blockIdx")
#pragma pop_macro("threadIdx")

template <class T> inline T __shfl_sync(unsigned, T v, int src) { return cuemu::shfl(v, src); }
inline unsigned __ballot_sync(unsigned, int p) { return cuemu::ballot(p); }
inline int __any_sync(unsigned, int p) { return cuemu::ballot(p) != 0; }
inline unsigned __reduce_max_sync(unsigned, unsigned v) { return cuemu::reduce_max(v); }
inline void __syncwarp() { cuemu::warp_barrier(); }
inline void __nanosleep(unsigned) { cuemu::yield(); cuemu::load_identity(cuemu::S().fibres[cuemu::S().current]); }
inline void fb_emu_yield() { __nanosleep(0); }
inline void __threadfence_block() {}
inline long long clock64() { return 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    const unsigned long long x = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((x >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
template <class T> inline T __ldcg(const T *p) { return *(const volatile T *)p; }
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 v; v.x = x; v.y = y; return v; }
inline int atomicAdd_emu(int *p, int v) { int o = *p; *p = o + v; return o; }
// shared-memory "addresses" are byte offsets into the block's emulated shared memory
inline unsigned __cvta_generic_to_shared(const void *p) { return (unsigned)((const unsigned char *)p - cuemu::S().dyn_smem); }
inline unsigned char *fb_emu_smem(unsigned addr) { return cuemu::S().dyn_smem + addr; }
