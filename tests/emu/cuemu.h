// TEST INFRASTRUCTURE ONLY.
//
// A tiny CUDA execution-model emulator: runs a kernel written for fuif_b200 (through the FB_* portability macros of
// fuif_b200/csrc/fb_port.h) on the CPU, one fibre per CUDA thread, so that the kernels' index arithmetic, barrier
// protocol and exactness logic can be checked against the oracle in the CPU-only test tier (there is no GPU in the
// build container).  It is NOT a fallback: nothing under fuif_b200/ links or includes it; only tests/emu builds it.
//
// Model: threads of a block (or, for cooperative launches, of the whole grid) are ucontext fibres scheduled round
// robin on one OS thread; a fibre runs until it reaches a barrier, where it yields until the barrier generation
// changes.  Block barriers (__syncthreads, __syncthreads_or, named bar.sync id,count) and the grid barrier are
// supported.  Global-memory atomics are plain operations (one OS thread).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <vector>

namespace cuemu {

struct Dim3 { unsigned x = 1, y = 1, z = 1; Dim3() {} Dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

struct Barrier { unsigned arrived = 0, gen = 0; int or_acc = 0, or_res = 0; };

struct Block {
    Barrier bars[16];
    unsigned char *smem = nullptr;
};

struct Fibre {
    ucontext_t uc;
    char *stack = nullptr;
    bool done = false;
    unsigned tid = 0, bid = 0;
    Block *blk = nullptr;
};

struct State {
    Dim3 threadIdx, blockIdx, blockDim, gridDim;
    unsigned char *dyn_smem = nullptr;
    Block *blk = nullptr;
    Barrier grid_bar;
    unsigned grid_threads = 0;
    std::vector<Fibre> fibres;
    int current = -1;
    ucontext_t sched;
    std::function<void()> body;
};
inline State &S() { static State s; return s; }

inline void fibre_entry() {
    State &s = S();
    s.body();
    s.fibres[s.current].done = true;
    swapcontext(&s.fibres[s.current].uc, &s.sched);
}

inline void yield() {
    State &s = S();
    swapcontext(&s.fibres[s.current].uc, &s.sched);
}

inline void load_identity(const Fibre &f) {
    State &s = S();
    s.threadIdx = Dim3(f.tid);
    s.blockIdx = Dim3(f.bid);
    s.blk = f.blk;
    s.dyn_smem = f.blk->smem;
}

// returns the OR of `pred` over the participants
inline int barrier_wait(Barrier &b, unsigned count, int pred) {
    b.or_acc |= pred;
    b.arrived++;
    const unsigned gen = b.gen;
    if (b.arrived == count) {
        b.arrived = 0;
        b.or_res = b.or_acc;
        b.or_acc = 0;
        b.gen++;
    } else {
        State &s = S();
        const int me = s.current;
        while (b.gen == gen) yield();
        load_identity(s.fibres[me]);
    }
    return b.or_res;
}

inline void syncthreads() { State &s = S(); barrier_wait(s.blk->bars[0], s.blockDim.x, 0); }
inline int syncthreads_or(int p) { State &s = S(); return barrier_wait(s.blk->bars[0], s.blockDim.x, p != 0) != 0; }
inline void bar_sync(int id, int count) { State &s = S(); barrier_wait(s.blk->bars[id], (unsigned)count, 0); }
inline void grid_sync() { State &s = S(); barrier_wait(s.grid_bar, s.grid_threads, 0); }

// Runs `body` once per CUDA thread.  cooperative: all blocks are resident together (grid barrier usable).
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, bool cooperative, std::function<void()> body) {
    State &s = S();
    s.body = body;
    s.gridDim = Dim3(grid);
    s.blockDim = Dim3(block);
    const size_t kStack = 96 * 1024;
    const unsigned blocks_at_once = cooperative ? grid : 1;
    s.grid_threads = grid * block;
    s.grid_bar = Barrier();
    std::vector<Block> blocks(blocks_at_once);
    for (auto &b : blocks) b.smem = (unsigned char *)aligned_alloc(128, (smem_bytes + 255) & ~(size_t)127);
    std::vector<char *> stacks(blocks_at_once * (size_t)block);
    for (auto &st : stacks) st = (char *)malloc(kStack);
    for (unsigned b0 = 0; b0 < grid; b0 += blocks_at_once) {
        const unsigned nb = std::min(blocks_at_once, grid - b0);
        s.fibres.assign((size_t)nb * block, Fibre());
        for (unsigned b = 0; b < nb; b++) {
            for (int k = 0; k < 16; k++) blocks[b].bars[k] = Barrier();
            memset(blocks[b].smem, 0xCD, smem_bytes);      // poison: uninitialised shared memory must not matter
            for (unsigned t = 0; t < block; t++) {
                Fibre &f = s.fibres[(size_t)b * block + t];
                f.tid = t; f.bid = b0 + b; f.blk = &blocks[b];
                f.stack = stacks[(size_t)b * block + t];
                getcontext(&f.uc);
                f.uc.uc_stack.ss_sp = f.stack;
                f.uc.uc_stack.ss_size = kStack;
                f.uc.uc_link = &s.sched;
                makecontext(&f.uc, (void (*)())fibre_entry, 0);
            }
        }
        // CUEMU_ORDER=reverse|shuffle changes the order in which the fibres get their turns: a kernel whose results depend on
        // which thread of a barrier interval runs first has a race that the default order hides
        static const char *order_env = getenv("CUEMU_ORDER");
        const bool reverse = order_env && order_env[0] == 'r', shuffle = order_env && order_env[0] == 's';
        unsigned long long lcg = 0x9E3779B97F4A7C15ull;
        size_t remaining = s.fibres.size();
        while (remaining) {
            const size_t nf = s.fibres.size();
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            const size_t rot = shuffle ? (size_t)(lcg >> 33) % nf : 0;
            for (size_t k = 0; k < nf; k++) {
                const size_t i = reverse ? nf - 1 - k : (k + rot) % nf;
                Fibre &f = s.fibres[i];
                if (f.done) continue;
                s.current = (int)i;
                load_identity(f);
                swapcontext(&s.sched, &f.uc);
                if (f.done) remaining--;
            }
        }
    }
    for (auto st : stacks) free(st);
    for (auto &b : blocks) free(b.smem);
    s.fibres.clear();
}

}  // namespace cuemu

// ---- the CUDA vocabulary the kernels use -------------------------------------------------------------------------
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
#define threadIdx (cuemu::S().threadIdx)
#define blockIdx (cuemu::S().blockIdx)
#define blockDim (cuemu::S().blockDim)
#define gridDim (cuemu::S().gridDim)
inline void __syncthreads() { cuemu::syncthreads(); }
inline int __syncthreads_or(int p) { return cuemu::syncthreads_or(p); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
template <class T> inline T __ldg(const T *p) { return *p; }
inline int atomicOr(int *p, int v) { int o = *p; *p = o | v; return o; }
inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; return o; }
inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v) { unsigned long long o = *p; if (o == cmp) *p = v; return o; }
inline int atomicExch(int *p, int v) { int o = *p; *p = v; return o; }
inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
using std::max;
using std::min;
