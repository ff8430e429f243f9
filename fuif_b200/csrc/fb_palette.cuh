// Palette (reference transform/palette.h).
// inv_palette (:32-68): an index plane + a palette (nb rows of ncolors samples) -> nb planes.
// One thread per sample: the index is clamped into the palette (palette.h:58), each output plane gets its row's entry; the
// index plane itself is output plane 0 (read before it is written, by the same thread).  The palette is a few KB and stays
// in L1/L2; the planes stream through HBM once (2 bytes in, 2*nb bytes out per sample).
//
// Compiled by nvcc (product) and by g++ -DFB_EMULATE (tests/emu: CPU execution-model emulator vs reference vectors).
#pragma once
#include "fb_port.h"

namespace pl {

constexpr int kMaxPlanes = 8;
struct Planes { int16_t *p[kMaxPlanes]; };

FB_KERNEL(256) k_palette_inv(Planes out, const int16_t *palette, size_t n, int ncolors, int nb) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int index = out.p[0][i];
    index = index < 0 ? 0 : (index > ncolors - 1 ? ncolors - 1 : index);
    for (int c = 0; c < nb; c++) out.p[c][i] = palette[(size_t)c * ncolors + index];
}

// ---- fwd_palette (palette.h:92-143) -------------------------------------------------------------------------------------
// The reference collects the colours in use in a std::set of vectors (lexicographic order) and gives up when there are more
// than the caller allows.  Here: (1) every sample's colour, packed into one 64-bit key whose integer order IS the
// lexicographic order of the signed tuple, goes into an open-addressing hash set in HBM (atomicCAS; the set is a few
// thousand slots and lives in L2); a counter of distinct colours stops the launch early when the limit is passed;
// (2) the host sorts the handful of keys (that is the palette); (3) every sample looks its key up by binary search.
// Up to four channels (16 bits each).

constexpr unsigned long long kEmptySlot = ~0ull;

// channel 0 most significant; +32768 maps int16 order onto unsigned order
FB_HD unsigned long long pack_colour(const int *v, int nb) {
    unsigned long long k = 0;
    for (int c = 0; c < 4; c++) k = (k << 16) | (c < nb ? (unsigned long long)(unsigned)(v[c] + 32768) & 0xffffu : 0ull);
    return k;
}
FB_HD int unpack_colour(unsigned long long k, int c) { return (int)((k >> (16 * (3 - c))) & 0xffffu) - 32768; }
FB_HD unsigned hash_colour(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}

struct Collect {
    unsigned long long *table;  // cap slots, all kEmptySlot before the launch
    unsigned cap_mask;          // cap - 1 (cap is a power of two)
    int limit;                  // more distinct colours than this: give up
    int *count;                 // distinct colours inserted (the all-ones key is counted through *has_allones)
    int *has_allones;           // the one key that equals the empty marker
    int *overflow;
};

FB_KERNEL(256) k_palette_collect(Planes in, size_t n, int nb, Collect C) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (*(volatile int *)C.overflow) return;
    int v[4] = {0, 0, 0, 0};
    for (int c = 0; c < nb; c++) v[c] = in.p[c][i];
    const unsigned long long key = pack_colour(v, nb);
    if (key == kEmptySlot) {
        if (atomicExch(C.has_allones, 1) == 0 && atomicAdd(C.count, 1) + 1 > C.limit) atomicExch(C.overflow, 1);
        return;
    }
    unsigned h = hash_colour(key) & C.cap_mask;
    for (unsigned probes = 0; probes <= C.cap_mask; probes++) {
        const unsigned long long cur = C.table[h];
        if (cur == key) return;
        if (cur == kEmptySlot) {
            const unsigned long long old = atomicCAS(&C.table[h], kEmptySlot, key);
            if (old == kEmptySlot) {
                if (atomicAdd(C.count, 1) + 1 > C.limit) atomicExch(C.overflow, 1);
                return;
            }
            if (old == key) return;
        }
        h = (h + 1) & C.cap_mask;
    }
    atomicExch(C.overflow, 1);          // table full: far more colours than the limit
}

// every sample's position in the sorted palette; written over plane 0 (the thread has read its own sample of every plane)
FB_KERNEL(256) k_palette_index(Planes in, size_t n, int nb, const unsigned long long *sorted, int count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v[4] = {0, 0, 0, 0};
    for (int c = 0; c < nb; c++) v[c] = in.p[c][i];
    const unsigned long long key = pack_colour(v, nb);
    int lo = 0, hi = count;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted[mid] < key) lo = mid + 1; else hi = mid; }
    in.p[0][i] = (int16_t)lo;
}

}  // namespace pl
