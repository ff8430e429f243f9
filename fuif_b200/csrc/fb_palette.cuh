// inv_palette (reference transform/palette.h:32-68): an index plane + a palette (nb rows of ncolors samples) -> nb planes.
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

}  // namespace pl
