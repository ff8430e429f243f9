// Approximate (reference transform/approximate.h): a channel is split into quotient and remainder of a floor division by
// q (forward, :83-113) and put back together as quotient * q + remainder with int16 wrap at both steps (inverse, :32-62).
// Elementwise, one sample per thread.
//
// Compiled by nvcc (product) and by g++ -DFB_EMULATE (tests/emu: CPU execution-model emulator vs reference vectors).
#pragma once
#include "fb_port.h"

namespace ap {

FB_DEV int s16w(int x) { return (int)(short)x; }

// chr == nullptr: the remainder channel was not decoded (partial decode): nothing is added (approximate.h:55)
FB_KERNEL(256) k_approx_inv(int16_t *ch, const int16_t *chr, size_t n, int q) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v = s16w(ch[i] * q);
    v = s16w(v + (chr ? (int)chr[i] : 0));
    ch[i] = (int16_t)v;
}

FB_KERNEL(256) k_approx_fwd(int16_t *ch, int16_t *chr, size_t n, int q) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = ch[i];
    int quotient = s16w(p / q), r = s16w(p % q);
    if (r < 0) { quotient = s16w(quotient - 1); r = s16w(r + q); }
    ch[i] = (int16_t)quotient;
    chr[i] = (int16_t)r;
}

}  // namespace ap
