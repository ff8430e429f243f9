// inv_match (reference transform/2dmatch.h:97-177).  Exact matches: in scanline order every sample whose match code z is not
// zero takes the already reconstructed sample at the 2D offset that z encodes -- a serial chain in the reference, because
// the source may itself be a matched sample.  On the GPU it is pointer jumping: every sample points at its source
// (k_match_parent), path doubling replaces "my source" by "my source's source" until everything points at a sample that is
// not matched (k_match_jump, log2(longest chain) rounds, double-buffered, stopped once nothing moves), and one gather per channel copies the roots'
// values (k_match_gather, out of place).
//
// parent[i] >= 0   : index of the sample this one copies (itself: not matched)
// parent[i] == -1  : the offset leaves the plane: the reference reads the channel's `zero` (Channel::value, image.h:82)
// parent[i] <= -2  : the offset points at sample -(parent+2) >= i, which the reference has not reconstructed yet when it
//                    reads it: the value is that sample's ORIGINAL content (degenerate streams only; terminal like -1)
// Channel::value is flat-indexed (r*w + c), so an offset that leaves the row lands in a neighbouring row: the source index
// is simply i + dy*w + dx.
//
// Compiled by nvcc (product) and by g++ -DFB_EMULATE (tests/emu: CPU execution-model emulator vs reference vectors).
#pragma once
#include "fb_port.h"

namespace mt {

// compute_offset, 2dmatch.h:52-77: the codes spiral outwards around the current sample in "onion layers" of 4, 8, 12, .. codes
FB_HD void match_offset(int code, int &dx, int &dy) {
    int layer = 0, size = 4;
    while (code > size) { code -= size; layer++; size += 4; }
    if (layer & 1) {
        if (code <= layer) { dx = 1 + layer; dy = -code; }
        else if (code <= 3 + 3 * layer) { dx = 2 + 2 * layer - code; dy = -1 - layer; }
        else { dx = -1 - layer; dy = -4 - 4 * layer + code; }
    } else {
        if (code <= 1 + layer) { dx = -1 - layer; dy = 1 - code; }
        else if (code <= 4 + 3 * layer) { dx = -3 - 2 * layer + code; dy = -1 - layer; }
        else { dx = 1 + layer; dy = -5 - 4 * layer + code; }
    }
}

// *bad is set when a code lies outside [0, maxcode] (the reference would index its offset table out of bounds)
FB_KERNEL(256) k_match_parent(const int16_t *m, int *parent, int n, int w, int maxcode, int *bad) {
    const int i = (int)((size_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n) return;
    const int z = m[i];
    if (z == 0) { parent[i] = i; return; }
    if (z < 0 || z > maxcode) { *bad = 1; parent[i] = i; return; }
    int dx, dy;
    match_offset(z, dx, dy);
    const long long src = (long long)i + (long long)dy * w + dx;
    if (src < 0 || src >= n) parent[i] = -1;
    else if (src >= i) parent[i] = -2 - (int)src;
    else parent[i] = (int)src;
}

// *changed is set when some sample moved: the host stops the rounds as soon as a whole batch of them changed nothing
FB_KERNEL(256) k_match_jump(const int *in, int *out, int n, int *changed) {
    const int i = (int)((size_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n) return;
    const int p = in[i];
    if (p < 0 || p == i) { out[i] = p; return; }
    const int g = in[p];
    const int q = (g == p) ? p : g;     // my source is a root: done; otherwise adopt what my source points at (an index or a terminal code)
    out[i] = q;
    if (q != p) *changed = 1;
}

FB_KERNEL(256) k_match_gather(const int16_t *src, int16_t *dst, const int *parent, int n, int zero) {
    const int i = (int)((size_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n) return;
    const int p = parent[i];
    dst[i] = p >= 0 ? src[p] : (p == -1 ? (int16_t)zero : src[-2 - p]);
}

// ---- soft matches (2dmatch.h:119-129): value[i] += value[source] instead of a copy ------------------------------------------
// The same chains, but what travels along them is a sum: acc[i] starts as the sample's own (residual) value, a round adds
// the source's acc and adopts the source's parent -- int16 wrap-around addition is associative, so the order of the
// additions does not matter.  Samples whose source is a terminal (outside the plane: `zero`; a forward reference: that
// sample's original content) are roots whose value already includes the terminal.  One channel at a time.
FB_KERNEL(256) k_match_soft_init(const int16_t *m, const int16_t *orig, int *parent, int16_t *acc, int n, int w, int maxcode, int zero, int *bad) {
    const int i = (int)((size_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n) return;
    const int z = m[i];
    const int own = orig[i];
    if (z == 0) { parent[i] = i; acc[i] = (int16_t)own; return; }
    if (z < 0 || z > maxcode) { *bad = 1; parent[i] = i; acc[i] = (int16_t)own; return; }
    int dx, dy;
    match_offset(z, dx, dy);
    const long long src = (long long)i + (long long)dy * w + dx;
    if (src < 0 || src >= n) { parent[i] = i; acc[i] = (int16_t)(own + zero); }
    else if (src >= i) { parent[i] = i; acc[i] = (int16_t)(own + orig[src]); }
    else { parent[i] = (int)src; acc[i] = (int16_t)own; }
}
FB_KERNEL(256) k_match_soft_jump(const int *pin, const int16_t *ain, int *pout, int16_t *aout, int n, int *changed) {
    const int i = (int)((size_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n) return;
    const int p = pin[i];
    const int a = ain[i];
    if (p == i) { pout[i] = p; aout[i] = (int16_t)a; return; }
    const int g = pin[p];
    if (g == p) { pout[i] = p; aout[i] = (int16_t)a; return; }         // my source is a root: its value is added at the end
    pout[i] = g;
    aout[i] = (int16_t)(a + ain[p]);
    *changed = 1;
}
FB_KERNEL(256) k_match_soft_finish(const int *parent, const int16_t *acc, int16_t *dst, int n) {
    const int i = (int)((size_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n) return;
    const int p = parent[i];
    dst[i] = p == i ? acc[i] : (int16_t)(acc[i] + acc[p]);
}

}  // namespace mt
