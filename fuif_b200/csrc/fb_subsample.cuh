// inv_subsample (reference transform/subsample.h:73-128): chroma upscaling.  Ratios <= 2: the reference's "fancy" separable
// filter -- horizontally (3*cur + left + 1) >> 2 for the even output column and (3*cur + right + 2) >> 2 for the odd one
// (edge samples repeat), the result stored as int16, then the same vertically over those rows.  Other ratios: box
// replication.  One thread per output sample: the (up to) six input samples it needs sit in two input rows that the
// neighbouring threads read as well, so the plane is fetched from HBM once; the writes are coalesced.
//
// Compiled by nvcc (product) and by g++ -DFB_EMULATE (tests/emu: CPU execution-model emulator vs the oracle).
#pragma once
#include "fb_port.h"

namespace sb {

FB_DEV int s16w(int x) { return (int)(short)x; }

// sample X of row y of the horizontally upscaled plane
FB_DEV int sub_h(const int16_t *in, int ow, int y, int X, int srh) {
    if (srh != 2) return in[(size_t)y * ow + X];
    const int x = X >> 1;
    const int cur = in[(size_t)y * ow + x];
    if (X & 1) return s16w((3 * cur + in[(size_t)y * ow + (x + 1 < ow ? x + 1 : x)] + 2) >> 2);
    return s16w((3 * cur + in[(size_t)y * ow + (x ? x - 1 : 0)] + 1) >> 2);
}

FB_KERNEL(256) k_inv_subsample(const int16_t *in, int16_t *out, int ow, int oh, int srh, int srv) {
    const int W = ow * srh, H = oh * srv;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)W * H) return;
    const int Y = (int)(i / W), X = (int)(i - (size_t)Y * W);
    int v;
    if (srh > 2 || srv > 2) v = in[(size_t)(Y / srv) * ow + X / srh];
    else if (srv != 2) v = sub_h(in, ow, Y, X, srh);
    else {
        const int y = Y >> 1;
        const int cur = sub_h(in, ow, y, X, srh);
        if (Y & 1) v = s16w((3 * cur + sub_h(in, ow, y + 1 < oh ? y + 1 : y, X, srh) + 2) >> 2);
        else v = s16w((3 * cur + sub_h(in, ow, y ? y - 1 : 0, X, srh) + 1) >> 2);
    }
    out[i] = (int16_t)v;
}

}  // namespace sb
