// "Direct" per-step inverse Squeeze kernels: register-resident chains, 128-bit global accesses, no shared-memory
// staging.  Reference semantics: transform/squeeze.h:81-132 (inv_hsqueeze), :173-224 (inv_vsqueeze), :61-77
// (smooth_tendency); optional epilogue transform/ycocg.h:51-56 (inv_YCoCg) + image/image.cpp:107-113 (final clamp).
//
// The inverse is a serial recurrence along the squeeze axis (pair p needs the reconstructed B of pair p-1), so every
// chain (row / column) is cut into segments that start 8 pairs early from a guessed state; the recurrence forgets its
// start within a few pairs.  A block owns WHOLE chains: after the first pass every segment compares the
// state it assumed with the true final state of its predecessor (through shared memory, a few bytes per segment) and
// is recomputed from the true state if they differ, until all agree.  Segment 0 starts at the true chain start, so by
// induction the output is bit-exact whatever the guesses were; the speculation only buys parallelism.
//
//   horizontal: thread = (row, segment of S pairs); a chunk = 8 pairs: one 16-byte load of averages, one of residuals,
//               two 16-byte stores (or, with the colour epilogue, Y is loaded and R, G, B are stored).  Up to two
//               planes of identical geometry per thread (Co and Cg) for ILP.
//   vertical:   thread = (8 adjacent columns, segment of S pairs): one 16-byte load per input row, eight independent
//               chains in registers (ILP 8), two 16-byte stores per pair of output rows.
//
// Compiled by nvcc (product) and by g++ -DFB_EMULATE (tests/emu: CPU execution-model emulator vs the oracle).
#pragma once
#include "fb_fused_squeeze.cuh"     // fq::unsqueeze_pair, clampi, FB_* macros

namespace dq {

using fq::clampi;
using fq::imin;

constexpr int kWarmPairs = 8;

struct HJob {                   // one horizontal step on np planes of identical geometry
    const int16_t *avg[2], *res[2];     // res[i] == nullptr: all-zero residual
    int16_t *out[2];
    int np;
    int wa, h;                  // averages wa x h, residuals wa x h, output 2*wa x h;  wa % 8 == 0
    int S, nseg;                // pairs per segment (multiple of 8), segments per row
    int R;                      // rows per block
    int blocks;
    int epilogue;               // fq::kEpNone / kEpClamp / kEpYCoCg (np == 2: planes are Co, Cg; out = G, B)
    const int16_t *yin;         // kEpYCoCg: the final Y plane (2*wa x h)
    int16_t *rout;              //           where R goes (a plane of its own: repairs re-read Y)
    int maxval, lo, hi, do_clamp;
};
struct HJobs { HJob j[3]; int n; };

struct VJob {                   // one vertical step on one plane
    const int16_t *avg, *res;
    int16_t *out;
    int w, ha;                  // averages w x ha, residuals w x ha, output w x 2*ha;  w % 8 == 0
    int S, nseg;                // pairs per segment, segments per column
    int CB;                     // 8-column groups per block
    int blocks;
    int do_clamp, lo, hi;       // final clamp folded in
};
struct VJobs { VJob j[4]; int n; };

FB_DEV int lo16(uint32_t w) { return (int)(short)(w & 0xffffu); }
FB_DEV int hi16(uint32_t w) { return (int)(short)(w >> 16); }
FB_DEV uint32_t pack16(int a, int b) { return (uint32_t)(uint16_t)a | ((uint32_t)(uint16_t)b << 16); }
FB_DEV uint4 ld16(const int16_t *p) { return *reinterpret_cast<const uint4 *>(p); }
FB_DEV void st16(int16_t *p, const uint4 &v) { *reinterpret_cast<uint4 *>(p) = v; }
FB_DEV uint4 zero4() { uint4 z; z.x = 0; z.y = 0; z.z = 0; z.w = 0; return z; }

// CLAMP(x, 0, hi) for hi >= 0 as max-then-min: two instructions where the conditional form costs three.  The two forms differ
// only for hi < 0, and the planner does not fuse the colour epilogue for such an image (fb_transforms.cu).
FB_DEV int clamp0(int x, int hi) { return fq::imin(fq::imax(x, 0), hi); }
// inv_YCoCg (+ clamp) on two samples packed in words, ycocg.h:51-56
FB_DEV void ycocg_word(uint32_t wy, uint32_t wo, uint32_t wg, int maxval, int lo, int hi, int do_clamp, uint32_t &r, uint32_t &g, uint32_t &b) {
    r = 0; g = 0; b = 0;
#pragma unroll
    for (int hlf = 0; hlf < 2; hlf++) {
        const int Yr = (int)(short)(wy >> (16 * hlf)), Co = (int)(short)(wo >> (16 * hlf)), Cg = (int)(short)(wg >> (16 * hlf));
        const int Y = clamp0(Yr, maxval);
        int G_ = clamp0(Y - ((-Cg) >> 1), maxval);
        int B_ = clamp0(Y + ((1 - Cg) >> 1) - (Co >> 1), maxval);
        int R_ = clamp0(Co + B_, maxval);
        if (do_clamp) { R_ = clampi(R_, lo, hi); G_ = clampi(G_, lo, hi); B_ = clampi(B_, lo, hi); }
        r |= (uint32_t)(uint16_t)R_ << (16 * hlf); g |= (uint32_t)(uint16_t)G_ << (16 * hlf); b |= (uint32_t)(uint16_t)B_ << (16 * hlf);
    }
}
FB_DEV uint32_t clamp_word(uint32_t w, int lo, int hi) { return pack16(clampi(lo16(w), lo, hi), clampi(hi16(w), lo, hi)); }

// ---------------------------------------------------------------------------------------------------------
// horizontal
// ---------------------------------------------------------------------------------------------------------

// Pairs [8*c_from, p1) of one row for NP planes, starting from state prev[] (chain_start: the first pair is pair 0 of
// the row and uses its own average as "left", squeeze.h:84-89).  Chunks >= c_store are written.  Returns the state
// before the first stored pair in bw[] and the final state in prev[].
template <int NP>
FB_DEV void h_run(const HJob &J, int row, int c_from, int c_store, int p1, bool chain_start, int *prev, int *bw) {
    const int wa = J.wa, wo = 2 * wa, nchunks = wa >> 3;
    const int c_end = (p1 + 7) >> 3;
    uint4 cur[NP], nxt[NP];
#pragma unroll
    for (int pl = 0; pl < NP; pl++) cur[pl] = ld16(J.avg[pl] + (size_t)row * wa + 8 * c_from);
    if (chain_start) {
#pragma unroll
        for (int pl = 0; pl < NP; pl++) prev[pl] = lo16(cur[pl].x);
    }
    for (int c = c_from; c < c_end; c++) {
        const bool more = c + 1 < nchunks;
        uint4 rs[NP];
#pragma unroll
        for (int pl = 0; pl < NP; pl++) {
            nxt[pl] = more ? ld16(J.avg[pl] + (size_t)row * wa + 8 * (c + 1)) : cur[pl];
            rs[pl] = J.res[pl] ? ld16(J.res[pl] + (size_t)row * wa + 8 * c) : zero4();
        }
        if (c == c_store) {
#pragma unroll
            for (int pl = 0; pl < NP; pl++) bw[pl] = prev[pl];
        }
        uint32_t ow[NP][8];
#pragma unroll
        for (int pl = 0; pl < NP; pl++) {
            const uint32_t aw[4] = {cur[pl].x, cur[pl].y, cur[pl].z, cur[pl].w};
            const uint32_t rw[4] = {rs[pl].x, rs[pl].y, rs[pl].z, rs[pl].w};
            int pv = prev[pl];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int av = (i & 1) ? hi16(aw[i >> 1]) : lo16(aw[i >> 1]);
                int nx;
                if (i < 7) nx = ((i + 1) & 1) ? hi16(aw[(i + 1) >> 1]) : lo16(aw[(i + 1) >> 1]);
                else nx = more ? lo16(nxt[pl].x) : av;         // last pair of the row: next average = own (squeeze.h:93)
                const int r = (i & 1) ? hi16(rw[i >> 1]) : lo16(rw[i >> 1]);
                int A, B;
                fq::unsqueeze_pair(pv, av, nx, r, A, B);
                ow[pl][i] = pack16(A, B);
                pv = B;
            }
            prev[pl] = pv;
        }
        if (c >= c_store) {
            const size_t o = (size_t)row * wo + 16 * c;
            if (J.epilogue == fq::kEpYCoCg) {       // NP == 2: ow[0] = Co, ow[1] = Cg
                const uint4 y0 = ld16(J.yin + o), y1 = ld16(J.yin + o + 8);
                const uint32_t yw[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
                uint32_t r[8], g[8], b[8];
#pragma unroll
                for (int i = 0; i < 8; i++) ycocg_word(yw[i], ow[0][i], ow[NP - 1][i], J.maxval, J.lo, J.hi, J.do_clamp, r[i], g[i], b[i]);
                uint4 v;
                v.x = r[0]; v.y = r[1]; v.z = r[2]; v.w = r[3]; st16(J.rout + o, v);
                v.x = r[4]; v.y = r[5]; v.z = r[6]; v.w = r[7]; st16(J.rout + o + 8, v);
                v.x = g[0]; v.y = g[1]; v.z = g[2]; v.w = g[3]; st16(J.out[0] + o, v);
                v.x = g[4]; v.y = g[5]; v.z = g[6]; v.w = g[7]; st16(J.out[0] + o + 8, v);
                v.x = b[0]; v.y = b[1]; v.z = b[2]; v.w = b[3]; st16(J.out[NP - 1] + o, v);
                v.x = b[4]; v.y = b[5]; v.z = b[6]; v.w = b[7]; st16(J.out[NP - 1] + o + 8, v);
            } else {
#pragma unroll
                for (int pl = 0; pl < NP; pl++) {
                    if (J.epilogue == fq::kEpClamp && J.do_clamp) {
#pragma unroll
                        for (int i = 0; i < 8; i++) ow[pl][i] = clamp_word(ow[pl][i], J.lo, J.hi);
                    }
                    uint4 v;
                    v.x = ow[pl][0]; v.y = ow[pl][1]; v.z = ow[pl][2]; v.w = ow[pl][3]; st16(J.out[pl] + o, v);
                    v.x = ow[pl][4]; v.y = ow[pl][5]; v.z = ow[pl][6]; v.w = ow[pl][7]; st16(J.out[pl] + o + 8, v);
                }
            }
        }
#pragma unroll
        for (int pl = 0; pl < NP; pl++) cur[pl] = nxt[pl];
    }
}

template <int NP>
FB_DEV void h_block(const HJob &J, int b, int16_t *bfS) {
    const int tid = (int)threadIdx.x;
    const int r = tid / J.nseg, s = tid - r * J.nseg;
    const int row = b * J.R + r;
    const bool active = r < J.R && row < J.h;
    const int p0 = s * J.S, p1 = imin(p0 + J.S, J.wa);
    int prev[NP], bw[NP];
    bool exact_start = (s == 0);
    if (active) {
        if (s == 0) h_run<NP>(J, row, 0, 0, p1, true, prev, bw);
        else {
            // guessed state: the average of the first warm-up pair (what a chain start would use)
#pragma unroll
            for (int pl = 0; pl < NP; pl++) prev[pl] = J.avg[pl][(size_t)row * J.wa + p0 - kWarmPairs];
            h_run<NP>(J, row, (p0 >> 3) - 1, p0 >> 3, p1, false, prev, bw);
        }
#pragma unroll
        for (int pl = 0; pl < NP; pl++) bfS[(r * J.nseg + s) * NP + pl] = (int16_t)prev[pl];
    }
    __syncthreads();
    for (;;) {
        bool bad = false;
        int want[NP];
        if (active && !exact_start) {
#pragma unroll
            for (int pl = 0; pl < NP; pl++) { want[pl] = bfS[(r * J.nseg + s - 1) * NP + pl]; bad = bad || (want[pl] != bw[pl]); }
        }
        if (!__syncthreads_or(bad)) break;
        if (bad) {
#pragma unroll
            for (int pl = 0; pl < NP; pl++) prev[pl] = want[pl];
            int bw2[NP];
            h_run<NP>(J, row, p0 >> 3, p0 >> 3, p1, false, prev, bw2);
#pragma unroll
            for (int pl = 0; pl < NP; pl++) { bw[pl] = want[pl]; bfS[(r * J.nseg + s) * NP + pl] = (int16_t)prev[pl]; }
        }
        __syncthreads();
    }
}

FB_KERNEL(512) k_inv_hsq_direct(const FB_GRID_CONSTANT HJobs jobs) {
    FB_DYN_SMEM(smraw);
    int b = (int)blockIdx.x, ji = 0;
    while (ji < jobs.n - 1 && b >= jobs.j[ji].blocks) { b -= jobs.j[ji].blocks; ji++; }
    const HJob &J = jobs.j[ji];
    if (J.np == 1) h_block<1>(J, b, reinterpret_cast<int16_t *>(smraw));
    else h_block<2>(J, b, reinterpret_cast<int16_t *>(smraw));
}

// ---------------------------------------------------------------------------------------------------------
// vertical
// ---------------------------------------------------------------------------------------------------------

// Pairs [q_from, q1) of 8 adjacent columns, from state prev[8]; rows of pairs >= q_store are written.
// The input rows of the next kVDepth pairs are requested while the current kVDepth pairs are computed (the loads do not
// depend on the chain), so a thread keeps 2 * kVDepth 16-byte loads in flight.
constexpr int kVDepth = 4;
FB_DEV void v_pair8(const VJob &J, int x, int q, bool store, const uint4 &cur, const uint4 &nxt, const uint4 &rs, int *prev) {
    const uint32_t aw[4] = {cur.x, cur.y, cur.z, cur.w}, nw[4] = {nxt.x, nxt.y, nxt.z, nxt.w}, rw[4] = {rs.x, rs.y, rs.z, rs.w};
    uint32_t oa[4], ob[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int A0, B0, A1, B1;
        fq::unsqueeze_pair(prev[2 * k], lo16(aw[k]), lo16(nw[k]), lo16(rw[k]), A0, B0);
        fq::unsqueeze_pair(prev[2 * k + 1], hi16(aw[k]), hi16(nw[k]), hi16(rw[k]), A1, B1);
        prev[2 * k] = B0; prev[2 * k + 1] = B1;
        oa[k] = pack16(A0, A1); ob[k] = pack16(B0, B1);
        if (J.do_clamp) { oa[k] = clamp_word(oa[k], J.lo, J.hi); ob[k] = clamp_word(ob[k], J.lo, J.hi); }
    }
    if (store) {
        uint4 v;
        v.x = oa[0]; v.y = oa[1]; v.z = oa[2]; v.w = oa[3]; st16(J.out + (size_t)(2 * q) * J.w + x, v);
        v.x = ob[0]; v.y = ob[1]; v.z = ob[2]; v.w = ob[3]; st16(J.out + (size_t)(2 * q + 1) * J.w + x, v);
    }
}
FB_DEV void v_run(const VJob &J, int x, int q_from, int q_store, int q1, bool chain_start, int *prev, int *bw) {
    const int w = J.w, last = J.ha - 1;
    uint4 A[kVDepth + 1], R[kVDepth];
#pragma unroll
    for (int i = 0; i <= kVDepth; i++) A[i] = ld16(J.avg + (size_t)imin(q_from + i, last) * w + x);
#pragma unroll
    for (int i = 0; i < kVDepth; i++) R[i] = J.res ? ld16(J.res + (size_t)imin(q_from + i, last) * w + x) : zero4();
    if (chain_start) {
        prev[0] = lo16(A[0].x); prev[1] = hi16(A[0].x); prev[2] = lo16(A[0].y); prev[3] = hi16(A[0].y);
        prev[4] = lo16(A[0].z); prev[5] = hi16(A[0].z); prev[6] = lo16(A[0].w); prev[7] = hi16(A[0].w);
    }
    for (int q0 = q_from; q0 < q1; q0 += kVDepth) {
        uint4 An[kVDepth], Rn[kVDepth];
#pragma unroll
        for (int i = 0; i < kVDepth; i++) {     // rows of the next round (addresses clamped into the plane: unused rows are harmless)
            An[i] = ld16(J.avg + (size_t)imin(q0 + kVDepth + 1 + i, last) * w + x);
            Rn[i] = J.res ? ld16(J.res + (size_t)imin(q0 + kVDepth + i, last) * w + x) : zero4();
        }
#pragma unroll
        for (int i = 0; i < kVDepth; i++) {
            const int q = q0 + i;
            if (q < q1) {
                if (q == q_store) {
#pragma unroll
                    for (int k = 0; k < 8; k++) bw[k] = prev[k];
                }
                // last pair of the column: next average = own (squeeze.h:201); A[i+1] then holds a clamped re-read of row ha-1 = A[i]
                v_pair8(J, x, q, q >= q_store, A[i], A[i + 1], R[i], prev);
            }
        }
        A[0] = A[kVDepth];
#pragma unroll
        for (int i = 0; i < kVDepth; i++) { A[i + 1] = An[i]; R[i] = Rn[i]; }
    }
}

FB_DEV void v_block(const VJob &J, int b, int16_t *bfS) {
    const int tid = (int)threadIdx.x;
    const int cb = tid % J.CB, s = tid / J.CB;          // lanes = adjacent column groups (coalesced), then segments
    const int x = (b * J.CB + cb) * 8;
    const bool active = s < J.nseg && x < J.w;
    const int q0 = s * J.S, q1 = imin(q0 + J.S, J.ha);
    int prev[8], bw[8];
    if (active) {
        if (s == 0) v_run(J, x, 0, 0, q1, true, prev, bw);
        else {
            const uint4 g = ld16(J.avg + (size_t)(q0 - kWarmPairs) * J.w + x);
            prev[0] = lo16(g.x); prev[1] = hi16(g.x); prev[2] = lo16(g.y); prev[3] = hi16(g.y);
            prev[4] = lo16(g.z); prev[5] = hi16(g.z); prev[6] = lo16(g.w); prev[7] = hi16(g.w);
            v_run(J, x, q0 - kWarmPairs, q0, q1, false, prev, bw);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) bfS[(s * J.CB + cb) * 8 + i] = (int16_t)prev[i];
    }
    __syncthreads();
    for (;;) {
        bool bad = false;
        int want[8];
        if (active && s > 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) { want[i] = bfS[((s - 1) * J.CB + cb) * 8 + i]; bad = bad || (want[i] != bw[i]); }
        }
        if (!__syncthreads_or(bad)) break;
        if (bad) {
            int bw2[8];
#pragma unroll
            for (int i = 0; i < 8; i++) { prev[i] = want[i]; bw[i] = want[i]; }
            v_run(J, x, q0, q0, q1, false, prev, bw2);
#pragma unroll
            for (int i = 0; i < 8; i++) bfS[(s * J.CB + cb) * 8 + i] = (int16_t)prev[i];
        }
        __syncthreads();
    }
}

FB_KERNEL(512) k_inv_vsq_direct(const FB_GRID_CONSTANT VJobs jobs) {
    FB_DYN_SMEM(smraw);
    int b = (int)blockIdx.x, ji = 0;
    while (ji < jobs.n - 1 && b >= jobs.j[ji].blocks) { b -= jobs.j[ji].blocks; ji++; }
    v_block(jobs.j[ji], b, reinterpret_cast<int16_t *>(smraw));
}

}  // namespace dq
