// Fused multi-level inverse Squeeze (+ colour inverse + final clamp) -- the tile kernel of the transform chain.
//
// Reference semantics: transform/squeeze.h:61-132, 173-224 (inv_hsqueeze / inv_vsqueeze with smooth_tendency),
// transform/ycocg.h:51-56 (inv_YCoCg), image/image.cpp:107-113 (final clamp).
//
// One CTA produces one tile of the FINAL planes of a run of up to kMaxLevels consecutive unsqueeze steps and keeps
// every intermediate level in shared memory, so each coefficient is read from HBM once and each output sample written
// once (the per-level kernels write and re-read every intermediate plane).
//
// The inverse step is a serial recurrence along its axis (pair p needs the reconstructed B of pair p-1), so a tile
// cannot simply start in the middle of a row / column.  Every chain that does not begin at the plane border starts
// kWarm pairs early from a guessed state (the recurrence forgets its start within a few pairs); the state
// it has reached at the first pair the tile OWNS is recorded ("est"), and the owner of the pair before it records the
// value it really produced ("act").  Ownership is a lattice: at the output of level k a tile owns the cell
// [ti*tw_k, (ti+1)*tw_k) x [tj*th_k, (tj+1)*th_k) with tw_k = TW >> (horizontal steps after k).  By induction over
// (level, tile index along the axis) all results are exact iff every est equals its act; k_fq_verify_fallback checks
// that after the last fused launch.  A tile of the LAST launch whose est failed is recomputed there in "exact mode"
// (no warm-up: its chains start from the recorded act values, exact by the same induction; if that changes any act the
// tile had published, the run is treated as failed).  A failure in an earlier launch, whose output later launches have
// consumed, makes the whole run be recomputed with plain serial chains (exact by construction).  So the output is
// bit-exact no matter how good the guesses were; the warm-up length only sets how rarely the repairs happen
// (failures drop ~7x per warm-up pair: ~1e-7 per chain start at 8 pairs).
//
// This header is compiled by nvcc (the product) and by g++ with -DFB_EMULATE (tests/emu: the CPU-only test tier runs
// the very same kernel source under an execution-model emulator and compares it with the oracle).
#pragma once
#include "fb_port.h"

namespace fq {

#ifndef FQ_WARM
#define FQ_WARM 8
#endif
#ifndef FQ_WARM_MID
#define FQ_WARM_MID 12
#endif
constexpr int kWarm = FQ_WARM;          // warm-up pairs of the last launch (failed tiles are repaired locally)
constexpr int kWarmMid = FQ_WARM_MID;   // warm-up pairs of earlier launches (a failure there costs a full serial recompute)
constexpr int kMaxLevels = 12;      // unsqueeze steps per gang in one launch
constexpr int kGP = 2;              // planes per gang (identical geometry: processed by one thread for ILP)
constexpr int kMaxGangs = 2;        // gangs per CTA (joined by the colour epilogue)
constexpr int kNoCheck = 0x7fffffff;

enum { kEpNone = 0, kEpClamp = 1, kEpYCoCg = 2 };

FB_HD int s16(int x) { return (int)(short)x; }
FB_HD int imin(int a, int b) { return a < b ? a : b; }
FB_HD int imax(int a, int b) { return a > b ? a : b; }
FB_HD int iabs(int a) { return a < 0 ? -a : a; }
FB_HD int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
// halfword pitch >= n whose 32-bit word count is odd (conflict-free column walks by 32 row-threads)
FB_HD int odd_pitch(int n) { int wds = (n + 1) >> 1; if (!(wds & 1)) wds++; return wds * 2; }

// One unsqueeze pair (squeeze.h:97-108) with smooth_tendency (:61-77) in closed form.  With t1 = B-a, t2 = a-n:
//   monotone triple: |tendency| = min((4|t1|+3|t2|+6)/12, 2|t1|+1, 2|t2|), sign of t1+t2   (both clamps fold into the min:
//   d-(d&1) > 2k <=> d >= 2k+2 and d+(d&1) > 2k <=> d >= 2k+1); otherwise 0.
// No int16 wrap can occur inside smooth_tendency for int16 inputs: t1+t2 = B-n <= 65535 bounds 4|t1|+3|t2|+6 by 262146,
// so the quotient is <= 21845, and the clamps only ever lower it.  The wraps of diff, A and B are kept.
FB_HD void unsqueeze_pair(int prev, int av, int nx, int rs, int &A, int &B) {
    const int t1 = prev - av, t2 = av - nx;
    const int a1 = iabs(t1), a2 = iabs(t2);
    const unsigned m = (unsigned)(4 * a1 + 3 * a2 + 6);
#if defined(__CUDA_ARCH__) || defined(FB_EMULATE)
    const int q = (int)(__umulhi(m, 0xAAAAAAABu) >> 3);        // m / 12
#else
    const int q = (int)(m / 12u);
#endif
    int d = imin(imin(q, 2 * a1 + 1), 2 * a2);
    const bool mono = ((t1 ^ t2) >= 0) | (t1 == 0);
    d = (t1 + t2) < 0 ? -d : d;
    const int tendency = mono ? d : 0;
    const int diff = s16(rs + tendency);
    A = s16(av + ((diff - (diff >> 31)) >> 1));                 // (2a + diff -+ (diff&1)) >> 1  ==  a + trunc(diff/2)
    B = s16(A - diff);
}

// smooth_tendency + pair exactly as the reference writes them (used by the serial fallback and by the tests)
FB_HD int smooth_tendency_literal(int B, int a, int n) {
    int diff = 0;
    if (B >= a && a >= n) {
        diff = s16((4 * B - 3 * n - a + 6) / 12);
        if (diff - (diff & 1) > 2 * (B - a)) diff = s16(2 * (B - a) + 1);
        if (diff + (diff & 1) > 2 * (a - n)) diff = s16(2 * (a - n));
    } else if (B <= a && a <= n) {
        diff = s16((4 * B - 3 * n - a - 6) / 12);
        if (diff + (diff & 1) < 2 * (B - a)) diff = s16(2 * (B - a) - 1);
        if (diff - (diff & 1) < 2 * (a - n)) diff = s16(2 * (a - n));
    }
    return diff;
}
FB_HD void unsqueeze_pair_literal(int prev, int avg, int next_avg, int res, int &A, int &B) {
    const int tendency = smooth_tendency_literal(prev, avg, next_avg);
    const int diff = s16(res + tendency);
    A = s16(((avg << 1) + diff + (diff > 0 ? -(diff & 1) : (diff & 1))) >> 1);
    B = s16(A - diff);
}

// ---------------------------------------------------------------------------------------------------------
// launch description (filled by the host planner, fb_fused_plan.h)
// ---------------------------------------------------------------------------------------------------------
struct Level {                  // one unsqueeze step applied to every plane of a gang
    int horizontal;
    int wa, ha, wr, hr;         // average / residual plane dims (horizontal: hr == ha, vertical: wr == wa)
    int wo, ho;                 // output dims
    int tw, th;                 // ownership lattice cell at this level's output
    int out_pitch, res_pitch;   // shared-memory pitches (halfwords)
    int est_cap;                // est entries per tile and plane
    int out_off[kGP], res_off[kGP];     // shared-memory offsets (halfwords from the CTA's base)
    const int16_t *res[kGP];    // residual planes in HBM (nullptr = all zero)
    int *est[kGP];              // [ntx*nty][est_cap] state each exact chain assumed (kNoCheck = started at the border)
    int16_t *act[kGP];          // horizontal: [ntx][ho], vertical: [nty][wo]: last owned B of every chain
};
struct Gang {
    int np, nlev;
    int w0, h0;                 // level-0 average planes
    int W, H;                   // final planes
    int TW, TH, ntx, nty;       // tile grid over the final planes
    int first_block;            // separate gangs: first CTA of this gang's tile grid
    int first_thread, nthreads, bar_id;
    int warm;                   // warm-up pairs of speculative chain starts
    int in_pitch;
    int in_off[kGP];
    const int16_t *in[kGP];
    int16_t *out[kGP];
    Level lv[kMaxLevels];
};
struct Task {
    int ngangs;
    int joined;                 // 1: every CTA runs all gangs on one tile (they share W, H and the tile grid; colour
                                //    epilogue possible); 0: every gang has its own tile grid and CTAs (first_block)
    int epilogue;               // kEpNone / kEpClamp / kEpYCoCg
    int maxval, lo, hi, do_clamp;
    int ycc_gang[3], ycc_plane[3];      // kEpYCoCg: where Y, Co, Cg live
    int geom_off;               // byte offset of the geometry scratch in shared memory
    int *counters;              // see VerifyParams::counters
    unsigned char *tile_bad;    // last launch: one flag per CTA, cleared by the CTA itself, set by the verification
    Gang g[kMaxGangs];
};

struct Geom {                   // what one tile computes at one level (absolute coordinates of the level's output)
    int x0, y0, x1, y1;         // stored output region
    int p_start, p_store, p_exact, p_end;       // pairs along the squeeze axis
    int tail;                   // the region includes the odd tail column / row (copy of the last average)
    int c0, c1, ce;             // chains (rows for horizontal, columns for vertical); chains >= ce must be exact
};
struct Region { int x0, y0, x1, y1; };

// Walks the levels backwards from the owned tile of the final planes; g[k] for k = 0..nlev-1, in = level-0 input.
// warm = 0 gives the exact-mode geometry (no halos: every chain starts at the first owned pair).
FB_HD void geometry(const Gang &G, int ti, int tj, int warm, Geom *g, Region &in) {
    int nx0 = ti * G.TW, nx1 = (ti == G.ntx - 1) ? G.W : (ti + 1) * G.TW;
    int ny0 = tj * G.TH, ny1 = (tj == G.nty - 1) ? G.H : (tj + 1) * G.TH;
    int xe = nx0, ye = ny0;
    for (int k = G.nlev - 1; k >= 0; k--) {
        const Level &L = G.lv[k];
        Geom &q = g[k];
        if (L.horizontal) {
            q.p_store = nx0 >> 1;
            q.p_exact = xe >> 1;
            q.p_start = imax(0, q.p_store - warm);
            q.p_end = imin((nx1 + 1) >> 1, L.wr);
            q.tail = (nx1 == L.wo) && (L.wo & 1);
            q.x0 = 2 * q.p_store; q.x1 = nx1; q.y0 = ny0; q.y1 = ny1;
            q.c0 = ny0; q.c1 = ny1; q.ce = ye;
            nx0 = q.p_start; nx1 = imin(q.p_end + 1, L.wa); xe = q.p_exact;
        } else {
            q.p_store = ny0 >> 1;
            q.p_exact = ye >> 1;
            q.p_start = imax(0, q.p_store - warm);
            q.p_end = imin((ny1 + 1) >> 1, L.hr);
            q.tail = (ny1 == L.ho) && (L.ho & 1);
            q.y0 = 2 * q.p_store; q.y1 = ny1; q.x0 = nx0; q.x1 = nx1;
            q.c0 = nx0; q.c1 = nx1; q.ce = xe;
            ny0 = q.p_start; ny1 = imin(q.p_end + 1, L.ha); ye = q.p_exact;
        }
    }
    in.x0 = nx0; in.x1 = nx1; in.y0 = ny0; in.y1 = ny1;
}

// ---------------------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------------------

// Copies the region [x0,x1) x [y0,y1) of a w-wide plane into shared memory; the buffer's origin is (x0 & ~7, y0)
// and its pitch covers whole 8-sample chunks.  src == nullptr stages zeros.
FB_DEV void stage_region(int16_t *dst, int dpitch, const int16_t *src, int w, int x0, int x1, int y0, int y1, int tid, int nthr) {
    const int ox = x0 & ~7;
    const int nchunk = (x1 - ox + 7) >> 3, rows = y1 - y0;
    if (nchunk <= 0 || rows <= 0) return;
    const bool vec = src && ((w & 7) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const int total = nchunk * rows;
    int r = tid / nchunk, c = tid - r * nchunk;
    const int dr = nthr / nchunk, dc = nthr - dr * nchunk;
    for (int i = tid; i < total; i += nthr) {
        const int gx = ox + 8 * c;
        int16_t *d16 = dst + r * dpitch + 8 * c;
        uint32_t *d = reinterpret_cast<uint32_t *>(d16);
        if (!src) {
            d[0] = 0; d[1] = 0; d[2] = 0; d[3] = 0;
        } else if (vec && gx + 8 <= w) {
            const uint4 v = *reinterpret_cast<const uint4 *>(src + (size_t)(y0 + r) * w + gx);
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        } else {
            const int16_t *s = src + (size_t)(y0 + r) * w;
            for (int j = 0; j < 8; j++) d16[j] = (gx + j < w) ? s[gx + j] : (int16_t)0;
        }
        c += dc; r += dr;
        if (c >= nchunk) { c -= nchunk; r++; }
    }
}

// All chains of one level for the planes of a gang.  H: chains are rows, V: chains are columns.
template <int NP, bool H>
FB_DEV void run_level(const Gang &G, int k, const Geom &q, const Geom *qprev, const Region &in, int16_t *sm, int ti, int tj,
                      int gtid, bool exact, int *counters) {
    const Level &L = G.lv[k];
    // source of the averages: the staged level-0 planes or the previous level's output
    int aox, aoy, apitch;
    const int *aoff;
    if (k == 0) { aox = in.x0 & ~7; aoy = in.y0; apitch = G.in_pitch; aoff = G.in_off; }
    else { aox = qprev->x0; aoy = qprev->y0; apitch = G.lv[k - 1].out_pitch; aoff = G.lv[k - 1].out_off; }
    const int rpitch = L.res_pitch, opitch = L.out_pitch;
    const int navg = H ? L.wa : L.ha, npair = H ? L.wr : L.hr;
    // the very last pair of a chain without an odd tail has no next average: it uses its own (squeeze.h:93, :201)
    const bool self_next = (q.p_end == npair) && (navg == npair) && (q.p_end > q.p_exact);
    const int p_plain_end = self_next ? q.p_end - 1 : q.p_end;
    const int last_tile_along = H ? (ti == G.ntx - 1) : (tj == G.nty - 1);
    const int own_end_along = H ? (ti + 1) * L.tw : (tj + 1) * L.th;       // only used when !last_tile_along
    const int own_c0 = H ? tj * L.th : ti * L.tw;
    const int own_c1 = H ? ((tj == G.nty - 1) ? L.ho : (tj + 1) * L.th) : ((ti == G.ntx - 1) ? L.wo : (ti + 1) * L.tw);
    const int dim_across = H ? L.ho : L.wo;
    const int tile_lin = ti * G.nty + tj;
    for (int c = q.c0 + gtid; c < q.c1; c += G.nthreads) {
        const int16_t *a[NP], *r[NP];
        int16_t *o[NP];
        int a_step, r_step;
#pragma unroll
        for (int pl = 0; pl < NP; pl++) {
            if (H) {
                a[pl] = sm + aoff[pl] + (c - aoy) * apitch - aox;                   // a[p]
                r[pl] = sm + L.res_off[pl] + (c - q.c0) * rpitch - (q.p_start & ~7);
                o[pl] = sm + L.out_off[pl] + (c - q.y0) * opitch - q.x0;            // o[x]
            } else {
                a[pl] = sm + aoff[pl] + (c - aox) - aoy * apitch;                   // a[p * apitch]
                r[pl] = sm + L.res_off[pl] + (c - (q.c0 & ~7)) - q.p_start * rpitch;
                o[pl] = sm + L.out_off[pl] + (c - q.x0) - q.y0 * opitch;            // o[y * opitch]
            }
        }
        a_step = H ? 1 : apitch;
        r_step = H ? 1 : rpitch;
        int prev[NP], av[NP], est[NP];
#pragma unroll
        for (int pl = 0; pl < NP; pl++) { av[pl] = a[pl][q.p_start * a_step]; prev[pl] = av[pl]; est[pl] = kNoCheck; }
        if (exact && q.p_start > 0) {       // repair: start from the value the owner of the previous pair produced
#pragma unroll
            for (int pl = 0; pl < NP; pl++) prev[pl] = L.act[pl][(size_t)((H ? ti : tj) - 1) * dim_across + c];
        }
        auto pairs = [&](int from, int to, bool store, bool own_next) {
            for (int p = from; p < to; p++) {
#pragma unroll
                for (int pl = 0; pl < NP; pl++) {
                    const int nx = own_next ? av[pl] : (int)a[pl][(p + 1) * a_step];
                    int A, B;
                    unsqueeze_pair(prev[pl], av[pl], nx, r[pl][p * r_step], A, B);
                    if (store) {
                        if (H) *reinterpret_cast<uint32_t *>(o[pl] + 2 * p) = (uint32_t)(uint16_t)A | ((uint32_t)(uint16_t)B << 16);
                        else { o[pl][(2 * p) * opitch] = (int16_t)A; o[pl][(2 * p + 1) * opitch] = (int16_t)B; }
                    }
                    prev[pl] = B;
                    av[pl] = nx;
                }
            }
        };
        pairs(q.p_start, imin(q.p_store, p_plain_end), false, false);
        pairs(q.p_store, imin(q.p_exact, p_plain_end), true, false);
        if (q.p_start > 0) {
#pragma unroll
            for (int pl = 0; pl < NP; pl++) est[pl] = prev[pl];
        }
        pairs(q.p_exact, p_plain_end, true, false);
        if (self_next) pairs(p_plain_end, q.p_end, true, true);
        if (q.tail) {
#pragma unroll
            for (int pl = 0; pl < NP; pl++) {
                const int16_t v = a[pl][(navg - 1) * a_step];
                if (H) o[pl][L.wo - 1] = v; else o[pl][(L.ho - 1) * opitch] = v;
            }
        }
        // verification records
        if (!exact && c >= q.ce && L.est[0]) {
#pragma unroll
            for (int pl = 0; pl < NP; pl++) L.est[pl][(size_t)tile_lin * L.est_cap + (c - q.ce)] = est[pl];
        }
        if (!last_tile_along && c >= own_c0 && c < own_c1 && L.act[0]) {
            const int idx = H ? ti : tj;
#pragma unroll
            for (int pl = 0; pl < NP; pl++) {
                const int16_t v = H ? o[pl][own_end_along - 1] : o[pl][(own_end_along - 1) * opitch];
                int16_t *slot = &L.act[pl][(size_t)idx * dim_across + c];
                if (exact) { if (*slot != v) atomicOr(counters + 2, 1); }  // a repair must not change what neighbours checked against
                else *slot = v;
            }
        }
    }
    // unused est slots of this tile
    if (!exact && L.est[0]) {
        for (int e = (q.c1 - q.ce) + gtid; e < L.est_cap; e += G.nthreads) {
#pragma unroll
            for (int pl = 0; pl < NP; pl++) L.est[pl][(size_t)tile_lin * L.est_cap + e] = kNoCheck;
        }
    }
}

template <int NP>
FB_DEV void run_gang(const Gang &G, const Geom *geo, const Region &in, int16_t *sm, int ti, int tj, int gtid, bool exact, int *counters) {
    // stage the level-0 averages and every level's residuals
    for (int pl = 0; pl < NP; pl++) stage_region(sm + G.in_off[pl], G.in_pitch, G.in[pl], G.w0, in.x0, in.x1, in.y0, in.y1, gtid, G.nthreads);
    for (int k = 0; k < G.nlev; k++) {
        const Level &L = G.lv[k];
        const Geom &q = geo[k];
        for (int pl = 0; pl < NP; pl++) {
            if (L.horizontal) stage_region(sm + L.res_off[pl], L.res_pitch, L.res[pl], L.wr, q.p_start, q.p_end, q.c0, q.c1, gtid, G.nthreads);
            else stage_region(sm + L.res_off[pl], L.res_pitch, L.res[pl], L.wr, q.c0, q.c1, q.p_start, q.p_end, gtid, G.nthreads);
        }
    }
    fb_bar_sync(G.bar_id, G.nthreads);
    for (int k = 0; k < G.nlev; k++) {
        if (G.lv[k].horizontal) run_level<NP, true>(G, k, geo[k], k ? &geo[k - 1] : nullptr, in, sm, ti, tj, gtid, exact, counters);
        else run_level<NP, false>(G, k, geo[k], k ? &geo[k - 1] : nullptr, in, sm, ti, tj, gtid, exact, counters);
        fb_bar_sync(G.bar_id, G.nthreads);
    }
}

// Final planes of the tile: shared memory -> HBM with the colour inverse / clamp applied, 8 samples per thread-step.
FB_DEV void epilogue(const Task &T, int g_first, int g_last, int16_t *sm, int ti, int tj) {
    const Gang &G0 = T.g[g_first];
    const int x0 = ti * G0.TW, y0 = tj * G0.TH;
    const int W = G0.W, H = G0.H;
    const int x1 = (ti == G0.ntx - 1) ? W : x0 + G0.TW, y1 = (tj == G0.nty - 1) ? H : y0 + G0.TH;
    const int nchunk = (x1 - x0 + 7) >> 3, rows = y1 - y0, total = nchunk * rows;
    const int nthr = (int)blockDim.x;
    // units of work: the YCoCg triple (if any) and every other plane on its own
    const int16_t *fin[kMaxGangs * kGP];
    int16_t *dstp[kMaxGangs * kGP];
    int pitch[kMaxGangs * kGP];
    int nfin = 0, ycc[3] = {-1, -1, -1};
    for (int gi = g_first; gi <= g_last; gi++) {
        const Gang &G = T.g[gi];
        const Level &L = G.lv[G.nlev - 1];
        for (int pl = 0; pl < G.np; pl++) {
            fin[nfin] = sm + L.out_off[pl];
            pitch[nfin] = L.out_pitch;
            dstp[nfin] = G.out[pl];
            if (T.epilogue == kEpYCoCg)
                for (int j = 0; j < 3; j++)
                    if (T.ycc_gang[j] == gi && T.ycc_plane[j] == pl) ycc[j] = nfin;
            nfin++;
        }
    }
    const bool vec = (W & 7) == 0;
    int r = (int)threadIdx.x / nchunk, c = (int)threadIdx.x - r * nchunk;
    const int dr = nthr / nchunk, dc = nthr - dr * nchunk;
    for (int i = (int)threadIdx.x; i < total; i += nthr) {
        const int gx = x0 + 8 * c, gy = y0 + r;
        const int n = imin(8, W - gx);
        const size_t goff = (size_t)gy * W + gx;
        if (T.epilogue == kEpYCoCg) {
            uint32_t wy[4], wo[4], wg[4];
            const uint32_t *sy = reinterpret_cast<const uint32_t *>(fin[ycc[0]] + r * pitch[ycc[0]] + 8 * c);
            const uint32_t *so = reinterpret_cast<const uint32_t *>(fin[ycc[1]] + r * pitch[ycc[1]] + 8 * c);
            const uint32_t *sg = reinterpret_cast<const uint32_t *>(fin[ycc[2]] + r * pitch[ycc[2]] + 8 * c);
#pragma unroll
            for (int j = 0; j < 4; j++) { wy[j] = sy[j]; wo[j] = so[j]; wg[j] = sg[j]; }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t ra = 0, rb = 0, rc = 0;
#pragma unroll
                for (int hlf = 0; hlf < 2; hlf++) {     // inv_YCoCg, ycocg.h:51-56
                    const int Yr = (int)(short)(wy[j] >> (16 * hlf)), Co = (int)(short)(wo[j] >> (16 * hlf)), Cg = (int)(short)(wg[j] >> (16 * hlf));
                    const int Y = clampi(Yr, 0, T.maxval);
                    int G_ = clampi(Y - ((-Cg) >> 1), 0, T.maxval);
                    int B_ = clampi(Y + ((1 - Cg) >> 1) - (Co >> 1), 0, T.maxval);
                    int R_ = clampi(Co + B_, 0, T.maxval);
                    if (T.do_clamp) { R_ = clampi(R_, T.lo, T.hi); G_ = clampi(G_, T.lo, T.hi); B_ = clampi(B_, T.lo, T.hi); }
                    ra |= (uint32_t)(uint16_t)R_ << (16 * hlf); rb |= (uint32_t)(uint16_t)G_ << (16 * hlf); rc |= (uint32_t)(uint16_t)B_ << (16 * hlf);
                }
                wy[j] = ra; wo[j] = rb; wg[j] = rc;
            }
            int16_t *d0 = dstp[ycc[0]] + goff, *d1 = dstp[ycc[1]] + goff, *d2 = dstp[ycc[2]] + goff;
            if (vec && n == 8) {
                uint4 v;
                v.x = wy[0]; v.y = wy[1]; v.z = wy[2]; v.w = wy[3]; *reinterpret_cast<uint4 *>(d0) = v;
                v.x = wo[0]; v.y = wo[1]; v.z = wo[2]; v.w = wo[3]; *reinterpret_cast<uint4 *>(d1) = v;
                v.x = wg[0]; v.y = wg[1]; v.z = wg[2]; v.w = wg[3]; *reinterpret_cast<uint4 *>(d2) = v;
            } else {
                for (int j = 0; j < n; j++) {
                    d0[j] = (int16_t)(wy[j >> 1] >> (16 * (j & 1)));
                    d1[j] = (int16_t)(wo[j >> 1] >> (16 * (j & 1)));
                    d2[j] = (int16_t)(wg[j >> 1] >> (16 * (j & 1)));
                }
            }
        }
        for (int f = 0; f < nfin; f++) {
            if (T.epilogue == kEpYCoCg && (f == ycc[0] || f == ycc[1] || f == ycc[2])) continue;
            const uint32_t *s = reinterpret_cast<const uint32_t *>(fin[f] + r * pitch[f] + 8 * c);
            uint32_t wv[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t v = s[j];
                if (T.epilogue != kEpNone && T.do_clamp) {
                    const int lo_ = clampi((int)(short)v, T.lo, T.hi), hi_ = clampi((int)(short)(v >> 16), T.lo, T.hi);
                    v = (uint32_t)(uint16_t)lo_ | ((uint32_t)(uint16_t)hi_ << 16);
                }
                wv[j] = v;
            }
            int16_t *d = dstp[f] + goff;
            if (vec && n == 8) {
                uint4 v; v.x = wv[0]; v.y = wv[1]; v.z = wv[2]; v.w = wv[3];
                *reinterpret_cast<uint4 *>(d) = v;
            } else {
                for (int j = 0; j < n; j++) d[j] = (int16_t)(wv[j >> 1] >> (16 * (j & 1)));
            }
        }
        c += dc; r += dr;
        if (c >= nchunk) { c -= nchunk; r++; }
    }
}

// One tile (CTA `block` of the launch): speculative (exact == false) or repair (exact == true) mode.
FB_DEV void tile_body(const Task &T, int block, bool exact, unsigned char *smraw) {
    int16_t *sm = reinterpret_cast<int16_t *>(smraw);
    Geom *geo = reinterpret_cast<Geom *>(smraw + T.geom_off);                  // [kMaxGangs][kMaxLevels]
    Region *inr = reinterpret_cast<Region *>(geo + kMaxGangs * kMaxLevels);    // [kMaxGangs]
    const int tid = (int)threadIdx.x;
    int gi, tile;
    if (T.joined) {
        gi = (T.ngangs > 1 && tid >= T.g[1].first_thread) ? 1 : 0;
        tile = block;
    } else {
        gi = (T.ngangs > 1 && block >= T.g[1].first_block) ? 1 : 0;
        tile = block - T.g[gi].first_block;
    }
    const Gang &G = T.g[gi];
    const int ti = tile % G.ntx, tj = tile / G.ntx;
    const int gtid = tid - G.first_thread;
    if (gtid == 0) geometry(G, ti, tj, exact ? 0 : G.warm, geo + gi * kMaxLevels, inr[gi]);
    if (tid == 0 && !exact && T.tile_bad) T.tile_bad[block] = 0;
    __syncthreads();
    if (gtid >= 0 && gtid < G.nthreads) {
        if (G.np == 1) run_gang<1>(G, geo + gi * kMaxLevels, inr[gi], sm, ti, tj, gtid, exact, T.counters);
        else run_gang<2>(G, geo + gi * kMaxLevels, inr[gi], sm, ti, tj, gtid, exact, T.counters);
    }
    __syncthreads();
    epilogue(T, T.joined ? 0 : gi, T.joined ? T.ngangs - 1 : gi, sm, ti, tj);
}

FB_KERNEL(512) k_fq_tiles(const FB_GRID_CONSTANT Task T) {
    FB_DYN_SMEM(smraw);
    tile_body(T, (int)blockIdx.x, false, smraw);
}

// ---------------------------------------------------------------------------------------------------------
// verification + exact serial fallback (one cooperative launch after the last fused launch of a run)
// ---------------------------------------------------------------------------------------------------------
struct Check {                  // one (level, plane) of a multi-tile launch
    const int *est;
    const int16_t *act;
    int horizontal, ntx, nty, est_cap, cell, dim_across;     // cell: lattice cell size across the axis
    int first_block;            // >= 0: a check of the LAST launch (CTA id = first_block + tj*ntx + ti): repairable
};
struct SerialOp {               // one unsqueeze step on one plane, as the per-level kernels see it
    const int16_t *avg, *res;
    int16_t *out;
    int wa, wr, ha, hr, horizontal, step;
};
constexpr int kMaxChecks = 96, kMaxSerialOps = 96;
struct VerifyParams {
    int nchecks, nops;
    int *counters;              // zeroed before every run: [0] a check of an early launch failed [1] tiles of the last launch
                                // to repair [2] a repair changed a published act;  statistics, never reset: [4] runs that
                                // were recomputed serially [5] tiles repaired
    int *bad_list;              // CTA ids of the last launch to repair
    int bad_cap;
    int force;                  // testing: 1 = behave as if an early check had failed, 2 = repair every tile of the last launch
    int epilogue, maxval, lo, hi, do_clamp;
    int16_t *ycc[3];            // kEpYCoCg: final Y/Co/Cg planes
    int16_t *other[4];          // planes that only get the clamp
    int nother;
    int W, H;
    int top_blocks;             // grid of the last launch
    Task top;                   // the last launch (for repairs)
    Check chk[kMaxChecks];
    SerialOp op[kMaxSerialOps];
};

FB_DEV void serial_chain(const SerialOp &o, int chain) {
    // inv_hsqueeze / inv_vsqueeze for one row / column, squeeze.h:81-132, 173-224
    const bool H = o.horizontal != 0;
    const int navg = H ? o.wa : o.ha, npair = H ? o.wr : o.hr;
    const int wo = H ? o.wa + o.wr : o.wa;
    const int16_t *a = H ? o.avg + (size_t)chain * o.wa : o.avg + chain;
    const int16_t *r = o.res ? (H ? o.res + (size_t)chain * o.wr : o.res + chain) : nullptr;
    int16_t *out = H ? o.out + (size_t)chain * wo : o.out + chain;
    const size_t as = H ? 1 : (size_t)o.wa, rs_ = H ? 1 : (size_t)o.wa, os = H ? 1 : (size_t)wo;
    int prev = 0;
    for (int p = 0; p < npair; p++) {
        const int av = a[p * as];
        const int nx = (p + 1 < navg) ? a[(p + 1) * as] : av;
        int A, B;
        unsqueeze_pair_literal(p == 0 ? av : prev, av, nx, r ? r[p * rs_] : 0, A, B);
        out[(size_t)(2 * p) * os] = (int16_t)A;
        out[(size_t)(2 * p + 1) * os] = (int16_t)B;
        prev = B;
    }
    if (navg > npair) out[(size_t)(navg + npair - 1) * os] = a[(size_t)(navg - 1) * as];
}

// Cooperative launch: blockDim = threads of the last fused launch, dynamic shared memory = its shared memory.
FB_KERNEL(512) k_fq_verify_fallback(const FB_GRID_CONSTANT VerifyParams P) {
    FB_DYN_SMEM(smraw);
    const int gthreads = (int)(gridDim.x * blockDim.x), gtid = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    int bad = P.force == 1;
    for (int ci = 0; ci < P.nchecks; ci++) {
        const Check &C = P.chk[ci];
        const int per_tile = C.est_cap, ntiles = C.ntx * C.nty;
        for (int i = gtid; i < ntiles * per_tile; i += gthreads) {
            const int tile = i / per_tile, e = i - tile * per_tile;
            const int ti = tile / C.nty, tj = tile - ti * C.nty;
            const int along = C.horizontal ? ti : tj, across = C.horizontal ? tj : ti;
            if (along == 0) continue;
            const int v = C.est[i];
            if (v == kNoCheck) continue;
            const int want = C.act[(size_t)(along - 1) * C.dim_across + across * C.cell + e];
            if (want != v) {
                if (C.first_block < 0) bad = 1;
                else {
                    const int cta = C.first_block + tj * C.ntx + ti;
                    if (P.top.tile_bad[cta] == 0) {         // benign race: duplicates are filtered by the exchange below
                        P.top.tile_bad[cta] = 1;
                        const int slot = atomicAdd(P.counters + 1, 1);
                        if (slot < P.bad_cap) P.bad_list[slot] = cta; else bad = 1;
                    }
                }
            }
        }
    }
    if (P.force == 2) {
        for (int i = gtid; i < P.top_blocks && i < P.bad_cap; i += gthreads) P.bad_list[i] = i;
        if (gtid == 0) P.counters[1] = imin(P.top_blocks, P.bad_cap);
    }
    if (bad) atomicOr(P.counters, 1);
    fb_grid_sync();
    // counters[0] and [1] are stable from here on (repairs report through [2])
    const int failed_early = *(volatile int *)P.counters;
    const int nbad = imin(*(volatile int *)(P.counters + 1), P.bad_cap);
    if (!failed_early && nbad == 0) return;
    // ---- local repair of tiles of the last launch (a tile may be listed twice after a lost race: harmless)
    if (!failed_early) {
        for (int i = (int)blockIdx.x; i < nbad; i += (int)gridDim.x) {
            tile_body(P.top, P.bad_list[i], true, smraw);
            __syncthreads();
        }
        if (gtid == 0) atomicAdd(P.counters + 5, nbad);
        fb_grid_sync();
        if (*(volatile int *)(P.counters + 2) == 0) return;
    }
    if (gtid == 0) atomicAdd(P.counters + 4, 1);
    // ---- exact recomputation of the whole run, one grid barrier per squeeze step
    int i0 = 0;
    while (i0 < P.nops) {
        int i1 = i0;
        while (i1 < P.nops && P.op[i1].step == P.op[i0].step) i1++;
        for (int oi = i0; oi < i1; oi++) {
            const SerialOp &o = P.op[oi];
            const int nchain = o.horizontal ? o.ha : o.wa;
            for (int c = gtid; c < nchain; c += gthreads) serial_chain(o, c);
        }
        fb_grid_sync();
        i0 = i1;
    }
    const size_t n = (size_t)P.W * P.H;
    if (P.epilogue == kEpYCoCg) {
        for (size_t i = gtid; i < n; i += gthreads) {
            const int Y = clampi(P.ycc[0][i], 0, P.maxval), Co = P.ycc[1][i], Cg = P.ycc[2][i];
            int G_ = clampi(Y - ((-Cg) >> 1), 0, P.maxval);
            int B_ = clampi(Y + ((1 - Cg) >> 1) - (Co >> 1), 0, P.maxval);
            int R_ = clampi(Co + B_, 0, P.maxval);
            if (P.do_clamp) { R_ = clampi(R_, P.lo, P.hi); G_ = clampi(G_, P.lo, P.hi); B_ = clampi(B_, P.lo, P.hi); }
            P.ycc[0][i] = (int16_t)R_; P.ycc[1][i] = (int16_t)G_; P.ycc[2][i] = (int16_t)B_;
        }
    }
    if (P.epilogue != kEpNone && P.do_clamp) {
        for (int k = 0; k < P.nother; k++)
            for (size_t i = gtid; i < n; i += gthreads) P.other[k][i] = (int16_t)clampi(P.other[k][i], P.lo, P.hi);
    }
}

}  // namespace fq
