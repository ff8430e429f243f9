// Transform-chain kernels of fuif_b200 (sm_100a).  Bit-exact with the reference's integer / double results.
//
// All sample arithmetic follows the C++ semantics of the reference with pixel_type == int16_t
// (reference image/image.h:35): a value assigned to a pixel_type wraps to 16 bits at that point (s16()).
// The DCT and YCbCr kernels use __dmul_rn/__dadd_rn/__fsub_rn so that nvcc can never contract a multiply
// and an add into an FMA: the reference is built for baseline x86-64 and rounds after every operation.
#include "fb_common.cuh"
#include "fb_fused_plan.h"
#include "fb_direct_plan.h"
#include <type_traits>
#include "fb_pk_plan.h"
#include "fb_idct_fused.cuh"
#include "fb_subsample.cuh"
#include "fb_approx.cuh"
#include "fb_palette.cuh"
#include "fb_match.cuh"

#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>

namespace {

__device__ __forceinline__ int s16(int x) { return (int)(short)x; }

// smooth_tendency, reference transform/squeeze.h:61-77
__device__ __forceinline__ int smooth_tendency(int B, int a, int n) {
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

// One unsqueeze pair (squeeze.h:97-108): given previous reconstructed sample `prev`, current average, next
// average and the stored residual, produce A and B.
__device__ __forceinline__ void unsqueeze_pair(int prev, int avg, int next_avg, int res, int &A, int &B) {
    int tendency = smooth_tendency(prev, avg, next_avg);
    int diff = s16(res + tendency);
    A = s16(((avg << 1) + diff + (diff > 0 ? -(diff & 1) : (diff & 1))) >> 1);
    B = s16(A - diff);
}

// ---------------------------------------------------------------------------------------------------------
// v1 unsqueeze kernels: one thread per chain (row for horizontal, column for vertical).
// ---------------------------------------------------------------------------------------------------------

// inv_hsqueeze, squeeze.h:81-132.  avg: wa x h, res: wr x h (nullptr = zeros), out: (wa+wr) x h
__global__ void k_inv_hsqueeze_rows(const int16_t *__restrict__ avg, const int16_t *__restrict__ res, int16_t *__restrict__ out,
                                    int wa, int wr, int h) {
    int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h) return;
    const int wo = wa + wr;
    const int16_t *a = avg + (size_t)y * wa;
    const int16_t *r = res ? res + (size_t)y * wr : nullptr;
    int16_t *o = out + (size_t)y * wo;
    if (wr == 0) {          // nothing to merge: the reference ends up copying the averages
        for (int x = 0; x < wa; x++) o[x] = a[x];
        return;
    }
    int prev = a[0];
    for (int x = 0; x < wr; x++) {
        int av = a[x];
        int nx = (x + 1 < wa) ? a[x + 1] : av;
        int rs = r ? r[x] : 0;
        int A, B;
        unsqueeze_pair(x == 0 ? av : prev, av, nx, rs, A, B);
        o[2 * x] = (int16_t)A;
        o[2 * x + 1] = (int16_t)B;
        prev = B;
    }
    if (wo & 1) o[wo - 1] = a[wa - 1];
}

// inv_vsqueeze, squeeze.h:173-224.  avg: w x ha, res: w x hr, out: w x (ha+hr)
__global__ void k_inv_vsqueeze_cols(const int16_t *__restrict__ avg, const int16_t *__restrict__ res, int16_t *__restrict__ out,
                                    int w, int ha, int hr) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    const int ho = ha + hr;
    if (hr == 0) {
        for (int y = 0; y < ha; y++) out[(size_t)y * w + x] = avg[(size_t)y * w + x];
        return;
    }
    int prev = 0;
    int av = avg[x];
    for (int y = 0; y < hr; y++) {
        int nx = (y + 1 < ha) ? avg[(size_t)(y + 1) * w + x] : av;
        int rs = res ? res[(size_t)y * w + x] : 0;
        int A, B;
        unsqueeze_pair(y == 0 ? av : prev, av, nx, rs, A, B);
        out[(size_t)(2 * y) * w + x] = (int16_t)A;
        out[(size_t)(2 * y + 1) * w + x] = (int16_t)B;
        prev = B;
        av = nx;
    }
    if (ho & 1) out[(size_t)(ho - 1) * w + x] = avg[(size_t)(ha - 1) * w + x];
}

// ---------------------------------------------------------------------------------------------------------
// Tiled unsqueeze kernels.
//
// The inverse is a serial recurrence along the squeeze axis: the tendency of pair x needs B of pair x-1
// (squeeze.h:104, :208).  The carried state is ONE int16, and the recurrence forgets its start within a few pairs
// (SURVEY F6), so every chain is cut into segments that start K pairs early from a guessed state ("warm-up").  A
// segment is exact iff the state it reached at its first pair equals the true last B of the segment before it; that
// is checked inside the block and, on a mismatch, the segment is recomputed from the true state (and its successors
// re-checked) until every segment agrees with its predecessor.  Segment 0 starts at the true chain start, so by
// induction the result is bit-exact regardless of how good the guess was.
// ---------------------------------------------------------------------------------------------------------

// unsqueeze_pair with the tendency in closed form.  With t1 = B-a, t2 = a-n, a1=|t1|, a2=|t2|:
//   monotone (t1, t2 of one sign):  |tendency| = min((4*a1+3*a2+6)/12, 2*a1+1, 2*a2), sign = sign(t1+t2)
// which is squeeze.h:63-75 with both clamps folded into a min (d-(d&1) > 2k  <=>  d >= 2k+2;  d+(d&1) > 2k  <=>  d >= 2k+1).
// The closed form assumes no int16 wrap inside smooth_tendency, guaranteed for a1, a2 <= 16383; anything larger takes
// the literal path.
__device__ __forceinline__ void unsqueeze_pair_fast(int prev, int av, int nx, int rs, int &A, int &B) {
    const int t1 = prev - av, t2 = av - nx;
    const int a1 = abs(t1), a2 = abs(t2);
    if ((a1 | a2) > 16383) { unsqueeze_pair(prev, av, nx, rs, A, B); return; }
    const unsigned m = (unsigned)(4 * a1 + 3 * a2 + 6);
    const int q = (int)(__umulhi(m, 0xAAAAAAABu) >> 3);        // m / 12
    int d = min(min(q, 2 * a1 + 1), 2 * a2);
    const bool mono = ((t1 ^ t2) >= 0) | (t1 == 0);
    d = (t1 + t2) < 0 ? -d : d;
    const int tendency = mono ? d : 0;
    const int diff = s16(rs + tendency);
    A = s16(av + ((diff - (diff >> 31)) >> 1));                 // (2a + diff -+ (diff&1)) >> 1  ==  a + trunc(diff/2)
    B = s16(A - diff);
}

struct SqJob {
    const int16_t *avg, *res;   // res == nullptr: all-zero residual
    int16_t *out;
    int wa, wr;                 // horizontal: widths of avg / res planes; vertical: wa = width, wr unused
    int ha, hr;                 // horizontal: ha = rows; vertical: heights of avg / res planes
    int blocks;                 // blocks assigned to this job
};
struct SqJobs { SqJob j[4]; int n; };

constexpr int kWarm = 8;        // warm-up pairs (the recurrence re-joins the exact chain within <= 7 pairs in practice;
                                // a wrong guess only costs a repair round, never exactness)

// Horizontal: a block owns R complete rows of one plane.  thread = (row, segment).
// Shared memory: staged averages [R][PA] (+1 sentinel = last average, so that "next average" needs no bounds test),
// staged residuals [R][PR], output words (A | B<<16) [R][wr], final state per segment [R][nseg].
__device__ __forceinline__ int sq_round8(int v) { return (v + 7) & ~7; }

__global__ void __launch_bounds__(1024) k_inv_hsqueeze_tiled(SqJobs jobs, int R, int threads_per_row, int kSegP) {
    extern __shared__ __align__(16) unsigned char smraw[];
    int b = blockIdx.x, ji = 0;
    while (ji < jobs.n - 1 && b >= jobs.j[ji].blocks) { b -= jobs.j[ji].blocks; ji++; }
    const SqJob J = jobs.j[ji];
    const int wa = J.wa, wr = J.wr, wo = wa + wr, h = J.ha;
    const int y0 = b * R;
    const int rows = min(R, h - y0);
    if (rows <= 0) return;
    const int PA = sq_round8(wa) + 8, PR = sq_round8(wr) + 8;       // halfword pitches (16-byte multiples)
    int16_t *avgS = reinterpret_cast<int16_t *>(smraw);
    int16_t *resS = avgS + R * PA;
    unsigned *outS = reinterpret_cast<unsigned *>(resS + R * PR);
    int16_t *bfS = reinterpret_cast<int16_t *>(outS + (size_t)R * wr);      // [R][nseg] final B of every segment
    const int nseg = (wr + kSegP - 1) / kSegP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // ---- stage the rows: one warp per row at a time, 16-byte loads when the row is 16-byte aligned
    for (int r = warp; r < rows; r += nwarps) {
        const int16_t *ga = J.avg + (size_t)(y0 + r) * wa;
        int16_t *sa = avgS + r * PA;
        if ((wa & 7) == 0) {
            for (int c = lane * 8; c < wa; c += 256) *reinterpret_cast<uint4 *>(sa + c) = *reinterpret_cast<const uint4 *>(ga + c);
        } else {
            for (int c = lane; c < wa; c += 32) sa[c] = ga[c];
        }
        int16_t *sr = resS + r * PR;
        if (J.res) {
            const int16_t *gr = J.res + (size_t)(y0 + r) * wr;
            if ((wr & 7) == 0) {
                for (int c = lane * 8; c < wr; c += 256) *reinterpret_cast<uint4 *>(sr + c) = *reinterpret_cast<const uint4 *>(gr + c);
            } else {
                for (int c = lane; c < wr; c += 32) sr[c] = gr[c];
            }
        } else {
            for (int c = lane; c < wr; c += 32) sr[c] = 0;
        }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < rows; r += blockDim.x) avgS[r * PA + wa] = avgS[r * PA + wa - 1];     // sentinel
    __syncthreads();

    const int r = threadIdx.x / threads_per_row, sgm = threadIdx.x % threads_per_row;
    const bool active = r < rows && sgm < nseg;
    const int xs = sgm * kSegP, xe = min(xs + kSegP, wr);
    const int16_t *a = avgS + r * PA, *rr = resS + r * PR;
    unsigned *o = outS + (size_t)r * wr;
    int bw = 0x7fffffff;    // state the first owned pair consumed (0x7fffffff: the true chain start, exact by construction)
    // main pass over [xs, xe) from a given state; used by the first pass and by repairs
    auto owned = [&](int prev, bool first_is_chain_start) {
        int av = a[xs];
        if (first_is_chain_start) prev = av;            // x == 0: left = avg (squeeze.h:84-89)
        for (int x = xs; x < xe; x++) {
            const int nx = a[x + 1];
            int A, B;
            unsqueeze_pair_fast(prev, av, nx, rr[x], A, B);
            o[x] = (unsigned)(uint16_t)A | ((unsigned)(uint16_t)B << 16);
            prev = B;
            av = nx;
        }
        return prev;
    };
    if (active) {
        int bf;
        if (xs == 0) bf = owned(0, true);
        else {
            const int from = xs - kWarm;                // xs >= kSegP > kWarm
            int av = a[from], prev = av;                // guessed state
            for (int x = from; x < xs; x++) {
                const int nx = a[x + 1];
                int A, B;
                unsqueeze_pair_fast(prev, av, nx, rr[x], A, B);
                prev = B;
                av = nx;
            }
            bw = prev;
            bf = owned(prev, false);
        }
        bfS[r * nseg + sgm] = (int16_t)bf;
    }
    __syncthreads();
    // ---- verify / repair until every segment started from its predecessor's true final state
    for (;;) {
        bool bad = false;
        int want = 0;
        if (active && sgm > 0) {
            want = bfS[r * nseg + sgm - 1];
            bad = (want != bw);
        }
        if (!__syncthreads_or(bad)) break;
        if (bad) {
            bw = want;
            bfS[r * nseg + sgm] = (int16_t)owned(want, false);
        }
        __syncthreads();
    }
    // ---- store (coalesced); odd tail column is a copy of the last average (squeeze.h:129)
    if ((wo & 7) == 0) {        // rows are contiguous and 16-byte aligned in both memories
        const uint4 *so = reinterpret_cast<const uint4 *>(outS);
        uint4 *go = reinterpret_cast<uint4 *>(J.out + (size_t)y0 * wo);
        const int n16 = rows * wr / 4;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) go[i] = so[i];
    } else {
        for (int q = warp; q < rows; q += nwarps) {
            int16_t *go = J.out + (size_t)(y0 + q) * wo;
            const unsigned *so = outS + (size_t)q * wr;
            for (int i = lane; i < 2 * wr; i += 32) go[i] = (int16_t)((so[i >> 1] >> ((i & 1) * 16)) & 0xffff);
            if ((wo & 1) && lane == 0) go[wo - 1] = avgS[q * PA + wa - 1];
        }
    }
}

// Vertical: a block owns 32 complete columns of one plane.  thread = (column, segment along y); lanes = columns.
__global__ void __launch_bounds__(1024) k_inv_vsqueeze_tiled(SqJobs jobs, int nseg, int segp) {
    extern __shared__ __align__(16) unsigned char smraw[];
    int16_t *bfS = reinterpret_cast<int16_t *>(smraw);          // [nseg][32]
    int b = blockIdx.x, ji = 0;
    while (ji < jobs.n - 1 && b >= jobs.j[ji].blocks) { b -= jobs.j[ji].blocks; ji++; }
    const SqJob J = jobs.j[ji];
    const int w = J.wa, ha = J.ha, hr = J.hr, ho = ha + hr;
    const int x = b * 32 + (threadIdx.x & 31);
    const int sgm = threadIdx.x >> 5;
    const bool active = x < w && sgm < nseg && sgm * segp < hr;
    const int ys = sgm * segp, ye = min(ys + segp, hr);
    const int16_t *a = J.avg + x, *rr = J.res ? J.res + x : nullptr;
    int16_t *o = J.out + x;
    int bw = 0x7fffffff;
    auto owned = [&](int prev, bool first_is_chain_start) {
        int av = a[(size_t)ys * w];
        if (first_is_chain_start) prev = av;
        // inputs of eight steps are fetched together (they do not depend on the chain), then consumed
        for (int yb = ys; yb < ye; yb += 8) {
            int nxv[8], rsv[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int y = yb + j;
                nxv[j] = (y < ye && y + 1 < ha) ? a[(size_t)(y + 1) * w] : 0;
                rsv[j] = (y < ye && rr) ? rr[(size_t)y * w] : 0;
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int y = yb + j;
                if (y < ye) {
                    const int nx = (y + 1 < ha) ? nxv[j] : av;
                    int A, B;
                    unsqueeze_pair_fast(prev, av, nx, rsv[j], A, B);
                    o[(size_t)(2 * y) * w] = (int16_t)A;
                    o[(size_t)(2 * y + 1) * w] = (int16_t)B;
                    prev = B;
                    av = nx;
                }
            }
        }
        return prev;
    };
    if (active) {
        int bf;
        if (ys == 0) bf = owned(0, true);
        else {
            const int from = ys - kWarm;
            int av = a[(size_t)from * w], prev = av;
            int nxv[kWarm], rsv[kWarm];
#pragma unroll
            for (int j = 0; j < kWarm; j++) {
                nxv[j] = a[(size_t)(from + j + 1) * w];     // from + j + 1 <= ys < hr <= ha
                rsv[j] = rr ? rr[(size_t)(from + j) * w] : 0;
            }
#pragma unroll
            for (int j = 0; j < kWarm; j++) {
                int A, B;
                unsqueeze_pair_fast(prev, av, nxv[j], rsv[j], A, B);
                prev = B;
                av = nxv[j];
            }
            bw = prev;
            bf = owned(prev, false);
        }
        bfS[sgm * 32 + (threadIdx.x & 31)] = (int16_t)bf;
    }
    __syncthreads();
    for (;;) {
        bool bad = false;
        int want = 0;
        if (active && sgm > 0) {
            want = bfS[(sgm - 1) * 32 + (threadIdx.x & 31)];
            bad = (want != bw);
        }
        if (!__syncthreads_or(bad)) break;
        if (bad) {
            bw = want;
            bfS[sgm * 32 + (threadIdx.x & 31)] = (int16_t)owned(want, false);
        }
        __syncthreads();
    }
    // odd tail row: copy of the last average row (squeeze.h:217-222)
    if ((ho & 1) && x < w && sgm == 0) o[(size_t)(ho - 1) * w] = a[(size_t)(ha - 1) * w];
}

// ---------------------------------------------------------------------------------------------------------
// Pyramid kernel: the coarse levels of the unsqueeze chain in ONE launch.
//
// Every plane of the image has its own chain of unsqueeze steps (a step never mixes planes), so one block walks one
// plane's chain level by level for as long as the level's planes fit in shared memory (output <= 64 Ki samples),
// with a block-wide barrier between levels.  This replaces ~12 dependent launches of tiny grids, each of which would
// cost a launch latency plus one serial chain latency.  Same segment / warm-up / verify scheme as the tiled kernels.
// ---------------------------------------------------------------------------------------------------------
struct PyrLevel {
    const int16_t *avg, *res;
    int16_t *out;
    int wa, wr, ha, hr;         // horizontal: avg wa x ha, res wr x ha;  vertical: avg wa x ha, res wa x hr
    int horizontal;
};
constexpr int kPyrMaxLevels = 56;
struct PyrParams {
    PyrLevel lv[kPyrMaxLevels];
    int chain_start[9];
    int nchains;
};
constexpr int kPyrSeg = 8, kPyrWarm = 8;
constexpr int kPyrMaxOut = 16384;          // samples of a level's output plane (one block = one SM works on it)
constexpr int kPyrBufHalfwords = 20480;    // each of the two ping-pong plane buffers in shared memory (output of a level incl. pitch padding)
constexpr int kPyrResHalfwords = 12288;    // the residual staging buffer
constexpr int kPyrMaxItems = 4096;         // (chain, segment) items of a level
// halfword pitch of a plane of width w in shared memory: for a plane that a HORIZONTAL level reads (lanes = rows) the word
// pitch is made odd so that the rows fall into different banks
__host__ __device__ __forceinline__ int pyr_pitch(int w, bool read_by_horizontal) {
    int P = (w + 1) & ~1;
    if (read_by_horizontal && ((P >> 1) & 1) == 0) P += 2;
    return P;
}

// One segment of one chain, inputs and outputs in shared memory.
// ch_stride / step strides let the same code run along x (horizontal) or y (vertical).
__device__ __forceinline__ int pyr_segment(const int16_t *a, const int16_t *rr, int a_step, int r_step, int16_t *o, int o_step, int n_avg,
                                           int from, int xs, int xe, int prev, bool chain_start, int &bw) {
    int av = a[from * a_step];
    if (chain_start) prev = av;
    for (int x = from; x < xe; x++) {
        const int nx = (x + 1 < n_avg) ? a[(x + 1) * a_step] : av;
        int A, B;
        if (x == xs) bw = prev;
        unsqueeze_pair_fast(prev, av, nx, rr[x * r_step], A, B);
        if (x >= xs) { o[(2 * x) * o_step] = (int16_t)A; o[(2 * x + 1) * o_step] = (int16_t)B; }
        prev = B;
        av = nx;
    }
    return prev;
}

// The coarse levels of one plane's chain, level after level in ONE block: the plane stays in shared memory between the levels
// (two ping-pong buffers), only the residuals come from HBM and only the last level's output goes back.
__global__ void __launch_bounds__(1024) k_inv_squeeze_pyramid(PyrParams P) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int c = blockIdx.x;
    const int l0 = P.chain_start[c], l1 = P.chain_start[c + 1];
    int *bwS = reinterpret_cast<int *>(smraw);                  // [items] state assumed at the segment start
    int *bfS = bwS + kPyrMaxItems;                              // [items] state after the segment
    int16_t *bufA = reinterpret_cast<int16_t *>(bfS + kPyrMaxItems), *bufB = bufA + kPyrBufHalfwords, *resS = bufB + kPyrBufHalfwords;
    int16_t *avgS = bufA, *outS = bufB;
    for (int l = l0; l < l1; l++) {
        const PyrLevel L = P.lv[l];
        const bool H = L.horizontal != 0;
        const int wa = L.wa, ha = L.ha;
        const int wr = H ? L.wr : wa, hr = H ? ha : L.hr;           // residual plane dims
        const int wo = H ? wa + wr : wa, ho = H ? ha : ha + hr;
        const int PA = pyr_pitch(wa, H), PR = pyr_pitch(wr, H);
        const bool last = l + 1 == l1;
        const int PO = pyr_pitch(wo, !last && P.lv[l + 1].horizontal != 0);
        if (l == l0)
            for (int i = threadIdx.x; i < wa * ha; i += blockDim.x) { const int r = i / wa, q = i - r * wa; avgS[r * PA + q] = L.avg[i]; }
        for (int i = threadIdx.x; i < wr * hr; i += blockDim.x) { const int r = i / wr, q = i - r * wr; resS[r * PR + q] = L.res ? L.res[i] : (int16_t)0; }
        __syncthreads();
        const int nchain = H ? ha : wa;                 // independent chains
        const int npair = H ? wr : hr;                  // pairs along a chain
        const int navg = H ? wa : ha;
        const int nseg = (npair + kPyrSeg - 1) / kPyrSeg;
        const int items = nchain * nseg;                // item = chain + nchain * seg  (lanes = chains)
        auto run_item = [&](int item, bool repair, int state) {
            const int chain = item % nchain, seg = item / nchain;
            const int xs = seg * kPyrSeg, xe = min(xs + kPyrSeg, npair);
            const int16_t *a = H ? avgS + chain * PA : avgS + chain;
            const int16_t *rr = H ? resS + chain * PR : resS + chain;
            int16_t *o = H ? outS + chain * PO : outS + chain;
            const int a_step = H ? 1 : PA, r_step = H ? 1 : PR, o_step = H ? 1 : PO;
            int bw = 0x7fffffff, bf;
            if (repair) bf = pyr_segment(a, rr, a_step, r_step, o, o_step, navg, xs, xs, xe, state, false, bw);
            else if (xs < kPyrWarm + 1) { bf = pyr_segment(a, rr, a_step, r_step, o, o_step, navg, 0, xs, xe, 0, true, bw); if (xs == 0) bw = 0x7fffffff; else bw = 0x7ffffffe; }
            else {
                const int from = xs - kPyrWarm;
                bf = pyr_segment(a, rr, a_step, r_step, o, o_step, navg, from, xs, xe, a[from * a_step], false, bw);
            }
            if (!repair) bwS[item] = bw;
            bfS[item] = bf;
        };
        if (items <= kPyrMaxItems) {
            for (int item = threadIdx.x; item < items; item += blockDim.x) run_item(item, false, 0);
            __syncthreads();
            for (;;) {
                bool any = false;
                for (int item = threadIdx.x; item < items; item += blockDim.x) {
                    if (item < nchain) continue;                                // segment 0 starts at the true chain start
                    const int bw = bwS[item];
                    if (bw == 0x7ffffffe) continue;                             // ran from the chain start: exact
                    const int want = bfS[item - nchain];
                    if (want != bw) any = true;
                }
                if (!__syncthreads_or(any)) break;
                for (int item = threadIdx.x; item < items; item += blockDim.x) {
                    if (item < nchain) continue;
                    const int bw = bwS[item];
                    if (bw == 0x7ffffffe) continue;
                    const int want = bfS[item - nchain];
                    if (want != bw) { bwS[item] = want; run_item(item, true, want); }
                }
                __syncthreads();
            }
        } else {        // cannot happen for planes the planner admits, kept as a safe serial path
            for (int chain = threadIdx.x; chain < nchain; chain += blockDim.x) {
                const int16_t *a = H ? avgS + chain * PA : avgS + chain;
                const int16_t *rr = H ? resS + chain * PR : resS + chain;
                int16_t *o = H ? outS + chain * PO : outS + chain;
                int bw;
                pyr_segment(a, rr, H ? 1 : PA, H ? 1 : PR, o, H ? 1 : PO, navg, 0, 0, npair, 0, true, bw);
            }
        }
        // odd tail: copy of the last average column / row (squeeze.h:129, :217-222)
        if (H) { if (wo & 1) for (int r = threadIdx.x; r < ha; r += blockDim.x) outS[r * PO + wo - 1] = avgS[r * PA + wa - 1]; }
        else { if (ho & 1) for (int q = threadIdx.x; q < wa; q += blockDim.x) outS[(ho - 1) * PO + q] = avgS[(ha - 1) * PA + q]; }
        __syncthreads();
        // the output of the last level of this launch is what the rest of the chain reads; the levels in between never leave the SM
        if (last)
            for (int i = threadIdx.x; i < wo * ho; i += blockDim.x) { const int r = i / wo, q = i - r * wo; L.out[i] = outS[r * PO + q]; }
        int16_t *t = avgS; avgS = outS; outS = t;
    }
}

// ---------------------------------------------------------------------------------------------------------
// forward squeeze: fully parallel, one thread per residual sample (squeeze.h:135-170, 227-263)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pair_avg(int A, int B) { return s16((A + B + (A > B)) >> 1); }

__global__ void k_fwd_hsqueeze(const int16_t *__restrict__ in, int16_t *__restrict__ avg, int16_t *__restrict__ res, int w, int h) {
    const int wa = (w + 1) / 2, wr = w - wa;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)wa * h) return;
    int y = (int)(idx / wa), x = (int)(idx % wa);
    const int16_t *row = in + (size_t)y * w;
    if (x >= wr) {      // odd tail column: plain copy (squeeze.h:164-167)
        avg[(size_t)y * wa + x] = row[2 * x];
        return;
    }
    int A = row[2 * x], B = row[2 * x + 1];
    int av = pair_avg(A, B);
    avg[(size_t)y * wa + x] = (int16_t)av;
    int diff = s16(A - B);
    int nx = av;
    if (x + 1 < wr) nx = pair_avg(row[2 * x + 2], row[2 * x + 3]);
    else if (w & 1) nx = row[2 * x + 2];
    int left = (x > 0) ? row[2 * x - 1] : av;
    res[(size_t)y * wr + x] = (int16_t)s16(diff - smooth_tendency(left, av, nx));
}

__global__ void k_fwd_vsqueeze(const int16_t *__restrict__ in, int16_t *__restrict__ avg, int16_t *__restrict__ res, int w, int h) {
    const int ha = (h + 1) / 2, hr = h - ha;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)w * ha) return;
    int y = (int)(idx / w), x = (int)(idx % w);
    if (y >= hr) {
        avg[(size_t)y * w + x] = in[(size_t)(2 * y) * w + x];
        return;
    }
    int A = in[(size_t)(2 * y) * w + x], B = in[(size_t)(2 * y + 1) * w + x];
    int av = pair_avg(A, B);
    avg[(size_t)y * w + x] = (int16_t)av;
    int diff = s16(A - B);
    int nx = av;
    if (y + 1 < hr) nx = pair_avg(in[(size_t)(2 * y + 2) * w + x], in[(size_t)(2 * y + 3) * w + x]);
    else if (h & 1) nx = in[(size_t)(2 * y + 2) * w + x];
    int top = (y > 0) ? in[(size_t)(2 * y - 1) * w + x] : av;
    res[(size_t)y * w + x] = (int16_t)s16(diff - smooth_tendency(top, av, nx));
}

// ---------------------------------------------------------------------------------------------------------
// colour transforms, quantisation, clamp
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

// inv_YCoCg (ycocg.h:51-56) optionally followed by the final clamp of undo_transforms (image.cpp:107-113)
__global__ void k_ycocg(int16_t *__restrict__ c0, int16_t *__restrict__ c1, int16_t *__restrict__ c2, size_t n, int maxval, int inverse,
                        int lo, int hi, int do_clamp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (inverse) {
        int Y = clampi(c0[i], 0, maxval);
        int Co = c1[i], Cg = c2[i];
        int G = clampi(Y - ((-Cg) >> 1), 0, maxval);
        int B = clampi(Y + ((1 - Cg) >> 1) - (Co >> 1), 0, maxval);
        int R = clampi(Co + B, 0, maxval);
        if (do_clamp) { R = clampi(R, lo, hi); G = clampi(G, lo, hi); B = clampi(B, lo, hi); }
        c0[i] = (int16_t)R; c1[i] = (int16_t)G; c2[i] = (int16_t)B;
    } else {
        int R = c0[i], G = c1[i], B = c2[i];
        int Y = (((R + B) >> 1) + G) >> 1;
        int Co = R - B;
        int Cg = G - ((R + B) >> 1);
        c0[i] = (int16_t)Y; c1[i] = (int16_t)Co; c2[i] = (int16_t)Cg;
    }
}

// 8 samples per thread, 128-bit loads / stores (n8 = number of 8-sample groups; planes are 256-byte aligned)
__global__ void k_ycocg_inv_vec8(int16_t *__restrict__ c0, int16_t *__restrict__ c1, int16_t *__restrict__ c2, size_t n8, int maxval, int lo, int hi,
                                 int do_clamp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    uint4 a = reinterpret_cast<const uint4 *>(c0)[i], b = reinterpret_cast<const uint4 *>(c1)[i], c = reinterpret_cast<const uint4 *>(c2)[i];
    unsigned *pa = &a.x, *pb = &b.x, *pc = &c.x;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        unsigned ra = 0, rb = 0, rc = 0;
#pragma unroll
        for (int hlf = 0; hlf < 2; hlf++) {
            const int Yr = (int)(short)(pa[k] >> (16 * hlf)), Co = (int)(short)(pb[k] >> (16 * hlf)), Cg = (int)(short)(pc[k] >> (16 * hlf));
            const int Y = clampi(Yr, 0, maxval);
            int G = clampi(Y - ((-Cg) >> 1), 0, maxval);
            int B = clampi(Y + ((1 - Cg) >> 1) - (Co >> 1), 0, maxval);
            int R = clampi(Co + B, 0, maxval);
            if (do_clamp) { R = clampi(R, lo, hi); G = clampi(G, lo, hi); B = clampi(B, lo, hi); }
            ra |= (unsigned)(uint16_t)R << (16 * hlf); rb |= (unsigned)(uint16_t)G << (16 * hlf); rc |= (unsigned)(uint16_t)B << (16 * hlf);
        }
        pa[k] = ra; pb[k] = rb; pc[k] = rc;
    }
    reinterpret_cast<uint4 *>(c0)[i] = a; reinterpret_cast<uint4 *>(c1)[i] = b; reinterpret_cast<uint4 *>(c2)[i] = c;
}

__device__ __forceinline__ int16_t clamp_trunc(double x, int lo, int hi) {
    // CLAMP(x, l, u) evaluated in double, then the truncating double -> pixel_type conversion (ycbcr.h:55-57)
    double v = (x < (double)lo) ? (double)lo : ((x > (double)hi) ? (double)hi : x);
    return (int16_t)__double2int_rz(v);
}

// inv_YCbCr / fwd_YCbCr (ycbcr.h:49-58 / 81-90): float loads, double arithmetic in source order, no FMA
__global__ void k_ycbcr(int16_t *__restrict__ c0, int16_t *__restrict__ c1, int16_t *__restrict__ c2, size_t n, int minval, int maxval,
                        int inverse) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float half = (float)((maxval + 1) / 2);
    if (inverse) {
        float yy = (float)c0[i];
        float cb = __fsub_rn((float)c1[i], half);
        float cr = __fsub_rn((float)c2[i], half);
        double dy = (double)yy, dcb = (double)cb, dcr = (double)cr;
        double r = __dadd_rn(__dadd_rn(dy, __dmul_rn(1.402, dcr)), 0.5);
        double g = __dadd_rn(__dsub_rn(__dsub_rn(dy, __dmul_rn(0.344136, dcb)), __dmul_rn(0.714136, dcr)), 0.5);
        double b = __dadd_rn(__dadd_rn(dy, __dmul_rn(1.772, dcb)), 0.5);
        c0[i] = clamp_trunc(r, minval, maxval);
        c1[i] = clamp_trunc(g, minval, maxval);
        c2[i] = clamp_trunc(b, minval, maxval);
    } else {
        double r = (double)(float)c0[i], g = (double)(float)c1[i], b = (double)(float)c2[i];
        double dh = (double)half;
        double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.299, r), __dmul_rn(0.587, g)), __dmul_rn(0.114, b));
        double cb = __dadd_rn(__dsub_rn(__dsub_rn(dh, __dmul_rn(0.168736, r)), __dmul_rn(0.331264, g)), __dmul_rn(0.5, b));
        double cr = __dsub_rn(__dsub_rn(__dadd_rn(dh, __dmul_rn(0.5, r)), __dmul_rn(0.418688, g)), __dmul_rn(0.081312, b));
        c0[i] = clamp_trunc(yy, minval, maxval);
        c1[i] = clamp_trunc(cb, minval, maxval);
        c2[i] = clamp_trunc(cr, minval, maxval);
    }
}

__global__ void k_quantize(int16_t *__restrict__ p, size_t n, int q, int inverse) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v = p[i];
    p[i] = (int16_t)(inverse ? v * q : v / q);      // quantize.h:42 / 63 (rounded_div == n/d, :54)
}

// 8 samples per thread, 16-byte accesses (planes are 256-byte aligned allocations); the last n % 8 samples by the scalar kernels
__global__ void k_clamp_vec8(int16_t *__restrict__ p, size_t n8, int lo, int hi) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    uint4 v = reinterpret_cast<uint4 *>(p)[i];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int a = clampi((int)(short)(w[k] & 0xffffu), lo, hi), b = clampi((int)(short)(w[k] >> 16), lo, hi);
        w[k] = (uint32_t)(uint16_t)a | ((uint32_t)(uint16_t)b << 16);
    }
    reinterpret_cast<uint4 *>(p)[i] = make_uint4(w[0], w[1], w[2], w[3]);
}
__global__ void k_quantize_vec8(int16_t *__restrict__ p, size_t n8, int q, int inverse) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    uint4 v = reinterpret_cast<uint4 *>(p)[i];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int a = (int)(short)(w[k] & 0xffffu), b = (int)(short)(w[k] >> 16);
        const int ra = inverse ? a * q : a / q, rb = inverse ? b * q : b / q;      // quantize.h:42 / 63, int16 wrap on the store
        w[k] = (uint32_t)(uint16_t)ra | ((uint32_t)(uint16_t)rb << 16);
    }
    reinterpret_cast<uint4 *>(p)[i] = make_uint4(w[0], w[1], w[2], w[3]);
}
__global__ void k_clamp(int16_t *__restrict__ p, size_t n, int lo, int hi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    p[i] = (int16_t)clampi(p[i], lo, hi);
}

// ---------------------------------------------------------------------------------------------------------
// 8x8 DCT in double (dct.h:60-107): out = 0.0; out += k*in, columns then rows
// ---------------------------------------------------------------------------------------------------------
// kDCTMatrix[8u+x] = 0.5*alpha(u)*cos((2x+1)u*pi/16) to 10 decimals (dct.h:60-77)
#define C0 0.3535533906
#define C1 0.4903926402
#define C2 0.4619397663
#define C3 0.4157348062
#define C5 0.2777851165
#define C6 0.1913417162
#define C7 0.0975451610
__constant__ double kDCT[64] = {
    C0,  C0,  C0,  C0,  C0,  C0,  C0,  C0,
    C1,  C3,  C5,  C7, -C7, -C5, -C3, -C1,
    C2,  C6, -C6, -C2, -C2, -C6,  C6,  C2,
    C3, -C7, -C1, -C5,  C5,  C1,  C7, -C3,
    C0, -C0, -C0,  C0,  C0, -C0, -C0,  C0,
    C5, -C1,  C7,  C3, -C3, -C7,  C1, -C5,
    C6, -C2,  C2, -C6, -C6,  C2, -C2,  C6,
    C7, -C5,  C3, -C1,  C1, -C3,  C5, -C7,
};

struct Planes64 { const int16_t *p[64]; };
struct Planes64W { int16_t *p[64]; };

__device__ __forceinline__ void dct_1d(const double *in, int stride, double *out, bool inverse) {
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double k = inverse ? kDCT[8 * u + x] : kDCT[8 * x + u];
            acc = __dadd_rn(acc, __dmul_rn(k, in[u * stride]));
        }
        out[x * stride] = acc;
    }
}

// inv_DCT body (dct.h:281-291): thread per 8x8 block
__global__ void k_inv_dct(Planes64 pl, int16_t *__restrict__ out, int bw, int bh, float dc_offset) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)bw * bh) return;
    int by = (int)(idx / bw), bx = (int)(idx % bw);
    double block[64], tmp[64];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        int v = pl.p[i] ? pl.p[i][idx] : 0;
        block[i] = (i == 0) ? (double)__fadd_rn((float)v, dc_offset) : (double)v;
    }
#pragma unroll
    for (int x = 0; x < 8; ++x) dct_1d(&block[x], 8, &tmp[x], true);
#pragma unroll
    for (int y = 0; y < 8; ++y) dct_1d(&tmp[8 * y], 1, &block[8 * y], true);
    const int ow = bw * 8;
#pragma unroll
    for (int y = 0; y < 8; y++) {
        int16_t v[8];
#pragma unroll
        for (int x = 0; x < 8; x++) v[x] = (int16_t)__double2int_rz(round(block[y * 8 + x]));
        int4 pk;
        pk.x = (uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16);
        pk.y = (uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16);
        pk.z = (uint16_t)v[4] | ((uint32_t)(uint16_t)v[5] << 16);
        pk.w = (uint16_t)v[6] | ((uint32_t)(uint16_t)v[7] << 16);
        *reinterpret_cast<int4 *>(out + (size_t)(by * 8 + y) * ow + bx * 8) = pk;
    }
}

// fwd_DCT body (dct.h:322-332)
__global__ void k_fwd_dct(const int16_t *__restrict__ in, int w, int h, Planes64W pl, int bw, int bh, float dc_offset) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)bw * bh) return;
    int by = (int)(idx / bw), bx = (int)(idx % bw);
    double block[64], tmp[64];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        int r = by * 8 + (i >> 3), c = bx * 8 + (i & 7);      // repeating_edge_value, image.h:86
        r = r >= h ? h - 1 : r;
        c = c >= w ? w - 1 : c;
        block[i] = (double)in[(size_t)r * w + c];
    }
#pragma unroll
    for (int x = 0; x < 8; ++x) dct_1d(&block[x], 8, &tmp[x], false);
#pragma unroll
    for (int y = 0; y < 8; ++y) dct_1d(&tmp[8 * y], 1, &block[8 * y], false);
#pragma unroll
    for (int i = 0; i < 64; i++) {
        double v = round(block[i]);
        if (i == 0) v = __dsub_rn(v, (double)dc_offset);
        pl.p[i][idx] = (int16_t)__double2int_rz(v);
    }
}

// ---------------------------------------------------------------------------------------------------------
// min/max and interleave
// ---------------------------------------------------------------------------------------------------------
__global__ void k_minmax(const int16_t *__restrict__ p, size_t n, int *out2) {
    int mn = 0x7FFF, mx = -0x7FFF;      // LARGEST_VAL / SMALLEST_VAL, image.h:36-37
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int v = p[i];
        mn = min(mn, v);
        mx = max(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&out2[0], mn); atomicMax(&out2[1], mx); }
}

struct PlanesN { const int16_t *p[8]; };
__global__ void k_interleave(PlanesN pl, int nch, size_t npix, int bps, uint8_t *__restrict__ dst) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    for (int c = 0; c < nch; c++) {
        int v = pl.p[c][i];
        if (bps == 1) dst[i * nch + c] = (uint8_t)(v & 0xFF);
        else { dst[(i * nch + c) * 2] = (uint8_t)((v >> 8) & 0xFF); dst[(i * nch + c) * 2 + 1] = (uint8_t)(v & 0xFF); }
    }
}

inline unsigned nblocks(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

#define FB_LAUNCH_CHECK(ctx)                                                                       \
    do {                                                                                           \
        (ctx)->launches++;                                                                         \
        (ctx)->mark(__func__);                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ != cudaSuccess) { (ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(e__); return FB_ERR_CUDA; } \
    } while (0)

int fb_launch_inv_hsqueeze(fb_ctx *ctx, const int16_t *avg, const int16_t *res, int16_t *out, int wa, int wr, int h) {
    if (h <= 0 || wa <= 0) return FB_OK;
    k_inv_hsqueeze_rows<<<nblocks(h, 64), 64, 0, ctx->stream>>>(avg, res, out, wa, wr, h);
    FB_LAUNCH_CHECK(ctx);
    return FB_OK;
}
int fb_launch_inv_vsqueeze(fb_ctx *ctx, const int16_t *avg, const int16_t *res, int16_t *out, int w, int ha, int hr) {
    if (w <= 0 || ha <= 0) return FB_OK;
    k_inv_vsqueeze_cols<<<nblocks(w, 64), 64, 0, ctx->stream>>>(avg, res, out, w, ha, hr);
    FB_LAUNCH_CHECK(ctx);
    return FB_OK;
}
// Batched launch: up to four planes of one squeeze step in one kernel.
int fb_launch_inv_squeeze_batch(fb_ctx *ctx, int horizontal, int n, const int16_t *const *avg, const int16_t *const *res, int16_t *const *out,
                                const int *wa, const int *wr, const int *ha, const int *hr) {
    if (n <= 0) return FB_OK;
    if (n > 4) return FB_ERR_INVALID;
    SqJobs jobs;
    jobs.n = 0;
    int total = 0;
    if (horizontal) {
        int maxwr = 0, maxwa = 0;
        for (int i = 0; i < n; i++) { maxwr = std::max(maxwr, wr[i]); maxwa = std::max(maxwa, wa[i]); }
        // segment length: short segments for small levels (latency bound: fewer serial steps per thread), long ones for
        // big levels (throughput bound: less warm-up redundancy).  Odd => conflict-free shared-memory columns.
        long long pairs_total = 0;
        for (int i = 0; i < n; i++) pairs_total += (long long)wr[i] * ha[i];
        int kSegP = 33;
        if (pairs_total / 33 < 148LL * 1024) kSegP = 17;
        if (pairs_total / 17 < 148LL * 1024) kSegP = 9;
        while ((maxwr + kSegP - 1) / kSegP > 1024) kSegP += 8;
        const int nseg = std::max(1, (maxwr + kSegP - 1) / kSegP);
        int tpr = nseg;                                     // threads per row
        if (tpr > 1024) { ctx->err = "plane too wide for the tiled unsqueeze"; return FB_ERR_UNSUPPORTED; }
        int R = std::max(1, 256 / tpr);
        auto smem_for = [&](int r) { return (size_t)r * (((maxwa + 7) & ~7) + 8 + ((maxwr + 7) & ~7) + 8) * 2 + (size_t)r * maxwr * 4 + (size_t)r * nseg * 2 + 16; };
        while (R > 1 && smem_for(R) > 100 * 1024) R--;
        if (smem_for(R) > 200 * 1024) { ctx->err = "plane too wide for the tiled unsqueeze"; return FB_ERR_UNSUPPORTED; }
        for (int i = 0; i < n; i++) {
            if (ha[i] <= 0 || wa[i] <= 0) continue;
            if (wr[i] == 0) {       // nothing to merge: the output is the average plane (see k_inv_hsqueeze_rows)
                k_inv_hsqueeze_rows<<<nblocks(ha[i], 64), 64, 0, ctx->stream>>>(avg[i], res[i], out[i], wa[i], wr[i], ha[i]);
                ctx->launches++;
                continue;
            }
            SqJob &J = jobs.j[jobs.n++];
            J.avg = avg[i]; J.res = res[i]; J.out = out[i]; J.wa = wa[i]; J.wr = wr[i]; J.ha = ha[i]; J.hr = 0;
            J.blocks = (ha[i] + R - 1) / R;
            total += J.blocks;
        }
        if (!jobs.n) return FB_OK;
        const size_t smem = smem_for(R);
        if (smem > 48 * 1024 && !(ctx->smem_optin & fb_ctx::kOptHsqTiled)) {
            FB_CUDA(ctx, cudaFuncSetAttribute(k_inv_hsqueeze_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ctx->smem_optin |= fb_ctx::kOptHsqTiled;
        }
        int threads = R * tpr;
        threads = (threads + 31) / 32 * 32;
        k_inv_hsqueeze_tiled<<<total, threads, smem, ctx->stream>>>(jobs, R, tpr, kSegP);
    } else {
        int maxhr = 0;
        for (int i = 0; i < n; i++) maxhr = std::max(maxhr, hr[i]);
        // segments along y: at most 32 per block (1024 threads), at least kSegP pairs each
        long long pairs_total = 0;
        for (int i = 0; i < n; i++) pairs_total += (long long)wa[i] * hr[i];
        int segp = 32;
        if (pairs_total / 32 < 148LL * 1024) segp = 16;
        if (pairs_total / 16 < 148LL * 1024) segp = 8;
        while ((maxhr + segp - 1) / segp > 32) segp += 8;
        const int nseg = std::max(1, (maxhr + segp - 1) / segp);
        for (int i = 0; i < n; i++) {
            if (ha[i] <= 0 || wa[i] <= 0) continue;
            if (hr[i] == 0) {
                k_inv_vsqueeze_cols<<<nblocks(wa[i], 64), 64, 0, ctx->stream>>>(avg[i], res[i], out[i], wa[i], ha[i], hr[i]);
                ctx->launches++;
                continue;
            }
            SqJob &J = jobs.j[jobs.n++];
            J.avg = avg[i]; J.res = res[i]; J.out = out[i]; J.wa = wa[i]; J.wr = 0; J.ha = ha[i]; J.hr = hr[i];
            J.blocks = (wa[i] + 31) / 32;
            total += J.blocks;
        }
        if (!jobs.n) return FB_OK;
        k_inv_vsqueeze_tiled<<<total, 32 * nseg, (size_t)nseg * 32 * 2 + 16, ctx->stream>>>(jobs, nseg, segp);
    }
    FB_LAUNCH_CHECK(ctx);
    return FB_OK;
}
// Executes a planned sequence of unsqueeze steps: coarse levels of every plane in one pyramid launch, the rest as one
// batched launch per step.
// use_direct: steps whose planes are eligible run on the direct kernels (fb_direct_squeeze.cuh), which can also carry
// the epilogue (inverse YCoCg on the step that produces Co and Cg, final clamp on every plane's last step).
// ---------------------------------------------------------------------------------------------------------
// packed unsqueeze kernels (fb_pk_squeeze.cuh): TMA descriptors + launches of one squeeze step
// ---------------------------------------------------------------------------------------------------------
namespace ps {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// box_w x box_h tiles of a row-major w x h int16 plane without padding: out-of-range elements load as zero, stores are clipped
bool ps_make_tilemap(TileMap *m, const void *base, int w, int h, int box_w, int box_h, int swizzle_bytes) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn || (reinterpret_cast<uintptr_t>(base) & 15) || (w & 7) || box_w > 256 || box_h > 256) return false;
    if (swizzle_bytes && swizzle_bytes != box_w * 2) return false;        // the swizzle span is the tile row
    // Planes come out of a per-context pool, so the same (address, shape, box) recurs with every image: the encoded
    // descriptor (a pure function of these values) is kept instead of calling into the driver again.
    struct Key { const void *b; int w, h, bw, bh, s; bool operator==(const Key &o) const { return b == o.b && w == o.w && h == o.h && bw == o.bw && bh == o.bh && s == o.s; } };
    struct KeyHash { size_t operator()(const Key &k) const { return std::hash<const void *>()(k.b) ^ ((size_t)k.w * 1000003u) ^ ((size_t)k.h * 10007u) ^ ((size_t)k.bw << 20) ^ ((size_t)k.bh << 28) ^ (size_t)k.s; } };
    static std::mutex mu;
    static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
    const Key key{base, w, h, box_w, box_h, swizzle_bytes};
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *m = it->second; return true; }
        if (cache.size() > 65536) cache.clear();
    }
    const CUtensorMapSwizzle swz = swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE);
    const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
    const cuuint64_t strides[1] = {(cuuint64_t)w * sizeof(int16_t)};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    if (fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    std::lock_guard<std::mutex> lk(mu);
    cache.emplace(key, *m);
    return true;
}
}  // namespace ps

template <int NP, int EP>
static cudaError_t pk_launch_h(fb_ctx *ctx, const ps::HLaunch &L) {
    // per device, not per process: a context on a second GPU needs the opt-in too (remembered per context, one bit per variant)
    const unsigned bit = (unsigned)fb_ctx::kOptPkH << ((NP - 1) * 3 + EP);
    if (!(ctx->smem_optin & bit)) {
        cudaError_t e = cudaFuncSetAttribute(ps::k_pk_hsq<NP, EP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
        if (e != cudaSuccess) return e;
        ctx->smem_optin |= bit;
    }
    ps::k_pk_hsq<NP, EP><<<L.grid, 32 * L.warps_per_block, L.smem, ctx->stream>>>(L.jobs, L.warps_per_block, L.smem_per_warp);
    return cudaGetLastError();
}

// Runs the planes of one squeeze step that the packed kernels take; `rest` receives the indices (into ops) they left.
static int pk_run_step(fb_ctx *ctx, const std::vector<ps::StepOp> &ops, bool horizontal, const ps::StepEpilogue &E, int lo, int hi,
                       std::vector<int> &rest, bool *epilogue_done) {
    if (!ctx->pk_stats) {
        FB_CUDA(ctx, cudaMalloc((void **)&ctx->pk_stats, 2 * sizeof(int)));
        FB_CUDA(ctx, cudaMemsetAsync(ctx->pk_stats, 0, 2 * sizeof(int), ctx->stream));
    }
    ps::StepPlan P = ps::plan_step(ops, horizontal, E, lo, hi, ctx->sm_count, ctx->pk_stats);
    rest = P.leftover;
    *epilogue_done = P.epilogue_done;
    if (P.h.empty() && P.v.empty()) return FB_OK;
    if (P.scratch_bytes > ctx->pk_scratch_bytes) {      // grown rarely: the largest step of an image comes last
        if (ctx->pk_scratch) { FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->pk_scratch); ctx->pk_scratch = nullptr; }
        const size_t want = P.scratch_bytes + P.scratch_bytes / 2;
        FB_CUDA(ctx, cudaMalloc((void **)&ctx->pk_scratch, want));
        ctx->pk_scratch_bytes = want;
    }
    if (P.counters > ctx->pk_counters_n) {
        if (ctx->pk_counters) { FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->pk_counters); ctx->pk_counters = nullptr; }
        const int want = P.counters * 2 + 64;
        FB_CUDA(ctx, cudaMalloc((void **)&ctx->pk_counters, want * sizeof(int)));
        FB_CUDA(ctx, cudaMemsetAsync(ctx->pk_counters, 0, want * sizeof(int), ctx->stream));
        ctx->pk_counters_n = want;
    }
    ps::relocate(P, ctx->pk_scratch, ctx->pk_counters);
    for (auto &L : P.h) {
        cudaError_t e;
        if (L.ep == fq::kEpYCoCg) e = pk_launch_h<2, fq::kEpYCoCg>(ctx, L);
        else if (L.np == 2 && L.ep == fq::kEpClamp) e = pk_launch_h<2, fq::kEpClamp>(ctx, L);
        else if (L.np == 2) e = pk_launch_h<2, fq::kEpNone>(ctx, L);
        else if (L.ep == fq::kEpClamp) e = pk_launch_h<1, fq::kEpClamp>(ctx, L);
        else e = pk_launch_h<1, fq::kEpNone>(ctx, L);
        if (e != cudaSuccess) { ctx->err = std::string("k_pk_hsq launch: ") + cudaGetErrorString(e); return FB_ERR_CUDA; }
        ctx->launches++;
        ctx->mark(L.ep == fq::kEpYCoCg ? "k_pk_hsq(ycocg)" : "k_pk_hsq", L.bytes);
    }
    for (auto &L : P.v) {
        static const int vdepth = getenv("FB_PK_VDEPTH") ? atoi(getenv("FB_PK_VDEPTH")) : ps::kVDepthDefault;
        if (vdepth == 8) ps::k_pk_vsq<8><<<L.grid, 32 * L.warps_per_block, 0, ctx->stream>>>(L.jobs, L.warps_per_block);
        else if (vdepth == 4) ps::k_pk_vsq<4><<<L.grid, 32 * L.warps_per_block, 0, ctx->stream>>>(L.jobs, L.warps_per_block);
        else ps::k_pk_vsq<ps::kVDepthDefault><<<L.grid, 32 * L.warps_per_block, 0, ctx->stream>>>(L.jobs, L.warps_per_block);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { ctx->err = std::string("k_pk_vsq launch: ") + cudaGetErrorString(e); return FB_ERR_CUDA; }
        ctx->launches++;
        ctx->mark("k_pk_vsq", L.bytes);
    }
    static const bool pk_debug = getenv("FB_PK_DEBUG") != nullptr;
    if (pk_debug) {        // development aid: repairs / range flags so far, after every step
        int v[2] = {0, 0};
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(v, ctx->pk_stats, sizeof(v), cudaMemcpyDeviceToHost);
        for (auto &L : P.h) for (int j = 0; j < L.jobs.n; j++)
            fprintf(stderr, "[pk] h np=%d ep=%d wa=%d h=%d S=%d nseg=%d nrb=%d\n", L.np, L.ep, L.jobs.j[j].wa, L.jobs.j[j].h, L.jobs.j[j].S, L.jobs.j[j].nseg, L.jobs.j[j].nrb);
        for (auto &L : P.v) for (int j = 0; j < L.jobs.n; j++)
            fprintf(stderr, "[pk] v w=%d ha=%d S=%d nseg=%d ncg=%d\n", L.jobs.j[j].w, L.jobs.j[j].ha, L.jobs.j[j].S, L.jobs.j[j].nseg, L.jobs.j[j].ncg);
        fprintf(stderr, "[pk]   -> repaired %d, range-flagged %d (cumulative)\n", v[0], v[1]);
    }
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// self-test of the packed primitives on the device (tests/test_gpu_pk_squeeze.py): the 16x2 pair and the 16x2 inverse
// YCoCg against their exact 32-bit forms on pseudo-random inputs inside the admitted range; counts mismatches
// ---------------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ unsigned st_rng(unsigned &s) { s = s * 1664525u + 1013904223u; return s >> 8; }
__global__ void k_pk_selftest(int which, unsigned seed, int scale, int maxval, int iters, unsigned one, int *mism) {
    const ps::PK K = ps::pk_consts(one);
    unsigned s = seed ^ (blockIdx.x * 9781u + threadIdx.x * 6271u + 1u);
    int bad = 0;
    for (int it = 0; it < iters; it++) {
        int v[2][4];
        for (int k = 0; k < 2; k++) {
            const int amax = min(4 * scale, ps::kMaxAvg), rmax = min(4 * scale, ps::kMaxRes);
            const int av = (int)(st_rng(s) % (2 * amax + 1)) - amax;
            int nx = av + (int)(st_rng(s) % (2 * scale + 1)) - scale;
            nx = max(-ps::kMaxAvg, min(ps::kMaxAvg, nx));
            int pv = av + (int)(st_rng(s) % (4 * scale + 1)) - 2 * scale;
            pv = max(-8189, min(8189, pv));
            const int rs = (int)(st_rng(s) % (2 * rmax + 1)) - rmax;
            v[k][0] = pv; v[k][1] = av; v[k][2] = nx; v[k][3] = rs;
        }
        auto pk = [&](int i) { return (uint32_t)(uint16_t)v[0][i] | ((uint32_t)(uint16_t)v[1][i] << 16); };
        if (which == 0) {
            uint32_t A, B;
            ps::pk_step(K, pk(0), pk(1), ps::pneg(pk(1), K), ps::pneg(pk(2), K), pk(3), A, B);
            for (int k = 0; k < 2; k++) {
                int A2, B2;
                fq::unsqueeze_pair(v[k][0], v[k][1], v[k][2], v[k][3], A2, B2);
                const int Ag = (int)(short)(k ? A >> 16 : A & 0xffff), Bg = (int)(short)(k ? B >> 16 : B & 0xffff);
                if (Ag != A2 || Bg != B2) bad++;
            }
        } else {
            // Y anywhere in int16, Co / Cg anywhere a checked step can leave them
            int y[2], co[2], cg[2];
            for (int k = 0; k < 2; k++) { y[k] = (int)(short)st_rng(s); co[k] = (int)(st_rng(s) % 16379) - 8189; cg[k] = (int)(st_rng(s) % 16379) - 8189; }
            uint32_t R, G, B;
            const uint32_t mv = (uint32_t)(uint16_t)maxval * 0x00010001u;
            ps::pk_ycocg(K, (uint32_t)(uint16_t)y[0] | ((uint32_t)(uint16_t)y[1] << 16), (uint32_t)(uint16_t)co[0] | ((uint32_t)(uint16_t)co[1] << 16),
                         (uint32_t)(uint16_t)cg[0] | ((uint32_t)(uint16_t)cg[1] << 16), mv, R, G, B);
            for (int k = 0; k < 2; k++) {
                int R2, G2, B2;
                ps::ycocg_exact(y[k], co[k], cg[k], maxval, 0, 0, 0, R2, G2, B2);
                const int Rg = (int)(short)(k ? R >> 16 : R & 0xffff), Gg = (int)(short)(k ? G >> 16 : G & 0xffff), Bg = (int)(short)(k ? B >> 16 : B & 0xffff);
                if (Rg != R2 || Gg != G2 || Bg != B2) bad++;
            }
        }
        // the range accumulator: inside the range it must stay quiet, one value outside must trip it
        if (which == 0) {
            uint32_t acc = 0;
            acc = ps::pk_chk(acc, pk(1), ps::kMaxAvg * 0x00010001u);
            acc = ps::pk_chk(acc, pk(2), ps::kMaxAvg * 0x00010001u);
            if (ps::pk_chk_bad(acc, ps::kMaxAvg)) bad++;
            const int out = (it & 1) ? ps::kMaxAvg + 1 + (int)(st_rng(s) % 30000) : -ps::kMaxAvg - 1 - (int)(st_rng(s) % 30000);
            acc = ps::pk_chk(acc, (uint32_t)(uint16_t)out << ((it & 2) ? 16 : 0), ps::kMaxAvg * 0x00010001u);
            if (!ps::pk_chk_bad(acc, ps::kMaxAvg)) bad++;
        }
    }
    if (bad) atomicAdd(mism, bad);
}
}  // namespace

extern "C" FB_API int fb_selftest_packed(fb_ctx *ctx, int which, unsigned seed, int scale, int maxval, long long *mismatches) {
    if (!ctx || !mismatches || which < 0 || which > 1 || scale < 1) return FB_ERR_INVALID;
    int *d = nullptr;
    FB_CUDA(ctx, cudaMalloc((void **)&d, sizeof(int)));
    FB_CUDA(ctx, cudaMemsetAsync(d, 0, sizeof(int), ctx->stream));
    k_pk_selftest<<<296, 256, 0, ctx->stream>>>(which, seed, scale, maxval, 64, 0x00010001u, d);
    ctx->launches++;
    int h = 0;
    FB_CUDA(ctx, cudaMemcpyAsync(&h, d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    *mismatches = h;
    return FB_OK;
}

static int run_inv_squeeze_plan_per_level(fb_ctx *ctx, const std::vector<FbSqOp> &ops, bool use_direct, const FbSqEpilogue *ep, int *epilogue_done) {
    const int n = (int)ops.size();
    if (epilogue_done) *epilogue_done = 0;
    if (!n) return FB_OK;
    // chains: op k continues the chain whose last op produced its `avg` plane
    std::vector<int> chain(n, -1), prev(n, -1);
    int nchains = 0;
    for (int k = 0; k < n; k++) {
        for (int q = k - 1; q >= 0; q--)
            if (ops[q].out == ops[k].avg) { chain[k] = chain[q]; prev[k] = q; break; }
        if (chain[k] < 0) chain[k] = nchains++;
    }
    // fused prefix of every chain
    std::vector<char> fused(n, 0);
    auto level_fits = [&](const FbSqOp &o) {
        const long long wo = o.horizontal ? o.wa + o.wr : o.wa, ho = o.horizontal ? o.ha : o.ha + o.hr;
        const long long npair = o.horizontal ? o.wr : o.hr, nchain = o.horizontal ? o.ha : o.wa;
        const long long wres = o.horizontal ? o.wr : o.wa, hres = o.horizontal ? o.ha : o.hr;
        if (npair <= 0 || wo * ho > kPyrMaxOut) return false;
        if (nchain * ((npair + kPyrSeg - 1) / kPyrSeg) > kPyrMaxItems) return false;
        // the planes with their shared-memory pitches (at most 3 halfwords of padding per row) must fit the buffers
        return (wo + 3) * ho <= kPyrBufHalfwords && (long long)(o.wa + 3) * o.ha <= kPyrBufHalfwords && (wres + 3) * hres <= kPyrResHalfwords;
    };
    PyrParams P;
    P.nchains = 0;
    int nlv = 0;
    if (nchains <= 8) {
        for (int cidx = 0; cidx < nchains; cidx++) {
            const int start = nlv;
            for (int k = 0; k < n; k++) {
                if (chain[k] != cidx) continue;
                if ((prev[k] >= 0 && !fused[prev[k]]) || !level_fits(ops[k]) || nlv >= kPyrMaxLevels) break;
                PyrLevel &L = P.lv[nlv++];
                L.avg = ops[k].avg; L.res = ops[k].res; L.out = ops[k].out;
                L.wa = ops[k].wa; L.wr = ops[k].wr; L.ha = ops[k].ha; L.hr = ops[k].hr; L.horizontal = ops[k].horizontal;
                fused[k] = 1;
            }
            if (nlv > start) { P.chain_start[P.nchains++] = start; }
        }
        P.chain_start[P.nchains] = nlv;
    }
    if (P.nchains > 0) {
        const size_t smem = 2 * kPyrMaxItems * sizeof(int) + (2 * (size_t)kPyrBufHalfwords + kPyrResHalfwords) * sizeof(int16_t) + 64;
        if (!(ctx->smem_optin & fb_ctx::kOptPyramid)) {
            FB_CUDA(ctx, cudaFuncSetAttribute(k_inv_squeeze_pyramid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->smem_optin |= fb_ctx::kOptPyramid;
        }
        k_inv_squeeze_pyramid<<<P.nchains, 1024, smem, ctx->stream>>>(P);
        FB_LAUNCH_CHECK(ctx);
    }
    // ---- can the epilogue ride on the direct kernels?  Every final plane's last op must be direct-eligible (and not
    // part of the pyramid launch); for YCoCg the very last step must be the horizontal step producing Co and Cg.
    auto as_step_op = [&](const FbSqOp &o) {
        dq::StepOp so;
        so.avg = o.avg; so.res = o.res; so.out = o.out; so.wa = o.wa; so.wr = o.wr; so.ha = o.ha; so.hr = o.hr; so.clamp = 0;
        return so;
    };
    std::vector<char> is_final(n, 1), clamp_op(n, 0);
    for (int k = 0; k < n; k++)
        for (int q = k + 1; q < n; q++) if (ops[q].avg == ops[k].out) is_final[k] = 0;
    // Measured on B200 (profiles/): the direct kernel wins every horizontal step, the tiled kernel (lanes = adjacent
    // columns, warps = segments) every vertical one; FB_SQUEEZE_DIRECT_V=1 sends vertical steps to the direct kernel too.
    static const bool direct_v = getenv("FB_SQUEEZE_DIRECT_V") != nullptr;
    bool ep_ok = use_direct && ep && ep->kind != 0 && !(ep->kind == 2 && ep->maxval < 0);     // dq::clamp0 wants maxval >= 0
    int ico = -1, icg = -1;
    if (ep_ok) {
        const int last_step = ops[n - 1].step;
        for (int k = 0; k < n && ep_ok; k++) {
            if (!is_final[k]) continue;
            const bool elig = !fused[k] && (ops[k].horizontal ? dq::h_eligible(as_step_op(ops[k])) : (direct_v && dq::v_eligible(as_step_op(ops[k]))));
            if (ep->kind == 2 && ops[k].out == ep->ycc[0]) { if (ops[k].step == last_step) ep_ok = false; continue; }   // Y: consumed by the epilogue, never clamped
            if (ep->kind == 2 && (ops[k].out == ep->ycc[1] || ops[k].out == ep->ycc[2])) {
                if (!elig || !ops[k].horizontal || ops[k].step != last_step) ep_ok = false;
                (ops[k].out == ep->ycc[1] ? ico : icg) = k;
                continue;
            }
            if (ep->do_clamp) { if (!elig) ep_ok = false; clamp_op[k] = 1; }
        }
        if (ep->kind == 2) {
            if (ico < 0 || icg < 0 || !ep->rout) ep_ok = false;
            else if (ops[ico].wa != ops[icg].wa || ops[ico].ha != ops[icg].ha) ep_ok = false;
            bool have_y = false;
            for (int k = 0; k < n; k++) if (is_final[k] && ops[k].out == ep->ycc[0]) have_y = true;
            if (!have_y) ep_ok = false;
        }
    }
    if (!ep_ok) std::fill(clamp_op.begin(), clamp_op.end(), 0);
    // ---- the remaining ops, step by step: direct kernels where eligible, the tiled kernels for the rest
    int k = 0;
    bool ep_applied = false;
    while (k < n) {
        if (fused[k]) { k++; continue; }
        const int step = ops[k].step, horizontal = ops[k].horizontal;
        std::vector<int> idx;
        int q = k;
        while (q < n && ops[q].step == step) { if (!fused[q]) idx.push_back(q); q++; }
        std::vector<int> rest = idx;
        static const bool pk_env_off = getenv("FB_SQUEEZE_PK") && atoi(getenv("FB_SQUEEZE_PK")) == 0;
        const bool last_step = step == ops[n - 1].step;
        bool pk_took_epilogue = false;
        if (use_direct && ctx->pk_mode && !pk_env_off && ctx->sq_maxval >= 0 && ctx->sq_maxval <= 1023) {
            // packed int16x2 kernels first (fb_pk_squeeze.cuh); what they refuse falls through to the older kernels
            std::vector<ps::StepOp> sops;
            for (int i : idx) {
                ps::StepOp so;
                so.avg = ops[i].avg; so.res = ops[i].res; so.out = ops[i].out; so.wa = ops[i].wa; so.wr = ops[i].wr; so.ha = ops[i].ha; so.hr = ops[i].hr;
                so.clamp = clamp_op[i];
                sops.push_back(so);
            }
            ps::StepEpilogue E;
            if (ep_ok && ep->kind == 2 && last_step && horizontal) {
                E.enabled = 1; E.yin = ep->ycc[0]; E.rout = ep->rout; E.co_out = ep->ycc[1]; E.cg_out = ep->ycc[2];
                E.maxval = ep->maxval; E.lo = ep->lo; E.hi = ep->hi;
                E.do_clamp = (ep->do_clamp && !(ep->lo <= 0 && ep->hi >= ep->maxval)) ? 1 : 0;
            }
            std::vector<int> left;
            int rc = pk_run_step(ctx, sops, horizontal != 0, E, ep ? ep->lo : 0, ep ? ep->hi : 0, left, &pk_took_epilogue);
            if (rc) return rc;
            std::vector<int> nidx;
            for (int i : left) nidx.push_back(idx[i]);
            idx = nidx;
            rest = idx;
            if (pk_took_epilogue) ep_applied = true;
        }
        if (idx.empty()) { k = q; continue; }
        if (use_direct && (horizontal || direct_v)) {
            std::vector<dq::StepOp> sops;
            for (int i : idx) { dq::StepOp so = as_step_op(ops[i]); so.clamp = clamp_op[i]; sops.push_back(so); }
            dq::StepEpilogue E;
            if (ep_ok && ep->kind == 2 && step == ops[n - 1].step && !pk_took_epilogue) {
                E.enabled = 1; E.yin = ep->ycc[0]; E.rout = ep->rout; E.co_out = ep->ycc[1]; E.cg_out = ep->ycc[2];
                E.maxval = ep->maxval; E.lo = ep->lo; E.hi = ep->hi;
                // inv_YCoCg leaves R, G, B in [0, maxval] (ycocg.h:51-56): a final clamp to a range that contains it is the identity
                E.do_clamp = (ep->do_clamp && !(ep->lo <= 0 && ep->hi >= ep->maxval)) ? 1 : 0;
            }
            dq::StepPlan SP = dq::plan_step(sops, horizontal != 0, E, ep ? ep->lo : 0, ep ? ep->hi : 0, ctx->sm_count);
            if (E.enabled && !SP.epilogue_done) { ctx->err = "internal: YCoCg epilogue planned but not placed"; return FB_ERR_INVALID; }
            if (!(ctx->smem_optin & fb_ctx::kOptDirect)) {
                FB_CUDA(ctx, cudaFuncSetAttribute(dq::k_inv_hsq_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
                FB_CUDA(ctx, cudaFuncSetAttribute(dq::k_inv_vsq_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
                ctx->smem_optin |= fb_ctx::kOptDirect;
            }
            if (SP.hj.n) {
                dq::k_inv_hsq_direct<<<SP.h_grid, SP.h_threads, SP.h_smem, ctx->stream>>>(SP.hj);
                ctx->launches++;
                ctx->mark(SP.epilogue_done ? "k_inv_hsq_direct(ycocg)" : "k_inv_hsq_direct", SP.bytes);
            }
            if (SP.vj.n) {
                dq::k_inv_vsq_direct<<<SP.v_grid, SP.v_threads, SP.v_smem, ctx->stream>>>(SP.vj);
                ctx->launches++;
                ctx->mark("k_inv_vsq_direct", SP.bytes);
            }
            cudaError_t e__ = cudaGetLastError();
            if (e__ != cudaSuccess) { ctx->err = std::string("direct unsqueeze launch: ") + cudaGetErrorString(e__); return FB_ERR_CUDA; }
            if (SP.epilogue_done) ep_applied = true;
            rest.clear();
            for (int i : SP.leftover) {
                if (clamp_op[idx[i]]) { ctx->err = "internal: clamp planned on an op the direct kernels refused"; return FB_ERR_INVALID; }
                rest.push_back(idx[i]);
            }
        }
        for (size_t r0 = 0; r0 < rest.size(); r0 += 4) {
            const int16_t *avgp[4], *resp[4];
            int16_t *outp[4];
            int wa[4], wr[4], ha[4], hr[4], m = 0;
            for (size_t r = r0; r < rest.size() && m < 4; r++, m++) {
                const FbSqOp &o = ops[rest[r]];
                avgp[m] = o.avg; resp[m] = o.res; outp[m] = o.out; wa[m] = o.wa; wr[m] = o.wr; ha[m] = o.ha; hr[m] = o.hr;
            }
            int rc = fb_launch_inv_squeeze_batch(ctx, horizontal, m, avgp, resp, outp, wa, wr, ha, hr);
            if (rc) return rc;
        }
        k = q;
    }
    if (ep_ok && epilogue_done) {
        if (ep->kind == 2) *epilogue_done = ep_applied ? 2 : 0;
        else *epilogue_done = 1;
        if (ep->kind == 2 && !ep_applied) { ctx->err = "internal: YCoCg epilogue not applied"; return FB_ERR_INVALID; }
    }
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Entry point of the Squeeze inverse.  Default: one launch per squeeze step on the direct kernels
// (fb_direct_squeeze.cuh), pyramid kernel for the coarse levels, tiled kernels for ineligible planes.  Optional
// (FB_OPT_SQUEEZE_MODE >= 2 / FB_SQUEEZE_MODE=fused): the multi-level fused tile kernels (fb_fused_squeeze.cuh) + one
// verification launch; tunables for that path: FB_FQ_TILE=WxH, FB_FQ_LEVELS, FB_FQ_COARSE, FB_FQ_THREADS,
// FB_FQ_FORCE_FALLBACK.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct FqTunables {
    int mode = 0;               // 0 direct per-step kernels (default), 1 tiled per-step kernels only, 4 fused tile kernels
    int force_fallback = 0;
    fq::PlanOptions opt;
    FqTunables() {
        if (const char *m = getenv("FB_SQUEEZE_MODE")) mode = std::string(m) == "perlevel" ? 1 : (std::string(m) == "fused" ? 4 : 0);
        if (const char *t = getenv("FB_FQ_TILE")) { int a = 0, b = 0; if (sscanf(t, "%dx%d", &a, &b) == 2) { opt.tile_w = a; opt.tile_h = b; } }
        if (const char *t = getenv("FB_FQ_LEVELS")) opt.levels_per_launch = atoi(t);
        if (const char *t = getenv("FB_FQ_COARSE")) opt.coarse_dim = atoi(t);
        if (const char *t = getenv("FB_FQ_THREADS")) opt.threads_per_gang = atoi(t);
        if (const char *t = getenv("FB_FQ_FORCE_FALLBACK")) force_fallback = atoi(t);
    }
};
const FqTunables &fq_tunables() { static FqTunables t; return t; }
}  // namespace

int fb_run_inv_squeeze_plan(fb_ctx *ctx, const std::vector<FbSqOp> &ops, const FbSqEpilogue *ep, int *epilogue_done) {
    if (epilogue_done) *epilogue_done = 0;
    if (ops.empty()) return FB_OK;
    const FqTunables &tun = fq_tunables();
    fq::Plan P;
    const int mode = ctx->fq_mode ? ctx->fq_mode : tun.mode;
    if (mode >= 2) {
        std::vector<fq::PlanOp> po(ops.size());
        for (size_t i = 0; i < ops.size(); i++) {
            po[i].step = ops[i].step; po[i].horizontal = ops[i].horizontal;
            po[i].avg = ops[i].avg; po[i].res = ops[i].res; po[i].out = ops[i].out;
            po[i].wa = ops[i].wa; po[i].wr = ops[i].wr; po[i].ha = ops[i].ha; po[i].hr = ops[i].hr;
        }
        fq::EpilogueSpec E;
        if (ep) {
            E.kind = ep->kind; E.maxval = ep->maxval; E.lo = ep->lo; E.hi = ep->hi; E.do_clamp = ep->do_clamp;
            for (int j = 0; j < 3; j++) E.ycc[j] = ep->ycc[j];
        }
        P = fq::make_plan(po, E, tun.opt);
        // an epilogue that cannot be fused is left to the caller; without it the plan is still good
        if (P.ok && ep && ep->kind != fq::kEpNone && !P.epilogue_fused) { E = fq::EpilogueSpec(); P = fq::make_plan(po, E, tun.opt); }
    }
    if (!P.ok) return run_inv_squeeze_plan_per_level(ctx, ops, mode != 1, mode != 1 ? ep : nullptr, epilogue_done);

    if (!(ctx->smem_optin & fb_ctx::kOptFq)) {
        FB_CUDA(ctx, cudaFuncSetAttribute(fq::k_fq_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        FB_CUDA(ctx, cudaFuncSetAttribute(fq::k_fq_verify_fallback, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ctx->smem_optin |= fb_ctx::kOptFq;
    }
    if (!ctx->fq_counters) {
        FB_CUDA(ctx, cudaMalloc((void **)&ctx->fq_counters, 8 * sizeof(int)));
        FB_CUDA(ctx, cudaMemsetAsync(ctx->fq_counters, 0, 8 * sizeof(int), ctx->stream));
    }
    const int force = tun.force_fallback ? tun.force_fallback : (mode == 2 ? 1 : (mode == 3 ? 2 : 0));
    const bool verify = P.need_verify || force;
    unsigned char *scratch = nullptr;
    if (P.scratch_bytes) FB_CUDA(ctx, cudaMallocAsync((void **)&scratch, P.scratch_bytes, ctx->stream));
    fq::relocate_scratch(P, scratch, ctx->fq_counters);
    if (verify) FB_CUDA(ctx, cudaMemsetAsync(ctx->fq_counters, 0, 4 * sizeof(int), ctx->stream));     // per-run part, see fq::VerifyParams
    for (size_t li = 0; li < P.launches.size(); li++) {
        auto &L = P.launches[li];
        fq::k_fq_tiles<<<L.grid, L.threads, L.smem, ctx->stream>>>(L.task);
        ctx->launches++;
        ctx->mark(li + 1 == P.launches.size() ? "k_fq_tiles(last)" : "k_fq_tiles", L.bytes);
        cudaError_t e__ = cudaGetLastError();
        if (e__ != cudaSuccess) { ctx->err = std::string("k_fq_tiles launch: ") + cudaGetErrorString(e__); return FB_ERR_CUDA; }
    }
    if (verify) {
        P.verify.force = force;
        if (force == 2 && !P.verify.bad_list) P.verify.force = 0;      // nothing speculative in this run
        void *args[] = {(void *)&P.verify};
        // one CTA per SM is always co-resident (<= 200 KiB of shared memory, <= 512 threads)
        FB_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)fq::k_fq_verify_fallback, dim3(ctx->sm_count), dim3(P.verify_threads), args,
                                                 P.verify_smem, ctx->stream));
        ctx->launches++;
        ctx->mark("k_fq_verify_fallback", 0);
    }
    if (scratch) cudaFreeAsync(scratch, ctx->stream);
    if (epilogue_done) *epilogue_done = P.epilogue_fused ? 1 : 0;
    return FB_OK;
}
int fb_launch_fwd_hsqueeze(fb_ctx *ctx, const int16_t *in, int16_t *avg, int16_t *res, int w, int h) {
    size_t n = (size_t)((w + 1) / 2) * h;
    if (!n) return FB_OK;
    k_fwd_hsqueeze<<<nblocks(n, 256), 256, 0, ctx->stream>>>(in, avg, res, w, h);
    FB_LAUNCH_CHECK(ctx);
    return FB_OK;
}
int fb_launch_fwd_vsqueeze(fb_ctx *ctx, const int16_t *in, int16_t *avg, int16_t *res, int w, int h) {
    size_t n = (size_t)w * ((h + 1) / 2);
    if (!n) return FB_OK;
    k_fwd_vsqueeze<<<nblocks(n, 256), 256, 0, ctx->stream>>>(in, avg, res, w, h);
    FB_LAUNCH_CHECK(ctx);
    return FB_OK;
}
int fb_launch_ycocg(fb_ctx *ctx, int16_t *c0, int16_t *c1, int16_t *c2, size_t n, int maxval, int inverse, int lo, int hi, int do_clamp) {
    if (!n) return FB_OK;
    const bool aligned = (((uintptr_t)c0 | (uintptr_t)c1 | (uintptr_t)c2) & 15) == 0;
    if (inverse && aligned && n >= 8) {
        const size_t n8 = n / 8;
        k_ycocg_inv_vec8<<<nblocks(n8, 256), 256, 0, ctx->stream>>>(c0, c1, c2, n8, maxval, lo, hi, do_clamp);
        const size_t done = n8 * 8;
        if (done < n) k_ycocg<<<1, 32, 0, ctx->stream>>>(c0 + done, c1 + done, c2 + done, n - done, maxval, inverse, lo, hi, do_clamp);
    } else {
        k_ycocg<<<nblocks(n, 256), 256, 0, ctx->stream>>>(c0, c1, c2, n, maxval, inverse, lo, hi, do_clamp);
    }
    FB_LAUNCH_CHECK(ctx);
    return FB_OK;
}
int fb_launch_ycbcr(fb_ctx *ctx, int16_t *c0, int16_t *c1, int16_t *c2, size_t n, int minval, int maxval, int inverse) {
    if (!n) return FB_OK;
    k_ycbcr<<<nblocks(n, 256), 256, 0, ctx->stream>>>(c0, c1, c2, n, minval, maxval, inverse);
    FB_LAUNCH_CHECK(ctx);
    return FB_OK;
}
int fb_launch_quantize(fb_ctx *ctx, int16_t *p, size_t n, int q, int inverse) {
    if (!n) return FB_OK;
    const size_t n8 = (reinterpret_cast<uintptr_t>(p) & 15) == 0 ? n / 8 : 0;
    if (n8) {
        k_quantize_vec8<<<nblocks(n8, 256), 256, 0, ctx->stream>>>(p, n8, q, inverse);
        FB_LAUNCH_CHECK(ctx);
    }
    if (n > 8 * n8) {
        k_quantize<<<nblocks(n - 8 * n8, 256), 256, 0, ctx->stream>>>(p + 8 * n8, n - 8 * n8, q, inverse);
        FB_LAUNCH_CHECK(ctx);
    }
    return FB_OK;
}
int fb_launch_clamp(fb_ctx *ctx, int16_t *p, size_t n, int lo, int hi) {
    if (!n) return FB_OK;
    const size_t n8 = (reinterpret_cast<uintptr_t>(p) & 15) == 0 ? n / 8 : 0;
    if (n8) {
        k_clamp_vec8<<<nblocks(n8, 256), 256, 0, ctx->stream>>>(p, n8, lo, hi);
        FB_LAUNCH_CHECK(ctx);
    }
    if (n > 8 * n8) {
        k_clamp<<<nblocks(n - 8 * n8, 256), 256, 0, ctx->stream>>>(p + 8 * n8, n - 8 * n8, lo, hi);
        FB_LAUNCH_CHECK(ctx);
    }
    return FB_OK;
}
int fb_launch_inv_dct(fb_ctx *ctx, const int16_t *const *planes64, int16_t *out, int bw, int bh, float dc_offset) {
    size_t n = (size_t)bw * bh;
    if (!n) return FB_OK;
    Planes64 pl;
    for (int i = 0; i < 64; i++) pl.p[i] = planes64[i];
    k_inv_dct<<<nblocks(n, 64), 64, 0, ctx->stream>>>(pl, out, bw, bh, dc_offset);
    FB_LAUNCH_CHECK(ctx);
    return FB_OK;
}
// dequantise + inverse DCT (+ inverse YCbCr + clamp) of up to three components in one launch (fb_idct_fused.cuh)
int fb_launch_idct_fused(fb_ctx *ctx, const int16_t *const (*planes)[64], const int (*q)[64], int16_t *const *out, int ncomp, int bw, int bh,
                         float dc_offset, int ycbcr, int minval, int maxval) {
    if (ncomp < 1 || ncomp > 3 || bw < 1 || bh < 1 || bh > 65535) { ctx->err = "fused inverse DCT: bad geometry"; return FB_ERR_INVALID; }
    idf::Job J;
    memset(&J, 0, sizeof(J));
    for (int c = 0; c < ncomp; c++) {
        for (int k = 0; k < 64; k++) { J.pl[c][k] = planes[c][k]; J.q[c][k] = q[c][k]; }
        J.out[c] = out[c];
    }
    J.ncomp = ncomp; J.bw = bw; J.bh = bh; J.dc_offset = dc_offset; J.ycbcr = ycbcr; J.minval = minval; J.maxval = maxval;
    const dim3 grid((unsigned)((bw + idf::kBlocksPerCta - 1) / idf::kBlocksPerCta), (unsigned)bh);
    idf::k_idct_ycbcr<<<grid, 256, 0, ctx->stream>>>(J);
    ctx->launches++;
    ctx->mark("k_idct_ycbcr", 4.0 * 64.0 * bw * bh * ncomp);      // every coefficient read once, every sample written once
    cudaError_t e__ = cudaGetLastError();
    if (e__ != cudaSuccess) { ctx->err = std::string("k_idct_ycbcr launch: ") + cudaGetErrorString(e__); return FB_ERR_CUDA; }
    return FB_OK;
}
int fb_launch_fwd_dct(fb_ctx *ctx, const int16_t *in, int w, int h, int16_t *const *planes64, int bw, int bh, float dc_offset) {
    size_t n = (size_t)bw * bh;
    if (!n) return FB_OK;
    Planes64W pl;
    for (int i = 0; i < 64; i++) pl.p[i] = planes64[i];
    k_fwd_dct<<<nblocks(n, 64), 64, 0, ctx->stream>>>(in, w, h, pl, bw, bh, dc_offset);
    FB_LAUNCH_CHECK(ctx);
    return FB_OK;
}
int fb_launch_minmax(fb_ctx *ctx, const int16_t *p, size_t n, int *out2_dev) {
    int init[2] = {0x7FFF, -0x7FFF};
    FB_CUDA(ctx, cudaMemcpyAsync(out2_dev, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    if (!n) return FB_OK;
    unsigned nb = nblocks(n, 256);
    if (nb > 1184) nb = 1184;
    k_minmax<<<nb, 256, 0, ctx->stream>>>(p, n, out2_dev);
    FB_LAUNCH_CHECK(ctx);
    return FB_OK;
}
int fb_launch_palette_inv(fb_ctx *ctx, int16_t *const *out_planes, int nb, const int16_t *palette, int ncolors, size_t n) {
    if (!n) return FB_OK;
    if (nb < 1 || nb > pl::kMaxPlanes || ncolors < 1) return FB_ERR_UNSUPPORTED;
    pl::Planes P;
    for (int c = 0; c < pl::kMaxPlanes; c++) P.p[c] = c < nb ? out_planes[c] : nullptr;
    pl::k_palette_inv<<<nblocks(n, 256), 256, 0, ctx->stream>>>(P, palette, n, ncolors, nb);
    ctx->launches++;
    ctx->mark("k_palette_inv", 2.0 * (double)n * (1 + nb));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ctx->err = std::string("kernel launch: ") + cudaGetErrorString(e); return FB_ERR_CUDA; }
    return FB_OK;
}
// fwd_palette, step 1: the distinct colours of nb planes.  Returns the sorted keys (pl::pack_colour order = the reference's
// std::set order) in `sorted`, or *too_many = 1 when there are more than `limit`.
int fb_palette_collect(fb_ctx *ctx, int16_t *const *planes, int nb, size_t n, int limit, std::vector<unsigned long long> &sorted, int *too_many) {
    *too_many = 0;
    sorted.clear();
    if (nb < 1 || nb > 4) return FB_ERR_UNSUPPORTED;
    if (limit < 0) limit = 0;
    size_t cap = 4096;
    while (cap < (size_t)limit * 4 + 4096) cap <<= 1;
    unsigned long long *table = nullptr;
    int *ctr = nullptr;
    FB_CUDA(ctx, cudaMallocAsync((void **)&table, cap * sizeof(unsigned long long), ctx->stream));
    FB_CUDA(ctx, cudaMallocAsync((void **)&ctr, 4 * sizeof(int), ctx->stream));
    FB_CUDA(ctx, cudaMemsetAsync(table, 0xFF, cap * sizeof(unsigned long long), ctx->stream));
    FB_CUDA(ctx, cudaMemsetAsync(ctr, 0, 4 * sizeof(int), ctx->stream));
    if (n) {
        pl::Planes P;
        for (int c = 0; c < pl::kMaxPlanes; c++) P.p[c] = c < nb ? planes[c] : nullptr;
        pl::Collect C;
        C.table = table; C.cap_mask = (unsigned)(cap - 1); C.limit = limit; C.count = ctr; C.has_allones = ctr + 1; C.overflow = ctr + 2;
        pl::k_palette_collect<<<nblocks(n, 256), 256, 0, ctx->stream>>>(P, n, nb, C);
        ctx->launches++;
        ctx->mark("k_palette_collect", 2.0 * (double)n * nb);
        FB_CUDA(ctx, cudaGetLastError());
    }
    int h[4];
    FB_CUDA(ctx, cudaMemcpyAsync(h, ctr, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h[2] || h[0] > limit) *too_many = 1;
    else {
        std::vector<unsigned long long> t(cap);
        FB_CUDA(ctx, cudaMemcpyAsync(t.data(), table, cap * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (unsigned long long k : t) if (k != pl::kEmptySlot) sorted.push_back(k);
        if (h[1]) sorted.push_back(pl::kEmptySlot);
        std::sort(sorted.begin(), sorted.end());
    }
    cudaFreeAsync(table, ctx->stream);
    cudaFreeAsync(ctr, ctx->stream);
    return FB_OK;
}
// fwd_palette, step 2: plane 0 <- position of every sample's colour in `sorted_dev` (count keys in HBM)
int fb_launch_palette_index(fb_ctx *ctx, int16_t *const *planes, int nb, size_t n, const unsigned long long *sorted_dev, int count) {
    if (!n) return FB_OK;
    pl::Planes P;
    for (int c = 0; c < pl::kMaxPlanes; c++) P.p[c] = c < nb ? planes[c] : nullptr;
    pl::k_palette_index<<<nblocks(n, 256), 256, 0, ctx->stream>>>(P, n, nb, sorted_dev, count);
    ctx->launches++;
    ctx->mark("k_palette_index", 2.0 * (double)n * (nb + 1));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ctx->err = std::string("kernel launch: ") + cudaGetErrorString(e); return FB_ERR_CUDA; }
    return FB_OK;
}
// inv_match (2dmatch.h:97-177, exact matches): resolves every sample's root; parent_out is a device array of n ints the
// caller frees with cudaFreeAsync.  *bad = 1: a match code outside [0, maxcode].
int fb_match_resolve(fb_ctx *ctx, const int16_t *m, int n, int w, int maxcode, int **parent_out, int *bad) {
    *parent_out = nullptr; *bad = 0;
    if (n <= 0) return FB_OK;
    int *a = nullptr, *b = nullptr, *flag = nullptr;
    FB_CUDA(ctx, cudaMallocAsync((void **)&a, (size_t)n * sizeof(int), ctx->stream));
    FB_CUDA(ctx, cudaMallocAsync((void **)&b, (size_t)n * sizeof(int), ctx->stream));
    FB_CUDA(ctx, cudaMallocAsync((void **)&flag, 2 * sizeof(int), ctx->stream));       // [0] bad code, [1] something moved in this batch
    FB_CUDA(ctx, cudaMemsetAsync(flag, 0, 2 * sizeof(int), ctx->stream));
    mt::k_match_parent<<<nblocks((size_t)n, 256), 256, 0, ctx->stream>>>(m, a, n, w, maxcode, flag);
    ctx->launches++;
    int rounds = 1;
    while ((1ll << rounds) < n) rounds++;           // a chain is at most n samples long
    // chains in real files are a few samples long: the rounds run in batches of three, and a batch in which nothing moved ends them
    int done = 0, hflag[2] = {0, 0};
    while (done < rounds) {
        FB_CUDA(ctx, cudaMemsetAsync(flag + 1, 0, sizeof(int), ctx->stream));
        for (int r = 0; r < 3 && done < rounds; r++, done++) {
            mt::k_match_jump<<<nblocks((size_t)n, 256), 256, 0, ctx->stream>>>(a, b, n, flag + 1);
            ctx->launches++;
            std::swap(a, b);
        }
        FB_CUDA(ctx, cudaMemcpyAsync(hflag, flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (!hflag[1]) break;
    }
    ctx->mark("k_match_parent+jump", (double)n * (2.0 + 8.0 * done));
    FB_CUDA(ctx, cudaGetLastError());
    FB_CUDA(ctx, cudaMemcpyAsync(bad, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(b, ctx->stream);
    cudaFreeAsync(flag, ctx->stream);
    *parent_out = a;
    return FB_OK;
}
// inv_match with soft matches on ONE channel: out = reconstructed plane (caller-allocated, n samples)
int fb_match_soft(fb_ctx *ctx, const int16_t *m, const int16_t *orig, int16_t *out, int n, int w, int maxcode, int zero, int *bad) {
    *bad = 0;
    if (n <= 0) return FB_OK;
    int *pa = nullptr, *pb = nullptr, *flag = nullptr;
    int16_t *aa = nullptr, *ab = nullptr;
    FB_CUDA(ctx, cudaMallocAsync((void **)&pa, (size_t)n * sizeof(int), ctx->stream));
    FB_CUDA(ctx, cudaMallocAsync((void **)&pb, (size_t)n * sizeof(int), ctx->stream));
    FB_CUDA(ctx, cudaMallocAsync((void **)&aa, (size_t)n * sizeof(int16_t), ctx->stream));
    FB_CUDA(ctx, cudaMallocAsync((void **)&ab, (size_t)n * sizeof(int16_t), ctx->stream));
    FB_CUDA(ctx, cudaMallocAsync((void **)&flag, 2 * sizeof(int), ctx->stream));
    FB_CUDA(ctx, cudaMemsetAsync(flag, 0, 2 * sizeof(int), ctx->stream));
    mt::k_match_soft_init<<<nblocks((size_t)n, 256), 256, 0, ctx->stream>>>(m, orig, pa, aa, n, w, maxcode, zero, flag);
    ctx->launches++;
    int rounds = 1;
    while ((1ll << rounds) < n) rounds++;
    int done = 0, hflag[2] = {0, 0};
    while (done < rounds) {
        FB_CUDA(ctx, cudaMemsetAsync(flag + 1, 0, sizeof(int), ctx->stream));
        for (int r = 0; r < 3 && done < rounds; r++, done++) {
            mt::k_match_soft_jump<<<nblocks((size_t)n, 256), 256, 0, ctx->stream>>>(pa, aa, pb, ab, n, flag + 1);
            ctx->launches++;
            std::swap(pa, pb);
            std::swap(aa, ab);
        }
        FB_CUDA(ctx, cudaMemcpyAsync(hflag, flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (!hflag[1]) break;
    }
    mt::k_match_soft_finish<<<nblocks((size_t)n, 256), 256, 0, ctx->stream>>>(pa, aa, out, n);
    ctx->launches++;
    ctx->mark("k_match_soft", (double)n * (8.0 + 12.0 * done));
    FB_CUDA(ctx, cudaGetLastError());
    *bad = hflag[0];
    cudaFreeAsync(pa, ctx->stream); cudaFreeAsync(pb, ctx->stream); cudaFreeAsync(aa, ctx->stream); cudaFreeAsync(ab, ctx->stream); cudaFreeAsync(flag, ctx->stream);
    return FB_OK;
}
int fb_launch_match_gather(fb_ctx *ctx, const int16_t *src, int16_t *dst, const int *parent, int n, int zero) {
    if (n <= 0) return FB_OK;
    mt::k_match_gather<<<nblocks((size_t)n, 256), 256, 0, ctx->stream>>>(src, dst, parent, n, zero);
    ctx->launches++;
    ctx->mark("k_match_gather", 8.0 * n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ctx->err = std::string("kernel launch: ") + cudaGetErrorString(e); return FB_ERR_CUDA; }
    return FB_OK;
}
int fb_launch_approximate(fb_ctx *ctx, int16_t *ch, int16_t *chr, size_t n, int q, int inverse) {
    if (!n) return FB_OK;
    if (inverse) ap::k_approx_inv<<<nblocks(n, 256), 256, 0, ctx->stream>>>(ch, chr, n, q);
    else ap::k_approx_fwd<<<nblocks(n, 256), 256, 0, ctx->stream>>>(ch, chr, n, q);
    ctx->launches++;
    ctx->mark(inverse ? "k_approx_inv" : "k_approx_fwd", 2.0 * (double)n * (chr ? 3 : 2));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ctx->err = std::string("kernel launch: ") + cudaGetErrorString(e); return FB_ERR_CUDA; }
    return FB_OK;
}
int fb_launch_inv_subsample(fb_ctx *ctx, const int16_t *in, int16_t *out, int ow, int oh, int srh, int srv) {
    const size_t n = (size_t)ow * srh * (size_t)oh * srv;
    if (!n) return FB_OK;
    sb::k_inv_subsample<<<nblocks(n, 256), 256, 0, ctx->stream>>>(in, out, ow, oh, srh, srv);
    ctx->launches++;
    ctx->mark("k_inv_subsample", 2.0 * ((double)ow * oh + (double)n));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ctx->err = std::string("kernel launch: ") + cudaGetErrorString(e); return FB_ERR_CUDA; }
    return FB_OK;
}
int fb_launch_interleave(fb_ctx *ctx, const int16_t *const *planes, int nch, size_t npix, int bps, void *dst) {
    if (!npix) return FB_OK;
    if (nch > 8) return FB_ERR_INVALID;
    PlanesN pl;
    for (int i = 0; i < 8; i++) pl.p[i] = i < nch ? planes[i] : nullptr;
    k_interleave<<<nblocks(npix, 256), 256, 0, ctx->stream>>>(pl, nch, npix, bps, (uint8_t *)dst);
    FB_LAUNCH_CHECK(ctx);
    return FB_OK;
}
