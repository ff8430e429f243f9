// Packed (int16x2) inverse Squeeze steps for sm_100a: TMA-fed horizontal step with the colour epilogue, coalesced
// vertical step.  These are the launches that move almost all the bytes of an inverse transform chain.
//
// Reference semantics: transform/squeeze.h:61-77 (smooth_tendency), :81-132 (inv_hsqueeze), :173-224 (inv_vsqueeze),
// transform/ycocg.h:51-56 (inv_YCoCg), image/image.cpp:107-113 (final clamp).
//
// Why packed: the unsqueeze recurrence is integer-issue-bound on B200, not HBM-bound (measured, tools/ubench_pipes.cu:
// LOP3 / SHF / PRMT / VIMNMX* issue every 2 cycles per SM sub-partition on the ALU pipe, VIADD.16x2 / IMAD every 2 cycles
// on the FMA pipe, a balanced mix ~1.45 cycles per instruction).  sm_100a has native 16x2 add (VIADD.16x2), min / max /
// min3 / add-min-relu (VIMNMX(3).S16x2, VIADDMNMX.S16x2.RELU): one instruction handles two chains.  A pair step costs ~38
// instructions for two chains where the 32-bit formulation needs ~28 per chain.
//
// Exactness: the 16x2 arithmetic is exact (no wrap inside smooth_tendency) while every average is within +-kMaxAvg and
// every residual within +-kMaxRes (8- and 10-bit images); the reference's own int16 wraps of diff / A / B are what VIADD.16x2
// does anyway.  Every loaded word is range-checked on the fly (one VIADDMNMX.U16x2 per word); a segment that saw a value
// outside the range is flagged and recomputed by the 32-bit routine (h_repair / v_repair), which is exact for all int16.
//
// Parallelism: a chain (row / column) is cut into segments that start kWarm pairs early from a guessed state (the
// recurrence forgets its start within <= 7 pairs in practice).  Every segment records the state it assumed at
// its first owned pair (est) and its final state (act).  Work items are independent warps; the warp that finishes LAST for
// a group of chains (atomic counter) walks the segments of those chains in order, and recomputes any segment whose est
// differs from the true state before it (or that was range-flagged) with the exact routine.  By induction over the segments
// the output is bit-exact whatever the guesses were.
//
//   horizontal (k_pk_hsq): warp = 64 rows x one segment; lane = two rows (l, l+32) packed in the two halves of a word;
//       planes Co and Cg side by side in one lane (ILP 2).  Input tiles (16 pairs x 64 rows of averages, residuals, Y)
//       arrive through TMA (cp.async.bulk.tensor.2d + mbarrier, 3-stage ring per warp), output tiles leave through TMA
//       stores: no address arithmetic, no uncoalesced LSU traffic.  Optional epilogue: inverse YCoCg + final clamp, so the
//       last launch of a 4096^2 RGB image reads 100 MB and writes 100 MB once.
//   vertical (k_pk_vsq): lane = 8 adjacent columns (4 packed words, ILP 4), 16-byte coalesced loads / stores straight
//       from / to HBM with a register prefetch pipeline.
//
// Compiled by nvcc (product) and by g++ -DFB_EMULATE (tests/emu: CPU execution-model emulator vs the oracle; TMA and
// mbarrier are replaced by synchronous copies there).
#pragma once
#include "fb_fused_squeeze.cuh"     // fq::unsqueeze_pair (exact 32-bit pair), FB_* macros

#if !defined(FB_EMULATE)
#include <cuda.h>
#endif

namespace ps {

constexpr int kMaxAvg = 2047;       // packed arithmetic is exact while |average| <= kMaxAvg and |residual| <= kMaxRes:
constexpr int kMaxRes = 4095;       // |B| <= 3*kMaxAvg + kMaxRes/2 + 1 = 8189, 4|P-a| + 3|a-n| + 6 <= 53232 < 65536
#ifndef PS_HCHUNK
#define PS_HCHUNK 16
#endif
constexpr int kHChunk = PS_HCHUNK;  // pairs per TMA tile (8: 16-byte average / residual rows, 32-byte output rows; 16: twice that)
constexpr int kHW = kHChunk / 2;    // 32-bit words of averages per tile row
constexpr int kHRows = 64;          // rows per warp tile
constexpr int kHWarm = 16;          // warm-up pairs of a horizontal segment (whole tiles, nothing stored)
constexpr int kHWarmChunks = kHWarm / kHChunk;
static_assert(kHChunk == 8 || kHChunk == 16, "tile widths the register staging is written for");
// Shared-memory swizzle of the tiles (TMA swizzle modes): a lane owns ROWS, so with dense tiles the 32 lanes of a 16-byte access
// would sit 32 / 64 bytes apart and hit the same few banks (measured: 16-way conflicts on the output stores, ~800 conflict cycles
// per tile).  With the 16-byte chunk index XORed with the row bits (address bits 7.. of 1024-byte aligned tiles) they spread out.
constexpr int kSwzA = kHChunk * 2 >= 32 ? kHChunk * 2 : 0;          // average / residual tiles: 32-byte rows -> SWIZZLE_32B
constexpr int kSwzO = 2 * kHChunk * 2 >= 32 ? 2 * kHChunk * 2 : 0;  // Y / output tiles: 64-byte rows -> SWIZZLE_64B
static_assert(kSwzA <= 64 && kSwzO <= 64, "swizzle spans");
constexpr int kVWarm = 12;          // warm-up pairs of a vertical segment (8 missed about once per 4096^2 image: a repair costs tens of microseconds)
constexpr int kVDepthDefault = 6;   // rows of register prefetch in the vertical kernel (template parameter of k_pk_vsq)
constexpr int kMaxHJobs = 8;
constexpr int kMaxVJobs = 24;

// ---------------------------------------------------------------------------------------------------------
// 16x2 primitives
// ---------------------------------------------------------------------------------------------------------
#if defined(FB_EMULATE)
inline uint32_t e_h(int lo, int hi) { return (uint32_t)(uint16_t)lo | ((uint32_t)(uint16_t)hi << 16); }
inline int e_lo(uint32_t w) { return (int)(short)(w & 0xffffu); }
inline int e_hi(uint32_t w) { return (int)(short)(w >> 16); }
inline uint32_t __vadd2(uint32_t a, uint32_t b) { return e_h(e_lo(a) + e_lo(b), e_hi(a) + e_hi(b)); }
inline uint32_t __vmins2(uint32_t a, uint32_t b) { return e_h(std::min(e_lo(a), e_lo(b)), std::min(e_hi(a), e_hi(b))); }
inline uint32_t __vmaxs2(uint32_t a, uint32_t b) { return e_h(std::max(e_lo(a), e_lo(b)), std::max(e_hi(a), e_hi(b))); }
inline uint32_t __viaddmin_s16x2_relu(uint32_t a, uint32_t b, uint32_t c) {
    const int l = std::max(std::min((int)(short)(e_lo(a) + e_lo(b)), e_lo(c)), 0), h = std::max(std::min((int)(short)(e_hi(a) + e_hi(b)), e_hi(c)), 0);
    return e_h(l, h);
}
inline uint32_t __vimin3_s16x2_relu(uint32_t a, uint32_t b, uint32_t c) {
    return e_h(std::max(std::min(std::min(e_lo(a), e_lo(b)), e_lo(c)), 0), std::max(std::min(std::min(e_hi(a), e_hi(b)), e_hi(c)), 0));
}
inline uint32_t __viaddmax_u16x2(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t l = std::max((uint32_t)(uint16_t)((a & 0xffff) + (b & 0xffff)), c & 0xffff), h = std::max((uint32_t)(uint16_t)((a >> 16) + (b >> 16)), c >> 16);
    return l | (h << 16);
}
inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t sel = (s >> (4 * i)) & 0xf;
        const uint32_t b = (uint32_t)(v >> (8 * (sel & 7))) & 0xff;      // the CUDA intrinsic ignores bit 3 of a selector nibble
        r |= b << (8 * i);
    }
    return r;
}
#endif

FB_DEV uint32_t padd(uint32_t a, uint32_t b) { return __vadd2(a, b); }                  // VIADD.16x2 (wraps per half)
// Packed constants come in through a register the compiler cannot see through (PK::k1 = 0x00010001 read from the job
// descriptor): ptxas 12.9 was caught splitting `add.u16x2 x, 0x00010001` into per-half pieces and dropping the +1 of the low
// half in one unrolled instance (VIADD.16x2 R, R, 0x0 + PRMT 0x7610; wrong low halves on the hardware, tests/test_sass_checks.py
// looks for the pattern).  With a register operand there is no immediate to split.
struct PK { uint32_t k1, k2, k6; };
FB_DEV PK pk_consts(uint32_t one) { PK k; k.k1 = one; k.k2 = __vadd2(one, one); k.k6 = __vadd2(__vadd2(k.k2, k.k2), k.k2); return k; }
FB_DEV uint32_t pneg(uint32_t a, const PK &K) { return __vadd2(~a, K.k1); }
// 0xffff where the half is negative: PRMT with the sign-replicate bit of the selector nibbles, which only PTX exposes
// (__byte_perm masks the selector to 3 bits per nibble)
#if defined(FB_EMULATE)
inline uint32_t psign(uint32_t a) { return ((a & 0x8000u) ? 0xffffu : 0u) | ((a & 0x80000000u) ? 0xffff0000u : 0u); }
#else
FB_DEV uint32_t psign(uint32_t a) {
    uint32_t r;
    asm("prmt.b32 %0, %1, 0, 0xbb99;" : "=r"(r) : "r"(a));
    return r;
}
#endif
FB_DEV uint32_t pasr1(uint32_t a) { return ((a >> 1) & 0x7fff7fffu) | (a & 0x80008000u); }     // arithmetic >> 1 per half
FB_DEV uint32_t plo(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x5410); }       // (a.lo, b.lo)
FB_DEV uint32_t phi(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x7632); }       // (a.hi, b.hi)

// floor(t / 12) per unsigned half: (t * 43691) >> 19 is exact for t < 2^16 (43691 * 12 = 2^19 + 4)
FB_DEV uint32_t pdiv12(uint32_t t) {
    const uint32_t ql = ((t & 0xffffu) * 43691u) >> 19, qh = ((t >> 16) * 43691u) >> 19;
    return ql | (qh << 16);
}

// One unsqueeze pair on two chains at once (squeeze.h:97-108 with smooth_tendency :61-77 in the closed form of
// fq::unsqueeze_pair).  P = previous reconstructed B, a = average, nega = -a, negn = -(next average), r = residual.
// With sigma = sign(P - n): a monotone triple has sigma*(P-a) >= 0 and sigma*(a-n) >= 0 and
//   tendency = sigma * min((4 s(P-a) + 3 s(a-n) + 6) / 12, 2 s(P-a) + 1, 2 s(a-n));
// otherwise one of 2 s(P-a) + 1, 2 s(a-n) is negative and the RELU of the min gives the reference's 0.
FB_DEV void pk_step(const PK &K, uint32_t P, uint32_t a, uint32_t nega, uint32_t negn, uint32_t r, uint32_t &A, uint32_t &B) {
    const uint32_t t1 = padd(P, nega), t2 = padd(a, negn), u = padd(P, negn);
    const uint32_t s = psign(u), s1 = s & K.k1;
    const uint32_t a1 = padd(t1 ^ s, s1), a2 = padd(t2 ^ s, s1);
    const uint32_t m1 = padd(a1, a1), m2 = padd(a2, a2);
    const uint32_t t = padd(padd(padd(m1, m1), padd(m2, a2)), K.k6);
    const uint32_t q = pdiv12(t);
    const uint32_t d = __viaddmin_s16x2_relu(m1, K.k1, __vmins2(q, m2));
    const uint32_t diff = padd(r, padd(d ^ s, s1));
    const uint32_t dd = padd(diff, (diff >> 15) & K.k1);             // + 1 where negative: >> 1 then truncates toward zero
    A = padd(a, pasr1(dd));
    B = padd(padd(A, ~diff), K.k1);
}

// inverse YCoCg on two pixels at once (ycocg.h:51-56): G = clamp(Y + ((Cg+1)>>1)), B = clamp(Y - (Cg>>1) - (Co>>1)), R = clamp(Co + B),
// all clamps to [0, maxval]; exact while |Co|, |Cg| <= 8189 (outputs of a range-checked step) and maxval >= 0
FB_DEV void pk_ycocg(const PK &K, uint32_t y, uint32_t co, uint32_t cg, uint32_t mv, uint32_t &R, uint32_t &G, uint32_t &B) {
    const uint32_t yc = __vimin3_s16x2_relu(y, mv, mv);
    G = __viaddmin_s16x2_relu(yc, pasr1(padd(cg, K.k1)), mv);
    const uint32_t nsum = padd(padd(pasr1(~cg), pasr1(~co)), K.k2);       // -(cg>>1) - (co>>1): ~(x>>1) = -(x>>1) - 1
    B = __viaddmin_s16x2_relu(yc, nsum, mv);
    R = __viaddmin_s16x2_relu(co, B, mv);
}
FB_DEV uint32_t pk_clamp(uint32_t x, uint32_t lo, uint32_t hi) { return __vmins2(__vmaxs2(x, lo), hi); }

// range accumulators: every half of `acc` stays <= 2*bound while all checked values are within +-bound
FB_DEV uint32_t pk_chk(uint32_t acc, uint32_t w, uint32_t bound2) { return __viaddmax_u16x2(w, bound2, acc); }
FB_DEV bool pk_chk_bad(uint32_t acc, int bound) { return (int)(acc & 0xffffu) > 2 * bound || (int)(acc >> 16) > 2 * bound; }

// ---------------------------------------------------------------------------------------------------------
// tile movement: TMA + mbarrier on the GPU, synchronous copies under the emulator
// ---------------------------------------------------------------------------------------------------------
#if defined(FB_EMULATE)
struct TileMap { int16_t *base; int w, h, box_w, box_h, swz; };      // swz: 0, 32 or 64 = the TMA shared-memory swizzle span in bytes
// 16-byte chunk index bits [4, 4+B) of a shared-memory offset XORed with bits [7, 7+B): CU_TENSOR_MAP_SWIZZLE_32B (B = 1) / _64B (B = 2)
inline int emu_swz(int off, int swz) { const int m = swz == 64 ? 3 : (swz == 32 ? 1 : 0); return off ^ (((off >> 7) & m) << 4); }
inline void fb_syncwarp() { cuemu::bar_sync(1 + (int)(threadIdx.x >> 5), 32); }
inline void fb_threadfence() {}
inline void tile_load(void *dst, const TileMap *m, int x, int y, uint64_t *, int) {
    int16_t *d = (int16_t *)dst;
    for (int r = 0; r < m->box_h; r++)
        for (int c = 0; c < m->box_w; c++) {
            const int gx = x + c, gy = y + r;
            d[emu_swz((r * m->box_w + c) * 2, m->swz) / 2] = (gx >= 0 && gx < m->w && gy >= 0 && gy < m->h) ? m->base[(size_t)gy * m->w + gx] : (int16_t)0;
        }
}
inline void tile_store(const TileMap *m, int x, int y, const void *src) {
    const int16_t *s = (const int16_t *)src;
    for (int r = 0; r < m->box_h; r++)
        for (int c = 0; c < m->box_w; c++) {
            const int gx = x + c, gy = y + r;
            if (gx >= 0 && gx < m->w && gy >= 0 && gy < m->h) m->base[(size_t)gy * m->w + gx] = s[emu_swz((r * m->box_w + c) * 2, m->swz) / 2];
        }
}
// the copies above are synchronous, so "the tile has landed" only needs lane 0 to have reached its issue point: a warp
// rendezvous stands in for the mbarrier wait (emu_issue_point is empty in the product)
inline void emu_issue_point() { fb_syncwarp(); }
inline void mbar_init(uint64_t *, int) {}
inline void mbar_expect(uint64_t *, int) {}
inline void mbar_wait(uint64_t *, int) {}
inline void fence_async_smem() {}
inline void store_commit() {}
inline void store_wait_read() {}
inline void store_wait_all() {}
inline void fence_mbar_init() {}
#else
typedef CUtensorMap TileMap;
FB_DEV void emu_issue_point() {}
FB_DEV void fb_syncwarp() { __syncwarp(); }
FB_DEV void fb_threadfence() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }     // release before / acquire after the arrival counter
FB_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
FB_DEV void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
FB_DEV void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
FB_DEV void mbar_expect(uint64_t *bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
FB_DEV void mbar_wait(uint64_t *bar, int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one 2-D tile global -> shared; out-of-range elements arrive as zero; completion is counted in bytes on `bar`
FB_DEV void tile_load(void *dst, const TileMap *m, int x, int y, uint64_t *bar, int) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(m), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
// one 2-D tile shared -> global (clipped at the plane border), bulk-group completion
FB_DEV void tile_store(const TileMap *m, int x, int y, const void *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(m), "r"(x), "r"(y), "r"(smem_u32(src)) : "memory");
}
FB_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
FB_DEV void store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
FB_DEV void store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
FB_DEV void store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
#endif

// value of lane 0 for the whole warp
#if defined(FB_EMULATE)
inline int warp_bcast0(int v, int lane) {
    static int slot[64];
    const int w = (int)(threadIdx.x >> 5);
    fb_syncwarp();
    if (lane == 0) slot[w] = v;
    fb_syncwarp();
    return slot[w];
}
#else
FB_DEV int warp_bcast0(int v, int) { return __shfl_sync(0xffffffffu, v, 0); }
#endif

// ---------------------------------------------------------------------------------------------------------
// horizontal step
// ---------------------------------------------------------------------------------------------------------
struct HJob {                   // one horizontal step on np planes of identical geometry
    TileMap tm_a[2], tm_r[2];   // averages / residuals: wa x h, box kHChunk x kHRows
    TileMap tm_o[3];            // outputs 2*wa x h, box 2*kHChunk x kHRows: planes [0..np), or R, G, B with the YCoCg epilogue
    TileMap tm_y;               // epilogue: the final Y plane
    const int16_t *avg[2], *res[2];     // the same planes for the exact repair routine
    int16_t *out[3];
    const int16_t *yin;
    int16_t *est[2], *act[2];   // [h][nsegp] per plane: act[row][g] = final state of segment g, est[row][g-1] = state segment g assumed at its
                                // first owned pair (so a row joins everywhere iff est[row][j] == act[row][j] for j < nseg-1: vector compares)
    int16_t *bad;               // [h][nsegp]: segment saw a value outside the packed range (it is recomputed by the exact routine)
    int *counter;               // [nrb] arrivals per row block (left at zero by the last arriver)
    int *stats;                 // [0] repaired segments, [1] range-flagged segments (diagnostics)
    int np, wa, h;
    int S, nseg, nrb;           // pairs per segment (multiple of kHChunk), segments per row, row blocks of kHRows
    int nsegp;                  // nseg rounded up to a multiple of 8 (16-byte rows of est / act / bad)
    int item0;                  // first work item of this job
    int epilogue;               // fq::kEpNone / kEpClamp / kEpYCoCg
    int maxval, lo, hi, do_clamp;
    uint32_t k1;                // 0x00010001 (see PK)
};
struct HJobs { HJob j[kMaxHJobs]; int n, items; };

FB_HD size_t h_smem_per_warp(int np, int epilogue) {
    const size_t stage = (size_t)np * 2 * kHRows * kHChunk * 2 + (epilogue == fq::kEpYCoCg ? (size_t)kHRows * 2 * kHChunk * 2 : 0);
    return 2 * stage + 1024;        // two slots (inputs, then the outputs in place) + mbarriers; slices stay 1024-byte aligned
}

// exact inverse YCoCg + final clamp of one pixel
FB_DEV void ycocg_exact(int Yr, int Co, int Cg, int maxval, int lo, int hi, int do_clamp, int &R, int &G, int &B) {
    const int Y = fq::clampi(Yr, 0, maxval);
    G = fq::clampi(Y - ((-Cg) >> 1), 0, maxval);
    B = fq::clampi(Y + ((1 - Cg) >> 1) - (Co >> 1), 0, maxval);
    R = fq::clampi(Co + B, 0, maxval);
    if (do_clamp) { R = fq::clampi(R, lo, hi); G = fq::clampi(G, lo, hi); B = fq::clampi(B, lo, hi); }
}

// Exact recomputation of segment g of one row from the true state before it (prev[]); plain global accesses.
FB_DEV void h_repair(const HJob &J, int row, int g, int *prev) {
    const int x0 = g * J.S, x1 = fq::imin(x0 + J.S, J.wa), wo = 2 * J.wa;
    for (int x = x0; x < x1; x++) {
        int A[2], B[2];
        for (int p = 0; p < J.np; p++) {
            const int16_t *ar = J.avg[p] + (size_t)row * J.wa;
            const int av = ar[x], nx = x + 1 < J.wa ? ar[x + 1] : av, rs = J.res[p][(size_t)row * J.wa + x];
            fq::unsqueeze_pair(x == 0 ? av : prev[p], av, nx, rs, A[p], B[p]);
            prev[p] = B[p];
        }
        const size_t o = (size_t)row * wo + 2 * x;
        if (J.epilogue == fq::kEpYCoCg) {
            for (int k = 0; k < 2; k++) {
                int R, G, Bl;
                ycocg_exact(J.yin[o + k], k ? B[0] : A[0], k ? B[1] : A[1], J.maxval, J.lo, J.hi, J.do_clamp, R, G, Bl);
                J.out[0][o + k] = (int16_t)R; J.out[1][o + k] = (int16_t)G; J.out[2][o + k] = (int16_t)Bl;
            }
        } else {
            for (int p = 0; p < J.np; p++) {
                if (J.epilogue == fq::kEpClamp && J.do_clamp) { A[p] = fq::clampi(A[p], J.lo, J.hi); B[p] = fq::clampi(B[p], J.lo, J.hi); }
                J.out[p][o] = (int16_t)A[p]; J.out[p][o + 1] = (int16_t)B[p];
            }
        }
    }
}

// halves [0, valid) of the four words starting at halfword 8*i of a row are meaningful
FB_DEV uint32_t tail_mask(int valid, int word) {
    const int v = valid - 2 * word;
    return v >= 2 ? 0xffffffffu : (v == 1 ? 0x0000ffffu : 0u);
}
// the last arriver of a row block: does every segment of its rows join its predecessor?  First with a handful of independent
// 16-byte loads per row (the common case: one memory round trip); only a row that does not is walked in order and repaired.
FB_DEV void h_verify(const HJob &J, int rb, int lane) {
    for (int half = 0; half < 2; half++) {
        const int row = rb * kHRows + lane + 32 * half;
        if (row >= J.h) continue;
        const size_t ro = (size_t)row * J.nsegp;
        uint32_t dirty = 0;
        for (int i = 0; i < J.nsegp / 8; i++) {
            const uint4 b = *reinterpret_cast<const uint4 *>(J.bad + ro + 8 * i);
            const int vb = J.nseg - 8 * i, ve = J.nseg - 1 - 8 * i;
            dirty |= (b.x & tail_mask(vb, 0)) | (b.y & tail_mask(vb, 1)) | (b.z & tail_mask(vb, 2)) | (b.w & tail_mask(vb, 3));
            for (int p = 0; p < J.np; p++) {
                const uint4 e = *reinterpret_cast<const uint4 *>(J.est[p] + ro + 8 * i), a = *reinterpret_cast<const uint4 *>(J.act[p] + ro + 8 * i);
                dirty |= ((e.x ^ a.x) & tail_mask(ve, 0)) | ((e.y ^ a.y) & tail_mask(ve, 1)) | ((e.z ^ a.z) & tail_mask(ve, 2)) | ((e.w ^ a.w) & tail_mask(ve, 3));
            }
        }
        if (!dirty) continue;
        int cur[2] = {0, 0};
        for (int g = 0; g < J.nseg; g++) {
            bool need = J.bad[ro + g] != 0;
            if (g > 0)
                for (int p = 0; p < J.np; p++) need = need || (J.est[p][ro + g - 1] != (int16_t)cur[p]);
            if (need) {
                h_repair(J, row, g, cur);
                for (int p = 0; p < J.np; p++) J.act[p][ro + g] = (int16_t)cur[p];
                atomicAdd(J.stats, 1);
            } else {
                for (int p = 0; p < J.np; p++) cur[p] = J.act[p][ro + g];
            }
        }
    }
}

// One work item: rows [64 rb, 64 rb + 64) x segment g of job J, by one warp.  sm = this warp's shared memory: two slots of
// kStage bytes + two mbarriers.  A slot receives the input tiles of a chunk (16 pairs x 64 rows of every plane); the lanes pull
// their two rows into registers, and the slot then becomes the staging area of the chunk's OUTPUT tiles (same size: a pair of
// averages + residuals makes a pair of samples), which a TMA store writes out while the other slot is being worked on.
template <int NP, int EP>
FB_DEV void h_item(const HJob &J, int rb, int g, unsigned char *sm, int lane) {
    constexpr bool kCol = EP == fq::kEpYCoCg;
    constexpr int kTileA = kHRows * kHChunk * 2;                // bytes of an average / residual tile (32-byte rows)
    constexpr int kTileO = kHRows * 2 * kHChunk * 2;            // bytes of an output / Y tile (64-byte rows)
    constexpr int kStage = NP * 2 * kTileA + (kCol ? kTileO : 0);
    constexpr int kNOut = kCol ? 3 : NP;
    static_assert(kNOut * kTileO == kStage, "output tiles reuse the input slot");
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + 2 * kStage);
    const int wa = J.wa, row0 = rb * kHRows;
    const int x_own = g * J.S, x_end = fq::imin(x_own + J.S, wa);
    const int x_first = g ? x_own - kHWarm : 0;
    const int nchunks = (x_end - x_first + kHChunk - 1) / kHChunk;
    const int rowA = lane * (kHChunk * 2), rowB = (lane + 32) * (kHChunk * 2);         // byte offsets of this lane's rows in a 32-byte-row tile
    // swizzle terms of this lane's rows (the same for row l and row l + 32): XOR them into the 16-byte chunk offset
    const int sA = kSwzA == 32 ? ((lane >> 2) & 1) << 4 : (kSwzA == 64 ? ((lane >> 1) & 3) << 4 : 0);
    const int sO = kSwzO == 64 ? ((lane >> 1) & 3) << 4 : (kSwzO == 32 ? ((lane >> 2) & 1) << 4 : 0);

    auto issue = [&](int c) {       // lane 0: the input tiles of chunk c into slot c & 1
        const int xc = x_first + c * kHChunk;
        unsigned char *base = sm + (c & 1) * kStage;
        const bool with_y = kCol && !(g && c < kHWarmChunks);
        mbar_expect(&bars[c & 1], NP * 2 * kTileA + (with_y ? kTileO : 0));
        for (int p = 0; p < NP; p++) {
            tile_load(base + p * 2 * kTileA, &J.tm_a[p], xc, row0, &bars[c & 1], 0);
            tile_load(base + p * 2 * kTileA + kTileA, &J.tm_r[p], xc, row0, &bars[c & 1], 0);
        }
        if (with_y) tile_load(base + NP * 2 * kTileA, &J.tm_y, 2 * xc, row0, &bars[c & 1], 0);
    };
    if (lane == 0) {
        issue(0);
        if (nchunks > 1) issue(1);
    }
    emu_issue_point();

    uint32_t P[NP], chk_a = 0, chk_r = 0;
    int ae_lo[NP], ae_hi[NP];           // the average right after the segment (the last owned pair's "next"): one direct load per row,
#pragma unroll                          // packed and range-checked only where it is used, so that nothing waits for these loads here
    for (int p = 0; p < NP; p++) {
        ae_lo[p] = 0; ae_hi[p] = 0;
        if (x_end < wa) {
            const int rA = fq::imin(row0 + lane, J.h - 1), rB = fq::imin(row0 + lane + 32, J.h - 1);
            ae_lo[p] = J.avg[p][(size_t)rA * wa + x_end];
            ae_hi[p] = J.avg[p][(size_t)rB * wa + x_end];
        }
    }
    const bool do_clamp = J.do_clamp != 0;
    const PK K = pk_consts(J.k1);
    const uint32_t mv = (uint32_t)(uint16_t)J.maxval * 0x00010001u, clo = (uint32_t)(uint16_t)J.lo * 0x00010001u, chi = (uint32_t)(uint16_t)J.hi * 0x00010001u;
    for (int c = 0; c < nchunks; c++) {
        const int xc = x_first + c * kHChunk;
        unsigned char *slot = sm + (c & 1) * kStage;
        const unsigned char *nslot = sm + ((c + 1) & 1) * kStage;
        const bool warm = g && c < kHWarmChunks, have_next = c + 1 < nchunks;
        const int nsteps = fq::imin(kHChunk, x_end - xc);       // a whole tile, or 8 at the end of a row whose width is 8 mod 16
        mbar_wait(&bars[c & 1], (c >> 1) & 1);
        // ---- the lane's two rows of every input tile into registers
        uint32_t aw[NP][2][kHW], rw[NP][2][kHW], yw[2][2 * kHW];
#pragma unroll
        for (int p = 0; p < NP; p++) {
            const unsigned char *ta = slot + p * 2 * kTileA, *tr = ta + kTileA;
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int ro = r ? rowB : rowA;
#pragma unroll
                for (int v = 0; v < kHW / 4; v++) {
                    const uint4 va = *reinterpret_cast<const uint4 *>(ta + ro + ((16 * v) ^ sA)), vr = *reinterpret_cast<const uint4 *>(tr + ro + ((16 * v) ^ sA));
                    aw[p][r][4 * v] = va.x; aw[p][r][4 * v + 1] = va.y; aw[p][r][4 * v + 2] = va.z; aw[p][r][4 * v + 3] = va.w;
                    rw[p][r][4 * v] = vr.x; rw[p][r][4 * v + 1] = vr.y; rw[p][r][4 * v + 2] = vr.z; rw[p][r][4 * v + 3] = vr.w;
                }
#pragma unroll
                for (int i = 0; i < kHW; i++) {
                    chk_a = pk_chk(chk_a, aw[p][r][i], kMaxAvg * 0x00010001u);
                    chk_r = pk_chk(chk_r, rw[p][r][i], kMaxRes * 0x00010001u);
                }
            }
        }
        if (kCol && !warm) {
            const unsigned char *ty = slot + NP * 2 * kTileA;
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int v = 0; v < kHW / 2; v++) {
                    const uint4 y = *reinterpret_cast<const uint4 *>(ty + 2 * (r ? rowB : rowA) + ((16 * v) ^ sO));
                    yw[r][4 * v] = y.x; yw[r][4 * v + 1] = y.y; yw[r][4 * v + 2] = y.z; yw[r][4 * v + 3] = y.w;
                }
        }
        if (c == 0) {
#pragma unroll
            for (int p = 0; p < NP; p++) P[p] = plo(aw[p][0][0], aw[p][1][0]);      // chain start: "left" = own average (squeeze.h:84-89); a segment start guesses the same
        }
        if (c == kHWarmChunks && g) {       // first owned pair: remember the state the warm-up reached
#pragma unroll
            for (int p = 0; p < NP; p++) {
                if (row0 + lane < J.h) J.est[p][(size_t)(row0 + lane) * J.nsegp + g - 1] = (int16_t)(P[p] & 0xffffu);
                if (row0 + lane + 32 < J.h) J.est[p][(size_t)(row0 + lane + 32) * J.nsegp + g - 1] = (int16_t)(P[p] >> 16);
            }
        }
        uint32_t a_last[NP];
        fb_syncwarp();              // every lane holds its inputs: the slot may now receive the outputs
#pragma unroll
        for (int k = 0; k < kHW; k++) {                         // two pairs per iteration: one word of averages per row
            if (k == kHW / 4 && c >= 1 && have_next) {
                // the other slot held chunk c-1: its output tiles have had a quarter of this chunk's computation to be read by the
                // store engine; now it takes chunk c+1, which has the remaining three quarters to arrive
                if (lane == 0) { store_wait_read(); issue(c + 1); }
                emu_issue_point();
            }
            if (k == kHW - 1) {
                // what the last pair of the chunk sees as its next average: the first average of the next chunk (its tiles have had
                // this chunk's computation to arrive in the other slot), or the average after the segment (direct load at the start)
                if (have_next) {
                    mbar_wait(&bars[(c + 1) & 1], ((c + 1) >> 1) & 1);
#pragma unroll
                    for (int p = 0; p < NP; p++)
                        a_last[p] = plo(*reinterpret_cast<const uint32_t *>(nslot + p * 2 * kTileA + rowA + sA), *reinterpret_cast<const uint32_t *>(nslot + p * 2 * kTileA + rowB + sA));
                } else {
#pragma unroll
                    for (int p = 0; p < NP; p++) {
                        a_last[p] = (uint32_t)(uint16_t)ae_lo[p] | ((uint32_t)(uint16_t)ae_hi[p] << 16);
                        chk_a = pk_chk(chk_a, a_last[p], kMaxAvg * 0x00010001u);
                    }
                }
            }
            if (2 * k < nsteps) {
                uint32_t oA[NP][2], oB[NP][2];                  // [plane][pair]: packed (row l, row l+32)
#pragma unroll
                for (int p = 0; p < NP; p++) {
                    const uint32_t a0 = plo(aw[p][0][k], aw[p][1][k]), a1 = phi(aw[p][0][k], aw[p][1][k]);
                    uint32_t a2;
                    if (k < kHW - 1) a2 = plo(aw[p][0][k + 1], aw[p][1][k + 1]);
                    else a2 = a_last[p];
                    if (xc + 2 * k + 2 >= wa) a2 = a1;          // last pair of the row
                    const uint32_t r0 = plo(rw[p][0][k], rw[p][1][k]), r1 = phi(rw[p][0][k], rw[p][1][k]);
                    const uint32_t n1 = pneg(a1, K);
                    pk_step(K, P[p], a0, pneg(a0, K), n1, r0, oA[p][0], oB[p][0]);
                    pk_step(K, oB[p][0], a1, n1, pneg(a2, K), r1, oA[p][1], oB[p][1]);
                    P[p] = oB[p][1];
                }
                if (!warm) {
                    const int ob = ((16 * (k >> 1)) ^ sO) + 8 * (k & 1);      // byte offset of output columns 4k .. 4k+3 in a (swizzled) output row
                    if (kCol) {
                        uint32_t R[4], G[4], Bc[4];
                        const uint32_t y4[4] = {plo(yw[0][2 * k], yw[1][2 * k]), phi(yw[0][2 * k], yw[1][2 * k]), plo(yw[0][2 * k + 1], yw[1][2 * k + 1]),
                                                phi(yw[0][2 * k + 1], yw[1][2 * k + 1])};
                        const uint32_t co4[4] = {oA[0][0], oB[0][0], oA[0][1], oB[0][1]}, cg4[4] = {oA[NP - 1][0], oB[NP - 1][0], oA[NP - 1][1], oB[NP - 1][1]};
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            pk_ycocg(K, y4[i], co4[i], cg4[i], mv, R[i], G[i], Bc[i]);
                            if (do_clamp) { R[i] = pk_clamp(R[i], clo, chi); G[i] = pk_clamp(G[i], clo, chi); Bc[i] = pk_clamp(Bc[i], clo, chi); }
                        }
                        uint2 v;
                        v.x = plo(R[0], R[1]); v.y = plo(R[2], R[3]); *reinterpret_cast<uint2 *>(slot + 2 * rowA + ob) = v;
                        v.x = phi(R[0], R[1]); v.y = phi(R[2], R[3]); *reinterpret_cast<uint2 *>(slot + 2 * rowB + ob) = v;
                        v.x = plo(G[0], G[1]); v.y = plo(G[2], G[3]); *reinterpret_cast<uint2 *>(slot + kTileO + 2 * rowA + ob) = v;
                        v.x = phi(G[0], G[1]); v.y = phi(G[2], G[3]); *reinterpret_cast<uint2 *>(slot + kTileO + 2 * rowB + ob) = v;
                        v.x = plo(Bc[0], Bc[1]); v.y = plo(Bc[2], Bc[3]); *reinterpret_cast<uint2 *>(slot + 2 * kTileO + 2 * rowA + ob) = v;
                        v.x = phi(Bc[0], Bc[1]); v.y = phi(Bc[2], Bc[3]); *reinterpret_cast<uint2 *>(slot + 2 * kTileO + 2 * rowB + ob) = v;
                    } else {
#pragma unroll
                        for (int p = 0; p < NP; p++) {
                            uint32_t o4[4] = {oA[p][0], oB[p][0], oA[p][1], oB[p][1]};
                            if (EP == fq::kEpClamp && do_clamp) {
#pragma unroll
                                for (int i = 0; i < 4; i++) o4[i] = pk_clamp(o4[i], clo, chi);
                            }
                            uint2 v;
                            v.x = plo(o4[0], o4[1]); v.y = plo(o4[2], o4[3]); *reinterpret_cast<uint2 *>(slot + p * kTileO + 2 * rowA + ob) = v;
                            v.x = phi(o4[0], o4[1]); v.y = phi(o4[2], o4[3]); *reinterpret_cast<uint2 *>(slot + p * kTileO + 2 * rowB + ob) = v;
                        }
                    }
                }
            }
        }
        if (!warm) {
            fence_async_smem();
            fb_syncwarp();
            if (lane == 0) {
                for (int p = 0; p < kNOut; p++) tile_store(&J.tm_o[p], 2 * xc, row0, slot + p * kTileO);
                store_commit();
            }
        } else {
            fb_syncwarp();          // the warm-up chunk stores nothing, but its slot is refilled by lane 0 next
        }
    }
    // final state + range flag of this segment
    const bool bad = pk_chk_bad(chk_a, kMaxAvg) || pk_chk_bad(chk_r, kMaxRes);
#pragma unroll
    for (int p = 0; p < NP; p++) {
        if (row0 + lane < J.h) J.act[p][(size_t)(row0 + lane) * J.nsegp + g] = (int16_t)(P[p] & 0xffffu);
        if (row0 + lane + 32 < J.h) J.act[p][(size_t)(row0 + lane + 32) * J.nsegp + g] = (int16_t)(P[p] >> 16);
    }
    if (row0 + lane < J.h) J.bad[(size_t)(row0 + lane) * J.nsegp + g] = bad ? 1 : 0;
    if (row0 + lane + 32 < J.h) J.bad[(size_t)(row0 + lane + 32) * J.nsegp + g] = bad ? 1 : 0;
    if (bad) atomicAdd(J.stats + 1, 1);
    if (lane == 0) store_wait_all();                            // this segment's tiles are in global memory
    fb_threadfence();
    fb_syncwarp();
    int last = 0;
    if (lane == 0) {
        const int seen = atomicAdd(J.counter + rb, 1);
        last = seen == J.nseg - 1;
        if (last) J.counter[rb] = 0;                            // ready for the next launch
    }
    last = warp_bcast0(last, lane);
    if (last) {
        fb_threadfence();
        h_verify(J, rb, lane);
    }
}

// All jobs of a launch have the same (NP, EP).  One warp per block (the blocks of an SM are independent pipelines that the
// block scheduler keeps refilling); its shared memory: two slots + the mbarriers.
template <int NP, int EP>
FB_KERNEL(32) k_pk_hsq(const FB_GRID_CONSTANT HJobs jobs, int warps_per_block, int smem_per_warp) {
    FB_DYN_SMEM(smraw);
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    unsigned char *sm = smraw + ((1024 - (int)(reinterpret_cast<uintptr_t>(smraw) & 1023)) & 1023) + (size_t)warp * smem_per_warp;     // swizzled tiles: 1024-byte aligned
    const int item = (int)blockIdx.x * warps_per_block + warp;
    if (item >= jobs.items) return;
    int ji = 0;
    while (ji < jobs.n - 1 && item >= jobs.j[ji + 1].item0) ji++;
    const HJob &J = jobs.j[ji];
    const int local = item - J.item0, rb = local / J.nseg, g = local - rb * J.nseg;
    constexpr int kStage = NP * 2 * kHRows * kHChunk * 2 + (EP == fq::kEpYCoCg ? kHRows * 2 * kHChunk * 2 : 0);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + 2 * kStage);
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    fb_syncwarp();
    h_item<NP, EP>(J, rb, g, sm, lane);
}

// ---------------------------------------------------------------------------------------------------------
// vertical step
// ---------------------------------------------------------------------------------------------------------
struct VJob {                   // one vertical step on one plane
    const int16_t *avg, *res;
    int16_t *out;
    int16_t *est, *act;         // [nseg][w]
    unsigned char *bad;         // [nseg][w / 8]
    int *counter;               // [ncg] arrivals per group of 256 columns
    int *stats;
    int w, ha;                  // averages w x ha, residuals w x ha, output w x 2*ha;  w % 8 == 0
    int S, nseg, ncg;           // pairs per segment, segments per column, column groups of 256
    int item0;
    int do_clamp, lo, hi;
    uint32_t k1;                // 0x00010001 (see PK)
};
struct VJobs { VJob j[kMaxVJobs]; int n, items; };

FB_DEV uint4 ld16(const int16_t *p) { return *reinterpret_cast<const uint4 *>(p); }
FB_DEV void st16(int16_t *p, const uint4 &v) { *reinterpret_cast<uint4 *>(p) = v; }

// exact recomputation of segment g of the 8 columns at x from the true state before it (16-byte accesses, the rows of the
// next four pairs requested before the current four are computed)
FB_DEV void v_repair(const VJob &J, int x, int g, int *prev) {
    const int q0 = g * J.S, q1 = fq::imin(q0 + J.S, J.ha), last = J.ha - 1, w = J.w;
    for (int qb = q0; qb < q1; qb += 4) {
        uint4 A[5], R[4];
#pragma unroll
        for (int i = 0; i < 5; i++) A[i] = ld16(J.avg + (size_t)fq::imin(qb + i, last) * w + x);
#pragma unroll
        for (int i = 0; i < 4; i++) R[i] = ld16(J.res + (size_t)fq::imin(qb + i, last) * w + x);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int q = qb + i;
            if (q >= q1) break;
            const uint32_t aw[4] = {A[i].x, A[i].y, A[i].z, A[i].w}, nw[4] = {A[i + 1].x, A[i + 1].y, A[i + 1].z, A[i + 1].w}, rw[4] = {R[i].x, R[i].y, R[i].z, R[i].w};
            uint32_t oa[4], ob[4];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int sh = 16 * (k & 1);
                const int av = (int)(short)(aw[k >> 1] >> sh), nx = (int)(short)(nw[k >> 1] >> sh), rs = (int)(short)(rw[k >> 1] >> sh);     // row ha-1: A[i+1] re-reads it, next = own
                int Av, Bv;
                fq::unsqueeze_pair(q == 0 ? av : prev[k], av, nx, rs, Av, Bv);
                prev[k] = Bv;
                if (J.do_clamp) { Av = fq::clampi(Av, J.lo, J.hi); Bv = fq::clampi(Bv, J.lo, J.hi); }
                if (k & 1) { oa[k >> 1] |= (uint32_t)(uint16_t)Av << 16; ob[k >> 1] |= (uint32_t)(uint16_t)Bv << 16; }
                else { oa[k >> 1] = (uint32_t)(uint16_t)Av; ob[k >> 1] = (uint32_t)(uint16_t)Bv; }
            }
            uint4 v;
            v.x = oa[0]; v.y = oa[1]; v.z = oa[2]; v.w = oa[3]; st16(J.out + (size_t)(2 * q) * w + x, v);
            v.x = ob[0]; v.y = ob[1]; v.z = ob[2]; v.w = ob[3]; st16(J.out + (size_t)(2 * q + 1) * w + x, v);
        }
    }
}

FB_DEV void v_verify(const VJob &J, int x) {
    int dirty = 0;
#pragma unroll 8
    for (int g = 0; g < J.nseg; g++) {          // independent loads: a few round trips when everything joins
        dirty |= J.bad[(size_t)g * (J.w >> 3) + (x >> 3)];
        if (g > 0) {
            const uint4 e = *reinterpret_cast<const uint4 *>(J.est + (size_t)g * J.w + x), a = *reinterpret_cast<const uint4 *>(J.act + (size_t)(g - 1) * J.w + x);
            dirty |= (int)((e.x ^ a.x) | (e.y ^ a.y) | (e.z ^ a.z) | (e.w ^ a.w));
        }
    }
    if (!dirty) return;
    int cur[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int g = 0; g < J.nseg; g++) {
        bool need = J.bad[(size_t)g * (J.w >> 3) + (x >> 3)] != 0;
        if (g > 0)
            for (int k = 0; k < 8; k++) need = need || (J.est[(size_t)g * J.w + x + k] != (int16_t)cur[k]);
        if (need) {
            v_repair(J, x, g, cur);
            for (int k = 0; k < 8; k++) J.act[(size_t)g * J.w + x + k] = (int16_t)cur[k];
            atomicAdd(J.stats, 1);
        } else {
            for (int k = 0; k < 8; k++) cur[k] = J.act[(size_t)g * J.w + x + k];
        }
    }
}

template <int kVDepth>
FB_DEV void v_item(const VJob &J, int cg, int g, int lane) {
    const int x = (cg * 32 + lane) * 8, w = J.w, last_row = J.ha - 1;
    const bool active = x < w;
    const int q_own = g * J.S, q_end = fq::imin(q_own + J.S, J.ha);
    const int q_first = g ? q_own - kVWarm : 0;
    const uint32_t clo = (uint32_t)(uint16_t)J.lo * 0x00010001u, chi = (uint32_t)(uint16_t)J.hi * 0x00010001u;
    const PK K = pk_consts(J.k1);
    if (active) {
        uint32_t P[4], chk_a = 0, chk_r = 0;
        uint4 A[kVDepth + 1], R[kVDepth];
#pragma unroll
        for (int i = 0; i <= kVDepth; i++) A[i] = ld16(J.avg + (size_t)fq::imin(q_first + i, last_row) * w + x);
#pragma unroll
        for (int i = 0; i < kVDepth; i++) R[i] = ld16(J.res + (size_t)fq::imin(q_first + i, last_row) * w + x);
        P[0] = A[0].x; P[1] = A[0].y; P[2] = A[0].z; P[3] = A[0].w;         // chain start: "top" = own average; a segment start guesses the same
        for (int q0 = q_first; q0 < q_end; q0 += kVDepth) {
            uint4 An[kVDepth], Rn[kVDepth];
#pragma unroll
            for (int i = 0; i < kVDepth; i++) {             // rows of the next round (clamped into the plane: unused rows are harmless)
                An[i] = ld16(J.avg + (size_t)fq::imin(q0 + kVDepth + 1 + i, last_row) * w + x);
                Rn[i] = ld16(J.res + (size_t)fq::imin(q0 + kVDepth + i, last_row) * w + x);
            }
#pragma unroll
            for (int i = 0; i < kVDepth; i++) {
                const int q = q0 + i;
                if (q < q_end) {
                    if (q == q_own && g) {
                        uint4 e; e.x = P[0]; e.y = P[1]; e.z = P[2]; e.w = P[3];
                        st16(J.est + (size_t)g * w + x, e);
                    }
                    // last pair of the column: next average = own (squeeze.h:201); A[i+1] is then a clamped re-read of row ha-1 = A[i]
                    const uint32_t aw[4] = {A[i].x, A[i].y, A[i].z, A[i].w}, nw[4] = {A[i + 1].x, A[i + 1].y, A[i + 1].z, A[i + 1].w};
                    const uint32_t rw[4] = {R[i].x, R[i].y, R[i].z, R[i].w};
                    uint32_t oa[4], ob[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        chk_a = pk_chk(chk_a, aw[k], kMaxAvg * 0x00010001u);
                        if (q == q_end - 1) chk_a = pk_chk(chk_a, nw[k], kMaxAvg * 0x00010001u);     // a row the next segment owns
                        chk_r = pk_chk(chk_r, rw[k], kMaxRes * 0x00010001u);
                        pk_step(K, P[k], aw[k], pneg(aw[k], K), pneg(nw[k], K), rw[k], oa[k], ob[k]);
                        P[k] = ob[k];
                        if (J.do_clamp) { oa[k] = pk_clamp(oa[k], clo, chi); ob[k] = pk_clamp(ob[k], clo, chi); }
                    }
                    if (q >= q_own) {
                        uint4 v;
                        v.x = oa[0]; v.y = oa[1]; v.z = oa[2]; v.w = oa[3]; st16(J.out + (size_t)(2 * q) * w + x, v);
                        v.x = ob[0]; v.y = ob[1]; v.z = ob[2]; v.w = ob[3]; st16(J.out + (size_t)(2 * q + 1) * w + x, v);
                    }
                }
            }
            A[0] = A[kVDepth];
#pragma unroll
            for (int i = 0; i < kVDepth; i++) { A[i + 1] = An[i]; R[i] = Rn[i]; }
        }
        const bool bad = pk_chk_bad(chk_a, kMaxAvg) || pk_chk_bad(chk_r, kMaxRes);
        uint4 e; e.x = P[0]; e.y = P[1]; e.z = P[2]; e.w = P[3];
        st16(J.act + (size_t)g * w + x, e);
        J.bad[(size_t)g * (w >> 3) + (x >> 3)] = bad ? 1 : 0;
        if (bad) atomicAdd(J.stats + 1, 1);
    }
    fb_threadfence();
    fb_syncwarp();
    int last = 0;
    if (lane == 0) {
        const int seen = atomicAdd(J.counter + cg, 1);
        last = seen == J.nseg - 1;
        if (last) J.counter[cg] = 0;
    }
    last = warp_bcast0(last, lane);
    if (last && active) {
        fb_threadfence();
        v_verify(J, x);
    }
}

template <int kVDepth>
FB_KERNEL(256) k_pk_vsq(const FB_GRID_CONSTANT VJobs jobs, int warps_per_block) {
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const int item = (int)blockIdx.x * warps_per_block + warp;
    if (item >= jobs.items) return;
    int ji = 0;
    while (ji < jobs.n - 1 && item >= jobs.j[ji + 1].item0) ji++;
    const VJob &J = jobs.j[ji];
    // consecutive items = consecutive column groups of the same segment (neighbouring warps touch neighbouring lines)
    const int local = item - J.item0, g = local / J.ncg, cg = local - g * J.ncg;
    v_item<kVDepth>(J, cg, g, lane);
}

}  // namespace ps
