// Host-side planning for the packed unsqueeze kernels (fb_pk_squeeze.cuh): which planes of one squeeze step they take,
// segment lengths sized so that one round of work items fills the SMs, scratch layout (est / act / bad), launch shapes.
// Pure C++ so that the CPU-only test tier drives the same code under the emulator; the including translation unit
// provides ps_make_tilemap (driver cuTensorMapEncodeTiled in the product, a plain descriptor under the emulator).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "fb_pk_squeeze.cuh"

namespace ps {

// box_w x box_h tiles of a w x h int16 plane; false = the plane cannot be described (unaligned base / pitch)
bool ps_make_tilemap(TileMap *m, const void *base, int w, int h, int box_w, int box_h, int swizzle_bytes);

struct StepOp {                 // one plane of one squeeze step
    const int16_t *avg, *res;
    int16_t *out;
    int wa, wr, ha, hr;
    int clamp;                  // fold the final clamp into this op (it produces a final plane)
};
struct StepEpilogue {           // inverse YCoCg riding on the step that produces Co and Cg
    int enabled = 0;
    const int16_t *yin = nullptr;
    int16_t *rout = nullptr;
    const int16_t *co_out = nullptr, *cg_out = nullptr;
    int maxval = 0, lo = 0, hi = 0, do_clamp = 0;
};
struct HLaunch { HJobs jobs; int np, ep, grid, warps_per_block, smem_per_warp; size_t smem; double bytes; };
struct VLaunch { VJobs jobs; int grid, warps_per_block; double bytes; };
struct StepPlan {
    std::vector<HLaunch> h;
    std::vector<VLaunch> v;
    std::vector<int> leftover;  // ops these kernels do not take
    bool epilogue_done = false;
    size_t scratch_bytes = 0;   // est / act / bad, offsets stored as pointers relative to nullptr until relocate()
    int counters = 0;           // ints of arrival counters
};

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline bool h_eligible(const StepOp &o) {
    return o.res && o.wr == o.wa && o.wa >= 64 && (o.wa & 7) == 0 && o.ha >= 16 && aligned16(o.avg) && aligned16(o.out) && aligned16(o.res);
}
inline bool v_eligible(const StepOp &o) {
    return o.res && o.hr == o.ha && o.ha >= 32 && o.wa >= 64 && (o.wa & 7) == 0 && aligned16(o.avg) && aligned16(o.out) && aligned16(o.res);
}

struct ScratchCursor {
    size_t off = 0;
    template <class T> T *take(size_t n) {
        off = (off + 15) & ~(size_t)15;
        T *p = reinterpret_cast<T *>(off);
        off += n * sizeof(T);
        return p;
    }
};

// warps of the packed horizontal kernel that fit one SM (shared memory is the limit)
inline int h_warps_per_sm(int np, int ep) {
    int w = (int)((226 * 1024) / (h_smem_per_warp(np, ep) + 1024 + 1024));     // + alignment slack + the per-block reservation
    return w > 16 ? 16 : w;
}

inline StepPlan plan_step(const std::vector<StepOp> &ops, bool horizontal, const StepEpilogue &ep, int lo, int hi, int sm_count, int *stats_dev) {
    StepPlan P;
    const int n = (int)ops.size();
    std::vector<char> taken(n, 0);
    ScratchCursor cur;
    if (horizontal) {
        struct Pend { int i0, i1, epk; };
        std::vector<Pend> pend;
        int ico = -1, icg = -1;
        if (ep.enabled)
            for (int i = 0; i < n; i++) { if (ops[i].out == ep.co_out) ico = i; if (ops[i].out == ep.cg_out) icg = i; }
        auto same = [&](const StepOp &a, const StepOp &b) { return a.wa == b.wa && a.ha == b.ha && a.clamp == b.clamp; };
        if (ico >= 0 && icg >= 0 && ico != icg && h_eligible(ops[ico]) && h_eligible(ops[icg]) && same(ops[ico], ops[icg]) && aligned16(ep.yin) &&
            aligned16(ep.rout) && ep.maxval >= 0) {
            pend.push_back(Pend{ico, icg, fq::kEpYCoCg});
            taken[ico] = taken[icg] = 1;
            P.epilogue_done = true;
        }
        // every other plane is a job of its own in ONE launch (one chain per lane, but twice the warps of a paired job fit an SM,
        // and luma and chroma of a step run side by side instead of one launch after the other)
        for (int i = 0; i < n; i++) {
            if (taken[i] || !h_eligible(ops[i])) continue;
            pend.push_back(Pend{i, -1, ops[i].clamp ? fq::kEpClamp : fq::kEpNone});
            taken[i] = 1;
        }
        // one launch per (np, epilogue) class
        for (int cls = 0; cls < 6; cls++) {
            const int np = 1 + (cls & 1), epk = cls >> 1;
            std::vector<Pend> mine;
            for (auto &pd : pend) if ((pd.i1 >= 0 ? 2 : 1) == np && pd.epk == epk) mine.push_back(pd);
            for (size_t j0 = 0; j0 < mine.size(); j0 += kMaxHJobs) {
                const size_t j1 = j0 + kMaxHJobs < mine.size() ? j0 + kMaxHJobs : mine.size();
                HLaunch L;
                memset(&L.jobs, 0, sizeof(L.jobs));
                L.np = np; L.ep = epk; L.bytes = 0;
                const int wps = h_warps_per_sm(np, epk);
                long long total_rb = 0;
                for (size_t j = j0; j < j1; j++) total_rb += (ops[mine[j].i0].ha + kHRows - 1) / kHRows;
                // one round of items: at most wps warps per SM fit, 12 (three per scheduler) keep the integer pipes busy through
                // the dependent-issue latency of the chains -- beyond that shorter segments only add warm-up work
                long long nseg_want = ((long long)sm_count * (wps < 12 ? wps : 12)) / (total_rb > 0 ? total_rb : 1);
                if (nseg_want < 1) nseg_want = 1;
                int items = 0;
                bool ok = true;
                for (size_t j = j0; j < j1; j++) {
                    const StepOp &a = ops[mine[j].i0];
                    HJob &J = L.jobs.j[L.jobs.n];
                    const int i1 = mine[j].i1;
                    int S = (int)((a.wa + nseg_want - 1) / nseg_want);
                    S = (S + 15) / 16 * 16;         // (a multiple of 16 whatever the tile width: only the last segment of a row may end on 8)
                    if (S < 16) S = 16;
                    J.np = np; J.wa = a.wa; J.h = a.ha; J.S = S; J.nseg = (a.wa + S - 1) / S; J.nrb = (a.ha + kHRows - 1) / kHRows;
                    J.nsegp = (J.nseg + 7) / 8 * 8;
                    J.item0 = items;
                    J.k1 = 0x00010001u;
                    items += J.nrb * J.nseg;
                    J.epilogue = epk; J.maxval = ep.maxval; J.lo = lo; J.hi = hi; J.do_clamp = epk == fq::kEpClamp ? 1 : 0;
                    const StepOp *pl[2] = {&a, i1 >= 0 ? &ops[i1] : nullptr};
                    for (int p = 0; p < np; p++) {
                        J.avg[p] = pl[p]->avg; J.res[p] = pl[p]->res;
                        ok = ok && ps_make_tilemap(&J.tm_a[p], pl[p]->avg, a.wa, a.ha, kHChunk, kHRows, kSwzA);
                        ok = ok && ps_make_tilemap(&J.tm_r[p], pl[p]->res, a.wa, a.ha, kHChunk, kHRows, kSwzA);
                        J.est[p] = cur.take<int16_t>((size_t)J.nsegp * a.ha);
                        J.act[p] = cur.take<int16_t>((size_t)J.nsegp * a.ha);
                    }
                    J.bad = cur.take<int16_t>((size_t)J.nsegp * a.ha);
                    J.counter = reinterpret_cast<int *>((size_t)P.counters * sizeof(int));
                    P.counters += J.nrb;
                    J.stats = stats_dev;
                    if (epk == fq::kEpYCoCg) {
                        J.yin = ep.yin; J.out[0] = ep.rout; J.out[1] = a.out; J.out[2] = ops[i1].out;
                        J.do_clamp = ep.do_clamp; J.lo = ep.lo; J.hi = ep.hi;
                        ok = ok && ps_make_tilemap(&J.tm_y, ep.yin, 2 * a.wa, a.ha, 2 * kHChunk, kHRows, kSwzO);
                        for (int k = 0; k < 3; k++) ok = ok && ps_make_tilemap(&J.tm_o[k], J.out[k], 2 * a.wa, a.ha, 2 * kHChunk, kHRows, kSwzO);
                    } else {
                        for (int p = 0; p < np; p++) {
                            J.out[p] = pl[p]->out;
                            ok = ok && ps_make_tilemap(&J.tm_o[p], pl[p]->out, 2 * a.wa, a.ha, 2 * kHChunk, kHRows, kSwzO);
                        }
                    }
                    L.bytes += (double)np * 8.0 * a.wa * a.ha + (epk == fq::kEpYCoCg ? 8.0 * a.wa * a.ha : 0.0);
                    L.jobs.n++;
                }
                if (!ok) {          // a plane TMA cannot describe: leave this class to the older kernels
                    for (size_t j = j0; j < j1; j++) { taken[mine[j].i0] = 0; if (mine[j].i1 >= 0) taken[mine[j].i1] = 0; }
                    if (epk == fq::kEpYCoCg) P.epilogue_done = false;
                    continue;
                }
                L.jobs.items = items;
                L.warps_per_block = 1;
                L.smem_per_warp = (int)h_smem_per_warp(np, epk);
                L.smem = (size_t)L.warps_per_block * L.smem_per_warp + 1024;
                L.grid = (items + L.warps_per_block - 1) / L.warps_per_block;
                P.h.push_back(L);
            }
        }
    } else {
        std::vector<int> mine;
        for (int i = 0; i < n; i++) if (v_eligible(ops[i])) { mine.push_back(i); taken[i] = 1; }
        for (size_t j0 = 0; j0 < mine.size(); j0 += kMaxVJobs) {
            const size_t j1 = j0 + kMaxVJobs < mine.size() ? j0 + kMaxVJobs : mine.size();
            VLaunch L;
            memset(&L.jobs, 0, sizeof(L.jobs));
            L.bytes = 0;
            long long total_cg = 0;
            for (size_t j = j0; j < j1; j++) total_cg += (ops[mine[j]].wa + 255) / 256;
            static const int v_warps = getenv("FB_PK_VWARPS") ? atoi(getenv("FB_PK_VWARPS")) : 8;    // resident warps per SM aimed at (4 chains each)
            long long nseg_want = ((long long)sm_count * v_warps) / (total_cg > 0 ? total_cg : 1);
            if (nseg_want < 1) nseg_want = 1;
            int items = 0;
            for (size_t j = j0; j < j1; j++) {
                const StepOp &a = ops[mine[j]];
                VJob &J = L.jobs.j[L.jobs.n++];
                int S = (int)((a.ha + nseg_want - 1) / nseg_want);
                S = (S + 7) / 8 * 8;
                if (S < 16) S = 16;
                static const int v_maxseg = getenv("FB_PK_VMAXSEG") ? atoi(getenv("FB_PK_VMAXSEG")) : 32;
                while ((a.ha + S - 1) / S > v_maxseg) S += 8;  // the verification pass reads two 16-byte words per segment and lane
                J.avg = a.avg; J.res = a.res; J.out = a.out; J.w = a.wa; J.ha = a.ha; J.S = S; J.nseg = (a.ha + S - 1) / S; J.ncg = (a.wa + 255) / 256;
                J.item0 = items;
                J.k1 = 0x00010001u;
                items += J.ncg * J.nseg;
                J.do_clamp = a.clamp; J.lo = lo; J.hi = hi;
                J.est = cur.take<int16_t>((size_t)J.nseg * a.wa);
                J.act = cur.take<int16_t>((size_t)J.nseg * a.wa);
                J.bad = cur.take<unsigned char>((size_t)J.nseg * (a.wa / 8));
                J.counter = reinterpret_cast<int *>((size_t)P.counters * sizeof(int));
                P.counters += J.ncg;
                J.stats = stats_dev;
                L.bytes += 8.0 * a.wa * a.ha;
            }
            L.jobs.items = items;
            L.warps_per_block = 2;
            L.grid = (items + L.warps_per_block - 1) / L.warps_per_block;
            P.v.push_back(L);
        }
    }
    P.scratch_bytes = cur.off + 64;
    for (int i = 0; i < n; i++) if (!taken[i]) P.leftover.push_back(i);
    return P;
}

// scratch pointers were laid out relative to address 0: move them into the real buffers
inline void relocate(StepPlan &P, unsigned char *scratch, int *counters) {
    auto mv = [&](auto *&p) { p = reinterpret_cast<std::remove_reference_t<decltype(p)>>(scratch + reinterpret_cast<size_t>(p)); };
    for (auto &L : P.h)
        for (int j = 0; j < L.jobs.n; j++) {
            HJob &J = L.jobs.j[j];
            for (int p = 0; p < J.np; p++) { mv(J.est[p]); mv(J.act[p]); }
            mv(J.bad);
            J.counter = counters + reinterpret_cast<size_t>(J.counter) / sizeof(int);
        }
    for (auto &L : P.v)
        for (int j = 0; j < L.jobs.n; j++) {
            VJob &J = L.jobs.j[j];
            mv(J.est); mv(J.act); mv(J.bad);
            J.counter = counters + reinterpret_cast<size_t>(J.counter) / sizeof(int);
        }
}

}  // namespace ps
