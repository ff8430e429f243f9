// Host-side planning for the direct per-step unsqueeze kernels (fb_direct_squeeze.cuh): which planes of one squeeze
// step are eligible (even output size, rows of whole 16-byte chunks, 16-byte aligned planes), how they are grouped
// into jobs, segment lengths and block shapes.  Pure C++ so that the CPU-only test tier drives the same code.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "fb_direct_squeeze.cuh"

namespace dq {

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
    const int16_t *co_out = nullptr, *cg_out = nullptr;     // identify the two ops
    int maxval = 0, lo = 0, hi = 0, do_clamp = 0;
};
struct StepPlan {
    HJobs hj; int h_grid = 0, h_threads = 0; size_t h_smem = 0;
    VJobs vj; int v_grid = 0, v_threads = 0; size_t v_smem = 0;
    std::vector<int> leftover;  // ops the direct kernels do not take
    bool epilogue_done = false;
    double bytes = 0;           // algorithmic HBM bytes of the direct launch
};

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline bool h_eligible(const StepOp &o) {
    return o.wr == o.wa && o.wa >= 8 && (o.wa & 7) == 0 && o.ha > 0 && aligned16(o.avg) && aligned16(o.out) && (!o.res || aligned16(o.res));
}
inline bool v_eligible(const StepOp &o) {
    return o.hr == o.ha && o.ha >= 1 && o.wa >= 8 && (o.wa & 7) == 0 && aligned16(o.avg) && aligned16(o.out) && (!o.res || aligned16(o.res));
}

inline StepPlan plan_step(const std::vector<StepOp> &ops, bool horizontal, const StepEpilogue &ep, int lo, int hi, int sm_count = 148) {
    StepPlan P;
    memset(&P.hj, 0, sizeof(P.hj));
    memset(&P.vj, 0, sizeof(P.vj));
    const int n = (int)ops.size();
    std::vector<char> taken(n, 0);
    if (horizontal) {
        long long pairs_total = 0;
        for (auto &o : ops) pairs_total += (long long)o.wa * o.ha;
        // 16 pairs per segment (+ 8 warm-up pairs): measured best on B200 at every level size (8, 24, 32, 64, 128 tried);
        // more, shorter segments beat less warm-up work because the kernel lives on the number of chains in flight
        int S = 16;
        if (pairs_total / 16 < (long long)sm_count * 256) S = 8;
        static const int env_hs = getenv("FB_DQ_HS") ? atoi(getenv("FB_DQ_HS")) : 0;
        if (env_hs > 0 && S == 16) S = env_hs;
        // the YCoCg pair first
        int ico = -1, icg = -1;
        if (ep.enabled)
            for (int i = 0; i < n; i++) { if (ops[i].out == ep.co_out) ico = i; if (ops[i].out == ep.cg_out) icg = i; }
        auto add_job = [&](int i0, int i1, bool with_ep) {
            HJob &J = P.hj.j[P.hj.n];
            const StepOp &a = ops[i0];
            int s = S;
            while ((a.wa + s - 1) / s > 512) s += 8;
            J.np = i1 >= 0 ? 2 : 1;
            J.avg[0] = a.avg; J.res[0] = a.res; J.out[0] = a.out;
            if (i1 >= 0) { J.avg[1] = ops[i1].avg; J.res[1] = ops[i1].res; J.out[1] = ops[i1].out; }
            J.wa = a.wa; J.h = a.ha; J.S = s; J.nseg = (a.wa + s - 1) / s;
            J.R = J.nseg >= 256 ? 1 : 256 / J.nseg;
            if (J.R > a.ha) J.R = a.ha;
            J.blocks = (a.ha + J.R - 1) / J.R;
            J.maxval = ep.maxval; J.lo = lo; J.hi = hi;
            if (with_ep) { J.epilogue = fq::kEpYCoCg; J.yin = ep.yin; J.rout = ep.rout; J.do_clamp = ep.do_clamp; J.lo = ep.lo; J.hi = ep.hi; }
            else if (a.clamp) { J.epilogue = fq::kEpClamp; J.do_clamp = 1; }
            else { J.epilogue = fq::kEpNone; J.do_clamp = 0; }
            const int threads = ((J.R * J.nseg + 31) / 32) * 32;
            if (threads > P.h_threads) P.h_threads = threads;
            const size_t smem = (size_t)J.R * J.nseg * J.np * sizeof(int16_t) + 16;
            if (smem > P.h_smem) P.h_smem = smem;
            P.h_grid += J.blocks;
            P.bytes += (double)J.np * (2.0 * a.wa * a.ha * (a.res ? 2 : 1) + 4.0 * a.wa * a.ha) + (with_ep ? 8.0 * a.wa * a.ha : 0.0);
            P.hj.n++;
            taken[i0] = 1;
            if (i1 >= 0) taken[i1] = 1;
        };
        auto same = [&](const StepOp &a, const StepOp &b) { return a.wa == b.wa && a.ha == b.ha && a.clamp == b.clamp; };
        if (ico >= 0 && icg >= 0 && ico != icg && h_eligible(ops[ico]) && h_eligible(ops[icg]) && same(ops[ico], ops[icg]) && aligned16(ep.yin) && aligned16(ep.rout)) {
            add_job(ico, icg, true);
            P.epilogue_done = true;
        }
        for (int i = 0; i < n && P.hj.n < 3; i++) {
            if (taken[i] || !h_eligible(ops[i])) continue;
            int mate = -1;
            for (int k = i + 1; k < n; k++) if (!taken[k] && h_eligible(ops[k]) && same(ops[i], ops[k])) { mate = k; break; }
            add_job(i, mate, false);
        }
    } else {
        long long groups_total = 0;
        for (auto &o : ops) groups_total += (long long)(o.wa / 8) * o.ha;
        int S = 32;
        if (groups_total / 32 < (long long)sm_count * 128) S = 16;
        if (groups_total / 16 < (long long)sm_count * 128) S = 8;
        static const int env_s = getenv("FB_DQ_VS") ? atoi(getenv("FB_DQ_VS")) : 0, env_t = getenv("FB_DQ_VT") ? atoi(getenv("FB_DQ_VT")) : 128;
        if (env_s > 0) S = env_s;
        for (int i = 0; i < n && P.vj.n < 4; i++) {
            const StepOp &a = ops[i];
            if (!v_eligible(a)) continue;
            VJob &J = P.vj.j[P.vj.n++];
            int s = S;
            while ((a.ha + s - 1) / s > 256) s += 8;
            J.avg = a.avg; J.res = a.res; J.out = a.out; J.w = a.wa; J.ha = a.ha; J.S = s; J.nseg = (a.ha + s - 1) / s;
            int cb = env_t / J.nseg;
            if (cb < 2) cb = 2;
            if (cb > 8) cb = 8;
            while (cb > 1 && cb * J.nseg > 512) cb--;
            if (cb > a.wa / 8) cb = a.wa / 8;
            J.CB = cb;
            J.blocks = (a.wa / 8 + cb - 1) / cb;
            J.do_clamp = a.clamp; J.lo = lo; J.hi = hi;
            const int threads = ((J.CB * J.nseg + 31) / 32) * 32;
            if (threads > P.v_threads) P.v_threads = threads;
            const size_t smem = (size_t)J.CB * J.nseg * 8 * sizeof(int16_t) + 16;
            if (smem > P.v_smem) P.v_smem = smem;
            P.v_grid += J.blocks;
            P.bytes += 2.0 * a.wa * a.ha * (a.res ? 2 : 1) + 4.0 * a.wa * a.ha;
            taken[i] = 1;
        }
    }
    for (int i = 0; i < n; i++) if (!taken[i]) P.leftover.push_back(i);
    return P;
}

}  // namespace dq
