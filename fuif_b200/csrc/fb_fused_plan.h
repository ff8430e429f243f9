// Host-side planner for the fused unsqueeze tile kernel (fb_fused_squeeze.cuh): cuts the step list of one
// Squeeze inverse (reference transform/squeeze.h:367-388, steps walked backwards) into a few launches, groups planes
// of identical geometry into gangs, sizes tiles / shared memory / verification scratch.  Pure C++ (no CUDA calls) so
// that the CPU-only test tier drives the same planner as the library.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string.h>

#include <algorithm>
#include <vector>

#include "fb_fused_squeeze.cuh"

namespace fq {

struct PlanOp {                 // = FbSqOp (fb_common.cuh): one unsqueeze step on one plane, execution order
    int step, horizontal;
    const int16_t *avg, *res;
    int16_t *out;
    int wa, wr, ha, hr;
};

struct EpilogueSpec {
    int kind = kEpNone;         // what follows the last step: nothing, clamp of every plane, inverse YCoCg (+ clamp)
    int maxval = 0, lo = 0, hi = 0, do_clamp = 0;
    const int16_t *ycc[3] = {nullptr, nullptr, nullptr};    // final planes holding Y, Co, Cg (kEpYCoCg)
};

struct PlanOptions {
    int tile_w = 64, tile_h = 64;       // tile of the final planes of a tiled launch (multiples of 16)
    int levels_per_launch = 4;          // steps fused per tiled launch
    int coarse_dim = 128;               // the first launch takes every step whose output is at most this wide and high
    int threads_per_gang = 128;
    int warm_last = kWarm, warm_mid = kWarmMid;
    size_t max_smem = 200 * 1024;
};

struct PlannedLaunch {
    Task task;
    int grid = 0, threads = 0;
    size_t smem = 0;
    double bytes = 0;           // algorithmic HBM bytes: level-0 averages + residuals read once, final planes written once
};

struct Plan {
    Plan() { memset(&verify, 0, sizeof(verify)); }
    bool ok = false;            // false: shape outside what the fused kernel handles -> caller uses the per-level kernels
    bool epilogue_fused = false;
    std::vector<PlannedLaunch> launches;
    bool need_verify = false;
    VerifyParams verify;        // counters pointer left for the caller
    int verify_threads = 256;   // block size / shared memory of the cooperative verification launch (= the last launch's)
    size_t verify_smem = 0;
    size_t scratch_bytes = 0;   // est / act scratch; pointers in tasks / checks are offsets (+1) until relocated
    size_t tile_bad_off = 0, bad_list_off = 0;
};

namespace detail {

inline bool same_geometry(const std::vector<int> &a, const std::vector<int> &b, const std::vector<PlanOp> &ops) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++) {
        const PlanOp &x = ops[a[i]], &y = ops[b[i]];
        if (x.step != y.step || x.horizontal != y.horizontal || x.wa != y.wa || x.wr != y.wr || x.ha != y.ha || x.hr != y.hr) return false;
    }
    return true;
}

}  // namespace detail

// Relocates the scratch offsets stored in a plan to a real allocation and wires the counters.
inline void relocate_scratch(Plan &P, unsigned char *base, int *counters) {
    auto fix_i = [&](int *&p) { if (p) p = reinterpret_cast<int *>(base + (reinterpret_cast<uintptr_t>(p) - 1)); };
    auto fix_s = [&](int16_t *&p) { if (p) p = reinterpret_cast<int16_t *>(base + (reinterpret_cast<uintptr_t>(p) - 1)); };
    for (auto &L : P.launches)
        for (int gi = 0; gi < L.task.ngangs; gi++)
            for (int k = 0; k < L.task.g[gi].nlev; k++)
                for (int pl = 0; pl < kGP; pl++) { fix_i(L.task.g[gi].lv[k].est[pl]); fix_s(L.task.g[gi].lv[k].act[pl]); }
    for (int i = 0; i < P.verify.nchecks; i++) {
        int *e = const_cast<int *>(P.verify.chk[i].est);
        int16_t *a = const_cast<int16_t *>(P.verify.chk[i].act);
        fix_i(e); fix_s(a);
        P.verify.chk[i].est = e; P.verify.chk[i].act = a;
    }
    for (auto &L : P.launches) { L.task.counters = counters; L.task.tile_bad = nullptr; }
    P.verify.counters = counters;
    if (!P.launches.empty()) {
        if (P.tile_bad_off) P.launches.back().task.tile_bad = base + (P.tile_bad_off - 1);
        P.verify.bad_list = P.bad_list_off ? reinterpret_cast<int *>(base + (P.bad_list_off - 1)) : nullptr;
        P.verify.top = P.launches.back().task;
    }
}

inline Plan make_plan(const std::vector<PlanOp> &ops, const EpilogueSpec &ep, const PlanOptions &opt) {
    Plan P;
    const int n = (int)ops.size();
    if (!n) return P;
    if (opt.tile_w % 16 || opt.tile_h % 16 || opt.tile_w < 16 || opt.tile_h < 16) return P;
    // ---- chains: op k continues the chain whose last op produced its average plane
    std::vector<int> chain(n, -1);
    std::vector<std::vector<int>> chains;
    for (int k = 0; k < n; k++) {
        const PlanOp &o = ops[k];
        if (o.wa <= 0 || o.ha <= 0 || !o.avg || !o.out) return P;
        if (o.horizontal ? (o.wr <= 0 || o.wr > o.wa || o.wa - o.wr > 1) : (o.hr <= 0 || o.hr > o.ha || o.ha - o.hr > 1)) return P;
        if (k && ops[k].step < ops[k - 1].step) return P;
        for (int q = k - 1; q >= 0; q--)
            if (ops[q].out == o.avg) { chain[k] = chain[q]; break; }
        if (chain[k] < 0) { chain[k] = (int)chains.size(); chains.emplace_back(); }
        chains[chain[k]].push_back(k);
    }
    // every chain must be a simple line (its ops consumed in order) -- guaranteed by construction above
    const int nsteps = ops.back().step + 1;
    // ---- cut the steps into launches: a coarse single-tile launch from step 0, then groups from the end backwards
    auto out_w = [&](const PlanOp &o) { return o.horizontal ? o.wa + o.wr : o.wa; };
    auto out_h = [&](const PlanOp &o) { return o.horizontal ? o.ha : o.ha + o.hr; };
    int coarse_end = 0;     // steps [0, coarse_end)
    {
        std::vector<int> cnt(chains.size(), 0);
        for (int s = 0; s < nsteps; s++) {
            bool fits = true;
            for (int k = 0; k < n; k++)
                if (ops[k].step == s) {
                    if (out_w(ops[k]) > opt.coarse_dim || out_h(ops[k]) > opt.coarse_dim) fits = false;
                    if (cnt[chain[k]] + 1 > kMaxLevels) fits = false;
                }
            if (!fits) break;
            for (int k = 0; k < n; k++) if (ops[k].step == s) cnt[chain[k]]++;
            coarse_end = s + 1;
        }
    }
    std::vector<std::pair<int, int>> ranges;    // [first, last) step ranges in execution order
    {
        std::vector<std::pair<int, int>> rev;
        int hi_ = nsteps;
        while (hi_ > coarse_end) {
            std::vector<int> cnt(chains.size(), 0);
            int lo_ = hi_;
            while (lo_ > coarse_end) {
                bool fits = true;
                for (int k = 0; k < n; k++) if (ops[k].step == lo_ - 1 && cnt[chain[k]] + 1 > opt.levels_per_launch) fits = false;
                if (!fits) break;
                for (int k = 0; k < n; k++) if (ops[k].step == lo_ - 1) cnt[chain[k]]++;
                lo_--;
            }
            if (lo_ == hi_) return P;
            rev.emplace_back(lo_, hi_);
            hi_ = lo_;
        }
        if (coarse_end > 0) ranges.emplace_back(0, coarse_end);
        for (int i = (int)rev.size() - 1; i >= 0; i--) ranges.push_back(rev[i]);
    }
    // ---- scratch allocator (offsets + 1 so that 0 stays "none")
    size_t scratch = 0;
    auto salloc = [&](size_t bytes) { size_t o = scratch; scratch += (bytes + 255) & ~(size_t)255; return o + 1; };
    P.verify.nchecks = 0;
    P.verify.nops = 0;
    // ---- one or more launches per range
    for (size_t ri = 0; ri < ranges.size(); ri++) {
        const int s0 = ranges[ri].first, s1 = ranges[ri].second;
        const bool last_range = ri + 1 == ranges.size();
        // chain segments inside the range
        std::vector<std::vector<int>> segs;
        for (auto &c : chains) {
            std::vector<int> sg;
            for (int k : c) if (ops[k].step >= s0 && ops[k].step < s1) sg.push_back(k);
            if (!sg.empty()) segs.push_back(sg);
        }
        // gangs: segments of identical geometry, at most kGP planes each
        std::vector<std::vector<int>> gangs;    // indices into segs
        for (int si = 0; si < (int)segs.size(); si++) {
            bool placed = false;
            for (auto &g : gangs)
                if ((int)g.size() < kGP && detail::same_geometry(segs[g[0]], segs[si], ops)) { g.push_back(si); placed = true; break; }
            if (!placed) gangs.push_back({si});
        }
        // tasks: with a colour epilogue every gang of the final range shares one CTA (they share W x H); otherwise
        // every gang is its own launch
        std::vector<std::vector<int>> tasks;    // gang indices
        const bool want_ep = last_range && ep.kind != kEpNone;
        bool join = false;
        if (want_ep) {
            join = (int)gangs.size() <= kMaxGangs;
            int W = -1, H = -1;
            for (auto &g : gangs) {
                const PlanOp &lo = ops[segs[g[0]].back()];
                if (W < 0) { W = out_w(lo); H = out_h(lo); }
                else if (W != out_w(lo) || H != out_h(lo)) join = false;
            }
            // every plane the epilogue touches must be produced by this range
            if (ep.kind == kEpYCoCg)
                for (int j = 0; j < 3; j++) {
                    bool found = false;
                    for (auto &sg : segs) if (ops[sg.back()].out == ep.ycc[j]) found = true;
                    if (!found) join = false;
                }
        }
        if (join) { std::vector<int> all; for (int gi = 0; gi < (int)gangs.size(); gi++) all.push_back(gi); tasks.push_back(all); }
        else
            for (int gi = 0; gi < (int)gangs.size(); gi++) {
                if (tasks.empty() || (int)tasks.back().size() >= kMaxGangs) tasks.emplace_back();
                tasks.back().push_back(gi);
            }
        if (last_range) P.epilogue_fused = join;

        for (auto &tk : tasks) {
            PlannedLaunch PL;
            Task &T = PL.task;
            memset(&T, 0, sizeof(T));
            T.ngangs = (int)tk.size();
            T.joined = join ? 1 : 0;
            const bool single = ri == 0 && coarse_end > 0;
            T.epilogue = join ? ep.kind : kEpNone;
            T.maxval = ep.maxval; T.lo = ep.lo; T.hi = ep.hi; T.do_clamp = join ? ep.do_clamp : 0;
            for (int j = 0; j < 3; j++) { T.ycc_gang[j] = -1; T.ycc_plane[j] = -1; }
            int thread0 = 0, block0 = 0;
            size_t smem_hw = 0, smem_hw_max = 0;     // halfwords
            for (int gq = 0; gq < T.ngangs; gq++) {
                Gang &G = T.g[gq];
                const std::vector<int> &members = gangs[tk[gq]];
                const std::vector<int> &seg0 = segs[members[0]];
                G.np = (int)members.size();
                G.nlev = (int)seg0.size();
                if (G.nlev > kMaxLevels) return Plan();
                G.w0 = ops[seg0[0]].wa; G.h0 = ops[seg0[0]].ha;
                const int W = out_w(ops[seg0.back()]), H = out_h(ops[seg0.back()]);
                G.W = W; G.H = H;
                G.TW = single ? ((W + 15) & ~15) : opt.tile_w;
                G.TH = single ? ((H + 15) & ~15) : opt.tile_h;
                G.ntx = (W + G.TW - 1) / G.TW;
                G.nty = (H + G.TH - 1) / G.TH;
                if (join && gq > 0 && (G.ntx != T.g[0].ntx || G.nty != T.g[0].nty || W != T.g[0].W || H != T.g[0].H)) return Plan();
                G.bar_id = 1 + gq;
                G.warm = last_range ? opt.warm_last : opt.warm_mid;
                if (join) { G.first_thread = thread0; G.nthreads = opt.threads_per_gang; thread0 += G.nthreads; G.first_block = 0; }
                else { G.first_thread = 0; G.nthreads = opt.threads_per_gang; thread0 = G.nthreads; G.first_block = block0; block0 += G.ntx * G.nty; smem_hw = 0; }
                Gang &TT = G;       // tile grid of this gang
                int nh_after = 0, nv_after = 0;
                for (int k = G.nlev - 1; k >= 0; k--) {
                    Level &L = G.lv[k];
                    const PlanOp &o = ops[seg0[k]];
                    L.horizontal = o.horizontal;
                    L.wa = o.wa; L.ha = o.ha;
                    L.wr = o.horizontal ? o.wr : o.wa; L.hr = o.horizontal ? o.ha : o.hr;
                    L.wo = out_w(o); L.ho = out_h(o);
                    if (TT.ntx > 1 && (TT.TW >> nh_after) < 2) return Plan();
                    if (TT.nty > 1 && (TT.TH >> nv_after) < 2) return Plan();
                    if (TT.ntx > 1 && (TT.TW % (1 << (nh_after + (o.horizontal ? 1 : 0))))) return Plan();
                    if (TT.nty > 1 && (TT.TH % (1 << (nv_after + (o.horizontal ? 0 : 1))))) return Plan();
                    L.tw = TT.TW >> nh_after; L.th = TT.TH >> nv_after;
                    for (int pl = 0; pl < G.np; pl++) L.res[pl] = ops[segs[members[pl]][k]].res;
                    if (o.horizontal) nh_after++; else nv_after++;
                }
                for (int pl = 0; pl < G.np; pl++) {
                    G.in[pl] = ops[segs[members[pl]][0]].avg;
                    G.out[pl] = ops[segs[members[pl]].back()].out;
                    if (T.epilogue == kEpYCoCg)
                        for (int j = 0; j < 3; j++) if (G.out[pl] == ep.ycc[j]) { T.ycc_gang[j] = gq; T.ycc_plane[j] = pl; }
                }
                // ---- buffer extents: maximum over all tiles (x extents depend on ti only, y extents on tj only)
                std::vector<int> out_w_(G.nlev, 0), out_h_(G.nlev, 0), res_w_(G.nlev, 0), res_h_(G.nlev, 0), est_n(G.nlev, 0);
                int in_w = 0, in_h = 0;
                std::vector<std::pair<int, int>> probes;
                for (int i = 0; i < G.ntx; i++) probes.emplace_back(i, 0);
                for (int j = 1; j < G.nty; j++) probes.emplace_back(0, j);
                Geom gm[kMaxLevels];
                Region inr;
                for (auto &pr : probes) {
                    {
                        const int ti = pr.first, tj = pr.second;
                        geometry(G, ti, tj, G.warm, gm, inr);
                        in_w = std::max(in_w, ((inr.x1 - (inr.x0 & ~7)) + 7) & ~7);
                        in_h = std::max(in_h, inr.y1 - inr.y0);
                        for (int k = 0; k < G.nlev; k++) {
                            const Geom &q = gm[k];
                            out_w_[k] = std::max(out_w_[k], (q.x1 - q.x0 + 7) & ~7);
                            out_h_[k] = std::max(out_h_[k], q.y1 - q.y0);
                            if (G.lv[k].horizontal) {
                                res_w_[k] = std::max(res_w_[k], ((q.p_end - (q.p_start & ~7)) + 7) & ~7);
                                res_h_[k] = std::max(res_h_[k], q.c1 - q.c0);
                            } else {
                                res_w_[k] = std::max(res_w_[k], ((q.c1 - (q.c0 & ~7)) + 7) & ~7);
                                res_h_[k] = std::max(res_h_[k], q.p_end - q.p_start);
                            }
                            est_n[k] = std::max(est_n[k], q.c1 - q.ce);
                        }
                    }
                }
                // ---- shared-memory layout: residuals back to back, level buffers alternate between two arenas
                auto take = [&](size_t hw) { size_t o = smem_hw; smem_hw += (hw + 7) & ~(size_t)7; return (int)o; };
                G.in_pitch = odd_pitch(in_w);
                size_t arena[2] = {0, 0};
                arena[0] = (size_t)G.in_pitch * in_h;
                for (int k = 0; k < G.nlev; k++) {
                    Level &L = G.lv[k];
                    L.out_pitch = odd_pitch(out_w_[k]);
                    L.res_pitch = odd_pitch(std::max(res_w_[k], 8));
                    arena[(k + 1) & 1] = std::max(arena[(k + 1) & 1], (size_t)L.out_pitch * out_h_[k]);
                    L.est_cap = std::max(est_n[k], 1);
                }
                for (int pl = 0; pl < G.np; pl++) {
                    const int a0 = take(arena[0]), a1 = take(arena[1]);
                    G.in_off[pl] = a0;
                    for (int k = 0; k < G.nlev; k++) G.lv[k].out_off[pl] = ((k + 1) & 1) ? a1 : a0;
                    for (int k = 0; k < G.nlev; k++) G.lv[k].res_off[pl] = take((size_t)G.lv[k].res_pitch * std::max(res_h_[k], 1));
                }
                // ---- verification scratch for the levels whose chains can start inside the plane
                for (int k = 0; k < G.nlev; k++) {
                    Level &L = G.lv[k];
                    const bool multi = L.horizontal ? G.ntx > 1 : G.nty > 1;
                    for (int pl = 0; pl < kGP; pl++) { L.est[pl] = nullptr; L.act[pl] = nullptr; }
                    if (!multi) continue;
                    for (int pl = 0; pl < G.np; pl++) {
                        L.est[pl] = reinterpret_cast<int *>(salloc((size_t)G.ntx * G.nty * L.est_cap * sizeof(int)));
                        const size_t nact = L.horizontal ? (size_t)G.ntx * L.ho : (size_t)G.nty * L.wo;
                        L.act[pl] = reinterpret_cast<int16_t *>(salloc(nact * sizeof(int16_t)));
                        if (P.verify.nchecks >= kMaxChecks) return Plan();
                        Check &C = P.verify.chk[P.verify.nchecks++];
                        C.est = L.est[pl]; C.act = L.act[pl];
                        C.horizontal = L.horizontal; C.ntx = G.ntx; C.nty = G.nty; C.est_cap = L.est_cap;
                        C.cell = L.horizontal ? L.th : L.tw;
                        C.dim_across = L.horizontal ? L.ho : L.wo;
                        C.first_block = -1 - G.first_block;     // provisional: turned into first_block for the last launch below
                        P.need_verify = true;
                    }
                }
                smem_hw_max = std::max(smem_hw_max, smem_hw);
            }
            if (T.epilogue == kEpYCoCg && (T.ycc_gang[0] < 0 || T.ycc_gang[1] < 0 || T.ycc_gang[2] < 0)) return Plan();
            T.geom_off = (int)(((smem_hw_max * 2) + 15) & ~(size_t)15);
            PL.smem = (size_t)T.geom_off + sizeof(Geom) * kMaxGangs * kMaxLevels + sizeof(Region) * kMaxGangs + 16;
            if (PL.smem > opt.max_smem) return Plan();
            for (int gq = 0; gq < T.ngangs; gq++) {
                const Gang &G = T.g[gq];
                for (int pl = 0; pl < G.np; pl++) {
                    PL.bytes += 2.0 * G.w0 * G.h0 + 2.0 * G.W * G.H;
                    for (int k = 0; k < G.nlev; k++) if (G.lv[k].res[pl]) PL.bytes += 2.0 * G.lv[k].wr * G.lv[k].hr;
                }
            }
            PL.grid = join ? T.g[0].ntx * T.g[0].nty : block0;
            PL.threads = thread0;
            P.launches.push_back(PL);
        }
    }
    // ---- the serial fallback recomputes every step after the (exact) single-tile launch
    {
        const bool exact_first = coarse_end > 0;
        const int from_step = exact_first ? coarse_end : 0;
        for (int k = 0; k < n; k++) {
            if (ops[k].step < from_step) continue;
            if (P.verify.nops >= kMaxSerialOps) return Plan();
            SerialOp &so = P.verify.op[P.verify.nops++];
            so.avg = ops[k].avg; so.res = ops[k].res; so.out = ops[k].out;
            so.wa = ops[k].wa; so.wr = ops[k].wr; so.ha = ops[k].ha; so.hr = ops[k].hr;
            so.horizontal = ops[k].horizontal; so.step = ops[k].step;
        }
        P.verify.force = 0;
        P.verify.counters = nullptr;
        P.verify.epilogue = P.epilogue_fused ? ep.kind : kEpNone;
        P.verify.maxval = ep.maxval; P.verify.lo = ep.lo; P.verify.hi = ep.hi; P.verify.do_clamp = P.epilogue_fused ? ep.do_clamp : 0;
        P.verify.nother = 0;
        P.verify.W = P.verify.H = 0;
        for (int j = 0; j < 3; j++) P.verify.ycc[j] = nullptr;
        if (P.epilogue_fused) {
            const PlannedLaunch &PL = P.launches.back();
            P.verify.W = PL.task.g[0].W; P.verify.H = PL.task.g[0].H;
            for (int gi = 0; gi < PL.task.ngangs; gi++)
                for (int pl = 0; pl < PL.task.g[gi].np; pl++) {
                    int16_t *o = PL.task.g[gi].out[pl];
                    bool is_ycc = false;
                    if (ep.kind == kEpYCoCg)
                        for (int j = 0; j < 3; j++) if (o == ep.ycc[j]) { P.verify.ycc[j] = o; is_ycc = true; }
                    if (!is_ycc) { if (P.verify.nother >= 4) return Plan(); P.verify.other[P.verify.nother++] = o; }
                }
        }
    }
    // ---- checks of the very last launch are repairable tile by tile
    {
        const PlannedLaunch &PL = P.launches.back();
        std::vector<const int *> last_est;
        for (int gi = 0; gi < PL.task.ngangs; gi++)
            for (int k = 0; k < PL.task.g[gi].nlev; k++)
                for (int pl = 0; pl < PL.task.g[gi].np; pl++) if (PL.task.g[gi].lv[k].est[pl]) last_est.push_back(PL.task.g[gi].lv[k].est[pl]);
        bool any = false;
        for (int i = 0; i < P.verify.nchecks; i++) {
            Check &C = P.verify.chk[i];
            const bool is_last = std::find(last_est.begin(), last_est.end(), C.est) != last_est.end();
            C.first_block = is_last ? (-1 - C.first_block) : -1;
            any = any || is_last;
        }
        P.verify.top_blocks = PL.grid;
        P.verify.bad_cap = PL.grid;
        if (any) {
            P.tile_bad_off = salloc((size_t)PL.grid);
            P.bad_list_off = salloc((size_t)PL.grid * sizeof(int));
        } else P.verify.bad_cap = 0;
        P.verify_threads = PL.threads;
        P.verify_smem = PL.smem;
    }
    P.scratch_bytes = scratch;
    P.ok = true;
    return P;
}

}  // namespace fq
