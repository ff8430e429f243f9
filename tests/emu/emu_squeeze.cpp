// TEST INFRASTRUCTURE ONLY: runs the fused unsqueeze kernel source of the product (fb_fused_squeeze.cuh, planned by
// fb_fused_plan.h) under the CPU execution-model emulator (cuemu.h), so that the CPU-only test tier can compare it with
// the oracle.  Built by tests/emu_util.py with  g++ -DFB_EMULATE.
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "fb_fused_plan.h"
#include "fb_subsample.cuh"
#include "fb_approx.cuh"
#include "fb_palette.cuh"
#include "fb_match.cuh"

extern "C" {

// the palette gather: planes[0] holds the indices (and receives row 0 of the palette), planes[1..nb-1] the other rows
void emu_palette_inv(int16_t **planes, int nb, const int16_t *palette, int ncolors, long long n) {
    if (n <= 0) return;
    pl::Planes P;
    for (int c = 0; c < pl::kMaxPlanes; c++) P.p[c] = c < nb ? planes[c] : nullptr;
    cuemu::launch((unsigned)((n + 255) / 256), 256, 0, false, [&]() { pl::k_palette_inv(P, palette, (size_t)n, ncolors, nb); });
}

// fwd_palette as the library runs it: hash-set collect kernel, sort of the keys, index kernel.  planes[0..nb-1] hold the
// channels; on success planes[0] holds the indices and palette[c * count + k] the colours.  Returns the number of colours
// or -1 when there are more than `limit`.  table_cap (a power of two) lets a test force collisions / a full table.
int emu_palette_fwd(int16_t **planes, int nb, long long n, int limit, int16_t *palette, int table_cap) {
    std::vector<unsigned long long> table((size_t)table_cap, pl::kEmptySlot);
    int ctr[4] = {0, 0, 0, 0};
    pl::Planes P;
    for (int c = 0; c < pl::kMaxPlanes; c++) P.p[c] = c < nb ? planes[c] : nullptr;
    pl::Collect C;
    C.table = table.data(); C.cap_mask = (unsigned)(table_cap - 1); C.limit = limit; C.count = ctr; C.has_allones = ctr + 1; C.overflow = ctr + 2;
    const unsigned nblk = (unsigned)((n + 255) / 256);
    if (n > 0) cuemu::launch(nblk, 256, 0, false, [&]() { pl::k_palette_collect(P, (size_t)n, nb, C); });
    if (ctr[2] || ctr[0] > limit) return -1;
    std::vector<unsigned long long> sorted;
    for (unsigned long long k : table) if (k != pl::kEmptySlot) sorted.push_back(k);
    if (ctr[1]) sorted.push_back(pl::kEmptySlot);
    std::sort(sorted.begin(), sorted.end());
    const int count = (int)sorted.size();
    for (int k = 0; k < count; k++) for (int c = 0; c < nb; c++) palette[(size_t)c * count + k] = (int16_t)pl::unpack_colour(sorted[(size_t)k], c);
    if (n > 0) cuemu::launch(nblk, 256, 0, false, [&]() { pl::k_palette_index(P, (size_t)n, nb, sorted.data(), count); });
    return count;
}

// inv_match as the library runs it: parents, log2(n) rounds of pointer jumping (double-buffered), one out-of-place gather per
// channel.  planes[c] are overwritten with the result.  Returns 1 when a match code is out of range.
int emu_match_inv(const int16_t *m, int16_t **planes, int nc, int w, int h, int maxcode, const int *zero) {
    const int n = w * h;
    if (n <= 0) return 0;
    std::vector<int> a((size_t)n), b((size_t)n);
    int bad = 0;
    const unsigned nblk = (unsigned)((n + 255) / 256);
    cuemu::launch(nblk, 256, 0, false, [&]() { mt::k_match_parent(m, a.data(), n, w, maxcode, &bad); });
    int rounds = 1;
    while ((1ll << rounds) < n) rounds++;
    for (int done = 0; done < rounds;) {            // batches of three rounds, as fb_match_resolve runs them
        int changed = 0;
        for (int r = 0; r < 3 && done < rounds; r++, done++) {
            cuemu::launch(nblk, 256, 0, false, [&]() { mt::k_match_jump(a.data(), b.data(), n, &changed); });
            a.swap(b);
        }
        if (!changed) break;
    }
    if (bad) return 1;
    for (int c = 0; c < nc; c++) {
        std::vector<int16_t> out((size_t)n);
        cuemu::launch(nblk, 256, 0, false, [&]() { mt::k_match_gather(planes[c], out.data(), a.data(), n, zero[c]); });
        memcpy(planes[c], out.data(), (size_t)n * sizeof(int16_t));
    }
    return 0;
}

// inv_match with soft matches on one channel, as fb_match_soft runs it; plane is overwritten.  Returns 1 on a bad code.
int emu_match_soft(const int16_t *m, int16_t *plane, int w, int h, int maxcode, int zero) {
    const int n = w * h;
    if (n <= 0) return 0;
    std::vector<int> pa((size_t)n), pb((size_t)n);
    std::vector<int16_t> aa((size_t)n), ab((size_t)n), out((size_t)n);
    int bad = 0;
    const unsigned nblk = (unsigned)((n + 255) / 256);
    cuemu::launch(nblk, 256, 0, false, [&]() { mt::k_match_soft_init(m, plane, pa.data(), aa.data(), n, w, maxcode, zero, &bad); });
    int rounds = 1;
    while ((1ll << rounds) < n) rounds++;
    for (int done = 0; done < rounds;) {
        int changed = 0;
        for (int r = 0; r < 3 && done < rounds; r++, done++) {
            cuemu::launch(nblk, 256, 0, false, [&]() { mt::k_match_soft_jump(pa.data(), aa.data(), pb.data(), ab.data(), n, &changed); });
            pa.swap(pb); aa.swap(ab);
        }
        if (!changed) break;
    }
    cuemu::launch(nblk, 256, 0, false, [&]() { mt::k_match_soft_finish(pa.data(), aa.data(), out.data(), n); });
    memcpy(plane, out.data(), (size_t)n * sizeof(int16_t));
    return bad;
}

// the Approximate kernels on one channel (+ its remainder channel; chr may be NULL for the inverse)
void emu_approximate(int16_t *ch, int16_t *chr, long long n, int q, int inverse) {
    if (n <= 0) return;
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (inverse) cuemu::launch(nb, 256, 0, false, [&]() { ap::k_approx_inv(ch, chr, (size_t)n, q); });
    else cuemu::launch(nb, 256, 0, false, [&]() { ap::k_approx_fwd(ch, chr, (size_t)n, q); });
}

// the chroma upscaling kernel on one plane: in (ow x oh) -> out (ow*srh x oh*srv)
void emu_inv_subsample(const int16_t *in, int16_t *out, int ow, int oh, int srh, int srv) {
    const size_t n = (size_t)ow * srh * (size_t)oh * srv;
    if (!n) return;
    cuemu::launch((unsigned)((n + 255) / 256), 256, 0, false, [&]() { sb::k_inv_subsample(in, out, ow, oh, srh, srv); });
}

// opdesc[nops][9] = step, horizontal, avg plane, res plane (-1: none), out plane, wa, wr, ha, hr
// ep[8] = kind, maxval, lo, hi, do_clamp, Y plane, Co plane, Cg plane
// opts[8] = tile_w, tile_h, levels_per_launch, coarse_dim, threads_per_gang, force (1 serial fallback, 2 repair all), warm_last, warm_mid
// stats[9] (out) = plan ok, launches, serial fallback ran, checks, failed comparisons, epilogue fused, max smem, total CTAs, tiles repaired
int emu_run_plan(int nplanes, int16_t **planes, int nops, const int *opdesc, const int *ep, const int *opts, int *stats) {
    (void)nplanes;
    std::vector<fq::PlanOp> ops(nops);
    for (int i = 0; i < nops; i++) {
        const int *d = opdesc + 9 * i;
        ops[i].step = d[0]; ops[i].horizontal = d[1];
        ops[i].avg = planes[d[2]]; ops[i].res = d[3] >= 0 ? planes[d[3]] : nullptr; ops[i].out = planes[d[4]];
        ops[i].wa = d[5]; ops[i].wr = d[6]; ops[i].ha = d[7]; ops[i].hr = d[8];
    }
    fq::EpilogueSpec E;
    E.kind = ep[0]; E.maxval = ep[1]; E.lo = ep[2]; E.hi = ep[3]; E.do_clamp = ep[4];
    for (int j = 0; j < 3; j++) E.ycc[j] = ep[5 + j] >= 0 ? planes[ep[5 + j]] : nullptr;
    fq::PlanOptions O;
    O.tile_w = opts[0]; O.tile_h = opts[1]; O.levels_per_launch = opts[2]; O.coarse_dim = opts[3]; O.threads_per_gang = opts[4];
    if (opts[6] > 0) O.warm_last = opts[6];
    if (opts[7] > 0) O.warm_mid = opts[7];
    fq::Plan P = fq::make_plan(ops, E, O);
    memset(stats, 0, 9 * sizeof(int));
    stats[0] = P.ok;
    if (!P.ok) return 1;
    std::vector<unsigned char> scratch(P.scratch_bytes + 256, 0xEE);
    int counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    fq::relocate_scratch(P, scratch.data(), counters);
    P.verify.force = opts[5];
    stats[1] = (int)P.launches.size();
    stats[5] = P.epilogue_fused;
    for (auto &L : P.launches) {
        const fq::Task T = L.task;
        stats[6] = std::max(stats[6], (int)L.smem);
        stats[7] += L.grid;
        cuemu::launch((unsigned)L.grid, (unsigned)L.threads, L.smem, false, [&]() { fq::k_fq_tiles(T); });
    }
    stats[3] = P.verify.nchecks;
    // count failed comparisons (diagnostics: the speculation failure rate), then run the real verify / fallback kernel
    for (int ci = 0; ci < P.verify.nchecks; ci++) {
        const fq::Check &C = P.verify.chk[ci];
        for (int tile = 0; tile < C.ntx * C.nty; tile++) {
            const int ti = tile / C.nty, tj = tile % C.nty;
            const int along = C.horizontal ? ti : tj, across = C.horizontal ? tj : ti;
            if (!along) continue;
            for (int e = 0; e < C.est_cap; e++) {
                const int v = C.est[(size_t)tile * C.est_cap + e];
                if (v == fq::kNoCheck) continue;
                if (C.act[(size_t)(along - 1) * C.dim_across + across * C.cell + e] != v) stats[4]++;
            }
        }
    }
    if (P.need_verify || P.verify.force) {
        const fq::VerifyParams V = P.verify;
        cuemu::launch(3, (unsigned)P.verify_threads, P.verify_smem, true, [&]() { fq::k_fq_verify_fallback(V); });
    }
    stats[2] = counters[4];
    stats[8] = counters[5];
    return 0;
}

// closed-form pair vs the reference's literal formulation on n random full-range inputs; returns mismatches
int emu_check_pair(const int16_t *prev, const int16_t *av, const int16_t *nx, const int16_t *rs, int n) {
    int bad = 0;
    for (int i = 0; i < n; i++) {
        int A, B, A2, B2;
        fq::unsqueeze_pair(prev[i], av[i], nx[i], rs[i], A, B);
        fq::unsqueeze_pair_literal(prev[i], av[i], nx[i], rs[i], A2, B2);
        if (A != A2 || B != B2) bad++;
    }
    return bad;
}
}

// ---------------------------------------------------------------------------------------------------------
// direct per-step kernels (fb_direct_squeeze.cuh + fb_direct_plan.h)
// ---------------------------------------------------------------------------------------------------------
#include "fb_direct_plan.h"

extern "C" {
// opdesc[nops][10] = step, horizontal, avg plane, res plane (-1), out plane, wa, wr, ha, hr, clamp
// ep[9] = enabled, Y plane, R-out plane, Co-out plane, Cg-out plane, maxval, lo, hi, do_clamp
// stats[4] (out) = direct launches, ops handled by the direct kernels, ops left to the serial code, epilogue done
int emu_run_direct(int nplanes, int16_t **planes, int nops, const int *opdesc, const int *ep, int lo, int hi, int *stats) {
    (void)nplanes;
    memset(stats, 0, 4 * sizeof(int));
    int i0 = 0;
    while (i0 < nops) {
        int i1 = i0;
        while (i1 < nops && opdesc[10 * i1] == opdesc[10 * i0]) i1++;
        const bool horizontal = opdesc[10 * i0 + 1] != 0;
        std::vector<dq::StepOp> ops;
        for (int i = i0; i < i1; i++) {
            const int *d = opdesc + 10 * i;
            dq::StepOp o;
            o.avg = planes[d[2]]; o.res = d[3] >= 0 ? planes[d[3]] : nullptr; o.out = planes[d[4]];
            o.wa = d[5]; o.wr = d[6]; o.ha = d[7]; o.hr = d[8]; o.clamp = d[9];
            ops.push_back(o);
        }
        dq::StepEpilogue E;
        if (ep[0] && i1 == nops) {
            E.enabled = 1; E.yin = planes[ep[1]]; E.rout = planes[ep[2]]; E.co_out = planes[ep[3]]; E.cg_out = planes[ep[4]];
            E.maxval = ep[5]; E.lo = ep[6]; E.hi = ep[7]; E.do_clamp = ep[8];
        }
        dq::StepPlan P = dq::plan_step(ops, horizontal, E, lo, hi, 4);
        if (P.hj.n) {
            const dq::HJobs J = P.hj;
            cuemu::launch((unsigned)P.h_grid, (unsigned)P.h_threads, P.h_smem, false, [&]() { dq::k_inv_hsq_direct(J); });
            stats[0]++;
            for (int j = 0; j < J.n; j++) stats[1] += J.j[j].np;
        }
        if (P.vj.n) {
            const dq::VJobs J = P.vj;
            cuemu::launch((unsigned)P.v_grid, (unsigned)P.v_threads, P.v_smem, false, [&]() { dq::k_inv_vsq_direct(J); });
            stats[0]++;
            stats[1] += J.n;
        }
        for (int k : P.leftover) {
            const dq::StepOp &o = ops[k];
            fq::SerialOp so;
            so.avg = o.avg; so.res = o.res; so.out = o.out; so.wa = o.wa; so.wr = o.wr; so.ha = o.ha; so.hr = o.hr;
            so.horizontal = horizontal; so.step = 0;
            const int nchain = horizontal ? o.ha : o.wa;
            for (int c = 0; c < nchain; c++) fq::serial_chain(so, c);
            if (o.clamp) {
                const size_t nn = (size_t)(horizontal ? o.wa + o.wr : o.wa) * (horizontal ? o.ha : o.ha + o.hr);
                for (size_t q = 0; q < nn; q++) o.out[q] = (int16_t)fq::clampi(o.out[q], lo, hi);
            }
            stats[2]++;
        }
        if (P.epilogue_done) stats[3] = 1;
        i0 = i1;
    }
    return 0;
}
}

// ---------------------------------------------------------------------------------------------------------
// packed per-step kernels (fb_pk_squeeze.cuh + fb_pk_plan.h): TMA / mbarrier replaced by synchronous copies
// ---------------------------------------------------------------------------------------------------------
#include <type_traits>
#include "fb_pk_plan.h"

namespace ps {
bool ps_make_tilemap(TileMap *m, const void *base, int w, int h, int box_w, int box_h, int swizzle_bytes) {
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (w & 7)) return false;
    m->base = (int16_t *)base; m->w = w; m->h = h; m->box_w = box_w; m->box_h = box_h; m->swz = swizzle_bytes;
    return true;
}
}  // namespace ps

extern "C" {
// same descriptors as emu_run_direct.  sm_count shapes the segment lengths.
// stats[6] (out) = packed launches, ops handled by the packed kernels, ops left to the serial code, epilogue done, repaired segments, range-flagged segments
int emu_run_pk(int nplanes, int16_t **planes, int nops, const int *opdesc, const int *ep, int lo, int hi, int sm_count, int *stats) {
    (void)nplanes;
    memset(stats, 0, 6 * sizeof(int));
    int dev_stats[2] = {0, 0};
    int i0 = 0;
    while (i0 < nops) {
        int i1 = i0;
        while (i1 < nops && opdesc[10 * i1] == opdesc[10 * i0]) i1++;
        const bool horizontal = opdesc[10 * i0 + 1] != 0;
        std::vector<ps::StepOp> ops;
        for (int i = i0; i < i1; i++) {
            const int *d = opdesc + 10 * i;
            ps::StepOp o;
            o.avg = planes[d[2]]; o.res = d[3] >= 0 ? planes[d[3]] : nullptr; o.out = planes[d[4]];
            o.wa = d[5]; o.wr = d[6]; o.ha = d[7]; o.hr = d[8]; o.clamp = d[9];
            ops.push_back(o);
        }
        ps::StepEpilogue E;
        if (ep[0] && i1 == nops) {
            E.enabled = 1; E.yin = planes[ep[1]]; E.rout = planes[ep[2]]; E.co_out = planes[ep[3]]; E.cg_out = planes[ep[4]];
            E.maxval = ep[5]; E.lo = ep[6]; E.hi = ep[7]; E.do_clamp = ep[8];
        }
        ps::StepPlan P = ps::plan_step(ops, horizontal, E, lo, hi, sm_count, dev_stats);
        std::vector<unsigned char> scratch(P.scratch_bytes + 64, 0xEE);
        std::vector<int> counters((size_t)P.counters + 1, 0);
        ps::relocate(P, scratch.data(), counters.data());
        for (auto &L : P.h) {
            const ps::HJobs J = L.jobs;
            const int wpb = L.warps_per_block, spw = L.smem_per_warp;
            auto run = [&](auto kern) { cuemu::launch((unsigned)L.grid, (unsigned)(32 * wpb), L.smem, false, kern); };
            if (L.ep == fq::kEpYCoCg) run([&]() { ps::k_pk_hsq<2, fq::kEpYCoCg>(J, wpb, spw); });
            else if (L.np == 2 && L.ep == fq::kEpClamp) run([&]() { ps::k_pk_hsq<2, fq::kEpClamp>(J, wpb, spw); });
            else if (L.np == 2) run([&]() { ps::k_pk_hsq<2, fq::kEpNone>(J, wpb, spw); });
            else if (L.ep == fq::kEpClamp) run([&]() { ps::k_pk_hsq<1, fq::kEpClamp>(J, wpb, spw); });
            else run([&]() { ps::k_pk_hsq<1, fq::kEpNone>(J, wpb, spw); });
            stats[0]++;
            for (int j = 0; j < J.n; j++) stats[1] += J.j[j].np;
        }
        for (auto &L : P.v) {
            const ps::VJobs J = L.jobs;
            const int wpb = L.warps_per_block;
            cuemu::launch((unsigned)L.grid, (unsigned)(32 * wpb), 0, false, [&]() { ps::k_pk_vsq<ps::kVDepthDefault>(J, wpb); });
            stats[0]++;
            stats[1] += J.n;
        }
        for (int c : counters) if (c != 0) return 2;        // every arrival counter must be back at zero
        for (int k : P.leftover) {
            const ps::StepOp &o = ops[k];
            fq::SerialOp so;
            so.avg = o.avg; so.res = o.res; so.out = o.out; so.wa = o.wa; so.wr = o.wr; so.ha = o.ha; so.hr = o.hr;
            so.horizontal = horizontal; so.step = 0;
            const int nchain = horizontal ? o.ha : o.wa;
            for (int c = 0; c < nchain; c++) fq::serial_chain(so, c);
            if (o.clamp) {
                const size_t nn = (size_t)(horizontal ? o.wa + o.wr : o.wa) * (horizontal ? o.ha : o.ha + o.hr);
                for (size_t q = 0; q < nn; q++) o.out[q] = (int16_t)fq::clampi(o.out[q], lo, hi);
            }
            stats[2]++;
        }
        if (P.epilogue_done) stats[3] = 1;
        i0 = i1;
    }
    stats[4] = dev_stats[0];
    stats[5] = dev_stats[1];
    return 0;
}

// the packed pair against the exact 32-bit pair on n inputs; returns mismatches
int emu_check_pk_pair(const int16_t *prev, const int16_t *av, const int16_t *nx, const int16_t *rs, int n) {
    int bad = 0;
    for (int i = 0; i + 1 < n; i += 2) {
        const uint32_t P = ps::e_h(prev[i], prev[i + 1]), a = ps::e_h(av[i], av[i + 1]), nn = ps::e_h(nx[i], nx[i + 1]), r = ps::e_h(rs[i], rs[i + 1]);
        uint32_t A, B;
        const ps::PK K = ps::pk_consts(0x00010001u);
        ps::pk_step(K, P, a, ps::pneg(a, K), ps::pneg(nn, K), r, A, B);
        for (int k = 0; k < 2; k++) {
            int A2, B2;
            fq::unsqueeze_pair(prev[i + k], av[i + k], nx[i + k], rs[i + k], A2, B2);
            const int Ag = k ? ps::e_hi(A) : ps::e_lo(A), Bg = k ? ps::e_hi(B) : ps::e_lo(B);
            if (Ag != A2 || Bg != B2) bad++;
        }
    }
    return bad;
}

// packed inverse YCoCg against the exact one; returns mismatches
int emu_check_pk_ycocg(const int16_t *y, const int16_t *co, const int16_t *cg, int n, int maxval) {
    int bad = 0;
    const uint32_t mv = (uint32_t)(uint16_t)maxval * 0x00010001u;
    for (int i = 0; i + 1 < n; i += 2) {
        uint32_t R, G, B;
        ps::pk_ycocg(ps::pk_consts(0x00010001u), ps::e_h(y[i], y[i + 1]), ps::e_h(co[i], co[i + 1]), ps::e_h(cg[i], cg[i + 1]), mv, R, G, B);
        for (int k = 0; k < 2; k++) {
            int R2, G2, B2;
            ps::ycocg_exact(y[i + k], co[i + k], cg[i + k], maxval, 0, 0, 0, R2, G2, B2);
            const int Rg = k ? ps::e_hi(R) : ps::e_lo(R), Gg = k ? ps::e_hi(G) : ps::e_lo(G), Bg = k ? ps::e_hi(B) : ps::e_lo(B);
            if (Rg != R2 || Gg != G2 || Bg != B2) bad++;
        }
    }
    return bad;
}
}
