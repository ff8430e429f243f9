// MANIAC entropy ENCODING on the GPU (sm_100a): channel groups are independent when encoding (every plane is known up front),
// so one warp encodes one group: the learning pass, the pruning, the tree, and the real pass.
//
// What runs here is fuif_encode_channels (reference encoding/encoding.cpp:74-207) with everything it inlines:
//   group header varints                           encoding.cpp:76-124
//   init_properties / predictors / properties      encoding/context_predict.h:67-206, 233-289
//   PropertySymbolCoder (tree learning, simplify)   maniac/compound_enc.h:243-518
//   CompoundSymbolBitCoder::updateChances           maniac/compound_enc.h:91-109 (virtual chances + cost estimates)
//   MetaPropertySymbolCoder::write_tree             maniac/compound_enc.h:523-552
//   writer<>, UniformSymbolCoder::write_int         maniac/symbol_enc.h:28-109
//   RacOutput                                       maniac/rac_enc.h:28-100
// The container (magic, transform list, responsive offsets, the "larger than uncompressed -> roll back" rule of
// encoding.cpp:455-573) is assembled by the host from the per-group byte strings.
//
// Mapping.  Lane p of the warp owns property p of the context vector (at most 31 properties; lane 31 owns the leaf's real
// chances).  The learning pass is where the reference spends 3/4 of its encode time, and it is data-parallel over the
// properties: for every binary decision of every learned symbol each property has a virtual chance to adapt and a virtual
// cost to accumulate -- one lane each -- and the cheapest virtual context is an arg-min over the lanes.  The decisions of
// one symbol touch distinct chances, so they need no ordering among themselves.  The rows the reference learns from are
// drawn with libc rand(): the host tabulates that sequence (it depends only on the plane heights) and every group starts
// at its own offset.
//
// STATUS (round 1): validated against the oracle encoder (itself byte-exact against the reference's files) under the CPU
// execution-model emulator together with the host side (fb_encode_host.h), tests/test_emu_maniac_enc.py: whole files are
// identical.  fb_encode() below is the C-ABI entry; it was written after the round's GPU time had run out, so its first run
// on a GPU is the round-end test tier (tests/test_zz_gpu_encode.py).
#ifdef FB_EMULATE
#include "maniac_emu_shim.h"
#else
#include "fb_common.cuh"
#endif

#include <stdlib.h>
#include <string.h>

namespace fbenc {

constexpr int kMaxNodes = 65536;
constexpr int kNonRef = 13;                     // context_predict.h:210
constexpr int kSplitThreshold = 5461 * 8 * 2;   // CONTEXT_TREE_SPLIT_THRESHOLD, config.h
constexpr int kMinSubtree = 10;                 // CONTEXT_TREE_MIN_SUBTREE_SIZE, config.h

struct EChan {                  // a plane as the encoder sees it (ranges tight: fuif_prepare_encode, encoding.cpp:737-743)
    int w, h, minval, maxval, zero, q, hshift, vshift;
    const int16_t *data;
};
struct TNode { short property; unsigned short child; int splitval; };   // PropertyDecisionNode, compound.h:41-51
// leaf of the learning pass: CompoundSymbolChances, compound_enc.h:29-59.  virt[s][i][p] is chance i of property p's virtual
// context s (s = 0 while property p > its running split value): property-minor, so that the 32 lanes of a decision touch
// at most two 64-byte runs instead of 32 cache lines
struct LLeaf {
    uint16_t real[32];
    uint16_t virt[2][32][32];
    unsigned long long realSize;
    unsigned long long virtSize[32];
    long long virtPropSum[32];
    int count;
    int best;
};
struct EGroup {
    int beginc, endc, predictor, compress;
    long long rand_off;         // index of this group's first rand() value
    // working memory
    TNode *nodes;               // kMaxNodes
    LLeaf *leaves; int leaf_cap;
    uint16_t *fleaves;          // (kMaxNodes / 2) x 32: leaves of the real pass
    int *stack;                 // 8 * (kMaxNodes / 2 + 2) ints: explicit stacks of write_tree / simplify / kill_children
    // output
    unsigned char *out; unsigned out_cap;
    unsigned out_len, header_len;       // header_len: bytes before the entropy-coded part (header_pos - before, encoding.cpp)
    unsigned attempt_len;               // length of the compressed form if it was rolled back (its tail stays in `out`), else 0
    int status;                         // 0 ok, 1 output buffer too small, 2 leaf pool exhausted, 3 bit depth
    int nnodes;
};
struct EParams {
    const EChan *ch; int nch;
    EGroup *groups; int ngroups;
    int max_properties;
    float nb_repeats;
    const uint16_t *table, *meta_table;     // newchance[4096][2]: (cutoff, alpha) and the tree coder's (2, 0xFFFFFFFF/19)
    const uint16_t *log4k;                  // [4097], chance.cpp:67-91
    const int *rnd; long long nrnd;         // libc rand() sequence from its initial state
};

// 5 x 32: per-lane values of the current symbol, exchanged through shared memory (one warp barrier instead of a shuffle per
// tree level): property, range lo, range hi, split value, virtual cost.  One warp per block, so one copy per block.
__shared__ long long s_scr[160];
// the chance successor table of the sample coder and the cost table, copied in at kernel start: every decision of every lane
// looks both up at a data-dependent index
__shared__ uint16_t s_table[4096 * 2];
__shared__ uint16_t s_log4k[4104];

__device__ __forceinline__ int s16(int x) { return (int)(short)x; }
__device__ __forceinline__ int ilog2u(unsigned l) { return l == 0 ? 0 : 31 - __clz((int)l); }
__device__ __forceinline__ int slog(int x16) {      // context_predict.h:54-61
    const int x = s16(x16);
    const int b = 32 - __clz(abs(x));
    return x < 0 ? -b : b;
}
__device__ __forceinline__ int fooabs(int x16) { int x = s16(x16); return s16(x < 0 ? -x : x); }
__device__ __forceinline__ int median3(int a, int b, int c) {
    if (a < b) { if (b < c) return b; return a < c ? c : a; }
    if (a < c) return a;
    return b < c ? c : b;
}

// ---- byte sink + range encoder (lane 0) ---------------------------------------------------------------------------------
struct Sink { unsigned char *p; unsigned cap, len; int overflow; };
__device__ __forceinline__ void sink_put(Sink &s, int c) { if (s.len < s.cap) s.p[s.len] = (unsigned char)c; else s.overflow = 1; s.len++; }
__device__ void sink_varint(Sink &s, unsigned long long number) {        // write_big_endian_varint, encoding.cpp:30-41
    unsigned char tmp[12];
    int n = 0;
    tmp[n++] = (unsigned char)(number & 127);
    number >>= 7;
    while (number) { tmp[n++] = (unsigned char)((number & 127) | 128); number >>= 7; }
    while (n) sink_put(s, tmp[--n]);
}
struct RacOut { unsigned long long range, low; int delayed_byte, delayed_count; };      // rac_enc.h:28-100 (rac_t is 64 bits wide there)
__device__ __forceinline__ void rac_init(RacOut &r) { r.range = 1u << 24; r.low = 0; r.delayed_byte = -1; r.delayed_count = 0; }
__device__ void rac_output(RacOut &r, Sink &s) {
    while (r.range <= (1u << 16)) {
        const int byte = (int)(r.low >> 16);
        if (r.delayed_byte < 0) r.delayed_byte = byte;
        else if (((r.low + r.range) >> 8) < (1u << 16)) {
            sink_put(s, r.delayed_byte);
            while (r.delayed_count) { sink_put(s, 0xFF); r.delayed_count--; }
            r.delayed_byte = byte;
        } else if ((r.low >> 8) >= (1u << 16)) {
            sink_put(s, r.delayed_byte + 1);
            while (r.delayed_count) { sink_put(s, 0); r.delayed_count--; }
            r.delayed_byte = byte & 0xFF;
        } else r.delayed_count++;
        r.low = (r.low & ((1u << 16) - 1)) << 8;
        r.range <<= 8;
    }
}
__device__ __forceinline__ void rac_put(RacOut &r, Sink &s, unsigned long long chance, int bit) {
    if (bit) { r.low += r.range - chance; r.range = chance; } else r.range -= chance;
    rac_output(r, s);
}
__device__ __forceinline__ void rac_write12(RacOut &r, Sink &s, unsigned b12, int bit) { rac_put(r, s, (r.range * (unsigned long long)b12 + 0x800) >> 12, bit); }
__device__ __forceinline__ void rac_write_bit(RacOut &r, Sink &s, int bit) { rac_put(r, s, r.range >> 1, bit); }
__device__ void rac_flush(RacOut &r, Sink &s) {
    r.low += (1u << 16) - 1;
    for (int k = 0; k < 4; k++) { r.range = (1u << 16) - 1; rac_output(r, s); }
}

// ---- symbol model -------------------------------------------------------------------------------------------------------
#define SC_ZERO 0
#define SC_SIGN 1
#define SC_EXP 2
#define SC_MANT 16
__device__ __forceinline__ uint16_t initial_chance(int idx, int zero_chance) {      // SymbolChance(zero_chance), symbol.h:115-138
    if (idx == SC_ZERO) return (uint16_t)zero_chance;
    if (idx == SC_SIGN) return 0x800;
    if (idx >= SC_MANT) return idx == 31 ? 0 : 1024;
    unsigned long long rp = 0x1000 - (unsigned long long)zero_chance;
    for (int i = 0;; i++) {
        if (rp < 0x100) rp = 0x100;
        if (rp > 0xf00) rp = 0xf00;
        if (i == idx - SC_EXP) return (uint16_t)(0x1000 - rp);
        rp = (rp * rp + 0x800) >> 12;
    }
}
// The binary decisions writer<15>(coder, min, max, value) makes (symbol_enc.h:58-109), as a list of (chance index << 1 | bit).
__device__ int symbol_decisions(int min, int max, int value, unsigned char *dec) {
    int n = 0;
    if (min == max) return 0;
    if (value == 0) { dec[n++] = (SC_ZERO << 1) | 1; return n; }
    dec[n++] = (SC_ZERO << 1) | 0;
    const int sign = value > 0 ? 1 : 0;
    if (max > 0 && min < 0) dec[n++] = (unsigned char)((SC_SIGN << 1) | sign);
    const int a = abs(value);
    const int e = ilog2u((unsigned)a);
    const int amax = sign ? abs(max) : abs(min);
    const int emax = ilog2u((unsigned)amax);
    int i = 0;
    while (i < emax) {
        if ((1 << (i + 1)) > amax) break;
        dec[n++] = (unsigned char)(((SC_EXP + i) << 1) | (i == e ? 1 : 0));
        if (i == e) break;
        i++;
    }
    int have = 1 << e;
    for (int pos = e; pos > 0;) {
        int bit = 1;
        --pos;
        const int minabs1 = have | (1 << pos);
        if (minabs1 > amax) bit = 0;
        else { bit = (a >> pos) & 1; dec[n++] = (unsigned char)(((SC_MANT + pos) << 1) | bit); }
        have |= bit << pos;
    }
    return n;
}
__device__ int symbol_decisions2(int min, int max, int value, unsigned char *dec) {      // write_int2, symbol.h:223-227
    if (min > 0) return symbol_decisions(0, max - min, value - min, dec);
    if (max < 0) return symbol_decisions(min - max, 0, value - max, dec);
    return symbol_decisions(min, max, value, dec);
}
// lane 0: codes the decisions with the chances `c` (SimpleSymbolBitCoder / FinalCompoundSymbolBitCoder::write)
__device__ void code_decisions(RacOut &rac, Sink &s, const uint16_t *__restrict__ table, uint16_t *c, const unsigned char *dec, int n) {
    for (int k = 0; k < n; k++) {
        const int idx = dec[k] >> 1, bit = dec[k] & 1;
        rac_write12(rac, s, c[idx], bit);
        c[idx] = table[c[idx] * 2 + bit];
    }
}

// ---- context ------------------------------------------------------------------------------------------------------------
struct GroupCtx {
    int nprops, nref, nrefchan;
    int refchan[16];
    int lo, hi;                 // this lane's property range (init_properties)
};
// init_properties, context_predict.h:67-120: every lane computes the whole table and keeps its own row
__device__ void init_properties(const EParams &P, int beginc, int endc, GroupCtx &G, int lane) {
    int pr[64][2];
    int n = 0, offset = 0;
    G.nrefchan = 0;
    for (int j = beginc - 1; j >= 0 && offset < P.max_properties; j--) {
        const EChan &cj = P.ch[j];
        if (cj.minval == cj.maxval) continue;
        if (cj.hshift < 0) continue;
        int minval = cj.minval; if (minval > 0) minval = 0;
        int maxval = cj.maxval; if (maxval < 0) maxval = 0;
        pr[n][0] = 0; pr[n][1] = fooabs(maxval > -minval ? maxval : minval); n++; offset++;
        pr[n][0] = slog(minval); pr[n][1] = slog(maxval); n++; offset++;
        if (G.nrefchan < 16) G.refchan[G.nrefchan] = j;
        G.nrefchan++;
    }
    int minval = 0x7FFF, maxval = -0x7FFF, maxh = 0, maxw = 0;
    for (int j = beginc; j <= endc; j++) {
        const EChan &cj = P.ch[j];
        if (cj.minval < minval) minval = cj.minval;
        if (cj.maxval > maxval) maxval = cj.maxval;
        if (cj.h > maxh) maxh = cj.h;
        if (cj.w > maxw) maxw = cj.w;
    }
    if (minval > 0) minval = 0;
    if (maxval < 0) maxval = 0;
    const int amax = max(fooabs(minval), fooabs(maxval));
    pr[n][0] = 0; pr[n][1] = amax; n++;
    pr[n][0] = 0; pr[n][1] = amax; n++;
    pr[n][0] = slog(minval); pr[n][1] = slog(maxval); n++;
    pr[n][0] = slog(minval); pr[n][1] = slog(maxval); n++;
    pr[n][0] = 0; pr[n][1] = maxh - 1; n++;
    pr[n][0] = 0; pr[n][1] = maxw - 1; n++;
    pr[n][0] = minval + minval - maxval; pr[n][1] = maxval + maxval - minval; n++;
    pr[n][0] = minval + minval - maxval; pr[n][1] = maxval + maxval - minval; n++;
    for (int k = 0; k < 5; k++) { pr[n][0] = slog(minval - maxval); pr[n][1] = slog(maxval - minval); n++; }
    G.nprops = n;
    G.nref = n - kNonRef;
    G.lo = lane < n ? pr[lane][0] : 0;
    G.hi = lane < n ? pr[lane][1] : 0;
}
// This lane's property of pixel (x, y) and the predictor's guess (predict_and_compute_properties + precompute_references,
// context_predict.h:125-168, 233-289).  Every lane reads the same few neighbours; lane p keeps property p.
__device__ __forceinline__ int property_and_guess(const EParams &P, const EChan &ch, const GroupCtx &G, int x, int y, int predictor, int lane, int &guess) {
    const int16_t *d = ch.data;
    const int w = ch.w;
    const int left = x ? d[(size_t)y * w + x - 1] : ch.zero;
    const int top = y ? d[(size_t)(y - 1) * w + x] : ch.zero;
    const int topleft = (x && y) ? d[(size_t)(y - 1) * w + x - 1] : left;
    const int topright = (x + 1 < w && y) ? d[(size_t)(y - 1) * w + x + 1] : top;
    const int leftleft = x > 1 ? d[(size_t)y * w + x - 2] : left;
    const int toptop = y > 1 ? d[(size_t)(y - 2) * w + x] : top;
    switch (predictor) {
    case 0: guess = ch.zero; break;
    case 1: guess = s16((left + top) / 2); break;
    case 3: guess = left; break;
    case 4: guess = top; break;
    case 5: guess = s16((left + topleft + top + topright) / 4); break;
    case 6: { const int g = left + top - topleft; guess = s16(g < ch.minval ? ch.minval : (g > ch.maxval ? ch.maxval : g)); break; }
    default: guess = median3(s16(left + top - topleft), left, top); break;
    }
    if (lane < G.nref) {
        const EChan &cj = P.ch[G.refchan[lane >> 1]];
        int ry = (y << ch.vshift) >> cj.vshift;
        if (ry >= cj.h) ry = cj.h - 1;
        int rx;
        if (ch.hshift == cj.hshift && w <= cj.w) rx = x;
        else if (ch.hshift < cj.hshift) {
            const int stepsize = (1 << cj.hshift) >> ch.hshift;
            rx = stepsize > 0 ? x / stepsize : cj.w - 1;
            if (rx > cj.w - 1) rx = cj.w - 1;
        } else {
            rx = (x << ch.hshift) >> cj.hshift;
            if (rx >= cj.w) rx = cj.w - 1;
        }
        const int v = cj.data[(size_t)ry * cj.w + rx];
        return (lane & 1) ? slog(v) : fooabs(v);
    }
    switch (lane - G.nref) {
    case 0: return fooabs(top);
    case 1: return fooabs(left);
    case 2: return slog(top);
    case 3: return slog(left);
    case 4: return y;
    case 5: return x;
    case 6: return left + top - topleft;
    case 7: return topleft + topright - top;
    case 8: return slog(left - topleft);
    case 9: return slog(topleft - top);
    case 10: return slog(top - topright);
    case 11: return slog(top - toptop);
    case 12: return slog(left - leftleft);
    default: return 0;
    }
}

// ---- learning pass ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int div_down(long long sum, int count) {     // compound_enc.h:256-260
    if (sum >= 0) return (int)(sum / count);
    return (int)-((-sum + count - 1) / count);
}
__device__ __forceinline__ int compute_splitval(long long sum, int count, int lo, int hi) {       // compound_enc.h:261-284
    if (lo < 0 && hi > 0) return 0;
    int splitval = div_down(sum, count);
    if (splitval >= hi) splitval = hi - 1;
    return splitval;
}
__device__ void leaf_init(LLeaf &l, int zero_chance, int lane) {
    l.real[lane] = initial_chance(lane, zero_chance);
    for (int i = 0; i < 32; i++) { l.virt[0][i][lane] = initial_chance(i, zero_chance); l.virt[1][i][lane] = l.virt[0][i][lane]; }
    l.virtSize[lane] = 0; l.virtPropSum[lane] = 0;
    if (lane == 0) { l.realSize = 0; l.count = 0; l.best = -1; }
    __syncwarp();
}
// One learned symbol: PropertySymbolCoder::write_int = find_leaf (+ split) then the decisions on that leaf's real and
// virtual chances (compound_enc.h:307-366, 91-109).  All lanes call it with their own property value.
__device__ void learn_symbol(const EParams &P, EGroup &g, const GroupCtx &G, int &nnodes, int &nleaves, int myprop, int mn, int mx, int value, int lane) {
    TNode *nodes = g.nodes;
    long long *prop = s_scr, *slo = s_scr + 32, *shi = s_scr + 64, *ssp = s_scr + 96, *ssz = s_scr + 128;
    prop[lane] = myprop;
    __syncwarp();
    // every lane walks the tree (same path); lane p narrows the range of property p on the way
    int cur_lo = G.lo, cur_hi = G.hi;
    int pos = 0;
    while (nodes[pos].property != -1) {
        const int p = nodes[pos].property, sv = nodes[pos].splitval;
        if ((int)prop[p] > sv) { if (lane == p) cur_lo = sv + 1; pos = nodes[pos].child; }
        else { if (lane == p) cur_hi = sv; pos = nodes[pos].child + 1; }
    }
    int li = nodes[pos].child;
    LLeaf *leaf = &g.leaves[li];
    // set_selection_and_update_property_sums (the count is read by everybody before lane 0 bumps it)
    const int count = leaf->count + 1;
    const int bp = leaf->best;
    const unsigned long long realSize0 = leaf->realSize;
    int sel = 0, my_split = 0;
    if (lane < G.nprops) {
        const long long sum = leaf->virtPropSum[lane] + myprop;
        leaf->virtPropSum[lane] = sum;
        my_split = compute_splitval(sum, count, cur_lo, cur_hi);
        sel = myprop > my_split;
    }
    slo[lane] = cur_lo; shi[lane] = cur_hi; ssp[lane] = my_split;
    __syncwarp();
    if (lane == 0) leaf->count = count;
    // split the leaf if some virtual context does (significantly) better
    if (bp != -1) {
        const unsigned long long vs = leaf->virtSize[bp];
        const int blo = (int)slo[bp], bhi = (int)shi[bp], bsplit = (int)ssp[bp], bprop = (int)prop[bp];
        if (realSize0 > vs + (unsigned long long)kSplitThreshold && nleaves < 0xFFFF && nnodes < 0xFFFF && blo < bhi) {
            if (nleaves >= g.leaf_cap) { if (lane == 0) g.status = 2; __syncwarp(); return; }
            const int new_inner = nnodes;
            __syncwarp();       // everybody has read the leaf's counters
            if (lane == 0) {
                nodes[new_inner] = nodes[pos];
                nodes[new_inner + 1] = nodes[pos];
                nodes[pos].splitval = bsplit;
                nodes[pos].property = (short)bp;
            }
            // resetCounters, then the new leaf is a copy
            leaf->virtPropSum[lane] = 0; leaf->virtSize[lane] = 0;
            if (lane == 0) { leaf->best = -1; leaf->realSize = 0; leaf->count = 0; }
            __syncwarp();
            const int new_leaf = nleaves;
            {
                const unsigned *src = reinterpret_cast<const unsigned *>(leaf);
                unsigned *dst = reinterpret_cast<unsigned *>(&g.leaves[new_leaf]);
                for (unsigned k = lane; k < sizeof(LLeaf) / 4; k += 32) dst[k] = src[k];
            }
            __syncwarp();
            if (lane == 0) {
                const int old_leaf = nodes[pos].child;
                nodes[pos].child = (unsigned short)new_inner;
                nodes[new_inner].child = (unsigned short)old_leaf;
                nodes[new_inner + 1].child = (unsigned short)new_leaf;
            }
            nnodes += 2; nleaves += 1;
            li = bprop > bsplit ? li : new_leaf;
            leaf = &g.leaves[li];
        }
    }
    __syncwarp();       // everybody has compared the leaf's sizes before they change
    // the decisions: each touches a different chance, so there is no order among them; lane p adapts property p's virtual
    // chance and adds its cost, lane 31 does the same for the real chances
    unsigned char dec[40];
    const int nd = symbol_decisions(mn, mx, value, dec);
    if (nd == 0) { __syncwarp(); return; }
    unsigned long long mysz = ~0ull;
    if (lane < G.nprops) {
        uint16_t (*vc)[32] = leaf->virt[sel ? 0 : 1];
        unsigned long long sz = leaf->virtSize[lane];
        for (int k = 0; k < nd; k++) {
            const int idx = dec[k] >> 1, bit = dec[k] & 1;
            const unsigned c = vc[idx][lane];
            sz += s_log4k[bit ? c : 4096 - c];
            vc[idx][lane] = s_table[c * 2 + bit];
        }
        leaf->virtSize[lane] = sz;
        mysz = sz;
    } else if (lane == 31) {
        unsigned long long sz = leaf->realSize;
        for (int k = 0; k < nd; k++) {
            const int idx = dec[k] >> 1, bit = dec[k] & 1;
            const unsigned c = leaf->real[idx];
            sz += s_log4k[bit ? c : 4096 - c];
            leaf->real[idx] = s_table[c * 2 + bit];
        }
        leaf->realSize = sz;
    }
    ssz[lane] = (long long)mysz;
    __syncwarp();
    // best_property: the cheapest virtual context if it beats the real one, lowest index among equals (compound_enc.h:97-107)
    if (lane == 0) {
        int best = -1;
        unsigned long long best_size = leaf->realSize;
        for (int j = 0; j < G.nprops; j++) if ((unsigned long long)ssz[j] < best_size) { best_size = (unsigned long long)ssz[j]; best = j; }
        leaf->best = best;
    }
    __syncwarp();
}
// simplify, compound_enc.h:430-496, recursion unrolled (lane 0).  stack frames: {pos, stage, acc}
__device__ void kill_children(TNode *nodes, int pos, int *stack) {
    int sp = 0;
    stack[sp++] = pos;
    while (sp) {
        const int q = stack[--sp];
        if (nodes[q].property == -1) nodes[q].property = 0; else stack[sp++] = nodes[q].child;
        if (nodes[q + 1].property == -1) nodes[q + 1].property = 0; else stack[sp++] = nodes[q + 1].child;
    }
}
__device__ void simplify(EGroup &g, int *stack) {
    TNode *nodes = g.nodes;
    // post-order: frame = {pos, stage (0 enter, 1 left done, 2 right done), sum so far}; results travel in `ret`
    constexpr int N = kMaxNodes / 2 + 2;
    long long *acc = reinterpret_cast<long long *>(stack + 2 * N);      // ints [2N, 4N): one partial sum per frame
    int *kstack = stack + 4 * N;                                        // ints [4N, 8N): kill_children's own stack
    int sp = 0;
    long long ret = 0;
    stack[0] = 0; stack[1] = 0;
    sp = 1;
    while (sp) {
        int *f = stack + 2 * (sp - 1);
        const int pos = f[0];
        if (nodes[pos].property == -1) {
            const int c = g.leaves[nodes[pos].child].count;
            ret = c == 0 ? -100 : c;
            sp--;
            continue;
        }
        if (f[1] == 0) { f[1] = 1; acc[sp - 1] = 0; stack[2 * sp] = nodes[pos].child; stack[2 * sp + 1] = 0; sp++; continue; }
        if (f[1] == 1) { acc[sp - 1] += ret; f[1] = 2; stack[2 * sp] = nodes[pos].child + 1; stack[2 * sp + 1] = 0; sp++; continue; }
        acc[sp - 1] += ret;
        ret = acc[sp - 1];
        if (ret < kMinSubtree) { nodes[pos].property = -1; kill_children(nodes, nodes[pos].child, kstack); }
        sp--;
    }
}

// MetaPropertySymbolCoder::write_subtree, compound_enc.h:523-546, recursion unrolled (lane 0).  frame = {pos, stage | p << 2, oldmin, oldmax}
__device__ void write_tree(RacOut &rac, Sink &s, const EParams &P, const TNode *nodes, int nprops, const int (*range)[2], int *stack) {
    int sub[64][2];
    for (int i = 0; i < nprops; i++) { sub[i][0] = range[i][0]; sub[i][1] = range[i][1]; }
    uint16_t coder[3][32];
    for (int k = 0; k < 3; k++) for (int i = 0; i < 32; i++) coder[k][i] = initial_chance(i, 1024);
    unsigned char dec[40];
    int sp = 1;
    stack[0] = 0; stack[1] = 0; stack[2] = 0; stack[3] = 0;
    while (sp) {
        int *f = stack + 4 * (sp - 1);
        const int pos = f[0], stage = f[1] & 3, p = f[1] >> 2;
        if (stage == 0) {
            const int pp = nodes[pos].property;
            int n = symbol_decisions2(0, nprops, pp + 1, dec);
            code_decisions(rac, s, P.meta_table, coder[0], dec, n);
            if (pp == -1) { sp--; continue; }
            const int oldmin = sub[pp][0], oldmax = sub[pp][1];
            n = symbol_decisions2(oldmin, oldmax - 1, nodes[pos].splitval, dec);
            code_decisions(rac, s, P.meta_table, coder[2], dec, n);
            sub[pp][0] = nodes[pos].splitval + 1;
            f[1] = 1 | (pp << 2); f[2] = oldmin; f[3] = oldmax;
            int *gq = stack + 4 * sp;
            gq[0] = nodes[pos].child; gq[1] = 0; gq[2] = 0; gq[3] = 0;
            sp++;
        } else if (stage == 1) {
            sub[p][0] = f[2];
            sub[p][1] = nodes[pos].splitval;
            f[1] = 2 | (p << 2);
            int *gq = stack + 4 * sp;
            gq[0] = nodes[pos].child + 1; gq[1] = 0; gq[2] = 0; gq[3] = 0;
            sp++;
        } else {
            sub[p][1] = f[3];
            sp--;
        }
    }
}
__device__ void uniform_write(RacOut &rac, Sink &s, int min, int max, int val) {     // UniformSymbolCoder::write_int, symbol_enc.h:28-47 (iterative)
    for (;;) {
        if (min != 0) { max -= min; val -= min; min = 0; }
        if (max == 0) return;
        const int med = max / 2;
        if (val > med) { rac_write_bit(rac, s, 1); min = med + 1; }
        else { rac_write_bit(rac, s, 0); min = 0; max = med; }
    }
}

__device__ bool check_bit_depth(int minv, int maxv, int predictor) {    // encoding.cpp:61-72
    int maxav = s16(abs(maxv));
    if (-minv > maxav) maxav = s16(-minv);
    if (predictor > 0 && maxv - minv > maxav) maxav = s16(maxv - minv);
    if (predictor > 0 && abs(minv - maxv) > maxav) maxav = s16(abs(minv - maxv));
    return ilog2u((unsigned)maxav) + 1 <= 15;
}

// fuif_encode_channels<learn, compress> for one group.  learn: no output, grows g.nodes.  Returns false on error.
__device__ bool encode_channels(const EParams &P, EGroup &g, Sink &s, bool learn, bool compress, int lane) {
    const int beginc = g.beginc, endc = g.endc, predictor = g.predictor;
    // ---- group header (lane 0 writes; every lane follows the control flow)
    int global_minv = 0x7FFF, global_maxv = -0x7FFF;
    for (int i = beginc; i <= endc; i++) {
        const EChan &ch = P.ch[i];
        if (ch.w * ch.h <= 0) continue;
        if (ch.minval < global_minv) global_minv = ch.minval;
        if (ch.maxval > global_maxv) global_maxv = ch.maxval;
    }
    int firstrealc = beginc;
    bool depth_ok = true;
    if (lane == 0 && !learn) {
        sink_varint(s, (unsigned long long)(((endc - beginc) << 4) + (predictor << 1) + (compress ? 1 : 0)));
        if (global_minv <= 0) sink_varint(s, (unsigned long long)(1 - global_minv));
        else { sink_varint(s, 0); sink_varint(s, (unsigned long long)global_minv); }
        sink_varint(s, (unsigned long long)(global_maxv - global_minv));
    }
    for (int i = beginc; i <= endc; i++) {
        const EChan &ch = P.ch[i];
        if (ch.w * ch.h <= 0) continue;
        const int minv = ch.minval, maxv = ch.maxval;
        if (lane == 0 && !learn && endc > beginc && global_minv < global_maxv) { sink_varint(s, (unsigned long long)(minv - global_minv)); sink_varint(s, (unsigned long long)(maxv - minv)); }
        if (minv == maxv) firstrealc++;
        if (!check_bit_depth(minv, maxv, predictor)) { depth_ok = false; break; }
        if (minv == 0 && maxv == 0) continue;
        if (lane == 0 && !learn) sink_varint(s, (unsigned long long)ch.q);
    }
    if (!depth_ok) { if (lane == 0) g.status = 3; return false; }
    if (!learn && lane == 0) g.header_len = s.len;
    if (firstrealc > endc) return true;

    GroupCtx G;
    init_properties(P, beginc, endc, G, lane);
    int predictability = 2048;
    if (predictor == 0 && compress) {
        const EChan &ch = P.ch[firstrealc];
        const size_t pixels = (size_t)ch.w * ch.h;
        unsigned long long zeroes = 0;
        for (size_t k = lane; k < pixels; k += 32) zeroes += ch.data[k] == 0;
        for (int off = 16; off > 0; off >>= 1) zeroes += __shfl_sync(0xffffffffu, zeroes, lane ^ off);
        int rounded = (int)(zeroes * 128 / pixels);
        if (rounded < 1) rounded = 1;
        if (rounded > 127) rounded = 127;
        if (lane == 0 && !learn) sink_varint(s, (unsigned long long)rounded);
        predictability = rounded * 32;
    }
    RacOut rac;
    rac_init(rac);
    if (!compress) {
        if (lane == 0)
            for (int i = beginc; i <= endc; i++) {
                const EChan &ch = P.ch[i];
                for (size_t k = 0; k < (size_t)ch.w * ch.h; k++) uniform_write(rac, s, ch.minval, ch.maxval, ch.data[k]);
            }
        if (lane == 0) rac_flush(rac, s);
        __syncwarp();
        return true;
    }
    TNode *nodes = g.nodes;
    int nnodes = g.nnodes, nleaves = 1;
    if (learn) {
        leaf_init(g.leaves[0], predictability, lane);
    } else {
        // ranges of all properties, gathered from the lanes (lane 0 serialises the tree)
        int range[64][2];
        for (int p = 0; p < G.nprops; p++) { range[p][0] = __shfl_sync(0xffffffffu, G.lo, p); range[p][1] = __shfl_sync(0xffffffffu, G.hi, p); }
        if (lane == 0) write_tree(rac, s, P, nodes, G.nprops, range, g.stack);
        // FinalPropertySymbolCoder ctor, compound.h:213-225: leaf numbering in node order
        const int nl = (nnodes + 1) / 2;
        for (int k = lane; k < nl * 32; k += 32) g.fleaves[k] = initial_chance(k & 31, predictability);
        if (lane == 0) { int leafID = 0; for (int k = 0; k < nnodes; k++) if (nodes[k].property == -1) nodes[k].child = (unsigned short)leafID++; }
        __syncwarp();
    }
    long long ri = g.rand_off;
    unsigned char dec[40];
    for (int i = beginc; i <= endc; i++) {
        const EChan &ch = P.ch[i];
        const int minv = ch.minval, maxv = ch.maxval;
        if (minv == maxv) continue;
        int rowslearned = 0;
        for (int y = 0; y < ch.h; y++) {
            if (learn) { if ((float)++rowslearned > P.nb_repeats * (float)ch.h) break; }
            if (learn) { y = (ri < P.nrnd ? P.rnd[ri] : 0) % ch.h; ri++; }
            for (int x = 0; x < ch.w; x++) {
                int guess;
                const int myprop = property_and_guess(P, ch, G, x, y, predictor, lane, guess);
                const int diff = s16(ch.data[(size_t)y * ch.w + x] - guess);
                const int mn = minv - guess, mx = maxv - guess;
                if (learn) {
                    learn_symbol(P, g, G, nnodes, nleaves, myprop, mn, mx, diff, lane);
                    if (g.status) return false;
                } else if (mn != mx) {
                    s_scr[lane] = myprop;
                    __syncwarp();
                    if (lane == 0) {
                        int pos = 0;
                        while (nodes[pos].property != -1) pos = (int)s_scr[nodes[pos].property] > nodes[pos].splitval ? nodes[pos].child : nodes[pos].child + 1;
                        const int nd = symbol_decisions(mn, mx, diff, dec);
                        code_decisions(rac, s, s_table, g.fleaves + (size_t)nodes[pos].child * 32, dec, nd);
                    }
                    __syncwarp();
                }
            }
            if (learn) y = 0;
        }
    }
    __syncwarp();
    if (learn) {
        if (lane == 0) { g.nnodes = nnodes; simplify(g, g.stack); }
        __syncwarp();
    } else if (lane == 0) rac_flush(rac, s);
    __syncwarp();
    return true;
}

// One warp per group: learn, then write; if the compressed form is not smaller than the estimate of the plain form, write
// the plain form instead (encoding.cpp:528-551).
#ifdef FB_EMULATE
inline void k_maniac_encode(EParams P) {
#else
__global__ void __launch_bounds__(32) k_maniac_encode(EParams P) {
#endif
    const int lane = threadIdx.x & 31;
    const int gi = blockIdx.x;
    if (gi >= P.ngroups) return;
    EGroup &g = P.groups[gi];
    Sink s;
    s.p = g.out; s.cap = g.out_cap; s.len = 0; s.overflow = 0;
    for (int k = lane; k < 4096 * 2; k += 32) s_table[k] = P.table[k];
    for (int k = lane; k < 4097; k += 32) s_log4k[k] = P.log4k[k];
    if (lane == 0) { g.status = 0; g.attempt_len = 0; g.nnodes = 1; g.nodes[0].property = -1; g.nodes[0].child = 0; g.nodes[0].splitval = 0; g.header_len = 0; }
    __syncwarp();
    bool ok = true;
    if (!g.compress) ok = encode_channels(P, g, s, false, false, lane);
    else {
        ok = encode_channels(P, g, s, true, true, lane);
        if (ok) ok = encode_channels(P, g, s, false, true, lane);
        if (ok) {
            // bits >= ubits ?  (float arithmetic as in the reference)
            const unsigned after = (unsigned)__shfl_sync(0xffffffffu, (int)s.len, 0), hlen = (unsigned)__shfl_sync(0xffffffffu, (int)g.header_len, 0);
            const float bits = (float)(after - hlen) * 8.0f;
            float ubits = 0.0f;
            for (int k = g.beginc; k <= g.endc; k++) {
                const float chpixels = (float)(P.ch[k].w * P.ch[k].h);
                const float ubpp = (float)(ilog2u((unsigned)(P.ch[k].maxval - P.ch[k].minval)) + 1);
                if (P.ch[k].maxval > P.ch[k].minval) ubits += chpixels * ubpp;
            }
            if (ubits > 0.0f) ubits += 16;
            if (bits >= ubits) {
                if (lane == 0) g.attempt_len = after;
                s.len = 0; s.overflow = 0;
                ok = encode_channels(P, g, s, false, false, lane);
            }
        }
    }
    if (lane == 0) {
        g.out_len = s.len;
        if (ok && s.overflow) g.status = 1;
    }
}

}  // namespace fbenc

#ifndef FB_EMULATE
// ---------------------------------------------------------------------------------------------------------------------------
// C ABI: fb_encode = fuif_prepare_encode + fuif_encode (reference encoding/encoding.cpp:737-743, 455-573)
// ---------------------------------------------------------------------------------------------------------------------------
#include <new>

#include "fb_encode_host.h"

namespace {
struct DevBuf {         // frees on scope exit
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 16); }
};
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
}  // namespace

extern "C" void fb_free(void *p) { free(p); }

extern "C" int fb_encode(fb_ctx *ctx, fb_image *img, const fb_encode_options *opts, uint8_t **bytes_out, size_t *nbytes_out, int64_t *group_index,
                         int32_t *group_first, int cap_groups, int *n_groups_out) {
    using namespace fbenc;
    namespace H = fbenc_host;
    if (!ctx || !img || !bytes_out || !nbytes_out || img->ctx != ctx) return FB_ERR_INVALID;
    *bytes_out = nullptr; *nbytes_out = 0;
    if (n_groups_out) *n_groups_out = 0;
    if (img->info.error) { ctx->err = "fb_encode: image carries an error"; return FB_ERR_INVALID; }
    cudaSetDevice(ctx->device);
    H::Options o;
    if (opts) {
        o.nb_repeats = opts->nb_repeats; o.max_properties = opts->max_properties; o.maniac_cutoff = opts->maniac_cutoff; o.maniac_alpha = opts->maniac_alpha;
        o.compress = opts->compress != 0; o.max_group = opts->max_group;
        if (opts->n_predictors > 0 && opts->predictor) o.predictor.assign(opts->predictor, opts->predictor + opts->n_predictors);
    }
    if (o.max_properties < 0 || o.max_properties > 18) { ctx->err = "fb_encode: max_properties > 18 is not supported (one lane per property)"; return FB_ERR_UNSUPPORTED; }
    if (!(o.nb_repeats >= 0.0f) || o.nb_repeats > 64.0f) { ctx->err = "fb_encode: nb_repeats out of range"; return FB_ERR_INVALID; }
    // fuif_prepare_encode: tight ranges (the downscale positions are computed by plan_groups / assemble)
    int rc = fb_image_recompute_minmax(img);
    if (rc) return rc;
    const int nch = (int)img->ch.size();
    std::vector<H::Plane> planes((size_t)nch);
    std::vector<EChan> ech((size_t)nch);
    for (int i = 0; i < nch; i++) {
        fb_plane_desc &d = img->ch[(size_t)i].d;
        const bool has = d.w > 0 && d.h > 0;
        if (has && !img->ch[(size_t)i].dev) { ctx->err = "fb_encode: a plane has no samples"; return FB_ERR_INVALID; }
        if (has && (long long)d.w * d.h > 0x7fffffffLL) { ctx->err = "fb_encode: plane too large"; return FB_ERR_UNSUPPORTED; }
        H::Plane &p = planes[(size_t)i];
        p.w = d.w; p.h = d.h; p.minval = d.minval; p.maxval = d.maxval; p.zero = d.zero; p.q = d.q; p.hshift = d.hshift; p.vshift = d.vshift;
        p.hcshift = d.hcshift; p.vcshift = d.vcshift;
        if (has && !(p.minval == 0 && p.maxval == 0)) { H::chan_setzero(p); d.zero = p.zero; }      // encoding.cpp:118 (zero is `mutable` there)
        EChan &e = ech[(size_t)i];
        e.w = p.w; e.h = p.h; e.minval = p.minval; e.maxval = p.maxval; e.zero = p.zero; e.q = p.q; e.hshift = p.hshift; e.vshift = p.vshift;
        e.data = img->ch[(size_t)i].dev;
    }
    H::ImageInfo info;
    info.w = img->info.w; info.h = img->info.h; info.maxval = img->info.maxval; info.colormodel = img->info.colormodel;
    info.real_nb_channels = img->info.real_nb_channels; info.nb_channels = img->info.nb_channels; info.nb_meta_channels = img->info.nb_meta_channels;
    std::vector<H::Transform> tr;
    for (const FbXform &t : img->tr) tr.push_back(H::Transform{t.id, t.p});
    long long nrand = 0;
    std::vector<H::Group> groups;
    if (info.real_nb_channels >= 1) groups = H::plan_groups(planes, info, o, &nrand);
    const int ng = (int)groups.size();
    std::vector<H::GroupBytes> gb((size_t)ng);
    std::vector<std::vector<unsigned char>> host_bytes((size_t)ng);
    if (ng) {
        // tables
        std::vector<uint16_t> table(4096 * 2), meta(4096 * 2), log4k(4097);
        H::build_chance_table(table.data(), (uint32_t)o.maniac_alpha, (unsigned)(4096 - o.maniac_cutoff));
        H::build_chance_table(meta.data(), 0xFFFFFFFFu / 19, 4096 - 2);
        H::build_log4k(log4k.data());
        std::vector<int> rnd((size_t)nrand + 1);
        H::glibc_rand(rnd.data(), nrand);
        // working memory of every group, carved from one allocation per kind
        std::vector<EGroup> eg((size_t)ng);
        std::vector<size_t> leaf_off((size_t)ng), out_off((size_t)ng);
        size_t leaf_total = 0, out_total = 0;
        for (int g = 0; g < ng; g++) {
            const H::Group &G = groups[(size_t)g];
            long long cap = G.learned + 1;          // a split needs a learned symbol
            if (cap > kMaxNodes / 2) cap = kMaxNodes / 2;
            if (cap < 2) cap = 2;
            leaf_off[(size_t)g] = leaf_total; leaf_total += (size_t)cap;
            long long tree_bytes = 24 * G.learned;
            if (tree_bytes > (512 << 10)) tree_bytes = 512 << 10;
            const size_t ocap = align_up((size_t)(4 * G.pixels + tree_bytes + 4096), 256);
            if (ocap > 0xfffffff0u) { ctx->err = "fb_encode: group too large"; return FB_ERR_UNSUPPORTED; }
            out_off[(size_t)g] = out_total; out_total += ocap;
            EGroup &E = eg[(size_t)g];
            memset(&E, 0, sizeof(E));
            E.beginc = G.beginc; E.endc = G.endc; E.predictor = G.predictor; E.compress = o.compress ? 1 : 0; E.rand_off = G.rand_off;
            E.leaf_cap = (int)cap; E.out_cap = (unsigned)ocap;
        }
        const size_t stack_ints = (size_t)8 * (kMaxNodes / 2 + 2);
        DevBuf d_ch, d_groups, d_nodes, d_leaves, d_fleaves, d_stack, d_out, d_tables, d_rnd;
        const size_t tables_bytes = (4096 * 2 * 2 + 4104) * sizeof(uint16_t);
        if (d_ch.alloc(sizeof(EChan) * (size_t)nch) != cudaSuccess || d_groups.alloc(sizeof(EGroup) * (size_t)ng) != cudaSuccess ||
            d_nodes.alloc(sizeof(TNode) * kMaxNodes * (size_t)ng) != cudaSuccess || d_leaves.alloc(sizeof(LLeaf) * leaf_total) != cudaSuccess ||
            d_fleaves.alloc(sizeof(uint16_t) * 32 * (kMaxNodes / 2) * (size_t)ng) != cudaSuccess || d_stack.alloc(sizeof(int) * stack_ints * (size_t)ng) != cudaSuccess ||
            d_out.alloc(out_total) != cudaSuccess || d_tables.alloc(tables_bytes) != cudaSuccess ||
            d_rnd.alloc(sizeof(int) * rnd.size()) != cudaSuccess) {
            cudaGetLastError();
            ctx->err = "fb_encode: out of device memory";
            return FB_ERR_NOMEM;
        }
        for (int g = 0; g < ng; g++) {
            EGroup &E = eg[(size_t)g];
            E.nodes = (TNode *)d_nodes.p + (size_t)kMaxNodes * g;
            E.leaves = (LLeaf *)d_leaves.p + leaf_off[(size_t)g];
            E.fleaves = (uint16_t *)d_fleaves.p + (size_t)32 * (kMaxNodes / 2) * g;
            E.stack = (int *)d_stack.p + stack_ints * g;
            E.out = (unsigned char *)d_out.p + out_off[(size_t)g];
        }
        uint16_t *dt = (uint16_t *)d_tables.p;
        FB_CUDA(ctx, cudaMemcpyAsync(d_ch.p, ech.data(), sizeof(EChan) * (size_t)nch, cudaMemcpyHostToDevice, ctx->stream));
        FB_CUDA(ctx, cudaMemcpyAsync(d_groups.p, eg.data(), sizeof(EGroup) * (size_t)ng, cudaMemcpyHostToDevice, ctx->stream));
        FB_CUDA(ctx, cudaMemcpyAsync(dt, table.data(), 8192 * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
        FB_CUDA(ctx, cudaMemcpyAsync(dt + 8192, meta.data(), 8192 * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
        FB_CUDA(ctx, cudaMemcpyAsync(dt + 16384, log4k.data(), 4097 * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
        FB_CUDA(ctx, cudaMemcpyAsync(d_rnd.p, rnd.data(), sizeof(int) * rnd.size(), cudaMemcpyHostToDevice, ctx->stream));
        // the tail a rolled-back attempt leaves must read as written, and unwritten bytes as zero
        FB_CUDA(ctx, cudaMemsetAsync(d_out.p, 0, out_total, ctx->stream));
        EParams P;
        P.ch = (const EChan *)d_ch.p; P.nch = nch; P.groups = (EGroup *)d_groups.p; P.ngroups = ng; P.max_properties = o.max_properties;
        P.nb_repeats = o.nb_repeats; P.table = dt; P.meta_table = dt + 8192; P.log4k = dt + 16384; P.rnd = (const int *)d_rnd.p; P.nrnd = nrand;
        {   // the kernel's frame (property tables, decision lists) is larger than the default per-thread stack
            cudaFuncAttributes fa;
            size_t lim = 0;
            FB_CUDA(ctx, cudaFuncGetAttributes(&fa, k_maniac_encode));
            FB_CUDA(ctx, cudaDeviceGetLimit(&lim, cudaLimitStackSize));
            if (lim < fa.localSizeBytes + 512) FB_CUDA(ctx, cudaDeviceSetLimit(cudaLimitStackSize, fa.localSizeBytes + 512));
        }
        k_maniac_encode<<<ng, 32, 0, ctx->stream>>>(P);
        FB_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
        ctx->mark("k_maniac_encode");
        FB_CUDA(ctx, cudaMemcpyAsync(eg.data(), d_groups.p, sizeof(EGroup) * (size_t)ng, cudaMemcpyDeviceToHost, ctx->stream));
        FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int g = 0; g < ng; g++) {
            const EGroup &E = eg[(size_t)g];
            if (E.status == 3) { ctx->err = "fb_encode: sample range needs more than 15 bits (check_bit_depth, encoding.cpp:61-72)"; return FB_ERR_INVALID; }
            if (E.status) { ctx->err = E.status == 1 ? "fb_encode: group output buffer too small" : "fb_encode: leaf pool exhausted"; return FB_ERR_UNSUPPORTED; }
            if (E.attempt_len > E.out_cap) { ctx->err = "fb_encode: rolled-back attempt exceeds the output buffer"; return FB_ERR_UNSUPPORTED; }
            const unsigned n = E.attempt_len > E.out_len ? E.attempt_len : E.out_len;
            host_bytes[(size_t)g].resize(n ? n : 1);
            if (n) FB_CUDA(ctx, cudaMemcpyAsync(host_bytes[(size_t)g].data(), E.out, n, cudaMemcpyDeviceToHost, ctx->stream));
            gb[(size_t)g].bytes = host_bytes[(size_t)g].data();
            gb[(size_t)g].out_len = E.out_len;
            gb[(size_t)g].attempt_len = E.attempt_len;
        }
        FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    std::vector<int64_t> offs;
    std::vector<uint8_t> file = H::assemble(info, tr, planes, o, groups, gb, &offs);
    uint8_t *outp = (uint8_t *)malloc(file.size() ? file.size() : 1);
    if (!outp) return FB_ERR_NOMEM;
    if (!file.empty()) memcpy(outp, file.data(), file.size());
    *bytes_out = outp; *nbytes_out = file.size();
    if (n_groups_out) *n_groups_out = ng;
    for (int g = 0; g < ng && g < cap_groups; g++) {
        if (group_index) group_index[g] = offs[(size_t)g];
        if (group_first) group_first[g] = groups[(size_t)g].beginc;
    }
    return FB_OK;
}
#endif  // !FB_EMULATE
