// Host-threads backend of the entropy stage: fuif_decode_channel (reference encoding/encoding.cpp:259-429) with what it inlines,
//   group header varints                      encoding.cpp:264-329
//   init_properties / predictors / properties encoding/context_predict.h:67-120, 125-206
//   precompute_references                     context_predict.h:233-289
//   24-bit range decoder                      maniac/rac.h:35-114
//   adaptive 12-bit chances                   maniac/chance.h:42-84, chance.cpp:31-65
//   zero/sign/exponent/mantissa integer coder maniac/symbol.h:72-185, uniform coder :44-56
//   MANIAC tree parse + leaf walk             maniac/compound.h:135-320
// run on CPU threads.  The bitstream is serial inside a channel group, so the unit of work is a stream:
// one group when the caller knows the groups' byte offsets, else one image.  Streams are claimed through an atomic ticket in index
// order; a stream only waits (row wavefront, `rows_done`) for planes of lower-numbered streams, which are running or finished.
// Same structure as the device code in fb_maniac.cu (one "lane"), written for a CPU: every property of a pixel is computed
// into a small array (what does not depend on the pixel to the left for a whole chunk of pixels ahead of the serial loop), the
// tree walk is a chain of L1/L2 loads with a branch-free child select, and the chance update takes both successors from one table row.
#include "fb_host_entropy.h"

#include <stdio.h>
#include <string.h>
#ifdef __linux__
#include <sched.h>
#endif

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define FBH_PAUSE() _mm_pause()
#else
#define FBH_PAUSE() ((void)0)
#endif

#include "../../include/fuif_b200.h"

namespace fbh {
namespace {

constexpr int NB_NONREF = 13;           // context_predict.h:210
constexpr int MAX_BIT_DEPTH = 15;       // config.h:5
constexpr int kMaxProps = 120;          // property index field of a packed node (7 bits, 127 = leaf)
constexpr int kLeafMark = 127;
constexpr size_t kMaxTreeNodes = ((size_t)1 << 25) - 2;

#define FBH_INLINE inline __attribute__((always_inline))

FBH_INLINE int s16(int x) { return (int)(short)x; }
FBH_INLINE int ilog2u(unsigned l) { return l == 0 ? 0 : 31 - __builtin_clz(l); }      // maniac/util.h:33-36
FBH_INLINE int slog_calc(int x16) {     // context_predict.h:54-61
    const int x = s16(x16);
    const unsigned a = (unsigned)(x < 0 ? -x : x);
    const int b = a ? 32 - __builtin_clz(a) : 0;
    return x < 0 ? -b : b;
}
// slog of every 16-bit value: one load instead of abs / bsr / select (the values met in practice are small: a few hot lines)
struct SlogTable {
    int8_t t[65536];
    SlogTable() { for (int i = 0; i < 65536; i++) t[i] = (int8_t)slog_calc(i); }
};
const SlogTable g_slog;
FBH_INLINE int slog(int x16) { return g_slog.t[(uint16_t)x16]; }
FBH_INLINE int fooabs(int x16) { const int x = s16(x16); return s16(x < 0 ? -x : x); }   // :63-65
FBH_INLINE int median3(int a, int b, int c) {       // util.h:9-23
    if (a < b) { if (b < c) return b; return a < c ? c : a; }
    if (a < c) return a;
    return b < c ? c : b;
}
FBH_INLINE int shl(int v, int s) { return (s < 0 || s > 30) ? 0 : (int)((unsigned)v << s); }
FBH_INLINE int shr(int v, int s) { return s < 0 ? v : (s > 30 ? 0 : v >> s); }

FBH_INLINE int ld_acquire(const int *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
FBH_INLINE void st_release(int *p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
void wait_until_ge(const int *flag, int want) {
    for (int spins = 0; ld_acquire(flag) < want; spins++) {
        if (spins < 256) FBH_PAUSE();
        else std::this_thread::yield();
    }
}

// ---- byte reader with FileIO end-of-stream rules (fileio.h:33-81): EOF is only known after a failed read
struct Reader {
    const uint8_t *p;
    unsigned long long n, pos, btl;
    bool eof;
    FBH_INLINE int get() {
        if (pos >= n) { eof = true; return -1; }
        return p[pos++];
    }
    FBH_INLINE bool stop() const { return eof || (btl && pos >= btl); }
    int varint() {                      // read_big_endian_varint, encoding.cpp:45-59
        int result = 0, bytes_read = 0;
        while (bytes_read++ < 10) {
            int number = get();
            if (number < 0) break;
            if (number < 128) return result + number;
            number -= 128;
            result += number;
            result = (int)((unsigned)result << 7);
        }
        return -1;
    }
};

// ---- range decoder (maniac/rac.h)
struct Rac {
    Reader io;
    uint32_t range, low;
    bool ones;      // a read past the end turned `low` into all-ones garbage (rac.h:64-69): every later bit reads as 1
    FBH_INLINE void byte_in() {
        const int c = io.get();
        if (c < 0) ones = true;
        low = (low << 8) | (uint32_t)(c & 0xFF);
    }
    void init(const Reader &r) {        // RacInput ctor, rac.h:97-104
        io = r; range = 1u << 24; low = 0; ones = false;
        byte_in(); byte_in(); byte_in();
    }
    FBH_INLINE int get(uint32_t chance) {       // rac.h:82-95 with input() :70-81
        const uint32_t thr = range - chance;
        int bit;
        if (ones || low >= thr) { low -= thr; range = chance; bit = 1; }
        else { range = thr; bit = 0; }
        if (range <= (1u << 16)) {
            range <<= 8; byte_in();
            if (range <= (1u << 16)) { range <<= 8; byte_in(); }
        }
        return bit;
    }
    FBH_INLINE int read12(uint32_t b12) { return get((uint32_t)(((uint64_t)range * b12 + 0x800ull) >> 12)); }     // rac.h:42-52, 107
    FBH_INLINE int read_bit() { return get(range >> 1); }      // rac.h:111
};

// build_table, maniac/chance.cpp:31-65: t[c][bit] = the chance that follows c after decoding `bit`
void build_table(uint16_t *t /*[4096][2]*/, uint32_t factor, unsigned max_p) {
    const int64_t one = 1LL << 32;
    const int size = 4096;
    memset(t, 0, sizeof(uint16_t) * size * 2);
    unsigned last_p8 = 0, p8;
    int64_t p = one / 2;
    for (unsigned i = 0; i < (unsigned)size / 2; i++) {
        p8 = (unsigned)((size * p + one / 2) >> 32);
        if (p8 <= last_p8) p8 = last_p8 + 1;
        if (last_p8 && last_p8 < (unsigned)size && p8 <= max_p) t[last_p8 * 2 + 1] = (uint16_t)p8;
        p += ((one - p) * factor + one / 2) >> 32;
        last_p8 = p8;
    }
    for (unsigned i = size - max_p; i <= max_p; i++) {
        if (t[i * 2 + 1]) continue;
        p = ((int64_t)i * one + size / 2) / size;
        p += ((one - p) * factor + one / 2) >> 32;
        p8 = (unsigned)((size * p + one / 2) >> 32);
        if (p8 <= i) p8 = i + 1;
        if (p8 > max_p) p8 = max_p;
        t[i * 2 + 1] = (uint16_t)p8;
    }
    for (unsigned i = 1; i < (unsigned)size; i++) t[i * 2 + 0] = (uint16_t)(size - t[(size - i) * 2 + 1]);
}

// ---- symbol coder (maniac/symbol.h).  leaf layout: [0]=zero [1]=sign [2..15]=exp[14] [16..30]=mant[15] [31]=pad
constexpr int SC_ZERO = 0, SC_SIGN = 1, SC_EXP = 2, SC_MANT = 16;

uint16_t initial_chance(int idx, int zero_chance) {      // SymbolChance(zero_chance), symbol.h:115-138
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

// FinalCompoundSymbolBitCoder::read (compound.h:90-95): decode with the leaf's chance, then move the chance along the table
FBH_INLINE int read_ctx(Rac &rac, const uint16_t *table, uint16_t *leaf, int idx) {
    const unsigned ch = leaf[idx];
    const int bit = rac.read12(ch);
    leaf[idx] = table[ch * 2 + bit];
    return bit;
}

// reader<15>(coder, min, max), symbol.h:154-185.  mant_base: index of bit_mant[0] in the leaf (16 in the full layout, 9 in the
// compact one: zero, sign, exp[7], mant[7] -- enough while no value of the group needs more than 7 exponent / mantissa bits)
FBH_INLINE int read_int(Rac &rac, const uint16_t *table, uint16_t *leaf, int mn, int mx, int mant_base = SC_MANT) {
    if (mn == mx) return mn;
    if (read_ctx(rac, table, leaf, SC_ZERO)) return 0;
    int sign;
    if (mn < 0) { if (mx > 0) sign = read_ctx(rac, table, leaf, SC_SIGN); else sign = 0; } else sign = 1;
    const int amax = sign ? mx : -mn;
    const int emax = ilog2u((unsigned)amax);
    int e = 0;
    for (; e < emax; e++) if (read_ctx(rac, table, leaf, SC_EXP + e)) break;
    int have = 1 << e;
    for (int pos = e; pos > 0;) {
        pos--;
        const int minabs1 = have | (1 << pos);
        if (minabs1 > amax) continue;
        if (read_ctx(rac, table, leaf, mant_base + pos)) have = minabs1;
    }
    return sign ? have : -have;
}
int read_int2(Rac &rac, const uint16_t *table, uint16_t *leaf, int mn, int mx) {     // symbol.h:232-236
    if (mn > 0) return read_int(rac, table, leaf, 0, mx - mn) + mn;
    if (mx < 0) return read_int(rac, table, leaf, mn - mx, 0) + mx;
    return read_int(rac, table, leaf, mn, mx);
}
FBH_INLINE int uniform_read(Rac &rac, int mn, int len) {    // UniformSymbolCoder::read_int, symbol.h:44-56
    while (len > 0) {       // len < 0 only comes out of a damaged header (the reference asserts); it must not loop forever
        const int med = len / 2;
        if (rac.read_bit()) { mn = mn + med + 1; len = len - (med + 1); }
        else len = med;
    }
    return mn;
}

bool check_bit_depth(int minv, int maxv, int predictor) {    // encoding.cpp:61-72
    int maxav = s16(abs(maxv));
    if (-minv > maxav) maxav = s16(-minv);
    if (predictor > 0 && maxv - minv > maxav) maxav = s16(maxv - minv);
    if (predictor > 0 && abs(minv - maxv) > maxav) maxav = s16(abs(minv - maxv));
    return ilog2u((unsigned)maxav) + 1 <= MAX_BIT_DEPTH;
}

void fill_plane(Chan &c, int value) {
    std::fill(c.data, c.data + (size_t)c.w * c.h, (int16_t)value);
    c.state = 1;
}

// A tree node in one 64-bit word: split value in the high half; low half = slot << 7 | property for an inner node (its children
// sit in slots slot, slot + 1: the first is taken when property > split, compound.h:142-153), leaf id << 7 | 127 for a leaf
// (leaves numbered in node order, compound.h:213-225).  Node i lives in slot i + 1, so sibling pairs are 16-byte aligned and
// both children come in with one cache line before the compare that picks one of them is done.
typedef uint64_t Node;
FBH_INLINE Node make_node(int32_t split, uint32_t lo) { return ((uint64_t)(uint32_t)split << 32) | lo; }
FBH_INLINE int32_t node_split(Node n) { return (int32_t)(n >> 32); }
FBH_INLINE unsigned node_prop(Node n) { return (unsigned)n & 127u; }
FBH_INLINE unsigned node_ref(Node n) { return (unsigned)n >> 7; }        // first child's slot, or the leaf id
FBH_INLINE bool node_is_leaf(Node n) { return ((unsigned)n & 127u) == 127u; }
// find_leaf, compound.h:142-153.  A select, not a branch: the direction taken at a node is close to random for the predictor
// (measured: the branchy form is 15 % slower), and the two loads of a level -- the property value and the child pair -- start together.
FBH_INLINE Node walk(const Node *nodes, const int *props, Node n) {
    while (!node_is_leaf(n)) {
        const Node *c = nodes + node_ref(n);
        const Node a = c[0], b = c[1];
        const uint64_t take_a = (uint64_t)0 - (uint64_t)(props[node_prop(n)] > node_split(n));      // mask arithmetic: the compiler turns ?: into a branch
        n = b ^ ((a ^ b) & take_a);
    }
    return n;
}

struct Tables {
    uint16_t table[4096 * 2];       // decode chances (cutoff, alpha)
    uint16_t meta[4096 * 2];        // tree coder: SimpleBitChanceTable(cut 2, alpha 0xFFFFFFFF / 19), chance.h:53
    // Look-ahead helper threads of large planes (decode_plane_ahead): allowed unless the caller asked for ONE thread, and started
    // only while fewer threads are busy than the machine has hardware threads -- they shorten the critical path of the last, largest
    // groups of a file, but cost about 30 % more CPU work in total, which would only slow things down while every core is busy.
    bool helpers, debug;
    int hw_threads;
    mutable std::atomic<int> busy;      // worker threads inside a stream + helper threads running
};

struct Scratch {        // per thread, reused from group to group
    std::vector<Node> nodes;
    std::vector<int> prop_of, child_of, split_of;      // tree as parsed
    std::vector<int> stack;
    std::vector<uint16_t> leaves;
    std::vector<int16_t> refrow;        // [nrefchan][w]: the co-located sample of every referenced plane for the current row
    std::vector<int> chunk;             // [kChunk][stride]: the property rows of the pixels of the current chunk
};

// init_properties, context_predict.h:67-120: property ranges; which earlier planes are referenced
int init_properties(int (*pr)[2], Image &img, int beginc, int endc, int *refchan, int &nrefchan) {
    int n = 0, offset = 0;
    nrefchan = 0;
    for (int j = beginc - 1; j >= 0 && offset < img.max_properties; j--) {
        wait_until_ge(&img.ch[j].hdr_done, 1);
        const Chan &cj = img.ch[j];
        const int cmin = cj.minval, cmax = cj.maxval;
        if (cmin == cmax) continue;
        if (cj.hshift < 0) continue;
        int minval = cmin; if (minval > 0) minval = 0;
        int maxval = cmax; if (maxval < 0) maxval = 0;
        pr[n][0] = 0; pr[n][1] = fooabs(maxval > -minval ? maxval : minval); n++; offset++;
        pr[n][0] = slog(minval); pr[n][1] = slog(maxval); n++; offset++;
        refchan[nrefchan++] = j;
    }
    int minval = 0x7FFF, maxval = -0x7FFF, maxh = 0, maxw = 0;
    for (int j = beginc; j <= endc; j++) {
        const Chan &cj = img.ch[j];
        if (cj.minval < minval) minval = cj.minval;
        if (cj.maxval > maxval) maxval = cj.maxval;
        if (cj.h > maxh) maxh = cj.h;
        if (cj.w > maxw) maxw = cj.w;
    }
    if (minval > 0) minval = 0;
    if (maxval < 0) maxval = 0;
    const int amax = std::max(fooabs(minval), fooabs(maxval));
    pr[n][0] = 0; pr[n][1] = amax; n++;
    pr[n][0] = 0; pr[n][1] = amax; n++;
    pr[n][0] = slog(minval); pr[n][1] = slog(maxval); n++;
    pr[n][0] = slog(minval); pr[n][1] = slog(maxval); n++;
    pr[n][0] = 0; pr[n][1] = maxh - 1; n++;
    pr[n][0] = 0; pr[n][1] = maxw - 1; n++;
    pr[n][0] = minval + minval - maxval; pr[n][1] = maxval + maxval - minval; n++;
    pr[n][0] = minval + minval - maxval; pr[n][1] = maxval + maxval - minval; n++;
    for (int k = 0; k < 5; k++) { pr[n][0] = slog(minval - maxval); pr[n][1] = slog(maxval - minval); n++; }
    return n;
}

// MetaPropertySymbolCoder::read_tree, compound.h:277-320; the recursion runs on an explicit stack (a damaged stream must not
// overflow the thread's).  Returns false for an invalid tree.
bool read_tree(Rac &rac, const uint16_t *mtable, int (*range)[2], int nprops, Scratch &S) {
    int sub[kMaxProps + 16][2];
    for (int i = 0; i < nprops; i++) { sub[i][0] = range[i][0]; sub[i][1] = range[i][1]; }
    uint16_t coder[3][32];
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 32; i++) coder[k][i] = initial_chance(i, 1024);     // SimpleSymbolCoder ctx(ZERO_CHANCE), symbol.h:219
    auto &P = S.prop_of; auto &C = S.child_of; auto &V = S.split_of; auto &st = S.stack;
    P.assign(1, -1); C.assign(1, 0); V.assign(1, 0);
    st.clear();
    // frame = {node, stage | property << 2, oldmin, oldmax}
    st.insert(st.end(), {0, 0, 0, 0});
    while (!st.empty()) {
        const size_t f = st.size() - 4;
        const int pos = st[f], stage = st[f + 1] & 3, p = st[f + 1] >> 2;
        if (stage == 0) {
            const int pp = read_int2(rac, mtable, coder[0], 0, nprops) - 1;
            P[pos] = pp;
            if (pp == -1) { st.resize(f); continue; }
            const int oldmin = sub[pp][0], oldmax = sub[pp][1];
            if (oldmin >= oldmax) return false;                                     // "Invalid tree", compound.h:285-288
            const int splitval = read_int2(rac, mtable, coder[2], oldmin, oldmax - 1);
            V[pos] = splitval;
            if (P.size() + 2 > kMaxTreeNodes) return false;
            const int child = (int)P.size();
            C[pos] = child;
            P.push_back(-1); P.push_back(-1); C.push_back(0); C.push_back(0); V.push_back(0); V.push_back(0);
            sub[pp][0] = splitval + 1;
            st[f + 1] = 1 | (pp << 2); st[f + 2] = oldmin; st[f + 3] = oldmax;
            st.insert(st.end(), {child, 0, 0, 0});
        } else if (stage == 1) {
            sub[p][0] = st[f + 2];
            sub[p][1] = V[pos];
            st[f + 1] = 2 | (p << 2);
            st.insert(st.end(), {C[pos] + 1, 0, 0, 0});
        } else {
            sub[p][1] = st[f + 3];
            st.resize(f);
        }
    }
    return true;
}

FBH_INLINE int predict(int predictor, int left, int top, int topleft, int topright, int zero, int cmin, int cmax) {     // context_predict.h:157-166
    switch (predictor) {
    case 0: return zero;
    case 1: return s16((left + top) / 2);
    case 2: return median3(s16(left + top - topleft), left, top);
    case 3: return left;
    case 4: return top;
    case 5: return s16((left + topleft + top + topright) / 4);
    case 6: { const int g = left + top - topleft; return s16(g < cmin ? cmin : (g > cmax ? cmax : g)); }
    default: return median3(s16(left + top - topleft), left, top);
    }
}

// precompute_references, context_predict.h:233-289: for every x of row y the co-located sample of a referenced plane
void reference_row(const Chan &ch, const Chan &cj, int y, int16_t *out) {
    if (cj.w <= 0 || cj.h <= 0) {       // an empty plane can be "referenced" (its default range is not a constant); the reference reads
        memset(out, 0, (size_t)ch.w * sizeof(int16_t));     // out of bounds there -- nothing to be exact about, so: zeros
        return;
    }
    int ry = shr(shl(y, ch.vshift), cj.vshift);
    if (ry >= cj.h) ry = cj.h - 1;
    const int16_t *src = cj.data + (size_t)ry * cj.w;
    const int w = ch.w;
    if (ch.hshift == cj.hshift && w <= cj.w) {
        memcpy(out, src, (size_t)w * sizeof(int16_t));
    } else if (ch.hshift < cj.hshift) {
        const int stepsize = shr(shl(1, cj.hshift), ch.hshift);     // all samples but the last are repeated stepsize times
        for (int x = 0; x < w; x++) {
            int rx = stepsize > 0 ? x / stepsize : cj.w - 1;
            if (rx > cj.w - 1) rx = cj.w - 1;
            out[x] = src[rx];
        }
    } else {
        for (int x = 0; x < w; x++) {
            int rx = shr(shl(x, ch.hshift), cj.hshift);
            if (rx >= cj.w) rx = cj.w - 1;
            out[x] = src[rx];
        }
    }
}

// One row of the slow track (encoding.cpp:388-421), every property computed per pixel: used for row 0, where `topleft` is the
// pixel to the left and nearly everything depends on it.
template <bool PRED0>
FBH_INLINE void decode_row(const Chan &ch, int y, int predictor, int nused, const int *used_ref, int nref, const int16_t *refrow, Rac &rac,
                           const uint16_t *table, const Node *nodes, uint16_t *leaves, int leaf_shift, int mant_base) {
    const int w = ch.w, zero = ch.zero, cmin = ch.minval, cmax = ch.maxval;
    int16_t *row = ch.data + (size_t)y * w;
    const int16_t *row1 = row - w, *row2 = row1 - w;
    int props[kMaxProps + 16];
    int *np = props + nref;
    np[4] = y;
    int left = zero, leftleft = zero, topleft_next = zero;
    for (int x = 0; x < w; x++) {
        int top = zero, topright = zero, toptop = zero, topleft = left;
        if (y) {
            top = row1[x];
            if (x) topleft = topleft_next;
            topright = (x + 1 < w) ? row1[x + 1] : top;
            toptop = (y > 1) ? row2[x] : top;
            topleft_next = top;
        }
        for (int k = 0; k < nused; k++) {       // only the referenced planes some node of this group's tree tests
            const int r = used_ref[k];
            const int rv = refrow[(size_t)r * w + x];
            props[2 * r] = fooabs(rv);
            props[2 * r + 1] = slog(rv);
        }
        np[0] = fooabs(top);
        np[1] = fooabs(left);
        np[2] = slog(top);
        np[3] = slog(left);
        np[5] = x;
        np[6] = left + top - topleft;
        np[7] = topleft + topright - top;
        np[8] = slog(left - topleft);
        np[9] = slog(topleft - top);
        np[10] = slog(top - topright);
        np[11] = slog(top - toptop);
        np[12] = slog(left - leftleft);
        const int guess = PRED0 ? zero : predict(predictor, left, top, topleft, topright, zero, cmin, cmax);
        const int mn = cmin - guess, mx = cmax - guess;
        int diff = mn;
        if (mn != mx) {
            const Node n = walk(nodes, props, nodes[1]);
            diff = read_int(rac, table, leaves + ((size_t)node_ref(n) << leaf_shift), mn, mx, mant_base);
        }
        const int val = s16(s16(diff) + guess);
        row[x] = (int16_t)val;
        leftleft = x ? left : val;          // next pixel: x > 1 ? value(x-2) : left   (context_predict.h:132)
        left = val;
    }
}

// Rows y >= 1, a chunk of kChunk pixels at a time: everything that only depends on the rows above and on the referenced planes
// (8 of the 13 local properties, all reference properties) is computed for the whole chunk first -- independent work the core
// overlaps freely -- into a property row per pixel; the serial loop then adds the 5 properties that need `left` (fooabs(left),
// slog(left), left + (top - topleft), slog(left - topleft), slog(left - leftleft)), walks the tree and decodes.
// At x == 0 topleft is `left`, which is `zero` there (context_predict.h:126-128), so it is known in advance as well.
// Measured on a 2048^2 file, one thread: 4.1 s -> 2.5 s.  (Also tried: walking the tree ahead of the serial loop as far as it
// only tests known properties -- 3.4 of 9.4 levels on average, 1.8 tests per path need `left` -- which lost more in mispredicted
// loop exits than it saved.)
constexpr int kChunk = 64;
template <bool PRED0>
__attribute__((noinline)) void decode_row_chunked(const Chan &ch, int y, int predictor, int nused, const int *used_ref, int nref, int stride, int *pc,
                                   const int16_t *refrow, Rac &rac, const uint16_t *table, const Node *nodes, uint16_t *leaves, int leaf_shift,
                                   int mant_base) {
    const int w = ch.w, zero = ch.zero, cmin = ch.minval, cmax = ch.maxval;
    int16_t *row = ch.data + (size_t)y * w;
    const int16_t *row1 = row - w, *row2 = (y > 1) ? row1 - w : row1;       // toptop = top on row 1
    int left = zero, leftleft = zero;
    int ctl[kChunk], ctop[kChunk], ctr[kChunk];
    for (int x0 = 0; x0 < w; x0 += kChunk) {
        const int cnt = std::min(kChunk, w - x0);
        for (int i = 0; i < cnt; i++) {
            const int x = x0 + i;
            int *p = pc + (size_t)i * stride, *np = p + nref;
            const int top = row1[x];
            const int tl = x ? row1[x - 1] : zero;
            const int tr = (x + 1 < w) ? row1[x + 1] : top;
            const int tt = row2[x];
            for (int k = 0; k < nused; k++) {
                const int r = used_ref[k];
                const int rv = refrow[(size_t)r * w + x];
                p[2 * r] = fooabs(rv);
                p[2 * r + 1] = slog(rv);
            }
            np[0] = fooabs(top);
            np[2] = slog(top);
            np[4] = y;
            np[5] = x;
            np[6] = top - tl;
            np[7] = tl + tr - top;
            np[9] = slog(tl - top);
            np[10] = slog(top - tr);
            np[11] = slog(top - tt);
            ctl[i] = tl; ctop[i] = top; ctr[i] = tr;
        }
        for (int i = 0; i < cnt; i++) {
            const int x = x0 + i;
            int *p = pc + (size_t)i * stride, *np = p + nref;
            const int tl = ctl[i];
            np[1] = fooabs(left);
            np[3] = slog(left);
            np[6] += left;
            np[8] = slog(left - tl);
            np[12] = slog(left - leftleft);
            const int guess = PRED0 ? zero : predict(predictor, left, ctop[i], tl, ctr[i], zero, cmin, cmax);
            const int mn = cmin - guess, mx = cmax - guess;
            int diff = mn;
            if (mn != mx) {
                const Node n = walk(nodes, p, nodes[1]);
                diff = read_int(rac, table, leaves + ((size_t)node_ref(n) << leaf_shift), mn, mx, mant_base);
            }
            const int val = s16(s16(diff) + guess);
            row[x] = (int16_t)val;
            leftleft = x ? left : val;
            left = val;
        }
    }
}

// ---- large planes: the look-ahead half of the chunked decoder on a helper thread ------------------------------------------------
// What decode_row_chunked computes ahead of its serial loop needs nothing of the current row, only the row above up to x + 1 --
// so for a large plane it moves to a second thread that runs about one row ahead of the decoder, and because it is then off the
// decoder's timeline it also walks the tree ahead: from the root to the first node that tests a `left`-dependent property (3.4 of
// 9.4 levels on average), and from both children of that node on to the next such node or leaf.  The decoder resolves that node
// with one compare and continues from where the pre-walk stopped (measured: 231 -> 200 cycles per symbol in the serial loop, and
// the 45-60 cycles of the look-ahead pass leave its timeline altogether).  Chunk slots live in a ring of (chunks per row + 3)
// entries; `main_px` (pixels decoded, published per chunk) and `ready` (chunks prepared) are the only shared words.
constexpr size_t kAheadMinSamples = 1u << 17;
struct Ahead {
    const Chan *ch = nullptr;
    const Image *img = nullptr;
    const int *refchan = nullptr, *used_ref = nullptr;
    const Node *nodes = nullptr;
    int nused = 0, nref = 0, stride = 0, cpr = 0, K = 0;
    std::vector<int> props, ctl, ctop, ctr;
    std::vector<Node> start, cont;
    std::vector<int16_t> refrow;
    alignas(64) std::atomic<long long> main_px{0};
    alignas(64) std::atomic<long long> ready{0};
    alignas(64) std::atomic<int> quit{0};
};

void ahead_run(Ahead *Ap, int y0) {
    Ahead &A = *Ap;
    const Chan &ch = *A.ch;
    const int w = ch.w, zero = ch.zero, stride = A.stride, nref = A.nref;
    const Node *nodes = A.nodes;
    bool stops[128] = {false};       // properties that depend on the pixel to the left, and the leaf mark: the pre-walk stops there
    stops[nref + 1] = stops[nref + 3] = stops[nref + 6] = stops[nref + 8] = stops[nref + 12] = true;
    stops[kLeafMark] = true;
    auto prewalk = [&](Node n, const int *p) {
        while (!stops[node_prop(n)]) {
            const Node *c = nodes + node_ref(n);
            const Node a = c[0], b = c[1];
            const uint64_t take_a = (uint64_t)0 - (uint64_t)(p[node_prop(n)] > node_split(n));
            n = b ^ ((a ^ b) & take_a);
        }
        return n;
    };
    long long id = 0;
    for (int y = y0; y < ch.h; y++) {
        for (int k = 0; k < A.nused; k++) {
            const int r = A.used_ref[k];
            const Chan &cj = A.img->ch[A.refchan[r]];
            int ry = shr(shl(y, ch.vshift), cj.vshift);
            if (ry >= cj.h) ry = cj.h - 1;
            for (int spins = 0; ld_acquire(&cj.rows_done) < ry + 1; spins++) {
                if (A.quit.load(std::memory_order_relaxed)) return;
                if (spins < 256) FBH_PAUSE(); else std::this_thread::yield();
            }
            reference_row(ch, cj, y, A.refrow.data() + (size_t)r * w);
        }
        const int16_t *row1 = ch.data + (size_t)(y - 1) * w, *row2 = (y > 1) ? row1 - w : row1;
        for (int c = 0; c < A.cpr; c++, id++) {
            const int x0 = c * kChunk, cnt = std::min(kChunk, w - x0);
            const long long need = (long long)(y - 1) * w + std::min(w, x0 + cnt + 1);      // the row above, through the last pixel's topright
            for (int spins = 0; A.main_px.load(std::memory_order_acquire) < need; spins++) {
                if (A.quit.load(std::memory_order_relaxed)) return;
                if (spins < 256) FBH_PAUSE(); else std::this_thread::yield();
            }
            const size_t slot = (size_t)(id % A.K) * kChunk;
            for (int i = 0; i < cnt; i++) {
                const int x = x0 + i;
                int *p = A.props.data() + (slot + i) * stride, *np = p + nref;
                const int top = row1[x];
                const int tl = x ? row1[x - 1] : zero;
                const int tr = (x + 1 < w) ? row1[x + 1] : top;
                const int tt = row2[x];
                for (int k = 0; k < A.nused; k++) {
                    const int r = A.used_ref[k];
                    const int rv = A.refrow[(size_t)r * w + x];
                    p[2 * r] = fooabs(rv);
                    p[2 * r + 1] = slog(rv);
                }
                np[0] = fooabs(top);
                np[2] = slog(top);
                np[4] = y;
                np[5] = x;
                np[6] = top - tl;
                np[7] = tl + tr - top;
                np[9] = slog(tl - top);
                np[10] = slog(top - tr);
                np[11] = slog(top - tt);
                A.ctl[slot + i] = tl; A.ctop[slot + i] = top; A.ctr[slot + i] = tr;
                const Node n = prewalk(nodes[1], p);
                A.start[slot + i] = n;
                if (!node_is_leaf(n)) {
                    const Node *cc = nodes + node_ref(n);
                    A.cont[2 * (slot + i)] = prewalk(cc[0], p);
                    A.cont[2 * (slot + i) + 1] = prewalk(cc[1], p);
                }
            }
            A.ready.store(id + 1, std::memory_order_release);
        }
    }
}

// The decoder's side.  Rows are decoded by decode_row / decode_row_chunked until a hardware thread is free (checked every few
// rows); from then on a helper prepares the chunks and the rows come chunk by chunk from its ring.  Returns when the plane is
// done or the stream has stopped.
template <bool PRED0>
__attribute__((noinline)) void decode_plane_ahead(const Image &img, Chan &ch, int predictor, int nused, const int *used_ref, const int *refchan,
                                                  int nref, int nprops, Rac &rac, const Tables &T, const Node *nodes, uint16_t *leaves,
                                                  int leaf_shift, int mant_base, Scratch &S) {
    const int w = ch.w, zero = ch.zero, cmin = ch.minval, cmax = ch.maxval;
    const uint16_t *table = T.table;
    Ahead A;
    A.ch = &ch; A.img = &img; A.refchan = refchan; A.used_ref = used_ref; A.nodes = nodes;
    A.nused = nused; A.nref = nref; A.stride = (nprops + 3) & ~3; A.cpr = (w + kChunk - 1) / kChunk; A.K = A.cpr + 3;
    const int stride = A.stride;
    S.chunk.resize((size_t)kChunk * stride);
    std::thread helper;
    bool helping = false;
    struct Stop {
        Ahead &a; std::thread &t; const Tables &T; bool &on;
        ~Stop() { a.quit.store(1); if (t.joinable()) t.join(); if (on) T.busy.fetch_sub(1, std::memory_order_relaxed); }
    } stop{A, helper, T, helping};
    long long id = 0;
    for (int y = 0; y < ch.h; y++) {
        if (rac.io.stop()) break;
        if (!helping && y >= 1 && (y & 7) == 1 && ch.h - y >= 16 && T.busy.load(std::memory_order_relaxed) < T.hw_threads) {
            T.busy.fetch_add(1, std::memory_order_relaxed);
            helping = true;
            const size_t npx = (size_t)A.K * kChunk;
            A.props.resize(npx * stride); A.ctl.resize(npx); A.ctop.resize(npx); A.ctr.resize(npx); A.start.resize(npx); A.cont.resize(2 * npx);
            A.refrow.resize((size_t)std::max(1, nused ? used_ref[nused - 1] + 1 : 1) * w);
            A.main_px.store((long long)y * w, std::memory_order_release);
            helper = std::thread(ahead_run, &A, y);
            if (T.debug) fprintf(stderr, "[host entropy] plane %dx%d: look-ahead helper from row %d on (%d threads busy)\n", w, ch.h, y, T.busy.load());
        }
        if (!helping) {
            for (int k = 0; k < nused; k++) {       // row wavefront on the planes this row back-references
                const int r = used_ref[k];
                const Chan &cj = img.ch[refchan[r]];
                int ry = shr(shl(y, ch.vshift), cj.vshift);
                if (ry >= cj.h) ry = cj.h - 1;
                wait_until_ge(&cj.rows_done, ry + 1);
                reference_row(ch, cj, y, S.refrow.data() + (size_t)r * w);
            }
            if (y) decode_row_chunked<PRED0>(ch, y, predictor, nused, used_ref, nref, stride, S.chunk.data(), S.refrow.data(), rac, table, nodes, leaves, leaf_shift, mant_base);
            else decode_row<PRED0>(ch, y, predictor, nused, used_ref, nref, S.refrow.data(), rac, table, nodes, leaves, leaf_shift, mant_base);
            st_release(&ch.rows_done, y + 1);
            continue;
        }
        int16_t *row = ch.data + (size_t)y * w;
        int left = zero, leftleft = zero;
        for (int c = 0; c < A.cpr; c++, id++) {
            const int x0 = c * kChunk, cnt = std::min(kChunk, w - x0);
            for (int spins = 0; A.ready.load(std::memory_order_acquire) <= id; spins++) {
                if (spins < 1024) FBH_PAUSE(); else std::this_thread::yield();
            }
            const size_t slot = (size_t)(id % A.K) * kChunk;
            for (int i = 0; i < cnt; i++) {
                const int x = x0 + i;
                int *p = A.props.data() + (slot + i) * stride, *np = p + nref;
                const int tl = A.ctl[slot + i];
                np[1] = fooabs(left);
                np[3] = slog(left);
                np[6] += left;
                np[8] = slog(left - tl);
                np[12] = slog(left - leftleft);
                const int guess = PRED0 ? zero : predict(predictor, left, A.ctop[slot + i], tl, A.ctr[slot + i], zero, cmin, cmax);
                const int mn = cmin - guess, mx = cmax - guess;
                int diff = mn;
                if (mn != mx) {
                    Node n = A.start[slot + i];
                    if (!node_is_leaf(n)) {
                        const Node a = A.cont[2 * (slot + i)], b = A.cont[2 * (slot + i) + 1];
                        const uint64_t take_a = (uint64_t)0 - (uint64_t)(p[node_prop(n)] > node_split(n));
                        n = walk(nodes, p, b ^ ((a ^ b) & take_a));
                    }
                    diff = read_int(rac, table, leaves + ((size_t)node_ref(n) << leaf_shift), mn, mx, mant_base);
                }
                const int val = s16(s16(diff) + guess);
                row[x] = (int16_t)val;
                leftleft = x ? left : val;
                left = val;
            }
            A.main_px.store((long long)y * w + x0 + cnt, std::memory_order_release);
        }
        st_release(&ch.rows_done, y + 1);
    }
}

bool corrupt_or_truncated(bool stopped, Chan &c) {      // encoding.cpp:209-219: true = "truncated, carry on", false = corruption
    if (stopped) { fill_plane(c, 0); return true; }
    return false;
}

// fuif_decode_channel, encoding.cpp:259-429.  Returns false on a hard error; `beginc` is advanced to the group's last plane.
bool decode_group(Image &img, Reader &io, int &beginc, int limit, const Tables &T, Scratch &S) {
    if (io.stop()) return true;
    const long long header_pos = (long long)io.pos;
    const int firstbyte = io.varint();
    if (io.stop()) return true;
    const int b0 = beginc;
    const int endc = beginc + (firstbyte >> 4);
    const bool compress = firstbyte & 1;
    const int predictor = (firstbyte & 14) >> 1;
    int global_minv = s16(1 - io.varint());
    if (io.stop()) return true;
    if (global_minv == 1) global_minv = s16(io.varint());
    if (io.stop()) return true;
    const int global_maxv = s16(global_minv + io.varint());
    if (io.stop()) return true;
    // `limit`: the planes of this stream end there.  A group that reaches beyond them (a group index that does not belong to the file, a
    // damaged header) would write planes another stream owns -- and could lower their `rows_done` after the owner released it, which
    // leaves every stream that waits for those rows spinning forever.  Corrupt, not garbage.
    if (endc >= img.nch || endc < beginc || endc >= limit) return false;
    img.ch[b0].group_off = header_pos;

    int firstrealc = beginc;
    bool early = false, early_result = true;
    for (int i = beginc; i <= endc; i++) {
        Chan &ch = img.ch[i];
        if (ch.w * ch.h <= 0) continue;
        ch.minval = global_minv; ch.maxval = global_maxv;
        if (endc > beginc && global_minv < global_maxv) {
            ch.minval = s16(ch.minval + io.varint());
            ch.maxval = s16(ch.minval + io.varint());
        }
        if (ch.minval == ch.maxval) { fill_plane(ch, ch.minval); firstrealc++; }
        if (ch.minval == 0 && ch.maxval == 0) continue;
        ch.q = io.varint();
        if (io.stop()) { early = true; early_result = corrupt_or_truncated(true, ch); break; }
        // an inverted range only comes out of a damaged header (the reference runs into its asserts there): corrupt, like a failed depth check
        if (ch.maxval < ch.minval || (compress && !check_bit_depth(ch.minval, ch.maxval, predictor))) { early = true; early_result = false; break; }
    }
    for (int i = beginc; i <= endc; i++) {
        Chan &ch = img.ch[i];
        if (ch.w * ch.h <= 0 || ch.minval == ch.maxval) continue;      // the reference calls setzero() only on planes it decodes
        if (ch.minval > 0) ch.zero = ch.minval; else if (ch.maxval < 0) ch.zero = ch.maxval; else ch.zero = 0;   // setzero, image.h:70-74
    }
    // the ranges of this group's planes are final from here on: let dependent streams read them
    for (int i = beginc; i <= endc; i++) st_release(&img.ch[i].hdr_done, 1);
    if (early) return early_result;
    if (firstrealc > endc) { beginc = endc; return true; }

    if (13 + img.max_properties + 2 > kMaxProps) { img.status = FB_ERR_UNSUPPORTED; return false; }
    int pr[kMaxProps + 16][2];
    int refchan[kMaxProps / 2 + 8], nrefchan = 0;
    const int nprops = init_properties(pr, img, beginc, endc, refchan, nrefchan);
    const int nref = nprops - NB_NONREF;

    int predictability = 2048;
    if (predictor == 0 && compress) {
        const int rounded = io.varint();
        if (rounded < 1 || rounded > 127) return corrupt_or_truncated(io.stop(), img.ch[std::min(firstrealc, img.nch - 1)]);
        predictability = rounded * 32;
    }

    Rac rac;
    rac.init(io);

    if (!compress) {        // encoding.cpp:334-354
        for (int i = beginc; i <= endc; i++) {
            Chan &ch = img.ch[i];
            if (ch.minval == ch.maxval) continue;
            fill_plane(ch, i < img.n_orig ? 0 : ch.zero);
            for (int y = 0; y < ch.h; y++) {
                if (rac.io.stop()) break;
                int16_t *row = ch.data + (size_t)y * ch.w;
                for (int x = 0; x < ch.w; x++) row[x] = (int16_t)uniform_read(rac, ch.minval, ch.maxval - ch.minval);
                st_release(&ch.rows_done, y + 1);
            }
            if (rac.io.stop()) break;
        }
        beginc = endc;
        io = rac.io;
        return true;
    }

    if (!read_tree(rac, T.meta, pr, nprops, S)) {
        const bool stopped = rac.io.stop();
        io = rac.io;
        return corrupt_or_truncated(stopped, img.ch[beginc]);
    }
    // FinalPropertySymbolCoder ctor, compound.h:213-225: leaf numbering in node order, all leaves start from zero_chance
    const int nnodes = (int)S.prop_of.size();
    if (T.debug) fprintf(stderr, "[host entropy] group at %lld: planes %d-%d, predictor %d, %d properties, tree of %d nodes\n", header_pos, beginc, endc, predictor, nprops, nnodes);
    S.nodes.resize((size_t)nnodes + 2);
    Node *nodes = (Node *)(((uintptr_t)S.nodes.data() + 15) & ~(uintptr_t)15);     // slot 0 unused; slots 2k, 2k + 1 share 16 bytes
    int nleaves = 0;
    for (int i = 0; i < nnodes; i++) {
        if (S.prop_of[i] < 0) nodes[i + 1] = make_node(0, ((uint32_t)nleaves++ << 7) | (uint32_t)kLeafMark);
        else nodes[i + 1] = make_node(S.split_of[i], ((uint32_t)(S.child_of[i] + 1) << 7) | (uint32_t)S.prop_of[i]);
    }
    // referenced planes whose properties no node tests cost nothing: no wavefront wait, no row fetch, no property
    int used_ref[kMaxProps / 2 + 8], nused = 0;
    {
        bool used[kMaxProps + 16] = {false};
        for (int i = 0; i < nnodes; i++) if (S.prop_of[i] >= 0) used[S.prop_of[i]] = true;
        for (int r = 0; r < nrefchan; r++) if (used[2 * r] || used[2 * r + 1]) used_ref[nused++] = r;
    }
    // leaf layout: 32 chances (64 bytes), or 16 (32 bytes: twice as many leaves per cache level) when every |residual| of the group
    // fits 8 bits.  With a predictor the residual range is up to twice the value range.
    int group_abs = 0;
    for (int i = beginc; i <= endc; i++) {
        const Chan &c = img.ch[i];
        if (c.minval == c.maxval) continue;
        const int span = predictor ? (c.maxval - c.minval) : std::max(std::abs(c.minval - c.zero), std::abs(c.maxval - c.zero));
        group_abs = std::max(group_abs, span);
    }
    const bool compact = group_abs <= 255;
    const int leaf_shift = compact ? 4 : 5, mant_base = compact ? 9 : SC_MANT;
    uint16_t proto[32];
    for (int e = 0; e < 32; e++) proto[e] = initial_chance(e, predictability);
    if (compact) for (int e = 0; e < 7; e++) proto[9 + e] = proto[SC_MANT + e];       // zero, sign, exp[0..6] stay where they are
    const size_t lstride = (size_t)1 << leaf_shift;
    S.leaves.resize((size_t)nleaves * lstride + 32);
    uint16_t *leaves = (uint16_t *)(((uintptr_t)S.leaves.data() + 63) & ~(uintptr_t)63);
    for (int l = 0; l < nleaves; l++) memcpy(leaves + (size_t)l * lstride, proto, lstride * sizeof(uint16_t));

    for (int i = beginc; i <= endc; i++) {
        Chan &ch = img.ch[i];
        if (ch.minval == ch.maxval) continue;
        // channel.resize(w,h): buffers made by meta_apply start out as `zero`, the Image constructor's as 0
        fill_plane(ch, i < img.n_orig ? 0 : ch.zero);
        if (nnodes == 1 && predictor == 0 && ch.zero == 0) {        // fast track, encoding.cpp:371-383
            for (int y = 0; y < ch.h; y++) {
                if (rac.io.stop()) break;
                int16_t *row = ch.data + (size_t)y * ch.w;
                for (int x = 0; x < ch.w; x++) row[x] = (int16_t)read_int(rac, T.table, leaves, ch.minval, ch.maxval, mant_base);
                st_release(&ch.rows_done, y + 1);
            }
        } else if (T.helpers && (size_t)ch.w * ch.h >= kAheadMinSamples && ch.w >= 2 * kChunk) {
            S.refrow.resize((size_t)std::max(1, nrefchan) * ch.w);
            if (predictor == 0) decode_plane_ahead<true>(img, ch, predictor, nused, used_ref, refchan, nref, nprops, rac, T, nodes, leaves, leaf_shift, mant_base, S);
            else decode_plane_ahead<false>(img, ch, predictor, nused, used_ref, refchan, nref, nprops, rac, T, nodes, leaves, leaf_shift, mant_base, S);
        } else {
            S.refrow.resize((size_t)std::max(1, nrefchan) * ch.w);
            const int stride = (nprops + 3) & ~3;
            S.chunk.resize((size_t)kChunk * stride);
            for (int y = 0; y < ch.h; y++) {
                if (rac.io.stop()) break;
                for (int k = 0; k < nused; k++) {       // row wavefront on the planes this row back-references
                    const int r = used_ref[k];
                    const Chan &cj = img.ch[refchan[r]];
                    int ry = shr(shl(y, ch.vshift), cj.vshift);
                    if (ry >= cj.h) ry = cj.h - 1;
                    wait_until_ge(&cj.rows_done, ry + 1);
                    reference_row(ch, cj, y, S.refrow.data() + (size_t)r * ch.w);
                }
                if (y) {
                    if (predictor == 0) decode_row_chunked<true>(ch, y, predictor, nused, used_ref, nref, stride, S.chunk.data(), S.refrow.data(), rac, T.table, nodes, leaves, leaf_shift, mant_base);
                    else decode_row_chunked<false>(ch, y, predictor, nused, used_ref, nref, stride, S.chunk.data(), S.refrow.data(), rac, T.table, nodes, leaves, leaf_shift, mant_base);
                } else if (predictor == 0) decode_row<true>(ch, y, predictor, nused, used_ref, nref, S.refrow.data(), rac, T.table, nodes, leaves, leaf_shift, mant_base);
                else decode_row<false>(ch, y, predictor, nused, used_ref, nref, S.refrow.data(), rac, T.table, nodes, leaves, leaf_shift, mant_base);
                st_release(&ch.rows_done, y + 1);
            }
        }
        if (rac.io.stop()) break;
    }
    beginc = endc;
    io = rac.io;
    return true;
}

// the channel loop of fuif_decode, encoding.cpp:708-718, for the planes of one stream
void run_stream(Image *images, const Stream &st, const Tables &T, Scratch &S) {
    Image &img = images[st.image];
    Reader io{img.bytes, img.nbytes, st.offset, img.bytes_to_load, false};
    int groups = 0;
    try {
        for (int i = st.first_channel; i < img.nch; i++) {
            if (st.max_groups >= 0 && groups >= st.max_groups) break;
            if ((img.bytes_to_load == 0 || io.pos < img.bytes_to_load) && !io.eof) {
                if (!img.ch[i].w || !img.ch[i].h) continue;
                const bool ok = decode_group(img, io, i, st.end_channel, T, S);
                groups++;
                if (!ok) {
                    int expected = 0;
                    __atomic_compare_exchange_n(&img.status, &expected, (int)FB_ERR_INVALID, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
                    break;
                }
            } else break;
        }
    } catch (...) {     // out of memory for a tree of a damaged file: an exception must not leave a worker thread
        __atomic_store_n(&img.status, (int)FB_ERR_NOMEM, __ATOMIC_RELAXED);
    }
    // whatever happened (truncation, corruption, an exception), nobody may wait forever on this stream's planes
    for (int c = st.first_channel; c < st.end_channel && c < img.nch; c++) {
        st_release(&img.ch[c].hdr_done, 1);
        st_release(&img.ch[c].rows_done, 0x7fffffff);
    }
}

// hardware threads this process may run on (its affinity mask: a cpuset-restricted container has fewer than the machine)
int hardware_threads() {
#ifdef __linux__
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) return CPU_COUNT(&set);
#endif
    return std::max(1, (int)std::thread::hardware_concurrency());
}

}  // namespace

int decode(Image *images, int nimages, const Stream *streams, int nstreams, int cutoff, uint32_t alpha, int threads) {
    (void)nimages;
    if (nstreams <= 0) return 0;
    const int hw = hardware_threads();
    std::vector<Tables> tables(1);
    build_table(tables[0].table, alpha, (unsigned)(4096 - cutoff));
    build_table(tables[0].meta, 0xFFFFFFFFu / 19, 4096 - 2);
    tables[0].helpers = threads != 1;
    tables[0].debug = getenv("FB_HOST_DEBUG") != nullptr;
    tables[0].hw_threads = hw;
    tables[0].busy.store(0);
    // Default: up to four threads per hardware thread.  Streams are claimed in index order (the dependency rule), which puts the
    // largest groups of a file last; with more threads than cores they are all claimed at once and the OS shares the cores out
    // until the small ones are gone (4096^2, 61 groups, 8 cores: 2.7 s with 8 threads, 2.1 s with 16 or 61).  Waiting threads yield.
    if (threads <= 0) threads = 4 * hw;
    threads = std::max(1, std::min(threads, nstreams));
    std::atomic<int> ticket{0};
    auto worker = [&]() {
        Scratch S;
        for (;;) {
            const int sid = ticket.fetch_add(1, std::memory_order_relaxed);
            if (sid >= nstreams) break;
            tables[0].busy.fetch_add(1, std::memory_order_relaxed);
            run_stream(images, streams[sid], tables[0], S);
            tables[0].busy.fetch_sub(1, std::memory_order_relaxed);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    return threads;
}

}  // namespace fbh
