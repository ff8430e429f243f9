// MANIAC entropy decoding + context model on the GPU (sm_100a).
//
// What runs here is fuif_decode_channel (reference encoding/encoding.cpp:259-429) with everything it inlines:
//   group header varints                      encoding.cpp:264-329
//   init_properties / predictors / properties encoding/context_predict.h:67-120, 125-206
//   precompute_references                     context_predict.h:233-289
//   24-bit range decoder                      maniac/rac.h:35-114
//   adaptive 12-bit chances                   maniac/chance.h:42-84, chance.cpp:31-65 (table built on the host)
//   zero/sign/exponent/mantissa integer coder maniac/symbol.h:72-185, uniform coder :44-56
//   MANIAC tree parse + leaf walk             maniac/compound.h:135-320
//
// Mapping.  The bitstream is strictly serial inside one channel group (adaptive chances + neighbour
// context), so the unit of parallelism is a *stream*: one channel group when the caller supplies
// the groups' byte offsets (group index), otherwise one whole image.  One warp decodes one stream: the serial
// decode is executed warp-uniformly (no divergence), while the data-parallel pieces (reference-property
// rows, leaf initialisation, constant fills) are spread over the 32 lanes.  Streams are handed out through
// an atomic ticket so that a stream only ever waits for lower-numbered streams (row wavefront on the
// planes it back-references), which are guaranteed to be running already.
#include "fb_common.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace {

constexpr int kMaxNodes = 65536;
constexpr int kMaxProps = 64;
constexpr int NB_NONREF = 13;           // context_predict.h:210
constexpr int MAX_BIT_DEPTH = 15;       // config.h:5

struct DChan {
    int w, h, minval, maxval, zero, q, hshift, vshift;
    int16_t *data;
    int state;          // 0 untouched, 1 holds samples
    int hdr_done;       // group header of this channel has been parsed (ranges valid)
    int rows_done;      // rows published so far (wavefront)
    long long group_off;    // byte offset of the group header if this channel starts a group, else -1
};

struct DImage {
    const uint8_t *bytes;
    unsigned long long nbytes, bytes_to_load;
    DChan *ch;
    int nch, max_properties;
    int n_orig;         // channels the Image constructor created with zero-filled buffers (encoding.cpp:637)
    int status;         // 0 ok, else FB_ERR_*
};

struct DStream {
    int image, first_channel, end_channel, max_groups;     // channels [first_channel, end_channel) belong to this stream
    unsigned long long offset;
};

struct TNode { short property; unsigned short child; int splitval; };   // PropertyDecisionNode, compound.h:41-51

struct WarpScratch {    // per resident warp, in global memory
    TNode *nodes;           // kMaxNodes
    uint16_t *leaves;       // kMaxNodes/2 * 32
    int *stack;             // tree-parse frames, 4 ints each, kMaxNodes/2+2 frames
    int16_t *refs;          // maxw * 12
};

struct Params {
    DImage *images;
    DStream *streams;
    int nstreams;
    int *ticket;
    const uint16_t *table;      // [4096][2] decode table (cutoff, alpha)
    const uint16_t *meta_table; // [4096][2] tree-coder table (cut 2, alpha 0xFFFFFFFF/19)
    WarpScratch *scratch;
    int maxw;
    int warp_smem;          // bytes of shared memory per stream slot (after the block's 16 KiB chance table)
    int helpers;            // walker warps per stream (0 or kMaxWalkers)
    int debug;              // FB_MANIAC_DEBUG=1: trace group headers from lane 0
};

__device__ __forceinline__ int s16(int x) { return (int)(short)x; }
__device__ __forceinline__ int ilog2u(unsigned l) { return l == 0 ? 0 : 31 - __clz(l); }      // maniac/util.h:33-36

__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- byte reader with FileIO end-of-stream rules (fileio.h:33-81) ------------------------------------------
// `pos` counts consumed bytes (what ftell() reports); `win` caches up to four upcoming bytes, MSB first.
struct Reader {
    const uint8_t *p;
    unsigned long long n, pos, btl;
    unsigned win;
    int avail;
    bool eof;
    __device__ __forceinline__ void seek(unsigned long long to) { pos = to; avail = 0; win = 0; }
    __device__ __forceinline__ void refill() {
        const unsigned long long addr = (unsigned long long)(p + pos);
        const unsigned w = __ldg((const unsigned *)(addr & ~3ull));
        const int skip = (int)(addr & 3ull);
        win = __byte_perm(w, 0, 0x0123) << (8 * skip);
        avail = 4 - skip;
    }
    __device__ __forceinline__ int get() {
        if (pos >= n) { eof = true; return -1; }
        if (avail == 0) refill();
        const int b = (int)(win >> 24);
        win <<= 8;
        avail--;
        pos++;
        return b;
    }
    __device__ __forceinline__ bool stop() const { return eof || (btl && pos >= btl); }
    __device__ int varint() {           // read_big_endian_varint, encoding.cpp:45-59
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

// ---- range decoder (maniac/rac.h) -----------------------------------------------------------------------------
struct Rac {
    Reader io;          // by value: the whole coder state lives in registers
    unsigned range, low;
    bool ones;      // a read past the end turned `low` into all-ones garbage (rac.h:64-69): every bit reads as 1
    __device__ __forceinline__ void byte_in() {
        int c = io.get();
        if (c < 0) ones = true;
        low = (low << 8) | (unsigned)(c & 0xFF);
    }
    __device__ __forceinline__ void init(const Reader &r) {           // RacInput ctor, rac.h:97-104
        io = r; range = 1u << 24; low = 0; ones = false;
        byte_in(); byte_in(); byte_in();
    }
    __device__ __forceinline__ void input() {   // rac.h:70-81
        if (range <= (1u << 16)) {
            range <<= 8; byte_in();
            if (range <= (1u << 16)) { range <<= 8; byte_in(); }
        }
    }
    __device__ __forceinline__ int get(unsigned chance) {       // rac.h:82-95
        int bit;
        const unsigned thr = range - chance;
        if (ones || low >= thr) { low -= thr; range = chance; bit = 1; }
        else { range = thr; bit = 0; }
        input();
        return bit;
    }
    __device__ __forceinline__ int read12(unsigned b12) {       // rac.h:42-52, 107 (64-bit product)
        return get((unsigned)(((unsigned long long)range * b12 + 0x800ull) >> 12));
    }
    __device__ __forceinline__ int read_bit() { return get(range >> 1); }   // rac.h:111
};

// ---- symbol coder (maniac/symbol.h) -------------------------------------------------------------------------------
// leaf layout: [0]=zero [1]=sign [2..15]=exp[14] [16..30]=mant[15] [31]=pad   (SymbolChance<.,15>, symbol.h:72-139)
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

// The serial coder runs in ONE lane (lane 0): adaptive chances are read-modify-written in place (shared memory),
// so there is exactly one reader/writer and no lockstep assumption between lanes.
//
// SymCtx implements FinalCompoundSymbolBitCoder::read (compound.h:90-95): decode with chance lp[idx], then
// lp[idx] = table[chance][bit].  Both successors of the chance are fetched as one 32-bit word right after the chance
// itself, so the table lookup overlaps the range arithmetic instead of following it.
struct SymCtx {
    const unsigned *tab32;      // newchance[4096][2] viewed as 4096 words: low half = successor after a 0, high half after a 1
    uint16_t *lp;
    __device__ __forceinline__ void begin(const uint16_t *t, uint16_t *leaf) { tab32 = reinterpret_cast<const unsigned *>(t); lp = leaf; }
    __device__ __forceinline__ int read(Rac &rac, int idx) {
        const unsigned ch = lp[idx];
        const unsigned both = tab32[ch];        // issued before the bit is known: off the coder's dependency chain
        const int bit = rac.read12(ch);
        lp[idx] = (uint16_t)(bit ? (both >> 16) : (both & 0xffffu));
        return bit;
    }
    __device__ __forceinline__ void end() {}
};

// reader<15>(coder, min, max), symbol.h:154-185
// mant_base: index of bit_mant[0] in the leaf (16 in the full layout, 9 in the compact one, see LeafStore)
__device__ __forceinline__ int read_int(Rac &rac, const uint16_t *__restrict__ table, uint16_t *leaf, int mn, int mx, int mant_base = SC_MANT) {
    if (mn == mx) return mn;
    SymCtx c;
    c.begin(table, leaf);
    int result;
    if (c.read(rac, SC_ZERO)) result = 0;
    else {
        int sign;
        if (mn < 0) { if (mx > 0) sign = c.read(rac, SC_SIGN); else sign = 0; } else sign = 1;
        const int amax = sign ? mx : -mn;
        const int emax = ilog2u((unsigned)amax);
        int e = 0;
        for (; e < emax; e++) if (c.read(rac, SC_EXP + e)) break;
        int have = 1 << e;
        for (int pos = e; pos > 0;) {
            pos--;
            const int minabs1 = have | (1 << pos);
            if (minabs1 > amax) continue;
            if (c.read(rac, mant_base + pos)) have = minabs1;
        }
        result = sign ? have : -have;
    }
    c.end();
    return result;
}
__device__ int read_int2(Rac &rac, const uint16_t *table, uint16_t *leaf, int mn, int mx) {     // symbol.h:232-236
    if (mn > 0) return read_int(rac, table, leaf, 0, mx - mn) + mn;
    if (mx < 0) return read_int(rac, table, leaf, mn - mx, 0) + mx;
    return read_int(rac, table, leaf, mn, mx);
}
__device__ int uniform_read(Rac &rac, int mn, int len) {    // UniformSymbolCoder::read_int, symbol.h:44-56
    while (len != 0) {
        int med = len / 2;
        if (rac.read_bit()) { mn = mn + med + 1; len = len - (med + 1); }
        else len = med;
    }
    return mn;
}

// ---- context model (encoding/context_predict.h) ---------------------------------------------------------------------
__device__ __forceinline__ int slog(int x16) {      // context_predict.h:54-61 (branch-free: 32 - clz(0) == 0)
    const int x = s16(x16);
    const int b = 32 - __clz(abs(x));
    return x < 0 ? -b : b;
}
__device__ __forceinline__ int fooabs(int x16) { int x = s16(x16); return s16(x < 0 ? -x : x); }   // :63-65

__device__ __forceinline__ int median3(int a, int b, int c) {       // util.h:9-23
    if (a < b) { if (b < c) return b; return a < c ? c : a; }
    if (a < c) return a;
    return b < c ? c : b;
}

__device__ bool check_bit_depth(int minv, int maxv, int predictor) {    // encoding.cpp:61-72
    int maxav = s16(abs(maxv));
    if (-minv > maxav) maxav = s16(-minv);
    if (predictor > 0 && maxv - minv > maxav) maxav = s16(maxv - minv);
    if (predictor > 0 && abs(minv - maxv) > maxav) maxav = s16(abs(minv - maxv));
    return ilog2u((unsigned)maxav) + 1 <= MAX_BIT_DEPTH;
}

// Poll with relaxed loads that bypass L1 (no cache invalidation per poll); one acquire once the value is there.
__device__ __forceinline__ void spin_until_ge(const int *flag, int want) {
    if (__ldcg(flag) >= want) { (void)ld_acquire(flag); return; }
    while (__ldcg(flag) < want) __nanosleep(64);
    (void)ld_acquire(flag);
}

// All lanes of the warp call this with identical arguments.
__device__ void fill_plane(DChan &c, int value, int lane) {
    const size_t n = (size_t)c.w * c.h;
    for (size_t i = lane; i < n; i += 32) c.data[i] = (int16_t)value;
    __syncwarp();
    c.state = 1;
}

// init_properties, context_predict.h:67-120
__device__ int init_properties(int (*pr)[2], DImage &img, int beginc, int endc, int *refchan, int &nrefchan) {
    int n = 0, offset = 0;
    nrefchan = 0;
    for (int j = beginc - 1; j >= 0 && offset < img.max_properties; j--) {
        spin_until_ge(&img.ch[j].hdr_done, 1);
        const DChan &cj = img.ch[j];
        const int cmin = __ldcg(&cj.minval), cmax = __ldcg(&cj.maxval);
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
        const DChan &cj = img.ch[j];
        if (cj.minval < minval) minval = cj.minval;
        if (cj.maxval > maxval) maxval = cj.maxval;
        if (cj.h > maxh) maxh = cj.h;
        if (cj.w > maxw) maxw = cj.w;
    }
    if (minval > 0) minval = 0;
    if (maxval < 0) maxval = 0;
    int amax = max(fooabs(minval), fooabs(maxval));
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

// MetaPropertySymbolCoder::read_tree, compound.h:277-320, recursion unrolled on an explicit stack.  Lane 0 only.
__device__ bool read_tree(Rac &rac, const uint16_t *__restrict__ mtable, int (*range)[2], int nprops, TNode *nodes, int &nnodes,
                          int *stack, uint16_t (*coder)[32]) {
    int sub[kMaxProps][2];
    for (int i = 0; i < nprops; i++) { sub[i][0] = range[i][0]; sub[i][1] = range[i][1]; }
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 32; i++) coder[k][i] = initial_chance(i, 1024);     // SimpleSymbolCoder ctx(ZERO_CHANCE), symbol.h:219
    nnodes = 1;
    nodes[0].property = -1; nodes[0].child = 0; nodes[0].splitval = 0;
    int sp = 0;
    // frame = {pos, stage | p<<2, oldmin, oldmax}; splitval lives in the node
    stack[0] = 0; stack[1] = 0; stack[2] = 0; stack[3] = 0;
    sp = 1;
    while (sp > 0) {
        int *f = stack + 4 * (sp - 1);
        const int pos = f[0], stage = f[1] & 3, p = f[1] >> 2;
        if (stage == 0) {
            int pp = read_int2(rac, mtable, coder[0], 0, nprops) - 1;
            nodes[pos].property = (short)pp;
            if (pp == -1) { sp--; continue; }
            int oldmin = sub[pp][0], oldmax = sub[pp][1];
            if (oldmin >= oldmax) return false;                                     // "Invalid tree", compound.h:285-288
            int splitval = read_int2(rac, mtable, coder[2], oldmin, oldmax - 1);
            nodes[pos].splitval = splitval;
            if (nnodes + 2 > 65535) return false;
            int child = nnodes;
            nodes[pos].child = (unsigned short)child;
            nodes[child].property = -1; nodes[child].child = 0; nodes[child].splitval = 0;
            nodes[child + 1] = nodes[child];
            nnodes += 2;
            sub[pp][0] = splitval + 1;
            f[1] = 1 | (pp << 2); f[2] = oldmin; f[3] = oldmax;
            int *g = stack + 4 * sp;
            g[0] = child; g[1] = 0; g[2] = 0; g[3] = 0;
            sp++;
        } else if (stage == 1) {
            sub[p][0] = f[2];
            sub[p][1] = nodes[pos].splitval;
            f[1] = 2 | (p << 2);
            int *g = stack + 4 * sp;
            g[0] = nodes[pos].child + 1; g[1] = 0; g[2] = 0; g[3] = 0;
            sp++;
        } else {
            sub[p][1] = f[3];
            sp--;
        }
    }
    return true;
}

__device__ __forceinline__ int predict(int predictor, int left, int top, int topleft, int topright, int zero, int cmin, int cmax) {     // context_predict.h:157-166
    switch (predictor) {
    case 0: return zero;
    case 1: return s16((left + top) / 2);
    case 2: return median3(s16(left + top - topleft), left, top);
    case 3: return left;
    case 4: return top;
    case 5: return s16((left + topleft + top + topright) / 4);
    case 6: { int g = left + top - topleft; return s16(g < cmin ? cmin : (g > cmax ? cmax : g)); }
    default: return median3(s16(left + top - topleft), left, top);
    }
}

__device__ __forceinline__ void publish_rows(DChan &c, int rows, int lane) {
    __syncwarp();
    if (lane == 0) st_release(&c.rows_done, rows);
}

// corrupt_or_truncated, encoding.cpp:209-219.  returns true = "truncated, carry on", false = corruption
__device__ bool corrupt_or_truncated(bool stopped, DChan &c, int lane) {
    if (stopped) { fill_plane(c, 0, lane); return true; }
    return false;
}

struct Mail;
// Per-warp shared memory.
constexpr int kPropStride = 37;     // 32 property slots (one per lane) + top, topleft, topright + padding; odd => conflict-free rows
struct Smem {
    const uint16_t *table;  // [4096][2], shared by the warps of the block
    uint16_t (*coder)[32];  // 3 x 32 tree-coder chances
    int *cprop;             // [32][kPropStride]: per pixel of the current chunk: properties by lane, then top / topleft / topright
    unsigned char *dyn;     // dynamic region: tree-node cache, then leaf chances (resident or direct-mapped cache)
    int dyn_bytes;
    struct Mail *mail;      // mailbox shared with the walker warps (nullptr: no walkers)
    int *ldrows;            // [walkers][32][kLdRowStride] per-lane property values of the walkers
    int nwalkers;
};

// Node cache entry (8 bytes): x = property << 16 | slot16 (uint4 index of the child pair) for inner nodes,
// x = 0xFFFF0000 | leaf id for leaves (sign bit set);  y = split value.  Slot i+1 holds node i so that sibling pairs
// (odd node index, next) share one 16-byte line.
// Leaf layout: full = 32 chances (zero, sign, exp[14], mant[15], pad); compact = 16 chances (zero, sign, exp[7], mant[7]) when
// no value of the group can need more than 7 exponent / mantissa bits (value range <= 255) -- twice as many leaves on-chip.
struct LeafStore {
    int shift;              // log2(chances per leaf): 5 or 4
    int mant_base;          // 16 or 9
    uint16_t *lines;        // shared memory: nlines x (1 << shift) chances
    int *tags;              // direct-mapped tags (nullptr when every leaf is resident)
    int mask;               // nlines - 1
    uint16_t *gleaves;      // global backing store
};

// Returns the shared-memory address of leaf `leaf`'s 32 chances; all lanes cooperate on a miss.
__device__ __forceinline__ uint16_t *leaf_lookup(const LeafStore &ls, int leaf, int lane) {
    if (!ls.tags) return ls.lines + ((size_t)leaf << ls.shift);
    const int slot = leaf & ls.mask;
    const int tag = ls.tags[slot];
    uint16_t *line = ls.lines + ((size_t)slot << ls.shift);
    const int words = 1 << (ls.shift - 1);
    if (tag != leaf) {
        __syncwarp();       // lane 0's chance updates of the line being evicted are visible to the copying lanes
        if (lane < words) {
            unsigned *s = reinterpret_cast<unsigned *>(line);
            if (tag >= 0) reinterpret_cast<unsigned *>(ls.gleaves + ((size_t)tag << ls.shift))[lane] = s[lane];
            s[lane] = reinterpret_cast<const unsigned *>(ls.gleaves + ((size_t)leaf << ls.shift))[lane];
        }
        if (lane == 0) ls.tags[slot] = leaf;
        __syncwarp();
    }
    return line;
}

// One row of a channel in the "slow track" (encoding.cpp:388-421), 32 pixels at a time.
// Lane roles: lane k < nref holds reference property k, lane nref+j holds non-reference property j (0..12).
// Returns through `rac` (meaningful in lane 0 only).
__device__ __forceinline__ uint4 lds128(unsigned addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// Chunk prologue (lanes = the 32 pixels x0..x0+31 of row y): neighbours from the rows above, reference properties
// (precompute_references, context_predict.h:233-289) and every property that does not depend on `left`, written to
// cprop[lane][...].  All lanes of the decoder warp call it.
__device__ __forceinline__ void chunk_prologue(DImage &img, const DChan &ch, int y, int x0, const int *refchan, int nrefchan, int nref,
                                               int *cprop, int lane) {
    const int w = ch.w, zero = ch.zero;
    const int x = x0 + lane;
    if (x < w) {
        const int16_t *row1 = ch.data + (size_t)(y - 1) * w, *row2 = row1 - w;
        int T1 = zero, TL = zero, TR = zero, TT = zero;
        if (y) {
            T1 = row1[x];
            TL = x ? row1[x - 1] : zero;
            TR = (x + 1 < w) ? row1[x + 1] : T1;
            TT = (y > 1) ? row2[x] : T1;
        }
        int *pp = cprop + lane * kPropStride;
        for (int r = 0; r < nrefchan; r++) {
            const DChan &cj = img.ch[refchan[r]];
            int ry = (y << ch.vshift) >> cj.vshift;
            if (ry >= cj.h) ry = cj.h - 1;
            int rx;
            if (ch.hshift == cj.hshift && w <= cj.w) rx = x;
            else if (ch.hshift < cj.hshift) {
                const int stepsize = (1 << cj.hshift) >> ch.hshift;     // all samples but the last are repeated stepsize times
                rx = stepsize > 0 ? x / stepsize : cj.w - 1;
                if (rx > cj.w - 1) rx = cj.w - 1;
            } else {
                rx = (x << ch.hshift) >> cj.hshift;
                if (rx >= cj.w) rx = cj.w - 1;
            }
            const int v = __ldcg(cj.data + (size_t)ry * cj.w + rx);
            pp[2 * r] = fooabs(v);
            pp[2 * r + 1] = slog(v);
        }
        pp[nref + 0] = fooabs(T1);
        pp[nref + 2] = slog(T1);
        pp[nref + 4] = y;
        pp[nref + 5] = x;
        pp[nref + 10] = slog(T1 - TR);
        pp[nref + 11] = slog(T1 - TT);
        pp[32] = T1; pp[33] = TL; pp[34] = TR;
    }
}

// ---- walker warps ------------------------------------------------------------------------------------------------
// The tree walk of pixel x+1 depends on pixel x only through `left` (and on x-1 through `leftleft`).  While lane 0 of
// the decoder warp is busy with the bits of pixel x, up to two walker warps evaluate the MANIAC tree of pixel x+1 for
// EVERY value pixel x can take (one candidate per lane: cmin + lane, cmin + 32 + lane), and leave the leaf ids in shared
// memory.  The decoder then needs one table lookup instead of property evaluation + tree walk on its serial chain.
// Used when the group has predictor 0, a tree of more than one node cached in shared memory and a value range <= 64.
struct Mail {
    volatile int cmd_seq;       // bumped by the decoder for every command
    int cmd;                    // 1 = row, 2 = exit
    int y, w, zero, cmin, cmax, nref, nwalk;
    unsigned nodes_saddr;
    volatile int go;            // candidates of pixels <= go may be computed
    volatile int done[8];       // walker k has published the candidates of pixels < done[k]
    int cval[64];               // decoded values of the current row, ring indexed by x & 63
    unsigned short cand[2][256];// candidate leaf ids of pixel x in cand[x & 1]
};
constexpr int kMailBytes = 1536;
constexpr int kMaxWalkers = 8;  // value ranges up to 256

#define COMPILER_FENCE() asm volatile("" ::: "memory")
__device__ __forceinline__ int lds32(unsigned addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
constexpr int kLdRowStride = 9;     // words per walker lane: its 7 left-dependent property values (+ padding)

// Shared-memory accesses of one warp are performed in program order and there is no cache between the warps of a block,
// so the flag protocol below needs compiler barriers only (no MEMBAR per pixel).
__device__ void walker_main(Mail *mail, const int *cprop2 /* [2][32][kPropStride] */, int *ldrows, int widx, int lane) {
    int seen = 0;
    int *myrow = ldrows + (widx * 32 + lane) * kLdRowStride;
    const unsigned ld_s = (unsigned)__cvta_generic_to_shared(myrow) - 64 * 4;      // biased: offsets 64.. address this row
    for (;;) {
        while (mail->cmd_seq == seen) __nanosleep(32);
        seen = mail->cmd_seq;
        __threadfence_block();
        if (mail->cmd == 2) return;
        const int y = mail->y, w = mail->w, cmin = mail->cmin, cmax = mail->cmax;
        if (widx >= mail->nwalk) continue;
        const unsigned nodes_saddr = mail->nodes_saddr;
        const int cl = cmin + 32 * widx + lane;         // this lane's candidate for `left`
        const bool valid = cl <= cmax;
        const int q1 = fooabs(cl), q3 = slog(cl);
        for (int j = 0; j < w; j++) {
            while (mail->go < j) { }
            COMPILER_FENCE();
            const int *pp = cprop2 + (((j >> 5) & 1) * 32 + (j & 31)) * kPropStride;
            const unsigned pp_s = (unsigned)__cvta_generic_to_shared(pp);
            const int top = pp[32], topright = pp[34];
            const int topleft = (j && y) ? pp[33] : cl;
            const int leftleft = (j > 1) ? ((volatile int *)mail->cval)[(j - 2) & 63] : cl;
            myrow[0] = q1; myrow[1] = q3; myrow[2] = cl + top - topleft; myrow[3] = topleft + topright - top;
            myrow[4] = slog(cl - topleft); myrow[5] = slog(topleft - top); myrow[6] = slog(cl - leftleft);
            COMPILER_FENCE();       // the asm loads below read these
            const uint4 r0 = lds128(nodes_saddr);
            uint2 cur = make_uint2(r0.z, r0.w);
            while ((int)cur.x >= 0) {
                const uint4 pair = lds128(nodes_saddr + ((cur.x & 0xffffu) << 4));
                const unsigned off = cur.x >> 16;       // < 64: word of the shared property row, >= 64: word of this lane's row
                const int v = lds32((off >= 64u ? ld_s : pp_s) + off * 4);
                cur = (v > (int)cur.y) ? make_uint2(pair.x, pair.y) : make_uint2(pair.z, pair.w);
            }
            if (valid) ((volatile unsigned short *)mail->cand[j & 1])[cl - cmin] = (unsigned short)(cur.x & 0xffffu);
            COMPILER_FENCE();
            __syncwarp();
            if (lane == 0) mail->done[widx] = j + 1;
        }
    }
}

// One row with walker warps (PRED0, nodes in shared memory).  Same results as decode_row.
__device__ __forceinline__ void decode_row_helped(DImage &img, DChan &ch, int y, const int *refchan, int nrefchan, int nref, Rac &rac,
                                                  const Smem &sm, const LeafStore &ls, Mail *mail, int nwalk, int lane) {
    const int w = ch.w;
    int16_t *row = ch.data + (size_t)y * w;
    const int zero = ch.zero, cmin = ch.minval, cmax = ch.maxval;
    const int mn = cmin - zero, mx = cmax - zero;           // predictor 0: guess = zero
    chunk_prologue(img, ch, y, 0, refchan, nrefchan, nref, sm.cprop, lane);
    if (w > 32) chunk_prologue(img, ch, y, 32, refchan, nrefchan, nref, sm.cprop + 32 * kPropStride, lane);
    __syncwarp();
    if (lane == 0) {
        mail->y = y; mail->w = w; mail->zero = zero; mail->cmin = cmin; mail->cmax = cmax; mail->nref = nref; mail->nwalk = nwalk;
        mail->go = 0; mail->cmd = 1;
        for (int k = 0; k < kMaxWalkers; k++) mail->done[k] = 0;
        __threadfence_block();
        mail->cmd_seq = mail->cmd_seq + 1;
    }
    int left = zero;
    for (int x0 = 0; x0 < w; x0 += 32) {
        if (x0 > 0 && x0 + 32 < w) {      // properties of the chunk after this one (its buffer is free: nobody reads chunk x0-32 any more)
            chunk_prologue(img, ch, y, x0 + 32, refchan, nrefchan, nref, sm.cprop + (((x0 >> 5) + 1) & 1) * 32 * kPropStride, lane);
            __syncwarp();
        }
        int outv = 0;
        const int cnt = min(32, w - x0);
        for (int i = 0; i < cnt; i++) {
            const int xx = x0 + i;
            COMPILER_FENCE();
            if (lane == 0) mail->go = xx + 1;                                    // candidates of pixel xx+1 may start (val(xx-1) is in cval)
            const int k = (left - cmin) >> 5;
            while (mail->done[k] < xx + 1) { }
            COMPILER_FENCE();
            const int leaf = ((volatile unsigned short *)mail->cand[xx & 1])[left - cmin];
            uint16_t *lp = leaf_lookup(ls, leaf, lane);
            int diff = mn;
            if (lane == 0) diff = read_int(rac, sm.table, lp, mn, mx, ls.mant_base);
            diff = __shfl_sync(0xffffffffu, diff, 0);
            const int val = s16(s16(diff) + zero);
            if (lane == 0) ((volatile int *)mail->cval)[xx & 63] = val;
            outv = (lane == i) ? val : outv;
            left = val;
        }
        if (x0 + lane < w) row[x0 + lane] = (int16_t)outv;
        __syncwarp();
    }
}

template <bool NODES_SMEM, bool PRED0>
__device__ __forceinline__ void decode_row(DImage &img, DChan &ch, int y, int predictor, const int *refchan, int nrefchan, int nref,
                                           Rac &rac, const Smem &sm, const uint2 *nodes2, const LeafStore &ls, int lane) {
    const unsigned nodes_saddr = NODES_SMEM ? (unsigned)__cvta_generic_to_shared(nodes2) : 0u;
    const int w = ch.w;
    int16_t *row = ch.data + (size_t)y * w;
    const int role = lane - nref;
    const int zero = ch.zero, cmin = ch.minval, cmax = ch.maxval;
    // lane role -> operands / function of its left-dependent property (roles 1,3,6,7,8,9,12)
    const bool is_ld = role == 1 || role == 3 || role == 6 || role == 7 || role == 8 || role == 9 || role == 12;
    const bool selA1 = role == 6, selA2 = role == 7, selA3 = role == 9;
    const bool selB1 = role == 6 || role == 8, selB2 = role == 7 || role == 9, selB3 = role == 12;
    const bool mode_abs = role == 1, mode_id = role == 6 || role == 7;
    int left = zero, leftleft = zero;
    for (int x0 = 0; x0 < w; x0 += 32) {
        const int x = x0 + lane;
        chunk_prologue(img, ch, y, x0, refchan, nrefchan, nref, sm.cprop, lane);
        __syncwarp();
        int outv = 0;
        const int cnt = min(32, w - x0);
        for (int i = 0; i < cnt; i++) {
            const int xx = x0 + i;
            const int *pp = sm.cprop + i * kPropStride;
            int mine = pp[lane];
            const int top = pp[32];
            const int topleft = (xx && y) ? pp[33] : left;         // context_predict.h:128
            const int topright = pp[34];
            // The seven properties that depend on the pixel just decoded (context_predict.h:135-154) all have the form
            // f(A - B): |left-0|, slog(left-0), (left+top)-topleft, (topleft+topright)-top, slog(left-topleft),
            // slog(topleft-top), slog(left-leftleft).  Each lane evaluates only ITS expression: operands are picked with
            // loop-invariant per-lane predicates, f with a per-lane mode.
            {
                const int u1 = left + top, u2 = topleft + topright;
                int A = selA1 ? u1 : left;  A = selA2 ? u2 : A;      A = selA3 ? topleft : A;
                int B = selB1 ? topleft : 0; B = selB2 ? top : B;     B = selB3 ? leftleft : B;
                const int t = A - B;
                const int t16 = s16(t), ab = abs(t16);
                int sl = 32 - __clz(ab);
                sl = t16 < 0 ? -sl : sl;
                int r_ = mode_abs ? s16(ab) : sl;
                r_ = mode_id ? t : r_;
                mine = is_ld ? r_ : mine;
            }
            const int guess = PRED0 ? zero : predict(predictor, left, top, topleft, topright, zero, cmin, cmax);
            const int mn = cmin - guess, mx = cmax - guess;
            int diff = mn;
            if (mn != mx) {
                // find_leaf (compound.h:142-153): node values are warp-uniform, the tested property comes from its lane
                uint2 cur;
                if (NODES_SMEM) { const uint4 r0 = lds128(nodes_saddr); cur = make_uint2(r0.z, r0.w); }     // node 0 lives in slot 1
                else cur = nodes2[1];
                while ((int)cur.x >= 0) {
                    uint4 pair;
                    if (NODES_SMEM) pair = lds128(nodes_saddr + ((cur.x & 0xffffu) << 4));
                    else pair = reinterpret_cast<const uint4 *>(nodes2)[cur.x & 0xffffu];
                    const int v = __shfl_sync(0xffffffffu, mine, (int)(cur.x >> 16));
                    cur = (v > (int)cur.y) ? make_uint2(pair.x, pair.y) : make_uint2(pair.z, pair.w);
                }
                uint16_t *lp = leaf_lookup(ls, (int)(cur.x & 0xffffu), lane);
                if (lane == 0) diff = read_int(rac, sm.table, lp, mn, mx, ls.mant_base);
                diff = __shfl_sync(0xffffffffu, diff, 0);
            }
            const int val = s16(s16(diff) + guess);
            outv = (lane == i) ? val : outv;
            leftleft = xx ? left : val;          // next pixel: x > 1 ? value(x-2) : left   (context_predict.h:132)
            left = val;
        }
        if (x < w) row[x] = (int16_t)outv;
        __syncwarp();
    }
}

// fuif_decode_channel, encoding.cpp:259-429.  Returns false on a hard error. `beginc` is advanced to the group's last channel.
// `io` is kept identical in all lanes on entry and on exit; in between only lane 0's copy (inside `rac`) advances.
__device__ bool decode_group(DImage &img, Reader &io, int &beginc, const Params &P, WarpScratch &ws, int lane, const Smem &sm) {
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
    if (P.debug && lane == 0) printf("[maniac] group at %lld: ch %d-%d compress %d pred %d range %d..%d\n", header_pos, beginc, endc, (int)compress, predictor, global_minv, global_maxv);
    if (endc >= img.nch || endc < beginc) return false;
    img.ch[b0].group_off = header_pos;

    int firstrealc = beginc;
    bool early = false, early_result = true;
    for (int i = beginc; i <= endc; i++) {
        DChan &ch = img.ch[i];
        if (ch.w * ch.h <= 0) continue;
        ch.minval = global_minv; ch.maxval = global_maxv;
        if (endc > beginc && global_minv < global_maxv) {
            ch.minval = s16(ch.minval + io.varint());
            ch.maxval = s16(ch.minval + io.varint());
        }
        if (ch.minval == ch.maxval) { fill_plane(ch, ch.minval, lane); firstrealc++; }
        if (ch.minval == 0 && ch.maxval == 0) continue;
        ch.q = io.varint();
        if (io.stop()) { early = true; early_result = corrupt_or_truncated(true, ch, lane); break; }
        if (compress && !check_bit_depth(ch.minval, ch.maxval, predictor)) { early = true; early_result = false; break; }
    }
    for (int i = beginc; i <= endc; i++) {
        DChan &ch = img.ch[i];
        if (ch.w * ch.h <= 0 || ch.minval == ch.maxval) continue;      // the reference calls setzero() only on channels it decodes
        if (ch.minval > 0) ch.zero = ch.minval; else if (ch.maxval < 0) ch.zero = ch.maxval; else ch.zero = 0;   // setzero, image.h:70-74
    }
    // ranges of this group's channels are final from here on: let dependent streams read them
    __syncwarp();
    if (lane == 0) for (int i = beginc; i <= endc; i++) st_release(&img.ch[i].hdr_done, 1);
    __syncwarp();
    if (early) return early_result;
    if (firstrealc > endc) { beginc = endc; return true; }

    int pr[kMaxProps][2];
    int refchan[kMaxProps / 2], nrefchan = 0;
    const int nprops = init_properties(pr, img, beginc, endc, refchan, nrefchan);
    const int nref = nprops - NB_NONREF;
    if (nprops > 32) { img.status = FB_ERR_UNSUPPORTED; return false; }     // one lane per property (max_properties <= 18)

    int predictability = 2048;
    if (predictor == 0 && compress) {
        int rounded = io.varint();
        if (rounded < 1 || rounded > 127) return corrupt_or_truncated(io.stop(), img.ch[min(firstrealc, img.nch - 1)], lane);
        predictability = rounded * 32;
    }

    // ---- from here on the serial coder lives in lane 0 (rac.io is the byte position)
    Rac rac;
    rac.init(io);
#define SYNC_IO()                                                                     \
    do {                                                                              \
        io = rac.io;                                                                  \
        io.pos = __shfl_sync(0xffffffffu, io.pos, 0);                                 \
        io.eof = __shfl_sync(0xffffffffu, (int)io.eof, 0) != 0;                       \
        io.avail = 0;                                                                 \
    } while (0)
#define STOPPED() (__shfl_sync(0xffffffffu, (int)rac.io.stop(), 0) != 0)

    if (!compress) {        // encoding.cpp:334-354
        for (int i = beginc; i <= endc; i++) {
            DChan &ch = img.ch[i];
            if (ch.minval == ch.maxval) continue;
            fill_plane(ch, i < img.n_orig ? 0 : ch.zero, lane);
            if (lane == 0) {
                for (int y = 0; y < ch.h; y++) {
                    if (rac.io.stop()) break;
                    for (int x = 0; x < ch.w; x++) ch.data[(size_t)y * ch.w + x] = (int16_t)uniform_read(rac, ch.minval, ch.maxval - ch.minval);
                    st_release(&ch.rows_done, y + 1);
                }
            }
            __syncwarp();
            if (STOPPED()) break;
        }
        beginc = endc;
        SYNC_IO();
        return true;
    }

    int nnodes = 0, tree_ok = 1;
    if (lane == 0) tree_ok = read_tree(rac, P.meta_table, pr, nprops, ws.nodes, nnodes, ws.stack, sm.coder) ? 1 : 0;
    tree_ok = __shfl_sync(0xffffffffu, tree_ok, 0);
    nnodes = __shfl_sync(0xffffffffu, nnodes, 0);
    if (!tree_ok) { const bool st = STOPPED(); SYNC_IO(); return corrupt_or_truncated(st, img.ch[beginc], lane); }

    // FinalPropertySymbolCoder ctor, compound.h:213-225: leaf numbering in node order, all leaves start from zero_chance
    const int nleaves = (nnodes + 1) / 2;
    // shared-memory plan for this group: node cache (if it fits), then leaf lines
    int dyn_off = 0;
    uint2 *snodes = nullptr;
    const int node_bytes = ((nnodes + 2) * 8 + 15) & ~15;
    if (node_bytes + 64 * 8 + 64 <= sm.dyn_bytes) { snodes = reinterpret_cast<uint2 *>(sm.dyn); dyn_off = node_bytes; }
    LeafStore ls;
    ls.gleaves = ws.leaves;
    const int rem = sm.dyn_bytes - dyn_off;
    int group_range = 0;
    for (int i = beginc; i <= endc; i++) group_range = max(group_range, img.ch[i].maxval - img.ch[i].minval + 1);
    const bool compact = group_range <= 255;
    ls.shift = compact ? 4 : 5;
    ls.mant_base = compact ? 9 : SC_MANT;
    const int lbytes = 2 << ls.shift;           // bytes per leaf
    // initial chance of entry e of a leaf in the chosen layout
    auto init_entry = [&](int e) -> uint16_t {
        if (!compact) return initial_chance(e, predictability);
        if (e < 2) return initial_chance(e, predictability);
        if (e < 9) return initial_chance(SC_EXP + (e - 2), predictability);
        return 1024;
    };
    if (nleaves * lbytes <= rem) {              // every leaf resident
        ls.lines = reinterpret_cast<uint16_t *>(sm.dyn + dyn_off); ls.tags = nullptr; ls.mask = 0;
        for (int i = lane; i < (nleaves << ls.shift); i += 32) ls.lines[i] = init_entry(i & ((1 << ls.shift) - 1));
    } else {                                    // direct-mapped cache over the global leaf array
        int nlines = 1;
        while (nlines * 2 * (lbytes + 4) <= rem) nlines *= 2;
        ls.tags = reinterpret_cast<int *>(sm.dyn + dyn_off);
        ls.lines = reinterpret_cast<uint16_t *>(sm.dyn + dyn_off + ((nlines * 4 + 15) & ~15));
        ls.mask = nlines - 1;
        for (int i = lane; i < nlines; i += 32) ls.tags[i] = -1;
        for (int i = lane; i < (nleaves << ls.shift); i += 32) ws.leaves[i] = init_entry(i & ((1 << ls.shift) - 1));
    }
    // node cache: packed entries, see struct LeafStore comment
    uint2 *gpacked = reinterpret_cast<uint2 *>(ws.stack);      // the parse stack is free again: reuse it for the packed nodes
    if (lane == 0) {
        int leafID = 0;
        for (int i = 0; i < nnodes; i++) {
            const TNode nd = ws.nodes[i];
            uint2 e;
            if (nd.property == -1) { e.x = 0xFFFF0000u | (unsigned)leafID++; e.y = 0; }
            else { e.x = ((unsigned)nd.property << 16) | (unsigned)((nd.child + 1) >> 1); e.y = (unsigned)nd.splitval; }
            gpacked[i + 1] = e;
        }
    }
    __syncwarp();
    const uint2 *nodes2 = gpacked;
    // Groups that can use the walker warps keep a walker-friendly copy in shared memory: the property field becomes a word
    // offset -- < 64: into the shared per-pixel property row, 64 + k: the k-th left-dependent value of the walker's lane.
    const bool walker_nodes = sm.mail && snodes && predictor == 0 && nnodes > 1;
    if (snodes) {
        for (int i = lane; i < nnodes; i += 32) {
            uint2 e = gpacked[i + 1];
            if (walker_nodes && (int)e.x >= 0) {
                const int pidx = (int)(e.x >> 16), role = pidx - nref;
                int off = pidx;
                if (role == 1) off = 64; else if (role == 3) off = 65; else if (role == 6) off = 66; else if (role == 7) off = 67;
                else if (role == 8) off = 68; else if (role == 9) off = 69; else if (role == 12) off = 70;
                e.x = ((unsigned)off << 16) | (e.x & 0xffffu);
            }
            snodes[i + 1] = e;
        }
        if (!walker_nodes) nodes2 = snodes;     // the decoder's own walk reads the property index, i.e. the plain encoding
    }
    __syncwarp();
    if (P.debug && lane == 0)
        printf("[maniac]   tree %d nodes, nprops %d nref %d, zero_chance %d, pos %llu, %dx%d, nodes %s, leaves %s\n", nnodes, nprops, nref, predictability,
               rac.io.pos, img.ch[beginc].w, img.ch[beginc].h, snodes ? "smem" : "global", ls.tags ? "cached" : "resident");

    for (int i = beginc; i <= endc; i++) {
        DChan &ch = img.ch[i];
        if (ch.minval == ch.maxval) continue;
        // channel.resize(w,h): buffers made by meta_apply start out as `zero`, the Image constructor's as 0
        fill_plane(ch, i < img.n_orig ? 0 : ch.zero, lane);
        if (nnodes == 1 && predictor == 0 && ch.zero == 0) {        // fast track, encoding.cpp:371-383
            uint16_t *lp = leaf_lookup(ls, 0, lane);
            if (lane == 0) {
                for (int y = 0; y < ch.h; y++) {
                    if (rac.io.stop()) break;
                    int16_t *row = ch.data + (size_t)y * ch.w;
                    for (int x = 0; x < ch.w; x++) row[x] = (int16_t)read_int(rac, sm.table, lp, ch.minval, ch.maxval, ls.mant_base);
                    st_release(&ch.rows_done, y + 1);
                }
            }
            __syncwarp();
        } else {
            const int range = ch.maxval - ch.minval + 1;
            const bool helped = walker_nodes && range <= 32 * sm.nwalkers && range >= 1;
            const int nwalk = (range + 31) / 32;
            if (helped && lane == 0) sm.mail->nodes_saddr = (unsigned)__cvta_generic_to_shared(snodes);
            const long long t_start = clock64();
            for (int y = 0; y < ch.h; y++) {
                if (STOPPED()) break;
                for (int r = 0; r < nrefchan; r++) {       // row wavefront on the planes this row back-references
                    const DChan &cj = img.ch[refchan[r]];
                    int ry = (y << ch.vshift) >> cj.vshift;
                    if (ry >= cj.h) ry = cj.h - 1;
                    spin_until_ge(&cj.rows_done, ry + 1);
                }
                if (helped) decode_row_helped(img, ch, y, refchan, nrefchan, nref, rac, sm, ls, sm.mail, nwalk, lane);
                else if (snodes && !walker_nodes) {
                    if (predictor == 0) decode_row<true, true>(img, ch, y, predictor, refchan, nrefchan, nref, rac, sm, nodes2, ls, lane);
                    else decode_row<true, false>(img, ch, y, predictor, refchan, nrefchan, nref, rac, sm, nodes2, ls, lane);
                } else {
                    if (predictor == 0) decode_row<false, true>(img, ch, y, predictor, refchan, nrefchan, nref, rac, sm, nodes2, ls, lane);
                    else decode_row<false, false>(img, ch, y, predictor, refchan, nrefchan, nref, rac, sm, nodes2, ls, lane);
                }
                publish_rows(ch, y + 1, lane);
            }
            if (P.debug && lane == 0)
                printf("[maniac]   ch %d %dx%d range %d nodes %d helped %d walkers %d: %.0f cycles/symbol\n", i, ch.w, ch.h, range, nnodes, (int)helped, nwalk,
                       (double)(clock64() - t_start) / ((double)ch.w * ch.h));
        }
        if (STOPPED()) break;
    }
    beginc = endc;
    SYNC_IO();
    return true;
#undef SYNC_IO
#undef STOPPED
}

__global__ void k_maniac_decode(Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wps = 1 + P.helpers;                       // warps per stream: decoder + walkers
    const int slot = warp / wps, wrole = warp % wps;
    uint16_t *s_table = reinterpret_cast<uint16_t *>(smem_raw);                      // 16 KiB, shared by the block's warps
    for (int i = threadIdx.x; i < 4096 * 2; i += blockDim.x) s_table[i] = P.table[i];
    __syncthreads();
    unsigned char *mine = smem_raw + 16384 + (size_t)slot * P.warp_smem;
    Smem sm;
    sm.table = s_table;
    sm.coder = reinterpret_cast<uint16_t(*)[32]>(mine);                               // 192 B
    const int cprop_bytes = P.helpers ? 2 * 4736 : 4736;         // chunk properties, double-buffered when walkers run ahead
    const int mail_bytes = P.helpers ? kMailBytes + P.helpers * 32 * kLdRowStride * 4 : 0;
    sm.cprop = reinterpret_cast<int *>(mine + 256);
    sm.mail = P.helpers ? reinterpret_cast<Mail *>(mine + 256 + cprop_bytes) : nullptr;
    sm.ldrows = P.helpers ? reinterpret_cast<int *>(mine + 256 + cprop_bytes + kMailBytes) : nullptr;
    sm.nwalkers = P.helpers;
    sm.dyn = mine + 256 + cprop_bytes + mail_bytes;
    sm.dyn_bytes = P.warp_smem - 256 - cprop_bytes - mail_bytes;
    if (P.helpers) {
        if (wrole == 0 && lane == 0) { sm.mail->cmd_seq = 0; sm.mail->cmd = 0; sm.mail->go = 0; }
        __syncthreads();
        if (wrole > 0) { walker_main(sm.mail, sm.cprop, sm.ldrows, wrole - 1, lane); return; }
    }
    WarpScratch ws = P.scratch[blockIdx.x * (blockDim.x >> 5) / wps + slot];
    for (;;) {
        int sid = 0;
        if (lane == 0) sid = atomicAdd(P.ticket, 1);
        sid = __shfl_sync(0xffffffffu, sid, 0);
        if (sid >= P.nstreams) break;
        const DStream st = P.streams[sid];
        DImage &img = P.images[st.image];
        Reader io;
        io.p = img.bytes; io.n = img.nbytes; io.btl = img.bytes_to_load; io.eof = false;
        io.seek(st.offset);
        int groups = 0;
        // the channel loop of fuif_decode, encoding.cpp:708-718
        for (int i = st.first_channel; i < img.nch; i++) {
            if (st.max_groups >= 0 && groups >= st.max_groups) break;
            if ((img.bytes_to_load == 0 || io.pos < img.bytes_to_load) && !io.eof) {
                if (!img.ch[i].w || !img.ch[i].h) continue;
                bool ok = decode_group(img, io, i, P, ws, lane, sm);
                groups++;
                if (!ok) { if (!img.status) img.status = FB_ERR_INVALID; break; }
            } else break;
        }
        // whatever happened (truncation, corruption), nobody may wait forever on this stream's channels
        __syncwarp();
        if (lane == 0)
            for (int c = st.first_channel; c < st.end_channel && c < img.nch; c++) {
                st_release(&img.ch[c].hdr_done, 1);
                st_release(&img.ch[c].rows_done, 0x7fffffff);
            }
        __syncwarp();
    }
    if (P.helpers && lane == 0) {       // release the walkers
        sm.mail->cmd = 2;
        __threadfence_block();
        sm.mail->cmd_seq = sm.mail->cmd_seq + 1;
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------

// build_table, maniac/chance.cpp:31-65
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

struct ManiacState {
    uint16_t *table_dev = nullptr, *meta_dev = nullptr;
    int cutoff = -1, alpha = -1;
    int nslots = 0, maxw = 0;
    WarpScratch *scratch_dev = nullptr;
    void *arena = nullptr;
    int *ticket_dev = nullptr;
};

int ensure_state(fb_ctx *ctx, int cutoff, int alpha, int nslots, int maxw) {
    ManiacState *st = (ManiacState *)ctx->maniac_state;
    if (!st) { st = new ManiacState(); ctx->maniac_state = st; }
    if (!st->table_dev) {
        FB_CUDA(ctx, cudaMalloc((void **)&st->table_dev, 4096 * 2 * sizeof(uint16_t)));
        FB_CUDA(ctx, cudaMalloc((void **)&st->meta_dev, 4096 * 2 * sizeof(uint16_t)));
        FB_CUDA(ctx, cudaMalloc((void **)&st->ticket_dev, sizeof(int)));
        std::vector<uint16_t> t(4096 * 2);
        build_table(t.data(), 0xFFFFFFFFu / 19, 4096 - 2);      // SimpleBitChanceTable(cut=2, alpha=0xFFFFFFFF/19), chance.h:53
        FB_CUDA(ctx, cudaMemcpy(st->meta_dev, t.data(), t.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
    if (st->cutoff != cutoff || st->alpha != alpha) {
        std::vector<uint16_t> t(4096 * 2);
        build_table(t.data(), (uint32_t)alpha, (unsigned)(4096 - cutoff));
        FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        FB_CUDA(ctx, cudaMemcpy(st->table_dev, t.data(), t.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        st->cutoff = cutoff; st->alpha = alpha;
    }
    if (nslots > st->nslots || maxw > st->maxw) {
        FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (st->arena) cudaFree(st->arena);
        if (st->scratch_dev) cudaFree(st->scratch_dev);
        st->arena = nullptr; st->scratch_dev = nullptr;
        nslots = std::max(nslots, st->nslots);
        maxw = std::max(maxw, st->maxw);
        const size_t nodes_b = sizeof(TNode) * kMaxNodes, leaves_b = sizeof(uint16_t) * 32 * (kMaxNodes / 2);
        const size_t stack_b = sizeof(int) * 4 * (kMaxNodes / 2 + 2);
        const size_t refs_b = ((sizeof(int16_t) * (size_t)maxw * 12) + 255) & ~(size_t)255;
        const size_t per = nodes_b + leaves_b + stack_b + refs_b;
        FB_CUDA(ctx, cudaMalloc(&st->arena, per * nslots));
        std::vector<WarpScratch> ws(nslots);
        for (int i = 0; i < nslots; i++) {
            char *base = (char *)st->arena + per * i;
            ws[i].nodes = (TNode *)base;
            ws[i].leaves = (uint16_t *)(base + nodes_b);
            ws[i].stack = (int *)(base + nodes_b + leaves_b);
            ws[i].refs = (int16_t *)(base + nodes_b + leaves_b + stack_b);
        }
        FB_CUDA(ctx, cudaMalloc((void **)&st->scratch_dev, sizeof(WarpScratch) * nslots));
        FB_CUDA(ctx, cudaMemcpy(st->scratch_dev, ws.data(), sizeof(WarpScratch) * nslots, cudaMemcpyHostToDevice));
        st->nslots = nslots; st->maxw = maxw;
    }
    return FB_OK;
}

// host copy of read_big_endian_varint for walking the group index
int host_varint(const uint8_t *p, size_t n, size_t &pos) {
    int result = 0, k = 0;
    while (k++ < 10) {
        if (pos >= n) return -1;
        int number = p[pos++];
        if (number < 128) return result + number;
        result += number - 128;
        result = (int)((unsigned)result << 7);
    }
    return -1;
}

}  // namespace

void fb_maniac_release(fb_ctx *ctx) {
    ManiacState *st = (ManiacState *)ctx->maniac_state;
    if (!st) return;
    cudaFree(st->table_dev); cudaFree(st->meta_dev); cudaFree(st->ticket_dev); cudaFree(st->arena); cudaFree(st->scratch_dev);
    delete st;
    ctx->maniac_state = nullptr;
}

int fb_maniac_decode(fb_ctx *ctx, std::vector<FbManiacJob> &jobs) {
    const int nimg = (int)jobs.size();
    if (!nimg) return FB_OK;
    int cutoff = jobs[0].cutoff, alpha = jobs[0].alpha;
    // ---- plane allocation + descriptors
    std::vector<DImage> himg(nimg);
    std::vector<std::vector<DChan>> hch(nimg);
    std::vector<DStream> streams;
    size_t total_bytes = 0, total_ch = 0;
    int maxw = 8;
    for (int b = 0; b < nimg; b++) {
        if (jobs[b].cutoff != cutoff || jobs[b].alpha != alpha) { ctx->err = "batch with mixed maniac options"; return FB_ERR_INVALID; }
        if (!jobs[b].bytes_dev) total_bytes += (jobs[b].nbytes + 255) & ~(size_t)255;
        total_ch += jobs[b].img->ch.size();
    }
    uint8_t *bytes_dev = nullptr;
    DChan *ch_dev = nullptr;
    DImage *img_dev = nullptr;
    DStream *streams_dev = nullptr;
    FB_CUDA(ctx, cudaMallocAsync((void **)&bytes_dev, std::max<size_t>(total_bytes, 256), ctx->stream));
    FB_CUDA(ctx, cudaMallocAsync((void **)&ch_dev, std::max<size_t>(total_ch, 1) * sizeof(DChan), ctx->stream));
    FB_CUDA(ctx, cudaMallocAsync((void **)&img_dev, nimg * sizeof(DImage), ctx->stream));
    size_t boff = 0, coff = 0;
    for (int b = 0; b < nimg; b++) {
        FbManiacJob &job = jobs[b];
        fb_image *img = job.img;
        if (job.bytes_dev) himg[b].bytes = job.bytes_dev;
        else {
            FB_CUDA(ctx, cudaMemcpyAsync(bytes_dev + boff, job.bytes_host, job.nbytes, cudaMemcpyHostToDevice, ctx->stream));
            himg[b].bytes = bytes_dev + boff;
        }
        himg[b].nbytes = job.nbytes;
        himg[b].bytes_to_load = job.bytes_to_load;
        himg[b].ch = ch_dev + coff;
        himg[b].nch = (int)img->ch.size();
        himg[b].max_properties = job.max_properties;
        himg[b].n_orig = img->info.real_nb_channels;
        himg[b].status = 0;
        hch[b].resize(img->ch.size());
        for (size_t i = 0; i < img->ch.size(); i++) {
            FbChan &c = img->ch[i];
            DChan &d = hch[b][i];
            memset(&d, 0, sizeof(d));
            d.w = c.d.w; d.h = c.d.h; d.minval = c.d.minval; d.maxval = c.d.maxval; d.zero = c.d.zero; d.q = c.d.q;
            d.hshift = c.d.hshift; d.vshift = c.d.vshift; d.group_off = -1;
            const size_t n = (c.d.w > 0 && c.d.h > 0) ? (size_t)c.d.w * c.d.h : 0;
            if (n) { int rc = fb_plane_alloc(ctx, n, &c.dev); if (rc) return rc; }
            d.data = c.dev;
            if (!n) { d.hdr_done = 1; d.rows_done = 0x7fffffff; }     // empty channels are skipped by the channel loop
            maxw = std::max(maxw, c.d.w);
        }
        if (!hch[b].empty())
            FB_CUDA(ctx, cudaMemcpyAsync(ch_dev + coff, hch[b].data(), hch[b].size() * sizeof(DChan), cudaMemcpyHostToDevice, ctx->stream));
        // ---- streams
        if (himg[b].nch > 0) {
            bool indexed = job.group_index && job.n_groups > 0;
            if (indexed) {
                // one stream per group: walk the channel list exactly like the loop of fuif_decode (encoding.cpp:708-718)
                int i = 0;
                std::vector<DStream> mine;
                for (int g = 0; g < job.n_groups && indexed; g++) {
                    while (i < himg[b].nch && (!img->ch[i].d.w || !img->ch[i].d.h)) i++;
                    size_t pos = (size_t)job.group_index[g];
                    if (i >= himg[b].nch || job.group_index[g] < (int64_t)job.body_pos || pos >= job.nbytes) { indexed = false; break; }
                    if (job.bytes_to_load && pos >= job.bytes_to_load) break;
                    if (job.group_first) {
                        i = job.group_first[g];
                        if (i < 0 || i >= himg[b].nch || (!mine.empty() && i <= mine.back().first_channel)) { indexed = false; break; }
                    } else {
                        if (job.bytes_dev) { indexed = false; break; }      // cannot read group headers of a device buffer on the host
                        int fb = host_varint(job.bytes_host, job.nbytes, pos);
                        if (fb < 0) { indexed = false; break; }
                    }
                    if (!mine.empty()) mine.back().end_channel = i;
                    mine.push_back(DStream{b, i, himg[b].nch, 1, (unsigned long long)job.group_index[g]});
                    if (!job.group_first) { size_t p2 = (size_t)job.group_index[g]; int fb = host_varint(job.bytes_host, job.nbytes, p2); i += (fb >> 4) + 1; }
                }
                if (indexed && !mine.empty()) streams.insert(streams.end(), mine.begin(), mine.end());
                else indexed = false;
            }
            if (!indexed) streams.push_back(DStream{b, 0, himg[b].nch, -1, (unsigned long long)job.body_pos});
        }
        if (!job.bytes_dev) boff += (job.nbytes + 255) & ~(size_t)255;
        coff += img->ch.size();
    }
    FB_CUDA(ctx, cudaMemcpyAsync(img_dev, himg.data(), nimg * sizeof(DImage), cudaMemcpyHostToDevice, ctx->stream));
    if (nimg > 1) {
        // group-major ticket order: the k-th stream of every image before any (k+1)-th stream.  A stream still only depends on
        // lower tickets (earlier groups of its own image), and the large late groups of all images end up running together
        // instead of trailing image by image.
        std::vector<int> ord(streams.size());
        std::vector<int> seen(nimg, 0);
        for (size_t k = 0; k < streams.size(); k++) ord[k] = seen[streams[k].image]++;
        std::vector<size_t> idx(streams.size());
        for (size_t k = 0; k < idx.size(); k++) idx[k] = k;
        std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return ord[a] < ord[b]; });
        std::vector<DStream> sorted(streams.size());
        for (size_t k = 0; k < idx.size(); k++) sorted[k] = streams[idx[k]];
        streams.swap(sorted);
    }
    const int nstreams = (int)streams.size();
    if (nstreams) {
        // streams are ordered image-major, groups ascending: dependencies always point to lower stream ids
        FB_CUDA(ctx, cudaMallocAsync((void **)&streams_dev, nstreams * sizeof(DStream), ctx->stream));
        FB_CUDA(ctx, cudaMemcpyAsync(streams_dev, streams.data(), nstreams * sizeof(DStream), cudaMemcpyHostToDevice, ctx->stream));
        const int nslots = std::min((nstreams + 1) / 2 * 2, ctx->sm_count * 2);        // scratch slots: one per stream in flight
        int rc = ensure_state(ctx, cutoff, alpha, nslots, maxw);
        if (rc) return rc;
        ManiacState *st = (ManiacState *)ctx->maniac_state;
        FB_CUDA(ctx, cudaMemsetAsync(st->ticket_dev, 0, sizeof(int), ctx->stream));
        Params P;
        P.images = img_dev; P.streams = streams_dev; P.nstreams = nstreams; P.ticket = st->ticket_dev;
        P.table = st->table_dev; P.meta_table = st->meta_dev; P.scratch = st->scratch_dev; P.maxw = st->maxw;
        P.debug = getenv("FB_MANIAC_DEBUG") ? 1 : 0;
        // Launch shape.  Few streams (one image): one warp per block and block per SM with ~200 KiB of shared memory, so
        // that the whole MANIAC tree and most leaf chances of a stream stay on-chip.  Many streams (batches): up to 8
        // warps share a block's 16 KiB chance table and up to two blocks share an SM.
        // Launch shape.  A block serves `spb` streams at a time, each with one decoder warp and `helpers` walker warps, and
        // owns one SM (~200 KiB of shared memory split between its streams: tree-node cache, leaf chances, per-chunk
        // property rows).  Few streams (one image) => one stream per SM with 8 walkers (value ranges up to 256); batches =>
        // two streams per SM with 6 walkers each (ranges up to 192; wider ranges fall back to the decoder's own walk).
        const int per_sm = (nstreams + ctx->sm_count - 1) / ctx->sm_count;
        const int wpb = std::max(1, std::min(2, per_sm));          // streams per block
        P.helpers = getenv("FB_MANIAC_NO_WALKERS") ? 0 : (wpb == 1 ? kMaxWalkers : 6);
        const int nblocks = std::min((nstreams + wpb - 1) / wpb, ctx->sm_count);
        const size_t block_smem = 200 * 1024;
        const size_t warp_smem = ((block_smem - 16384) / wpb) & ~(size_t)15;
        P.warp_smem = (int)warp_smem;
        const size_t smem_bytes = 16384 + warp_smem * wpb;
        FB_CUDA(ctx, cudaFuncSetAttribute(k_maniac_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        k_maniac_decode<<<nblocks, 32 * wpb * (1 + P.helpers), smem_bytes, ctx->stream>>>(P);
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { ctx->err = std::string("maniac launch: ") + cudaGetErrorString(e); return FB_ERR_CUDA; }
    }
    // ---- read back channel descriptors (ranges, q, which planes hold data) and the per-image status
    std::vector<DChan> back(total_ch);
    if (total_ch) FB_CUDA(ctx, cudaMemcpyAsync(back.data(), ch_dev, total_ch * sizeof(DChan), cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(ctx, cudaMemcpyAsync(himg.data(), img_dev, nimg * sizeof(DImage), cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(bytes_dev, ctx->stream); cudaFreeAsync(ch_dev, ctx->stream); cudaFreeAsync(img_dev, ctx->stream);
    if (streams_dev) cudaFreeAsync(streams_dev, ctx->stream);
    coff = 0;
    int rc = FB_OK;
    for (int b = 0; b < nimg; b++) {
        fb_image *img = jobs[b].img;
        for (size_t i = 0; i < img->ch.size(); i++) {
            const DChan &d = back[coff + i];
            FbChan &c = img->ch[i];
            c.d.minval = d.minval; c.d.maxval = d.maxval; c.d.zero = d.zero; c.d.q = d.q;
            if (d.state) c.d.decoded = 1;
            else { fb_plane_free(ctx, c.dev); c.dev = nullptr; c.d.decoded = 0; }
            if (d.group_off >= 0) { img->group_off.push_back(d.group_off); img->group_first.push_back((int32_t)i); }
        }
        if (himg[b].status == FB_ERR_UNSUPPORTED) { ctx->err = "max_properties > 18 is not supported by the GPU context model"; rc = FB_ERR_UNSUPPORTED; }
        else if (himg[b].status) { ctx->err = "corrupt FUIF stream (image " + std::to_string(b) + ")"; rc = FB_ERR_INVALID; }
        coff += img->ch.size();
    }
    return rc;
}
