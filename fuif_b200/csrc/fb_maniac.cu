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
// the groups' byte offsets (group index), otherwise one whole image.  A stream gets one SM-resident team of warps:
// a decoder warp that runs the serial coder (its pixel loop is executed redundantly by all lanes on warp-uniform
// values, see RowState), walker warps that evaluate the MANIAC tree of upcoming pixels for every candidate value of
// `left` several pixels ahead of the decoder, and a warp that keeps the per-pixel property rows ahead of both
// (see walker_main).  Groups the walkers cannot serve fall back to decode_row (one warp, lane k owns property k).
// Streams are handed out through an atomic ticket so that a stream only ever waits for lower-numbered streams (row
// wavefront on the planes it back-references), which are guaranteed to be running already.
// The device part of this file is also compiled, unchanged, for the CPU execution-model emulator of the test tier
// (tests/emu/emu_maniac.cpp, -DFB_EMULATE): the few places where CUDA-only constructs need a stand-in are marked FB_EMULATE.
#ifdef FB_EMULATE
#include "maniac_emu_shim.h"
#define FB_SPIN() fb_emu_yield()
#else
#include "fb_common.cuh"
#include "fb_host_entropy.h"
#define FB_SPIN()
#define FB_DYN_SMEM_DECL(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace {

constexpr int kMaxNodes = 65536;
constexpr int kStatusTreeTooLarge = 100;   // internal image status: a MANIAC tree with more than 65535 nodes (the reference has no limit: std::vector)
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
    int helpers;            // extra warps per stream (0, 7 or 15); every fourth one stays idle, the others are walkers
    int debug;              // FB_MANIAC_DEBUG=1: trace group headers from lane 0
    int walker_sleep;       // ns a walker sleeps between polls of the decoder's progress
    int walkers_used;       // tuning: use at most this many walkers
    int prefetch;           // walkers prefetch leaf lines
};

__device__ __forceinline__ int s16(int x) { return (int)(short)x; }
__device__ __forceinline__ int ilog2u(unsigned l) { return l == 0 ? 0 : 31 - __clz(l); }      // maniac/util.h:33-36

#ifdef FB_EMULATE
__device__ __forceinline__ void st_release(int *p, int v) { *(volatile int *)p = v; }
__device__ __forceinline__ int ld_acquire(const int *p) { return *(const volatile int *)p; }
#else
__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#endif

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
    while (len > 0) {       // len < 0 only comes out of a damaged header (the reference asserts); it must not loop forever
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
            if (nnodes + 2 > 65535) { nnodes = -1; return false; }       // a valid but larger tree than this decoder holds (child ids are 16 bit): reported as unsupported, not as corrupt
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
    int walker_sleep, prefetch;
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
    int mask;               // nlines - 1 (power-of-two line count: leaf & mask), or
    int nlines;             // any line count for the run-ahead path (leaf % nlines); 0 = use mask
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
#ifdef FB_EMULATE
__device__ __forceinline__ uint4 lds128(unsigned addr) { uint4 v; memcpy(&v, fb_emu_smem(addr), 16); return v; }
#else
__device__ __forceinline__ uint4 lds128(unsigned addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
#endif
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
        // every load first (one memory latency for all of them), then the arithmetic
        int rv[16];
#pragma unroll
        for (int r = 0; r < 16; r++) {
            rv[r] = 0;
            if (r < nrefchan) {
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
                rv[r] = __ldcg(cj.data + (size_t)ry * cj.w + rx);
            }
        }
#pragma unroll
        for (int r = 0; r < 16; r++) {
            if (r < nrefchan) {
                pp[2 * r] = fooabs(rv[r]);
                pp[2 * r + 1] = slog(rv[r]);
            }
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

// ---- run-ahead walker warps ---------------------------------------------------------------------------------------
// The leaf of pixel x+1 depends on pixel x only through `left` (and on pixel x-1 only through `leftleft`, property 12).
// Walker warps evaluate the MANIAC tree of upcoming pixels for EVERY value `left` can take (one candidate per lane; a
// "unit" = one pixel x one block of 32 candidates) and leave the result -- a leaf id, or the inner node at which the walk
// met a test of property 12 -- in a ring in shared memory.  Nothing a walker needs depends on a pixel of the current row
// that is still to be decoded, so the walkers run up to K pixels AHEAD of the decoder and the tree walk is a
// throughput problem spread over many warps instead of a latency on the decoder's serial chain.  The decoder (one lane)
// then needs one shared-memory lookup per pixel; the rare walks that stopped at property 12 it finishes itself.
// Used when the group has predictor 0, a value range <= 256 and a tree whose inner nodes fit in shared memory.
//
// Compact inner-node array (8 bytes per INNER node, leaves are not stored):
//   x = splitval << 8 | off      off < 64: word of the shared per-pixel property row, 64 + k: k-th left-dependent value
//   y = ref(child taken when value > splitval) | ref(other child) << 16,   ref = 0x8000 | leaf id, or inner-node index
// Handshake words carry a 16-bit tag = (y & 1) << 15 | (x + 1) in their upper half, so a stale entry of the previous
// row (or of a pixel K / 64 positions back) can never be taken for the one that is awaited.  Walker (r, b) owns candidate
// block b of the pixels x = r mod g; K is a multiple of g, so a ring slot + block is only ever written by ONE warp, in order.
constexpr int kCandSlotsMax = 9;    // ring slots: the walkers run up to K <= 9 pixels ahead of the decoder
constexpr int kMaxCand = 256;       // value ranges up to 256
constexpr int kMaxWalkers = 12;
constexpr int kOffLeftLeft = 70;    // `off` of property 12 (slog(left - leftleft)): the only one that needs pixel x-2
// Everything the decoder's pixel loop needs, handed from lane 0 to ALL lanes through shared memory.  The loop is executed by
// the 32 lanes redundantly on identical values: its operands then provably come from uniform shared-memory loads, so the
// compiler emits plain branches for it (no convergence-barrier bookkeeping per decision, which costs more than the
// arithmetic when a single lane runs under `if (lane == 0)`).  Stores of the redundant lanes hit the same address with
// the same value.
struct RowState {
    int w, y, zero, cmin, mn, mx, emax_pos, emax_neg;
    unsigned mant_off, tagrow, kslots, inner_s, lines_s, line_shift, cached, mask, nlines, inv_nlines, tags_s, tab_s, cprop_s;
    unsigned range, low, ones, pos, n;
    const uint8_t *p;
    uint16_t *gleaves;
    // for the chunk prologues and the row store
    DImage *img;
    DChan *ch;
    int *cprop;
    int refchan[16], nrefchan, nref;
};

struct Mail {
    volatile int cmd_seq;       // bumped by the decoder for every command
    int cmd;                    // 1 = row, 2 = exit
    int y, w, cmin, nb, g, K;   // nb candidate blocks per pixel, g pixels in flight, K = ring slots (a multiple of g)
    unsigned inner_saddr, tagrow;
    const uint16_t *pf_leaves;  // non-null: leaf chances live in global memory behind a cache: walkers prefetch the lines they find
    int pf_shift;
    int prol_widx;              // the walker warp that computes the chunk prologues (property rows) of chunks >= 2
    // Monotone progress counters of the current channel (never reset inside a channel, so a walker that is still busy with
    // a candidate block nobody needed when the next row begins can never wait for something that was overwritten):
    volatile int progress;      // pixels decoded so far (row-major): pixel (y, x) is done when progress > y * w + x
    volatile int prol_ready;    // property rows of chunk (y, c) are in place when prol_ready >= y * nchunks + c
    volatile int ack[kMaxWalkers];              // last command each walker has read
    volatile int done[kMaxWalkers];             // last command each walker has finished
    volatile unsigned cval[64];                 // tag << 16 | decoded value & 0xffff, ring indexed by x & 63
    // slot x % K, column left - cmin:  x = tag << 16 | T & 0xffff,  y = kind << 30 | B << 15 | A
    //   kind 0: leaf A;  kind 1: the walk forked at a test of property 12: leaf A if leftleft <= T, else leaf B;
    //   kind 2: the walk met a second test of property 12: the decoder finishes it from inner node A
    volatile uint2 cand[kCandSlotsMax][kMaxCand];
    int dpriv[8];
    RowState row;
};
constexpr int kMailBytes = 19968;
// block configuration: shared address of stream slot 0's mailbox, log2(warps per stream), bytes per stream slot
__shared__ unsigned s_cfg[4];
static_assert(sizeof(Mail) <= kMailBytes, "Mail layout");

#define COMPILER_FENCE() asm volatile("" ::: "memory")
#ifdef FB_EMULATE
__device__ __forceinline__ int lds32(unsigned addr) { int v; memcpy(&v, fb_emu_smem(addr), 4); return v; }
__device__ __forceinline__ unsigned lds32v(unsigned addr) { fb_emu_yield(); unsigned v; memcpy(&v, fb_emu_smem(addr), 4); return v; }   // polled words
__device__ __forceinline__ uint2 lds64(unsigned addr) { uint2 v; memcpy(&v, fb_emu_smem(addr), 8); return v; }
__device__ __forceinline__ unsigned lds16(unsigned addr) { unsigned short v; memcpy(&v, fb_emu_smem(addr), 2); return v; }
__device__ __forceinline__ void sts16(unsigned addr, unsigned v) { const unsigned short h = (unsigned short)v; memcpy(fb_emu_smem(addr), &h, 2); }
__device__ __forceinline__ void sts32v(unsigned addr, unsigned v) { memcpy(fb_emu_smem(addr), &v, 4); }
__device__ __forceinline__ uint2 lds64v(unsigned addr) { fb_emu_yield(); uint2 v; memcpy(&v, fb_emu_smem(addr), 8); return v; }
__device__ __forceinline__ void sts64v(unsigned addr, unsigned x, unsigned y) { const uint2 v = {x, y}; memcpy(fb_emu_smem(addr), &v, 8); }
#else
__device__ __forceinline__ int lds32(unsigned addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds32v(unsigned addr) {     // polled words
    unsigned v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint2 lds64(unsigned addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds16(unsigned addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts16(unsigned addr, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ void sts32v(unsigned addr, unsigned v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint2 lds64v(unsigned addr) {
    uint2 v;
    asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts64v(unsigned addr, unsigned x, unsigned y) { asm volatile("st.volatile.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(x), "r"(y) : "memory"); }
#endif
constexpr int kLdRowStride = 9;     // words per walker lane: its 7 left-dependent property values (+ padding)

// Shared-memory accesses of one warp are performed in program order and there is no cache between the warps of a block,
// so the tag protocol needs compiler barriers only (no MEMBAR per pixel).
__device__ void walker_main(Mail *mail, const int *cprop2 /* [2][32][kPropStride] */, int *ldrows, int widx, int lane, int sleep_ns) {
    int seen = 0;
    int *myrow = ldrows + (widx * 32 + lane) * kLdRowStride;
    const unsigned ld_s = (unsigned)__cvta_generic_to_shared(myrow) - 64 * 4;      // biased: offsets 64.. address this row
    const unsigned cval_s = (unsigned)__cvta_generic_to_shared((const void *)mail->cval);
    const unsigned cand_s = (unsigned)__cvta_generic_to_shared((const void *)mail->cand);
    for (;;) {
        while (mail->cmd_seq == seen) __nanosleep(32);
        seen = mail->cmd_seq;
        __threadfence_block();
        if (mail->cmd == 2) return;
        const int y = mail->y, w = mail->w, cmin = mail->cmin, nb = mail->nb, g = mail->g, K = mail->K;
        const unsigned inner = mail->inner_saddr, tagrow = mail->tagrow;
        const uint16_t *pf_leaves = mail->pf_leaves;
        const int pf_shift = mail->pf_shift;
        COMPILER_FENCE();
        __syncwarp();
        if (lane == 0) mail->ack[widx] = seen;      // the command's fields may be overwritten from here on
        if (widx == mail->prol_widx) {
            // This warp keeps the per-pixel property rows ahead of everybody: chunk c (pixels 32c ..) goes into buffer c & 1
            // as soon as the decoder has begun chunk c-1 (every walk of chunk c-2 that matters is over by then).
            const RowState &S = mail->row;
            DImage &img = *S.img;
            DChan &ch = *S.ch;
            int *cprop = S.cprop;
            const int nrefchan = S.nrefchan, nref = S.nref;
            int refchan[16];
#pragma unroll
            for (int k = 0; k < 16; k++) refchan[k] = S.refchan[k];
            const int nchunks = (w + 31) >> 5;
            for (int c = 2; c < nchunks; c++) {
                while (mail->progress < y * w + 32 * (c - 1)) __nanosleep(sleep_ns);      // the decoder has begun chunk c-1
                __threadfence_block();
                chunk_prologue(img, ch, y, 32 * c, refchan, nrefchan, nref, cprop + (c & 1) * 32 * kPropStride, lane);
                __syncwarp();
                __threadfence_block();
                if (lane == 0) mail->prol_ready = y * nchunks + c;
            }
            __syncwarp();
            if (lane == 0) mail->done[widx] = seen;
            continue;
        }
        const int r = widx / nb, b = widx - r * nb;
        if (r >= g) { if (lane == 0) mail->done[widx] = seen; continue; }
        const int cl = cmin + 32 * b + lane;        // this lane's candidate for `left`
        const int nchunks_w = (w + 31) >> 5;
        int slot = r;                               // r < g <= K
        for (int j = r; j < w; j += g) {
            if (j >= K) while (mail->progress <= y * w + j - K) __nanosleep(sleep_ns);     // the ring slot is free once pixel j - K has been decoded
            while (mail->prol_ready < y * nchunks_w + (j >> 5)) { FB_SPIN(); }
            __threadfence_block();          // acquire: the property rows read below were written before prol_ready was posted
            COMPILER_FENCE();
            const int *pp = cprop2 + (((j >> 5) & 1) * 32 + (j & 31)) * kPropStride;
            const unsigned pp_s = (unsigned)__cvta_generic_to_shared(pp);
            const int top = pp[32], topright = pp[34];
            const int topleft = (j && y) ? pp[33] : cl;
            myrow[0] = fooabs(cl); myrow[1] = slog(cl); myrow[2] = cl + top - topleft; myrow[3] = topleft + topright - top;
            myrow[4] = slog(cl - topleft); myrow[5] = slog(topleft - top); myrow[6] = 0;   // x <= 1: leftleft = left
            COMPILER_FENCE();       // the asm loads below read these
            // Property 12 = slog(left - leftleft) needs pixel j-2, which is not decoded yet.  slog is monotone, so a test
            // "slog(cl - leftleft) > s" is "leftleft <= T" for a threshold T this lane can compute: the walk FORKS there, follows
            // both children to their leaves and leaves (A, B, T) for the decoder.  A second such test on either path is rare;
            // then the decoder finishes the walk itself from the first one.
            const bool stop12 = j > 1;
            unsigned ref = 0, phase = 0, leafA = 0, pendB = 0, forkref = 0, ry;
            int T = 0;
            for (;;) {
                const uint2 n = lds64(inner + ref * 8u);
                const unsigned off = n.x & 0xffu;
                unsigned c;
                if (stop12 && off == (unsigned)kOffLeftLeft) {
                    if (phase != 0u) { ry = (2u << 30) | forkref; break; }
                    const int sv = (int)n.x >> 8;
                    int dmin;       // slog(d) > sv  <=>  d >= dmin   (|d| <= 255 here)
                    if (sv >= 0) dmin = sv >= 9 ? 512 : (1 << sv);
                    else { const int t = -(sv + 1); dmin = t >= 9 ? -511 : -((1 << t) - 1); }
                    T = cl - dmin;
                    forkref = ref; phase = 1u; pendB = n.y >> 16;
                    c = n.y & 0xffffu;
                } else {
                    const int v = lds32((off >= 64u ? ld_s : pp_s) + off * 4u);
                    c = (v > ((int)n.x >> 8)) ? (n.y & 0xffffu) : (n.y >> 16);
                }
                if (c & 0x8000u) {
                    if (phase == 0u) { ry = c & 0x7fffu; break; }
                    if (phase == 1u) {
                        leafA = c & 0x7fffu; phase = 2u; c = pendB;
                        if (c & 0x8000u) { ry = (1u << 30) | ((c & 0x7fffu) << 15) | leafA; break; }
                    } else { ry = (1u << 30) | ((c & 0x7fffu) << 15) | leafA; break; }
                }
                ref = c;
            }
            sts64v(cand_s + (unsigned)(slot * kMaxCand + 32 * b + lane) * 8u, ((tagrow | (unsigned)(j + 1)) << 16) | ((unsigned)T & 0xffffu), ry);
            if (pf_leaves && (ry >> 30) < 2u) {     // pull the chances of the leaves this candidate leads to into L1 (a load whose
#ifndef FB_EMULATE
                unsigned dummy;                       // result nobody waits for), ahead of the decoder's cache miss
                asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(dummy) : "l"(pf_leaves + ((size_t)(ry & 0x7fffu) << pf_shift)));
                if (ry >> 30) asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(dummy) : "l"(pf_leaves + ((size_t)((ry >> 15) & 0x7fffu) << pf_shift)));
#endif
            }
            slot += g;
            if (slot >= K) slot -= K;
        }
        __syncwarp();
        if (lane == 0) mail->done[widx] = seen;
    }
}

// ---- lean single-lane coder for the run-ahead path ------------------------------------------------------------------
// Same arithmetic as Rac / read_int above, written without data-dependent branches inside a binary decision (the
// renormalisation is predicated), with shared-memory addresses and the leaf's first eight chances held in registers.
struct FRac {
    unsigned range, low, ones;
    const uint8_t *p;
    unsigned pos, n;            // byte offsets (helped groups require a file < 2 GiB)
};
// input(), rac.h:70-81 (the rare part of a decision: about one in ten).  After the first read past the end `low` is all
// ones for good (rac.h:64-69 ORs a sign-extended EOS into a 64-bit `low`): every later decision reads 1, and `low` is
// topped up again here before it could ever drop below a threshold (the thresholds between two calls sum to < 2^24).
__device__ __forceinline__ void fr_input(FRac &r) {
#pragma unroll 1
    for (int k = 0; k < 2 && r.range <= 0x10000u; k++) {
        unsigned c = 0xffu;
        if (r.pos < r.n) c = __ldg(r.p + r.pos); else r.ones = 1u;
        r.pos++;
        r.low = (r.low << 8) | c;
        r.range <<= 8;
    }
    if (r.ones) r.low = 0xffffffffu;
}
// one binary decision with chance `ch` (FinalCompoundSymbolBitCoder::read, compound.h:90-95 + RacInput::get, rac.h:82-95);
// the adapted chance goes back to shared memory at `slot`
__device__ __forceinline__ unsigned fr_step(FRac &r, unsigned ch, unsigned slot, unsigned tab_s) {
    const unsigned both = (unsigned)lds32(tab_s + ch * 4u);
    const unsigned chance = (unsigned)(((unsigned long long)r.range * ch + 0x800ull) >> 12);
    const unsigned thr = r.range - chance;
    const bool b = r.low >= thr;
    r.low = b ? r.low - thr : r.low;
    r.range = b ? chance : thr;
    sts16(slot, both >> (b ? 16 : 0));
    if (r.range <= 0x10000u) fr_input(r);
    return (unsigned)b;
}
struct SymConsts { int mn, mx, emax_pos, emax_neg; unsigned mant_off; };      // mant_off: byte offset of bit_mant[0] in a leaf
// reader<15>(coder, min, max), symbol.h:154-185.  SIGN_MODE 0: the sign is coded (min < 0 < max), 1: always positive, 2: always negative
template <int SIGN_MODE>
__device__ __forceinline__ int fread_int(FRac &r, unsigned leaf_s, const uint4 &L /* zero, sign, exp[0..5] of the leaf */, unsigned tab_s, const SymConsts &K) {
    if (fr_step(r, L.x & 0xffffu, leaf_s, tab_s)) return 0;
    unsigned sign;
    if (SIGN_MODE == 0) sign = fr_step(r, L.x >> 16, leaf_s + 2u, tab_s);
    else sign = SIGN_MODE == 1 ? 1u : 0u;
    const int amax = sign ? K.mx : -K.mn;
    const int emax = sign ? K.emax_pos : K.emax_neg;
    int e;
    {
        const unsigned ew[3] = {L.y, L.z, L.w};
        e = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            if (k >= emax) goto exp_done;
            if (fr_step(r, (k & 1) ? (ew[k >> 1] >> 16) : (ew[k >> 1] & 0xffffu), leaf_s + 4u + 2u * k, tab_s)) goto exp_done;
            e = k + 1;
        }
        for (; e < emax; e++)
            if (fr_step(r, lds16(leaf_s + 4u + 2u * e), leaf_s + 4u + 2u * e, tab_s)) break;
    }
exp_done:;
    int have = 1 << e;
    const unsigned mb = leaf_s + K.mant_off;
    // mantissa, top position first (symbol.h:174-182): a jump on e into a fall-through sequence with static positions
#define FB_MSTEP(p)                                                                                        \
    {                                                                                                      \
        const int minabs1 = have | (1 << (p));                                                             \
        if (minabs1 <= amax && fr_step(r, lds16(mb + 2u * (p)), mb + 2u * (p), tab_s)) have = minabs1;     \
    }
    switch (e) {
    case 14: FB_MSTEP(13)
    case 13: FB_MSTEP(12)
    case 12: FB_MSTEP(11)
    case 11: FB_MSTEP(10)
    case 10: FB_MSTEP(9)
    case 9: FB_MSTEP(8)
    case 8: FB_MSTEP(7)
    case 7: FB_MSTEP(6)
    case 6: FB_MSTEP(5)
    case 5: FB_MSTEP(4)
    case 4: FB_MSTEP(3)
    case 3: FB_MSTEP(2)
    case 2: FB_MSTEP(1)
    case 1: FB_MSTEP(0)
    default: break;
    }
#undef FB_MSTEP
    return sign ? have : -have;
}

__device__ __forceinline__ uint4 ldg128(const void *p) { return *reinterpret_cast<const uint4 *>(p); }
#ifdef FB_EMULATE
__device__ __forceinline__ void sts128(unsigned addr, const uint4 &v) { memcpy(fb_emu_smem(addr), &v, 16); }
#else
__device__ __forceinline__ void sts128(unsigned addr, const uint4 &v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
#endif
// shared-memory address of a leaf's chances (resident array, or a direct-mapped write-back cache over the global array) and its
// first eight chances; with the cache, the chances are loaded together with the tag and loaded again only after a miss
__device__ __forceinline__ unsigned leaf_addr_uniform(const RowState &R, unsigned leaf, uint4 &L) {
    if (!R.cached) { const unsigned a = R.lines_s + (leaf << R.line_shift); L = lds128(a); return a; }
    const unsigned slot = R.nlines ? leaf - __umulhi(leaf, R.inv_nlines) * R.nlines : (leaf & R.mask);
    const unsigned ta = R.tags_s + slot * 4u;
    const unsigned line = R.lines_s + (slot << R.line_shift);
    const int tag = lds32(ta);
    L = lds128(line);
    if (tag != (int)leaf) {
        // write the evicted leaf back and fetch the new one: 32 bytes (compact leaves) or 64
        const bool big = R.line_shift == 6u;
        if (tag >= 0) {
            uint4 *old = reinterpret_cast<uint4 *>(R.gleaves + ((size_t)tag << (R.line_shift - 1)));
            old[0] = L; old[1] = lds128(line + 16u);
            if (big) { old[2] = lds128(line + 32u); old[3] = lds128(line + 48u); }
        }
        const uint4 *src = reinterpret_cast<const uint4 *>(R.gleaves + ((size_t)leaf << (R.line_shift - 1)));
        const uint4 v0 = ldg128(src), v1 = ldg128(src + 1);
        sts128(line, v0); sts128(line + 16u, v1);
        if (big) { const uint4 v2 = ldg128(src + 2), v3 = ldg128(src + 3); sts128(line + 32u, v2); sts128(line + 48u, v3); }
        sts32v(ta, leaf);
        L = lds128(line);       // (not `L = v0`: a value that comes out of a generic load counts as divergent for the compiler, see RowState)
    }
    return line;
}

// the rest of a walk that a walker left at inner node `ref` (a test of property 12), with the real left / leftleft
__device__ __forceinline__ unsigned finish_walk(unsigned inner, unsigned ref, unsigned pp_s, int left, int leftleft, int x, int y) {
    const int top = lds32(pp_s + 32 * 4), topright = lds32(pp_s + 34 * 4);
    const int topleft = (x && y) ? lds32(pp_s + 33 * 4) : left;
    const int p0 = fooabs(left), p1 = slog(left), p2 = left + top - topleft, p3 = topleft + topright - top;
    const int p4 = slog(left - topleft), p5 = slog(topleft - top), p6 = slog(left - leftleft);
    for (;;) {
        const uint2 n = lds64(inner + ref * 8u);
        const unsigned off = n.x & 0xffu;
        int v;
        if (off < 64u) v = lds32(pp_s + off * 4u);
        else { const unsigned k = off - 64u; v = k == 0 ? p0 : k == 1 ? p1 : k == 2 ? p2 : k == 3 ? p3 : k == 4 ? p4 : k == 5 ? p5 : p6; }
        const unsigned c = (v > ((int)n.x >> 8)) ? (n.y & 0xffffu) : (n.y >> 16);
        if (c & 0x8000u) return c & 0x7fffu;
        ref = c;
    }
}

// The pixel loop of one row, a function of its own with NO arguments: everything comes out of shared memory, found through
// the block's configuration words, so that the compiler analyses the loop in isolation and sees only uniform operands.
template <int SIGN_MODE>
__device__ __noinline__ void row_ahead_loop() {
    const int lane = threadIdx.x & 31;
    const unsigned warp = __reduce_max_sync(0xffffffffu, threadIdx.x >> 5);
    const unsigned mail_s = s_cfg[0] + (warp >> s_cfg[1]) * s_cfg[2];
    RowState R;
    {
        const unsigned rs = mail_s + (unsigned)offsetof(Mail, row);
        unsigned *dst = reinterpret_cast<unsigned *>(&R);
#pragma unroll
        for (int k = 0; k < (int)(sizeof(RowState) / 4); k++) dst[k] = (unsigned)lds32(rs + 4u * k);
    }
    const unsigned cval_s = mail_s + (unsigned)offsetof(Mail, cval), cand_s = mail_s + (unsigned)offsetof(Mail, cand);
    const unsigned prol_s = mail_s + (unsigned)offsetof(Mail, prol_ready), prog_s = mail_s + (unsigned)offsetof(Mail, progress);
    const unsigned tab_s = R.tab_s, cprop_s = R.cprop_s;
    SymConsts K;
    K.mn = R.mn; K.mx = R.mx; K.emax_pos = R.emax_pos; K.emax_neg = R.emax_neg; K.mant_off = R.mant_off;
    FRac fr;
    fr.range = R.range; fr.low = R.low; fr.ones = R.ones; fr.p = R.p; fr.pos = R.pos; fr.n = R.n;
    const int w = R.w, zero = R.zero, cmin = R.cmin, y = R.y;
    int16_t *row = R.ch->data + (size_t)y * w;
    const int rowbase = y * w;
    int left = zero, leftleft = zero;
    unsigned slot = 0;
    for (int x0 = 0; x0 < w; x0 += 32) {
        const int cnt = min(32, w - x0);
#ifdef FB_EMULATE
        if (lane == 0)      // no SIMT lockstep in the emulator: the redundant lanes would see each other's chance updates
#endif
        for (int i = 0; i < cnt; i++) {         // all lanes, identical values
            const int xx = x0 + i;
            const unsigned want = R.tagrow | (unsigned)(xx + 1);
            const unsigned ca = cand_s + (slot * kMaxCand + (unsigned)(left - cmin)) * 8u;
            slot = slot + 1u == R.kslots ? 0u : slot + 1u;
            uint2 e;
            do { e = lds64v(ca); } while ((e.x >> 16) != want);
            unsigned res = e.y & 0x7fffu;
            const unsigned kind = e.y >> 30;
            if (kind == 1u) res = leftleft <= (int)(short)(e.x & 0xffffu) ? res : ((e.y >> 15) & 0x7fffu);
            else if (kind == 2u) {
                while ((int)lds32v(prol_s) < y * ((w + 31) >> 5) + (xx >> 5)) { }
                res = finish_walk(R.inner_s, res, cprop_s + (unsigned)((((xx >> 5) & 1) * 32 + (xx & 31)) * kPropStride) * 4u, left, leftleft, xx, y);
            }
#ifdef FB_EMU_TRACE
            if (getenv("FB_EMU_TRACE") && atoi(getenv("FB_EMU_TRACE")) == w * 1000 + R.ch->h) printf("[ahead] y %d x %d left %d leftleft %d kind %u leaf %u T %d\n", y, xx, left, leftleft, kind, res, (int)(short)(e.x & 0xffffu));
#endif
            uint4 L;
            const unsigned leaf_s = leaf_addr_uniform(R, res, L);
            const int diff = fread_int<SIGN_MODE>(fr, leaf_s, L, tab_s, K);
            const int val = s16(s16(diff) + zero);
            sts32v(cval_s + (unsigned)(xx & 63) * 4u, (want << 16) | ((unsigned)val & 0xffffu));
            sts32v(prog_s, (unsigned)(rowbase + xx + 1));
            leftleft = xx ? left : val;          // next pixel: x > 1 ? value(x-2) : left   (context_predict.h:132)
            left = val;
        }
        __syncwarp();
        if (x0 + lane < w) row[x0 + lane] = (int16_t)(lds32v(cval_s + (unsigned)((x0 + lane) & 63) * 4u) & 0xffffu);
        __syncwarp();
    }
    // the coder's state goes back the way it came
    const unsigned rs = mail_s + (unsigned)offsetof(Mail, row);
#ifdef FB_EMULATE
    if (lane == 0)      // (only lane 0 ran the loop there)
#endif
    {
    sts32v(rs + (unsigned)offsetof(RowState, range), fr.range);
    sts32v(rs + (unsigned)offsetof(RowState, low), fr.low);
    sts32v(rs + (unsigned)offsetof(RowState, ones), fr.ones);
    sts32v(rs + (unsigned)offsetof(RowState, pos), fr.pos);
    }
    __syncwarp();
}

// One row with run-ahead walkers (predictor 0).  Same results as decode_row.
__device__ __forceinline__ void decode_row_ahead(int sign_mode, DImage &img, DChan &ch, int y, const int *refchan, int nrefchan, int nref, Rac &rac,
                                                 const Smem &sm, const LeafStore &ls, unsigned inner_s, int nb, int lane) {
    Mail *mail = sm.mail;
    const int w = ch.w;
    chunk_prologue(img, ch, y, 0, refchan, nrefchan, nref, sm.cprop, lane);
    if (w > 32) chunk_prologue(img, ch, y, 32, refchan, nrefchan, nref, sm.cprop + 32 * kPropStride, lane);
    // every walker has read the previous command
    if (lane < sm.nwalkers) { const int cur = mail->cmd_seq; while (mail->ack[lane] != cur) { FB_SPIN(); } }
    __syncwarp();
    if (lane == 0) {
        const int zero = ch.zero, cmin = ch.minval, cmax = ch.maxval;
        const int g = max(1, min(8, (sm.nwalkers - 1) / nb)), kslots = g >= 5 ? g : g * ((8 + g - 1) / g);
        RowState &S = mail->row;
        S.w = w; S.y = y; S.zero = zero; S.cmin = cmin;
        S.mn = cmin - zero; S.mx = cmax - zero;                 // predictor 0: guess = zero
        S.emax_pos = ilog2u((unsigned)max(S.mx, 0)); S.emax_neg = ilog2u((unsigned)max(-S.mn, 0));
        S.mant_off = 2u * (unsigned)ls.mant_base;
        S.tagrow = (unsigned)(y & 1) << 15; S.kslots = (unsigned)kslots; S.inner_s = inner_s;
        S.lines_s = (unsigned)__cvta_generic_to_shared(ls.lines); S.line_shift = (unsigned)ls.shift + 1u;
        S.tab_s = (unsigned)__cvta_generic_to_shared(sm.table); S.cprop_s = (unsigned)__cvta_generic_to_shared(sm.cprop);
        S.cached = ls.tags ? 1u : 0u; S.mask = (unsigned)ls.mask; S.nlines = (unsigned)ls.nlines; S.inv_nlines = ls.nlines ? 0xffffffffu / (unsigned)ls.nlines + 1u : 0u; S.tags_s = ls.tags ? (unsigned)__cvta_generic_to_shared(ls.tags) : 0u;
        S.range = rac.range; S.low = rac.ones ? 0xffffffffu : rac.low; S.ones = rac.ones ? 1u : 0u;
        S.pos = (unsigned)rac.io.pos; S.n = (unsigned)rac.io.n; S.p = rac.io.p; S.gleaves = ls.gleaves;
        S.img = &img; S.ch = &ch; S.cprop = sm.cprop; S.nrefchan = nrefchan; S.nref = nref;
        for (int k = 0; k < 16; k++) S.refchan[k] = k < nrefchan ? refchan[k] : 0;
        mail->y = y; mail->w = w; mail->cmin = cmin; mail->nb = nb; mail->g = g; mail->K = kslots;
        mail->inner_saddr = inner_s; mail->tagrow = S.tagrow; mail->cmd = 1;
        mail->pf_leaves = (ls.tags && sm.prefetch) ? ls.gleaves : nullptr; mail->pf_shift = ls.shift;
        mail->prol_widx = sm.nwalkers - 1; mail->prol_ready = y * ((w + 31) >> 5) + 1;       // chunks 0 and 1 were computed above
        __threadfence_block();
        mail->cmd_seq = mail->cmd_seq + 1;
    }
    __syncwarp();
    if (sign_mode == 0) row_ahead_loop<0>();
    else if (sign_mode == 1) row_ahead_loop<1>();
    else row_ahead_loop<2>();
    if (lane == 0) {
        const RowState &S = mail->row;
        const unsigned ones = *(volatile const unsigned *)&S.ones;
        rac.range = *(volatile const unsigned *)&S.range; rac.low = *(volatile const unsigned *)&S.low; rac.ones = ones != 0u;
        rac.io.pos = *(volatile const unsigned *)&S.pos; rac.io.eof = rac.io.eof || (ones != 0u); rac.io.avail = 0; rac.io.win = 0;
    }
    __syncwarp();
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
#ifdef FB_EMU_TRACE
                if (lane == 0 && getenv("FB_EMU_TRACE") && atoi(getenv("FB_EMU_TRACE")) == w * 1000 + ch.h) printf("[plain] y %d x %d left %d leftleft %d leaf %u\n", y, xx, left, leftleft, cur.x & 0xffffu);
#endif
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
__device__ bool decode_group(DImage &img, Reader &io, int &beginc, int limit, const Params &P, WarpScratch &ws, int lane, const Smem &sm) {
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
    // `limit`: the planes of this stream end there.  A group that reaches beyond them (a group index that does not belong to the file, a
    // damaged header) would write planes another stream owns -- and could lower their `rows_done` after the owner released it, which
    // leaves every stream that waits for those rows spinning forever.  Corrupt, not garbage.
    if (endc >= img.nch || endc < beginc || endc >= limit) return false;
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
        // an inverted range only comes out of a damaged header (the reference runs into its asserts there): corrupt, like a failed depth check
        if (ch.maxval < ch.minval || (compress && !check_bit_depth(ch.minval, ch.maxval, predictor))) { early = true; early_result = false; break; }
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
    if (lane == 0) {
        tree_ok = read_tree(rac, P.meta_table, pr, nprops, ws.nodes, nnodes, ws.stack, sm.coder) ? 1 : 0;
        if (!tree_ok && nnodes < 0) img.status = kStatusTreeTooLarge;
    }
    tree_ok = __shfl_sync(0xffffffffu, tree_ok, 0);
    nnodes = __shfl_sync(0xffffffffu, nnodes, 0);
    if (!tree_ok) { const bool st = STOPPED(); SYNC_IO(); return corrupt_or_truncated(st, img.ch[beginc], lane); }

    // FinalPropertySymbolCoder ctor, compound.h:213-225: leaf numbering in node order, all leaves start from zero_chance
    const int nleaves = (nnodes + 1) / 2;
    int group_range = 0, group_maxw = 0, group_lo = 0, group_hi = 0;
    for (int i = beginc; i <= endc; i++) {
        group_range = max(group_range, img.ch[i].maxval - img.ch[i].minval + 1); group_maxw = max(group_maxw, img.ch[i].w);
        group_lo = min(group_lo, img.ch[i].minval); group_hi = max(group_hi, img.ch[i].maxval);
    }
    // Run-ahead walkers (see walker_main): predictor 0, value range <= 256 with one walker per block of 32 candidates, the
    // compact inner-node array in shared memory (+ room for some leaves), 24-bit split values, 15-bit x tags, 32-bit offsets.
    const int ninner = nnodes / 2;
    const int inner_bytes = (ninner * 8 + 15) & ~15;
    bool ahead = sm.mail && predictor == 0 && nnodes > 1 && group_range >= 1 && (group_range + 31) / 32 <= sm.nwalkers - 1 && group_range <= kMaxCand &&
                 inner_bytes + 16384 <= sm.dyn_bytes && group_maxw <= 32766 && img.nbytes < 0x7fff0000ull && group_lo >= -32000 && group_hi <= 32000;
    if (ahead) {
        bool wide = false;
        for (int i = lane; i < nnodes; i += 32) { const int sv = ws.nodes[i].splitval; wide = wide || sv < -(1 << 23) || sv >= (1 << 23); }
        if (__any_sync(0xffffffffu, wide)) ahead = false;
    }
    // shared-memory plan for this group: node cache (if it fits), then leaf lines
    int dyn_off = 0;
    uint2 *snodes = nullptr;
    const int node_bytes = ((nnodes + 2) * 8 + 15) & ~15;
    if (ahead) dyn_off = inner_bytes;
    else if (node_bytes + 64 * 8 + 64 <= sm.dyn_bytes) { snodes = reinterpret_cast<uint2 *>(sm.dyn); dyn_off = node_bytes; }
    LeafStore ls;
    ls.gleaves = ws.leaves;
    const int rem = sm.dyn_bytes - dyn_off;
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
        ls.lines = reinterpret_cast<uint16_t *>(sm.dyn + dyn_off); ls.tags = nullptr; ls.mask = 0; ls.nlines = 0;
        for (int i = lane; i < (nleaves << ls.shift); i += 32) ls.lines[i] = init_entry(i & ((1 << ls.shift) - 1));
    } else {                                    // direct-mapped cache over the global leaf array
        int nlines = 1;
        while (nlines * 2 * (lbytes + 4) <= rem) nlines *= 2;
        ls.nlines = 0;
        if (ahead) { nlines = (rem - 16) / (lbytes + 4); ls.nlines = nlines; }     // leaf % nlines: no power-of-two rounding
        ls.tags = reinterpret_cast<int *>(sm.dyn + dyn_off);
        ls.lines = reinterpret_cast<uint16_t *>(sm.dyn + dyn_off + ((nlines * 4 + 15) & ~15));
        ls.mask = nlines - 1;
        for (int i = lane; i < nlines; i += 32) ls.tags[i] = -1;
        for (int i = lane; i < (nleaves << ls.shift); i += 32) ws.leaves[i] = init_entry(i & ((1 << ls.shift) - 1));
    }
    uint2 *gpacked = reinterpret_cast<uint2 *>(ws.stack);      // the parse stack is free again: reuse it for the packed nodes
    const uint2 *nodes2 = gpacked;
    unsigned inner_s = 0;
    if (ahead) {
        // compact inner-node array (layout: see walker_main).  idx[i] = 0x8000 | leaf id (node order, compound.h:213-225) or inner index
        int *idx = ws.stack;
        int leaf_base = 0, inner_base = 0;
        for (int i0 = 0; i0 < nnodes; i0 += 32) {
            const int i = i0 + lane;
            const bool valid = i < nnodes;
            const bool isleaf = valid && ws.nodes[i].property == -1;
            const unsigned lm = __ballot_sync(0xffffffffu, isleaf), vm = __ballot_sync(0xffffffffu, valid);
            const unsigned im = vm & ~lm, lt = (1u << lane) - 1u;
            if (valid) idx[i] = isleaf ? (0x8000 | (leaf_base + __popc(lm & lt))) : (inner_base + __popc(im & lt));
            leaf_base += __popc(lm); inner_base += __popc(im);
        }
        __syncwarp();
        uint2 *sinner = reinterpret_cast<uint2 *>(sm.dyn);
        for (int i = lane; i < nnodes; i += 32) {
            const TNode nd = ws.nodes[i];
            if (nd.property == -1) continue;
            const int pidx = nd.property, role = pidx - nref;
            int off = pidx;
            if (role == 1) off = 64; else if (role == 3) off = 65; else if (role == 6) off = 66; else if (role == 7) off = 67;
            else if (role == 8) off = 68; else if (role == 9) off = 69; else if (role == 12) off = kOffLeftLeft;
            uint2 e;
            e.x = ((unsigned)nd.splitval << 8) | (unsigned)off;
            e.y = (unsigned)idx[nd.child] | ((unsigned)idx[nd.child + 1] << 16);
            sinner[idx[i]] = e;
        }
        inner_s = (unsigned)__cvta_generic_to_shared(sinner);
    } else {
        // node cache: packed entries, see struct LeafStore comment
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
        if (snodes) {
            for (int i = lane; i < nnodes; i += 32) snodes[i + 1] = gpacked[i + 1];
            nodes2 = snodes;
        }
    }
    __syncwarp();
    if (P.debug && lane == 0)
        printf("[maniac]   tree %d nodes, nprops %d nref %d, zero_chance %d, pos %llu, %dx%d, nodes %s, leaves %s\n", nnodes, nprops, nref, predictability,
               rac.io.pos, img.ch[beginc].w, img.ch[beginc].h, ahead ? "run-ahead walkers" : (snodes ? "smem" : "global"), ls.tags ? "cached" : "resident");

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
            const bool helped = ahead;
            const int nwalk = (range + 31) / 32;
            const int sign_mode = (ch.minval - ch.zero < 0) ? ((ch.maxval - ch.zero > 0) ? 0 : 2) : 1;
            if (helped) {       // handshake rings start out with tag 0 (never awaited); the walkers are idle here
                if (lane == 0) { sm.mail->progress = 0; sm.mail->prol_ready = 0; }
                for (int k = lane; k < 64; k += 32) sm.mail->cval[k] = 0;
                for (int k = lane; k < kCandSlotsMax * kMaxCand; k += 32) { (&sm.mail->cand[0][0])[k].x = 0; (&sm.mail->cand[0][0])[k].y = 0; }
                __syncwarp();
            }
            const long long t_start = clock64();
            for (int y = 0; y < ch.h; y++) {
                if (STOPPED()) break;
                for (int r = 0; r < nrefchan; r++) {       // row wavefront on the planes this row back-references
                    const DChan &cj = img.ch[refchan[r]];
                    int ry = (y << ch.vshift) >> cj.vshift;
                    if (ry >= cj.h) ry = cj.h - 1;
                    spin_until_ge(&cj.rows_done, ry + 1);
                }
                if (helped) {
                    decode_row_ahead(sign_mode, img, ch, y, refchan, nrefchan, nref, rac, sm, ls, inner_s, nwalk, lane);
                } else if (snodes) {
                    if (predictor == 0) decode_row<true, true>(img, ch, y, predictor, refchan, nrefchan, nref, rac, sm, nodes2, ls, lane);
                    else decode_row<true, false>(img, ch, y, predictor, refchan, nrefchan, nref, rac, sm, nodes2, ls, lane);
                } else {
                    if (predictor == 0) decode_row<false, true>(img, ch, y, predictor, refchan, nrefchan, nref, rac, sm, nodes2, ls, lane);
                    else decode_row<false, false>(img, ch, y, predictor, refchan, nrefchan, nref, rac, sm, nodes2, ls, lane);
                }
                publish_rows(ch, y + 1, lane);
            }
            if (helped) {       // stragglers (walks of candidate blocks nobody needed) must be over before the tree or the rings change
                if (lane < sm.nwalkers) { const int cur = sm.mail->cmd_seq; while (sm.mail->done[lane] != cur) { FB_SPIN(); } }
                __syncwarp();
            }
            if (P.debug && lane == 0) {
                const double ns = (double)ch.w * ch.h;
                printf("[maniac]   ch %d %dx%d range %d nodes %d helped %d walkers %d: %.0f cycles/symbol\n", i, ch.w, ch.h, range, nnodes, (int)helped, nwalk,
                       (double)(clock64() - t_start) / ns);
            }
        }
        if (STOPPED()) break;
    }
    beginc = endc;
    SYNC_IO();
    return true;
#undef SYNC_IO
#undef STOPPED
}

__global__ void __launch_bounds__(512, 1) k_maniac_decode(Params P) {
    FB_DYN_SMEM_DECL(smem_raw);
    // the warp index goes through a warp reduction so that the compiler knows it (and every shared-memory address derived
    // from it) is uniform across the warp: see RowState
    const int lane = threadIdx.x & 31, warp = (int)__reduce_max_sync(0xffffffffu, threadIdx.x >> 5);
    const int wps = 1 + P.helpers;                       // warps per stream: decoder + walkers
    const int slot = warp / wps;
    // Role of this warp inside its stream.  Warps are bound to schedulers by (warp id mod 4): the decoder of stream slot s sits
    // on scheduler s mod 4 and the other warps of its stream on that scheduler stay idle, so every decoder has an issue
    // port to itself (two streams per SM: schedulers 0 and 1).
    const int drole = slot & 3;
    const int wraw = warp % wps;
    // 0 = decoder, -1 = idle, k > 0 = walker k-1 (walkers numbered in warp order, skipping the warps of the decoder's scheduler)
    const int wrole = wraw == drole ? 0 : ((wraw & 3) == drole ? -1 : 1 + wraw - (wraw + 3 - drole) / 4);
    uint16_t *s_table = reinterpret_cast<uint16_t *>(smem_raw);                      // 16 KiB, shared by the block's warps
    for (int i = threadIdx.x; i < 4096 * 2; i += blockDim.x) s_table[i] = P.table[i];
    if (threadIdx.x == 0) {
        s_cfg[0] = (unsigned)__cvta_generic_to_shared(smem_raw) + 16384u + 256u + (P.helpers ? 2u * 4736u : 4736u);
        s_cfg[1] = P.helpers == 15 ? 4u : (P.helpers == 7 ? 3u : 0u);
        s_cfg[2] = (unsigned)P.warp_smem;
    }
    __syncthreads();
    unsigned char *mine = smem_raw + 16384 + (size_t)slot * P.warp_smem;
    Smem sm;
    sm.table = s_table;
    sm.coder = reinterpret_cast<uint16_t(*)[32]>(mine);                               // 192 B
    const int cprop_bytes = P.helpers ? 2 * 4736 : 4736;         // chunk properties, double-buffered when walkers run ahead
    const int nwalkers = P.helpers - P.helpers / 4;
    const int mail_bytes = P.helpers ? kMailBytes + nwalkers * 32 * kLdRowStride * 4 : 0;
    sm.cprop = reinterpret_cast<int *>(mine + 256);
    sm.mail = P.helpers ? reinterpret_cast<Mail *>(mine + 256 + cprop_bytes) : nullptr;
    sm.ldrows = P.helpers ? reinterpret_cast<int *>(mine + 256 + cprop_bytes + kMailBytes) : nullptr;
    sm.nwalkers = min(nwalkers, P.walkers_used);
    sm.prefetch = P.prefetch;
    sm.dyn = mine + 256 + cprop_bytes + mail_bytes;
    sm.dyn_bytes = P.warp_smem - 256 - cprop_bytes - mail_bytes;
    if (P.helpers) {
        if (wrole == 0 && lane == 0) {
            sm.mail->cmd_seq = 0; sm.mail->cmd = 0;
            for (int k = 0; k < kMaxWalkers; k++) { sm.mail->ack[k] = 0; sm.mail->done[k] = 0; }
        }
        __syncthreads();
        // warps 4, 8, 12 of a stream stay idle: the decoder warp keeps its scheduler (warp id mod 4) to itself
        if (wrole != 0) { if (wrole > 0) walker_main(sm.mail, sm.cprop, sm.ldrows, wrole - 1, lane, P.walker_sleep); return; }
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
                bool ok = decode_group(img, io, i, st.end_channel, P, ws, lane, sm);
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

#ifdef FB_EMULATE
}  // namespace  (the emulator harness, which includes this file, takes it from here)
#else
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


// The streams of one file: one per channel group when the caller supplied the groups' byte offsets (the group index), walking
// the channel list exactly like the loop of fuif_decode (encoding.cpp:708-718); otherwise the whole file as one stream.
// S = DStream (GPU backend) or fbh::Stream (host-threads backend): same fields.
template <class S>
void build_streams(const FbManiacJob &job, int b, std::vector<S> &streams) {
    const fb_image *img = job.img;
    const int nch = (int)img->ch.size();
    if (nch <= 0) return;
    bool indexed = job.group_index && job.n_groups > 0;
    if (indexed) {
        int i = 0;
        std::vector<S> mine;
        for (int g = 0; g < job.n_groups && indexed; g++) {
            while (i < nch && (!img->ch[i].d.w || !img->ch[i].d.h)) i++;
            size_t pos = (size_t)job.group_index[g];
            if (i >= nch || job.group_index[g] < (int64_t)job.body_pos || pos >= job.nbytes) { indexed = false; break; }
            if (job.bytes_to_load && pos >= job.bytes_to_load) break;
            if (job.group_first) {
                // the first stream must start at the first plane there is: planes before it would belong to no stream, nobody would
                // ever release them, and a stream that references them would wait forever
                if (mine.empty() && job.group_first[g] != i) { indexed = false; break; }
                i = job.group_first[g];
                if (i < 0 || i >= nch || (!mine.empty() && i <= mine.back().first_channel)) { indexed = false; break; }
            } else {
                if (job.bytes_dev) { indexed = false; break; }      // cannot read group headers of a device buffer on the host
                int fb = host_varint(job.bytes_host, job.nbytes, pos);
                if (fb < 0) { indexed = false; break; }
            }
            if (!mine.empty()) mine.back().end_channel = i;
            mine.push_back(S{b, i, nch, 1, (unsigned long long)job.group_index[g]});
            if (!job.group_first) { size_t p2 = (size_t)job.group_index[g]; int fb = host_varint(job.bytes_host, job.nbytes, p2); i += (fb >> 4) + 1; }
        }
        if (indexed && !mine.empty()) streams.insert(streams.end(), mine.begin(), mine.end());
        else indexed = false;
    }
    if (!indexed) streams.push_back(S{b, 0, nch, -1, (unsigned long long)job.body_pos});
}

// group-major ticket order for batches: the k-th stream of every image before any (k+1)-th stream.  A stream still only depends
// on lower tickets (earlier groups of its own image), and the large late groups of all images end up running together instead
// of trailing image by image.
template <class S>
void group_major_order(std::vector<S> &streams, int nimg) {
    if (nimg <= 1) return;
    std::vector<int> ord(streams.size());
    std::vector<int> seen(nimg, 0);
    for (size_t k = 0; k < streams.size(); k++) ord[k] = seen[streams[k].image]++;
    std::vector<size_t> idx(streams.size());
    for (size_t k = 0; k < idx.size(); k++) idx[k] = k;
    std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return ord[a] < ord[b]; });
    std::vector<S> sorted(streams.size());
    for (size_t k = 0; k < idx.size(); k++) sorted[k] = streams[idx[k]];
    streams.swap(sorted);
}

// what every backend leaves in the image: ranges / q / zero of the planes, which planes hold data, where the groups start
template <class C>
int finish_images(fb_ctx *ctx, std::vector<FbManiacJob> &jobs, const std::vector<C> &back, const std::vector<int> &status) {
    size_t coff = 0;
    int rc = FB_OK;
    for (size_t b = 0; b < jobs.size(); b++) {
        fb_image *img = jobs[b].img;
        for (size_t i = 0; i < img->ch.size(); i++) {
            const C &d = back[coff + i];
            FbChan &c = img->ch[i];
            c.d.minval = d.minval; c.d.maxval = d.maxval; c.d.zero = d.zero; c.d.q = d.q;
            if (d.state) c.d.decoded = 1;
            else { fb_plane_free(ctx, c.dev); c.dev = nullptr; c.host = nullptr; c.d.decoded = 0; }
            if (d.group_off >= 0) { img->group_off.push_back(d.group_off); img->group_first.push_back((int32_t)i); }
        }
        if (status[b] == FB_ERR_UNSUPPORTED) { ctx->err = "max_properties > 18 is not supported by the GPU context model"; rc = FB_ERR_UNSUPPORTED; }
        else if (status[b] == kStatusTreeTooLarge) { ctx->err = "a MANIAC tree of this file has more than 65535 nodes: not supported by this decoder (image " + std::to_string(b) + ")"; rc = FB_ERR_UNSUPPORTED; }
        else if (status[b] == FB_ERR_NOMEM) { ctx->err = "out of memory while decoding image " + std::to_string(b); rc = FB_ERR_NOMEM; }
        else if (status[b]) { ctx->err = "corrupt FUIF stream (image " + std::to_string(b) + ")"; rc = FB_ERR_INVALID; }
        coff += img->ch.size();
    }
    return rc;
}

// ---- host-threads backend (FB_OPT_ENTROPY_BACKEND = FB_ENTROPY_HOST) ---------------------------------------
// fb_host_entropy.cpp decodes into host memory -- pinned staging of the context, or the image's own block for a host-only image --
// one thread per channel group (with the group index) or per image; the planes are then copied to HBM, where the transform
// chain runs as with the GPU backend.  100 MB of planes of a 4096^2 image cross PCIe in a few ms against >= 1 s of decoding.
// Three phases; the middle one is pure CPU work and touches nothing of the context.  (Tried on top of this split: a hybrid
// backend that gave part of a batch's images to k_maniac_decode and the rest to the host threads at the same time.  No gain on
// 64 x 1080p -- the GPU side takes 1.8 s for 35 images as for 64, its time is the longest single stream -- so it was dropped:
// profiles/r02_hybrid_sweep_cfg4_experiment.txt.)
struct HostRun {
    std::vector<std::vector<uint8_t>> file_copy;
    std::vector<fbh::Image> himg;
    std::vector<fbh::Chan> hch;
    std::vector<fbh::Stream> streams;
    int cutoff = 0, threads_used = 0;
    uint32_t alpha = 0;
    bool gpu = true;
};

int host_prepare(fb_ctx *ctx, std::vector<FbManiacJob> &jobs, HostRun &R) {
    const int nimg = (int)jobs.size();
    R.gpu = ctx->device >= 0;
    R.file_copy.resize(nimg);
    R.himg.resize(nimg);
    R.cutoff = jobs[0].cutoff; R.alpha = (uint32_t)jobs[0].alpha;
    std::vector<size_t> plane_off;      // per plane: offset (in samples) into its block
    size_t total_ch = 0, total_samples = 0;
    for (int b = 0; b < nimg; b++) {
        FbManiacJob &job = jobs[b];
        if (job.cutoff != jobs[0].cutoff || job.alpha != jobs[0].alpha) { ctx->err = "batch with mixed maniac options"; return FB_ERR_INVALID; }
        if (job.bytes_dev) {        // the file lives in HBM: this backend reads it on the host
            R.file_copy[b].resize(job.nbytes);
            FB_CUDA(ctx, cudaMemcpyAsync(R.file_copy[b].data(), job.bytes_dev, job.nbytes, cudaMemcpyDeviceToHost, ctx->stream));
            FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            job.bytes_host = R.file_copy[b].data();
            job.bytes_dev = nullptr;
        }
        if (!R.gpu) total_samples = 0;     // host-only images own their block: offsets restart per image
        for (auto &c : job.img->ch) {
            const size_t n = (c.d.w > 0 && c.d.h > 0) ? (size_t)c.d.w * c.d.h : 0;
            plane_off.push_back(total_samples);
            total_samples += (n + 31) & ~(size_t)31;
        }
        if (!R.gpu) { job.img->host_block.assign(total_samples + 32, 0); job.img->on_host = true; }
        total_ch += job.img->ch.size();
    }
    int16_t *stage = nullptr;
    if (R.gpu) {
        const size_t need = std::max<size_t>(total_samples * sizeof(int16_t), 4096);
        if (need > ctx->host_stage_bytes) {
            FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (ctx->host_stage) cudaFreeHost(ctx->host_stage);
            ctx->host_stage = nullptr; ctx->host_stage_bytes = 0;
            FB_CUDA(ctx, cudaHostAlloc(&ctx->host_stage, need, cudaHostAllocDefault));
            ctx->host_stage_bytes = need;
        }
        stage = (int16_t *)ctx->host_stage;
    }
    R.hch.resize(total_ch);
    size_t coff = 0;
    for (int b = 0; b < nimg; b++) {
        FbManiacJob &job = jobs[b];
        fb_image *img = job.img;
        fbh::Image &hi = R.himg[b];
        hi.bytes = job.bytes_host; hi.nbytes = job.nbytes; hi.bytes_to_load = job.bytes_to_load;
        hi.ch = R.hch.data() + coff; hi.nch = (int)img->ch.size(); hi.max_properties = job.max_properties;
        hi.n_orig = img->info.real_nb_channels; hi.status = 0;
        int16_t *base = R.gpu ? stage : img->host_block.data();
        for (size_t i = 0; i < img->ch.size(); i++) {
            FbChan &c = img->ch[i];
            fbh::Chan &d = R.hch[coff + i];
            memset(&d, 0, sizeof(d));
            d.w = c.d.w; d.h = c.d.h; d.minval = c.d.minval; d.maxval = c.d.maxval; d.zero = c.d.zero; d.q = c.d.q;
            d.hshift = c.d.hshift; d.vshift = c.d.vshift; d.group_off = -1;
            d.data = base + plane_off[coff + i];
            if (!(c.d.w > 0 && c.d.h > 0)) { d.hdr_done = 1; d.rows_done = 0x7fffffff; }     // empty planes are skipped by the channel loop
        }
        build_streams(job, b, R.streams);
        coff += img->ch.size();
    }
    group_major_order(R.streams, nimg);
    return FB_OK;
}

void host_run(HostRun &R, int threads) {
    R.threads_used = fbh::decode(R.himg.data(), (int)R.himg.size(), R.streams.data(), (int)R.streams.size(), R.cutoff, R.alpha, threads);
}

// planes to HBM (or they stay where they are, for a host-only image), then what every backend leaves in the image
int host_finish(fb_ctx *ctx, std::vector<FbManiacJob> &jobs, HostRun &R) {
    const int nimg = (int)jobs.size();
    ctx->host_threads_used = R.threads_used;
    size_t coff = 0;
    for (int b = 0; b < nimg; b++) {
        fb_image *img = jobs[b].img;
        for (size_t i = 0; i < img->ch.size(); i++) {
            FbChan &c = img->ch[i];
            const fbh::Chan &d = R.hch[coff + i];
            if (!d.state) continue;
            if (!R.gpu) { c.host = d.data; continue; }
            const size_t n = (size_t)d.w * d.h;
            int rc = fb_plane_alloc(ctx, n, &c.dev);
            if (rc) return rc;
            FB_CUDA(ctx, cudaMemcpyAsync(c.dev, d.data, n * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
        }
        coff += img->ch.size();
    }
    if (R.gpu) FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // the staging is free for the next call
    std::vector<int> status(nimg);
    for (int b = 0; b < nimg; b++) status[b] = R.himg[b].status;
    int rc = finish_images(ctx, jobs, R.hch, status);
    if (rc == FB_ERR_UNSUPPORTED) ctx->err = "more properties than the host context model holds (max_properties > 100)";
    return rc;
}

int maniac_decode_host(fb_ctx *ctx, std::vector<FbManiacJob> &jobs) {
    HostRun R;
    int rc = host_prepare(ctx, jobs, R);
    if (rc) return rc;
    host_run(R, ctx->host_threads);
    return host_finish(ctx, jobs, R);
}

int maniac_decode_gpu(fb_ctx *ctx, std::vector<FbManiacJob> &jobs);

}  // namespace

void fb_maniac_release(fb_ctx *ctx) {
    ManiacState *st = (ManiacState *)ctx->maniac_state;
    if (!st) return;
    cudaFree(st->table_dev); cudaFree(st->meta_dev); cudaFree(st->ticket_dev); cudaFree(st->arena); cudaFree(st->scratch_dev);
    delete st;
    ctx->maniac_state = nullptr;
}

int fb_maniac_decode(fb_ctx *ctx, std::vector<FbManiacJob> &jobs) {
    if (jobs.empty()) return FB_OK;
    if (ctx->entropy_backend == FB_ENTROPY_HOST || ctx->device < 0) return maniac_decode_host(ctx, jobs);
    return maniac_decode_gpu(ctx, jobs);
}

namespace {
int maniac_decode_gpu(fb_ctx *ctx, std::vector<FbManiacJob> &jobs) {
    const int nimg = (int)jobs.size();
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
        build_streams(job, b, streams);
        if (!job.bytes_dev) boff += (job.nbytes + 255) & ~(size_t)255;
        coff += img->ch.size();
    }
    FB_CUDA(ctx, cudaMemcpyAsync(img_dev, himg.data(), nimg * sizeof(DImage), cudaMemcpyHostToDevice, ctx->stream));
    group_major_order(streams, nimg);
    const int nstreams = (int)streams.size();
    if (nstreams) {
        // streams are ordered image-major, groups ascending: dependencies always point to lower stream ids
        FB_CUDA(ctx, cudaMallocAsync((void **)&streams_dev, nstreams * sizeof(DStream), ctx->stream));
        FB_CUDA(ctx, cudaMemcpyAsync(streams_dev, streams.data(), nstreams * sizeof(DStream), cudaMemcpyHostToDevice, ctx->stream));
        // Throughput shape for big batches: 8 streams per block, each ONE warp on the one-warp decode path (no walkers).  A stream is
        // slower (no run-ahead) but four times as many are in flight; measured on 64 x 1080p (3520 streams): 41.5 -> 71.1 Mpx/s
        // (4 per block: 61.4, 16 per block: 47.3 -- 13 KB of shared memory per stream is too little for the leaf cache).  Taken when
        // the batch has at least 12 streams per SM; FB_MANIAC_SPB=n forces n (3..16) from sm_count * n streams on, 0 / 1 disables.
        static const int env_spb = getenv("FB_MANIAC_SPB") ? atoi(getenv("FB_MANIAC_SPB")) : -1;
        int spb_many = 0;
        if (env_spb > 2 && env_spb <= 16) { if (nstreams >= ctx->sm_count * env_spb) spb_many = env_spb; }
        else if (env_spb < 0 && nstreams >= ctx->sm_count * 12) spb_many = 8;
        const int nslots = spb_many ? ctx->sm_count * spb_many : std::min((nstreams + 1) / 2 * 2, ctx->sm_count * 2);        // scratch slots: one per stream in flight
        int rc = ensure_state(ctx, cutoff, alpha, nslots, maxw);
        if (rc) return rc;
        ManiacState *st = (ManiacState *)ctx->maniac_state;
        FB_CUDA(ctx, cudaMemsetAsync(st->ticket_dev, 0, sizeof(int), ctx->stream));
        Params P;
        P.images = img_dev; P.streams = streams_dev; P.nstreams = nstreams; P.ticket = st->ticket_dev;
        P.table = st->table_dev; P.meta_table = st->meta_dev; P.scratch = st->scratch_dev; P.maxw = st->maxw;
        P.walker_sleep = getenv("FB_MANIAC_WSLEEP") ? atoi(getenv("FB_MANIAC_WSLEEP")) : 100;
        P.walkers_used = getenv("FB_MANIAC_WUSED") ? atoi(getenv("FB_MANIAC_WUSED")) : 64;
        P.prefetch = getenv("FB_MANIAC_NO_PREFETCH") ? 0 : 1;
        P.debug = getenv("FB_MANIAC_DEBUG") ? std::max(1, atoi(getenv("FB_MANIAC_DEBUG"))) : 0;
        // Launch shape.  Few streams (one image): one warp per block and block per SM with ~200 KiB of shared memory, so
        // that the whole MANIAC tree and most leaf chances of a stream stay on-chip.  Many streams (batches): up to 8
        // warps share a block's 16 KiB chance table and up to two blocks share an SM.
        // Launch shape.  A block serves `spb` streams at a time, each with one decoder warp and `helpers` walker warps, and
        // owns one SM (~200 KiB of shared memory split between its streams: tree-node cache, leaf chances, per-chunk
        // property rows).  Few streams (one image) => one stream per SM with 8 walkers (value ranges up to 256); batches =>
        // two streams per SM with 6 walkers each (ranges up to 192; wider ranges fall back to the decoder's own walk).
        const int per_sm = (nstreams + ctx->sm_count - 1) / ctx->sm_count;
        const int wpb = spb_many ? spb_many : std::max(1, std::min(2, per_sm));          // streams per block
        P.helpers = (getenv("FB_MANIAC_NO_WALKERS") || spb_many) ? 0 : (wpb == 1 ? 15 : 7);     // 16 / 8 warps per stream: decoder, 12 / 6 walkers, idle warps
        const int nblocks = std::min((nstreams + wpb - 1) / wpb, ctx->sm_count);
        const size_t block_smem = 226 * 1024;
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
    std::vector<int> status(nimg);
    for (int b = 0; b < nimg; b++) status[b] = himg[b].status;
    return finish_images(ctx, jobs, back, status);
}
}  // namespace
#endif  // FB_EMULATE
