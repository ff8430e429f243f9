// Host side of fuif_encode (reference encoding/encoding.cpp:455-573): everything around the per-group entropy coding, which
// runs on the GPU (fb_maniac_enc.cu).  Plain C++ with no CUDA in it, so that the CPU test tier can run exactly this code
// around the emulated kernel (tests/emu/emu_maniac_enc.cpp).
//
//   plan_groups        which channels share a group, their predictor, and where in the rand() sequence each group's
//                      learning pass starts (encoding.cpp:500-527, 180-186; Image::recompute_downscales, image.cpp:124-136)
//   glibc_rand         the row order of the learning pass: the reference calls libc rand() from its initial state
//                      (encoding.cpp:185), i.e. glibc's TYPE_3 additive feedback generator seeded with 1
//   build_chance_table build_table, maniac/chance.cpp:31-65;  build_log4k: Log4kTable, chance.cpp:67-91
//   assemble           container: magic, header varints, transform list, the group byte strings in order, the five
//                      responsive offsets -- including BlobIO's "bytes_used runs one byte ahead" behaviour
//                      (fileio.h:245-251) and the tail a rolled-back compressed attempt leaves behind, so that the file
//                      is the reference's byte for byte
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

namespace fbenc_host {

struct Plane {          // Channel, image/image.h:54-91, without the samples
    int w, h, minval, maxval, zero, q, hshift, vshift, hcshift, vcshift;
};
struct Options {        // fuif_options, encoding/encoding.h:32-59 (encode side)
    float nb_repeats = 0.5f;
    int max_properties = 12;
    int maniac_cutoff = 6;
    int maniac_alpha = 0x0d000000;
    bool compress = true;
    int max_group = -1;
    std::vector<int> predictor;
};
struct Group {
    int beginc, endc, predictor;
    long long rand_off;         // index of the first rand() value of this group's learning pass
    long long learned;          // symbols its learning pass visits (an upper bound on the leaves it can create)
    long long pixels;
};
struct GroupBytes {             // what the kernel leaves for a group
    const unsigned char *bytes; // max(out_len, attempt_len) valid bytes
    unsigned out_len;           // the group's final form
    unsigned attempt_len;       // length of the compressed form when it was rolled back in favour of the plain one, else 0
};
struct Transform { int id; std::vector<int> params; };
struct ImageInfo { int w, h, maxval, colormodel, real_nb_channels, nb_channels, nb_meta_channels; };

inline void chan_setzero(Plane &p) {        // Channel::setzero, image/image.h:70-74
    if (p.minval > 0) p.zero = p.minval;
    else if (p.maxval < 0) p.zero = p.maxval;
    else p.zero = 0;
}

// Image::recompute_downscales, image/image.cpp:124-136
inline void recompute_downscales(const std::vector<Plane> &ch, const ImageInfo &info, int ds[6]) {
    const int n = (int)ch.size();
    ds[0] = info.nb_meta_channels + info.nb_channels - 1;
    for (int s = 1; s < 6; s++) {
        ds[s] = n - 1;
        for (int k = ds[s - 1]; k < n; k++) {
            const int rs = 32 >> s;
            if ((1 << ch[k].hcshift) < rs || (1 << ch[k].vcshift) < rs) break;
            if ((1 << ch[k].hcshift) == rs && (1 << ch[k].vcshift) == rs) ds[s] = k;
        }
    }
}

// rand() calls of the learning loop for a plane of h rows (encoding.cpp:180-203): the loop runs while
// (float)++rowslearned <= nb_repeats * h -- and, because of how it resets its row variable, at most once when h == 1
inline long long learn_rows(int h, float nb_repeats) {
    long long n = 0;
    int rows = 0;
    for (int y = 0; y < h; y = 1) {         // the loop variable is reset to 0 after every learned row and then incremented
        if ((float)++rows > nb_repeats * (float)h) break;
        n++;
    }
    return n;
}

// the groups in stream order; zeroes/ranges must be final (fuif_prepare_encode has run)
inline std::vector<Group> plan_groups(const std::vector<Plane> &ch, const ImageInfo &info, const Options &o, long long *nrand) {
    std::vector<Group> out;
    int ds[6];
    recompute_downscales(ch, info, ds);
    const int n = (int)ch.size();
    long long roff = 0;
    for (int i = 0; i < n; i++) {
        if (!ch[i].w || !ch[i].h) continue;
        int predictor = 0;
        if ((int)o.predictor.size() > i) predictor = o.predictor[i];
        else if (!o.predictor.empty()) predictor = o.predictor.back();
        int j = i;
        if (o.compress) {
            for (int s = 1; s < 5; s++) if (j > ds[s] && j < ds[s + 1]) j = ds[s + 1];
            for (int k = i + 1; k <= j; k++) if (ch[i].w != ch[k].w || ch[i].h != ch[k].h) { j = k - 1; break; }
            if (o.max_group > 0 && j > i + o.max_group - 1) j = i + o.max_group - 1;
        }
        Group g;
        g.beginc = i; g.endc = j; g.predictor = predictor; g.rand_off = roff; g.learned = 0; g.pixels = 0;
        for (int k = i; k <= j; k++) {
            g.pixels += (long long)ch[k].w * ch[k].h;
            if (o.compress && ch[k].minval != ch[k].maxval) {
                const long long rows = learn_rows(ch[k].h, o.nb_repeats);
                roff += rows;
                g.learned += rows * ch[k].w;
            }
        }
        out.push_back(g);
        i = j;
    }
    if (nrand) *nrand = roff;
    return out;
}

// glibc random_r TYPE_3 (degree 31, separation 3) as srand(1) / rand() run it
inline void glibc_rand(int *out, long long n) {
    std::vector<uint32_t> r((size_t)(n + 344 + 34));
    int32_t word = 1;
    r[0] = 1;
    for (int i = 1; i < 31; i++) {
        const long hi = word / 127773, lo = word % 127773;
        word = (int32_t)(16807 * lo - 2836 * hi);
        if (word < 0) word += 2147483647;
        r[(size_t)i] = (uint32_t)word;
    }
    for (int i = 31; i < 34; i++) r[(size_t)i] = r[(size_t)(i - 31)];
    for (size_t i = 34; i < r.size(); i++) r[i] = r[i - 31] + r[i - 3];
    for (long long k = 0; k < n; k++) out[k] = (int)(r[(size_t)(k + 344)] >> 1);
}

// build_table, maniac/chance.cpp:31-65.  t = newchance[4096][2]: [c][1] after a one, [c][0] after a zero
inline void build_chance_table(uint16_t *t, uint32_t factor, unsigned max_p) {
    const uint64_t one = 1ull << 32;
    const unsigned size = 4096;
    memset(t, 0, sizeof(uint16_t) * size * 2);
    unsigned last_p8 = 0;
    uint64_t p = one / 2;
    for (unsigned i = 0; i < size / 2; i++) {
        unsigned p8 = (unsigned)((size * p + one / 2) >> 32);
        if (p8 <= last_p8) p8 = last_p8 + 1;
        if (last_p8 && last_p8 < size && p8 <= max_p) t[last_p8 * 2 + 1] = (uint16_t)p8;
        p += ((one - p) * factor + one / 2) >> 32;
        last_p8 = p8;
    }
    for (unsigned i = size - max_p; i <= max_p; i++) {
        if (t[i * 2 + 1]) continue;
        p = (i * one + size / 2) / size;
        p += ((one - p) * factor + one / 2) >> 32;
        unsigned p8 = (unsigned)((size * p + one / 2) >> 32);
        if (p8 <= i) p8 = i + 1;
        if (p8 > max_p) p8 = max_p;
        t[i * 2 + 1] = (uint16_t)p8;
    }
    for (unsigned i = 1; i < size; i++) t[i * 2] = (uint16_t)(size - t[(size - i) * 2 + 1]);
}

// Log4kTable, maniac/chance.cpp:67-91: cost of a decision with chance i/4096, in 1/5461 bit
inline void build_log4k(uint16_t *out /*[4097]*/) {
    const uint32_t base = (65535u << 16) / 12;
    out[0] = 0;
    for (uint32_t i = 1; i <= 4096; i++) {
        int bits = 0;
        while ((i >> bits) != 0) bits++;
        uint64_t y = (uint64_t)i << (32 - bits);
        uint32_t res = base * (uint32_t)(13 - bits);
        uint32_t add = base;
        while (add > 1 && (y & 0x7FFFFFFF) != 0) {
            y = (y * y + 0x40000000) >> 31;
            add >>= 1;
            if ((y >> 32) != 0) { res -= add; y >>= 1; }
        }
        out[i] = (uint16_t)((res + (1u << 15)) >> 16);
    }
}

// BlobIO as fuif_encode uses it (fileio.h:148-272): bytes_used runs one byte ahead of the last byte written, seeking back
// does not shrink it, and fresh memory reads as zero here (the reference leaves it uninitialised)
struct Blob {
    std::vector<uint8_t> data;
    size_t pos = 0, used = 0;
    void putc(int c) {
        if (pos + 2 > data.size()) data.resize((pos + 2) * 3 / 2 + 4096, 0);
        data[pos++] = (uint8_t)c;
        if (used < pos) used = pos + 1;
    }
    void write(const unsigned char *p, size_t n) {
        if (!n) return;
        if (pos + n + 1 > data.size()) data.resize((pos + n + 1) * 3 / 2 + 4096, 0);
        memcpy(data.data() + pos, p, n);
        // what n calls of putc do to bytes_used: it jumps to pos + 1 whenever it has fallen behind pos, i.e. every other byte
        const size_t end = pos + n;
        size_t first = pos + 1 > used + 1 ? pos + 1 : used + 1;     // first position that finds bytes_used behind
        if (first <= end) used = first + 2 * ((end - first) / 2) + 1;
        pos = end;
    }
    void varint(size_t number) {        // write_big_endian_varint, encoding.cpp:30-41
        unsigned char tmp[12];
        int n = 0;
        tmp[n++] = (unsigned char)(number & 127);
        number >>= 7;
        while (number) { tmp[n++] = (unsigned char)((number & 127) | 128); number >>= 7; }
        while (n) putc(tmp[--n]);
    }
};

inline bool transform_has_parameters(int id) {      // Transform::has_parameters, transform/transform.h:85-104
    return id == 3 || id == 4 || id == 6 || id == 7 || id == 8 || id == 9 || id == 10;
}

// The file.  group_offsets (optional): byte offset of every group's header in the file -- the sidecar index with which
// fb_decode() decodes the groups concurrently.
inline std::vector<uint8_t> assemble(const ImageInfo &info, const std::vector<Transform> &tr, const std::vector<Plane> &ch, const Options &o,
                                     const std::vector<Group> &groups, const std::vector<GroupBytes> &gb, std::vector<int64_t> *group_offsets) {
    Blob real;      // FileIO: plain appends
    for (const char *m = "FUIF"; *m; m++) real.putc(*m);
    real.varint((size_t)(info.real_nb_channels + '0'));
    int bit_depth = 1, maxval = 1;
    while (maxval < info.maxval) { bit_depth++; maxval = maxval * 2 + 1; }
    real.varint((size_t)(bit_depth + '&'));
    real.varint((size_t)(info.w - 1));
    real.varint((size_t)(info.h - 1));
    real.varint((size_t)info.colormodel);
    real.varint((size_t)o.max_properties);
    Blob io;
    std::vector<int64_t> before_of;
    if (info.real_nb_channels >= 1) {
        io.varint(tr.size());
        for (const Transform &t : tr) {
            const size_t np = transform_has_parameters(t.id) ? t.params.size() : 0;
            io.varint((np << 4) + (size_t)t.id);
            for (size_t j = 0; j < np; j++) io.varint((size_t)t.params[j]);
        }
        int ds[6];
        recompute_downscales(ch, info, ds);
        long long responsive[5] = {-1, -1, -1, -1, -1};
        for (size_t g = 0; g < groups.size(); g++) {
            const size_t before = io.pos;
            before_of.push_back((int64_t)before);
            if (gb[g].attempt_len) {        // the compressed form was written first, then the position went back
                io.write(gb[g].bytes, gb[g].attempt_len > gb[g].out_len ? gb[g].attempt_len : gb[g].out_len);
                io.pos = before + gb[g].out_len;
            } else io.write(gb[g].bytes, gb[g].out_len);
            if (o.compress)
                for (int s = 0; s < 5; s++) if (ds[s] >= groups[g].beginc && ds[s] <= groups[g].endc) responsive[s] = (long long)io.pos;
        }
        long long relative = 0;
        for (int s = 0; s < 5; s++) {
            if (responsive[s] < 0) responsive[s] = (long long)io.pos;
            real.varint((size_t)(responsive[s] - relative));        // TRUNCATION_OFFSET_RESOLUTION == 1
            relative = responsive[s];
        }
    }
    std::vector<uint8_t> out(real.pos + io.used);
    if (real.pos) memcpy(out.data(), real.data.data(), real.pos);
    if (io.used) memcpy(out.data() + real.pos, io.data.data(), io.used);
    if (group_offsets) {
        group_offsets->clear();
        for (int64_t b : before_of) group_offsets->push_back(b + (int64_t)real.pos);
    }
    return out;
}

}  // namespace fbenc_host
