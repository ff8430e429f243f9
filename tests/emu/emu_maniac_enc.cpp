// TEST INFRASTRUCTURE ONLY: runs the MANIAC ENCODE kernel source (fuif_b200/csrc/fb_maniac_enc.cu) under the CPU
// execution-model emulator, one warp per channel group, so that the CPU-only test tier can compare its per-group byte
// strings with the oracle encoder (which is byte-exact against the reference).  Built by tests/test_emu_maniac_enc.py.
#define FB_EMULATE 1
#include "../../fuif_b200/csrc/fb_maniac_enc.cu"
#include "../../fuif_b200/csrc/fb_encode_host.h"

extern "C" {

// chdesc[nch][8] = w, h, minval, maxval, zero, q, hshift, vshift;  planes[nch] = samples (NULL for empty planes)
// gdesc[ngroups][4] = beginc, endc, predictor, rand offset;  out[ngroups] = byte buffers of out_cap bytes each
// glen[ngroups][3] (out) = bytes written, header bytes, status
int emu_maniac_encode(int nch, const int *chdesc, int16_t **planes, int ngroups, const long long *gdesc, unsigned char **out, unsigned out_cap,
                      int *glen, int max_properties, float nb_repeats, int compress, const uint16_t *table, const uint16_t *meta_table,
                      const uint16_t *log4k, const int *rnd, long long nrnd, int leaf_cap) {
    using namespace fbenc;
    std::vector<EChan> ch((size_t)nch);
    for (int i = 0; i < nch; i++) {
        EChan &c = ch[(size_t)i];
        const int *d = chdesc + 8 * i;
        c.w = d[0]; c.h = d[1]; c.minval = d[2]; c.maxval = d[3]; c.zero = d[4]; c.q = d[5]; c.hshift = d[6]; c.vshift = d[7];
        c.data = planes[i];
    }
    std::vector<EGroup> groups((size_t)ngroups);
    std::vector<std::vector<TNode>> nodes((size_t)ngroups);
    std::vector<std::vector<LLeaf>> leaves((size_t)ngroups);
    std::vector<std::vector<uint16_t>> fleaves((size_t)ngroups);
    std::vector<std::vector<int>> stacks((size_t)ngroups);
    for (int g = 0; g < ngroups; g++) {
        EGroup &G = groups[(size_t)g];
        memset(&G, 0, sizeof(G));
        G.beginc = (int)gdesc[4 * g]; G.endc = (int)gdesc[4 * g + 1]; G.predictor = (int)gdesc[4 * g + 2]; G.rand_off = gdesc[4 * g + 3];
        G.compress = compress;
        nodes[(size_t)g].resize(kMaxNodes);
        leaves[(size_t)g].resize((size_t)leaf_cap);
        fleaves[(size_t)g].resize((size_t)(kMaxNodes / 2) * 32);
        stacks[(size_t)g].resize((size_t)8 * (kMaxNodes / 2 + 2));
        G.nodes = nodes[(size_t)g].data(); G.leaves = leaves[(size_t)g].data(); G.leaf_cap = leaf_cap;
        G.fleaves = fleaves[(size_t)g].data(); G.stack = stacks[(size_t)g].data();
        G.out = out[g]; G.out_cap = out_cap;
    }
    EParams P;
    P.ch = ch.data(); P.nch = nch; P.groups = groups.data(); P.ngroups = ngroups; P.max_properties = max_properties; P.nb_repeats = nb_repeats;
    P.table = table; P.meta_table = meta_table; P.log4k = log4k; P.rnd = rnd; P.nrnd = nrnd;
    cuemu::launch((unsigned)ngroups, 32, 0, false, [&]() { k_maniac_encode(P); });
    int rc = 0;
    for (int g = 0; g < ngroups; g++) {
        glen[3 * g] = (int)groups[(size_t)g].out_len; glen[3 * g + 1] = (int)groups[(size_t)g].header_len; glen[3 * g + 2] = groups[(size_t)g].status;
        if (groups[(size_t)g].status) rc = groups[(size_t)g].status;
    }
    return rc;
}

// The whole of fb_encode() with the kernel emulated: the host side is the product's own (fb_encode_host.h), only the CUDA
// memory plumbing of fb_maniac_enc.cu is replaced by host vectors.
// chdesc[nch][10] = w, h, minval, maxval, zero, q, hshift, vshift, hcshift, vcshift (ranges tight);  info[7] = w, h, maxval,
// colormodel, real_nb_channels, nb_channels, nb_meta_channels;  tdesc = for every transform: id, nparams, params...
// Returns the file length (or -status); copies at most out_cap bytes; goffs[cap_groups] gets the groups' file offsets.
long long emu_fuif_encode(int nch, const int *chdesc, int16_t **planes_in, const int *info_in, int ntr, const int *tdesc, float nb_repeats, int max_properties,
                          int cutoff, int alpha, int compress, int max_group, int npred, const int *pred, unsigned char *out, long long out_cap,
                          long long *goffs, int *gfirst, int cap_groups, int *ngroups_out) {
    using namespace fbenc;
    namespace H = fbenc_host;
    H::Options o;
    o.nb_repeats = nb_repeats; o.max_properties = max_properties; o.maniac_cutoff = cutoff; o.maniac_alpha = alpha; o.compress = compress != 0; o.max_group = max_group;
    o.predictor.assign(pred, pred + npred);
    std::vector<H::Plane> planes((size_t)nch);
    std::vector<EChan> ch((size_t)nch);
    for (int i = 0; i < nch; i++) {
        const int *d = chdesc + 10 * i;
        H::Plane &p = planes[(size_t)i];
        p.w = d[0]; p.h = d[1]; p.minval = d[2]; p.maxval = d[3]; p.zero = d[4]; p.q = d[5]; p.hshift = d[6]; p.vshift = d[7]; p.hcshift = d[8]; p.vcshift = d[9];
        if (p.w > 0 && p.h > 0 && !(p.minval == 0 && p.maxval == 0)) H::chan_setzero(p);
        EChan &c = ch[(size_t)i];
        c.w = p.w; c.h = p.h; c.minval = p.minval; c.maxval = p.maxval; c.zero = p.zero; c.q = p.q; c.hshift = p.hshift; c.vshift = p.vshift;
        c.data = planes_in[i];
    }
    H::ImageInfo info;
    info.w = info_in[0]; info.h = info_in[1]; info.maxval = info_in[2]; info.colormodel = info_in[3]; info.real_nb_channels = info_in[4];
    info.nb_channels = info_in[5]; info.nb_meta_channels = info_in[6];
    std::vector<H::Transform> tr;
    for (int i = 0, k = 0; i < ntr; i++) {
        H::Transform t;
        t.id = tdesc[k]; const int np = tdesc[k + 1];
        t.params.assign(tdesc + k + 2, tdesc + k + 2 + np);
        k += 2 + np;
        tr.push_back(t);
    }
    long long nrand = 0;
    std::vector<H::Group> groups;
    if (info.real_nb_channels >= 1) groups = H::plan_groups(planes, info, o, &nrand);
    const int ng = (int)groups.size();
    std::vector<uint16_t> table(8192), meta(8192), log4k(4097);
    H::build_chance_table(table.data(), (uint32_t)alpha, (unsigned)(4096 - cutoff));
    H::build_chance_table(meta.data(), 0xFFFFFFFFu / 19, 4096 - 2);
    H::build_log4k(log4k.data());
    std::vector<int> rnd((size_t)nrand + 1);
    H::glibc_rand(rnd.data(), nrand);
    std::vector<EGroup> eg((size_t)ng);
    std::vector<std::vector<TNode>> nodes((size_t)ng);
    std::vector<std::vector<LLeaf>> leaves((size_t)ng);
    std::vector<std::vector<uint16_t>> fleaves((size_t)ng);
    std::vector<std::vector<int>> stacks((size_t)ng);
    std::vector<std::vector<unsigned char>> outs((size_t)ng);
    for (int g = 0; g < ng; g++) {
        const H::Group &G = groups[(size_t)g];
        long long cap = G.learned + 1;
        if (cap > kMaxNodes / 2) cap = kMaxNodes / 2;
        if (cap < 2) cap = 2;
        long long tree_bytes = 24 * G.learned;
        if (tree_bytes > (512 << 10)) tree_bytes = 512 << 10;
        EGroup &E = eg[(size_t)g];
        memset(&E, 0, sizeof(E));
        E.beginc = G.beginc; E.endc = G.endc; E.predictor = G.predictor; E.compress = o.compress ? 1 : 0; E.rand_off = G.rand_off;
        nodes[(size_t)g].resize(kMaxNodes); leaves[(size_t)g].resize((size_t)cap); fleaves[(size_t)g].resize((size_t)(kMaxNodes / 2) * 32);
        stacks[(size_t)g].resize((size_t)8 * (kMaxNodes / 2 + 2));
        // cudaMalloc does not clear memory: the kernel must not depend on zeroed working buffers (the output buffer is cleared by fb_encode)
        memset((void *)nodes[(size_t)g].data(), 0xCD, nodes[(size_t)g].size() * sizeof(TNode));
        memset((void *)leaves[(size_t)g].data(), 0xCD, leaves[(size_t)g].size() * sizeof(LLeaf));
        memset((void *)fleaves[(size_t)g].data(), 0xCD, fleaves[(size_t)g].size() * sizeof(uint16_t));
        memset((void *)stacks[(size_t)g].data(), 0xCD, stacks[(size_t)g].size() * sizeof(int));
        outs[(size_t)g].assign((size_t)(4 * G.pixels + tree_bytes + 4096), 0);
        E.nodes = nodes[(size_t)g].data(); E.leaves = leaves[(size_t)g].data(); E.leaf_cap = (int)cap; E.fleaves = fleaves[(size_t)g].data();
        E.stack = stacks[(size_t)g].data(); E.out = outs[(size_t)g].data(); E.out_cap = (unsigned)outs[(size_t)g].size();
    }
    EParams P;
    P.ch = ch.data(); P.nch = nch; P.groups = eg.data(); P.ngroups = ng; P.max_properties = max_properties; P.nb_repeats = nb_repeats;
    P.table = table.data(); P.meta_table = meta.data(); P.log4k = log4k.data(); P.rnd = rnd.data(); P.nrnd = nrand;
    if (ng) cuemu::launch((unsigned)ng, 32, 0, false, [&]() { k_maniac_encode(P); });
    std::vector<H::GroupBytes> gb((size_t)ng);
    for (int g = 0; g < ng; g++) {
        const EGroup &E = eg[(size_t)g];
        if (E.status) return -(long long)E.status;
        if (E.attempt_len > E.out_cap) return -9;
        gb[(size_t)g].bytes = E.out; gb[(size_t)g].out_len = E.out_len; gb[(size_t)g].attempt_len = E.attempt_len;
    }
    std::vector<int64_t> offs;
    std::vector<uint8_t> file = H::assemble(info, tr, planes, o, groups, gb, &offs);
    if (ngroups_out) *ngroups_out = ng;
    for (int g = 0; g < ng && g < cap_groups; g++) { goffs[g] = offs[(size_t)g]; gfirst[g] = groups[(size_t)g].beginc; }
    const long long n = (long long)file.size();
    memcpy(out, file.data(), (size_t)(n < out_cap ? n : out_cap));
    return n;
}

// the first n values of the rand() restatement (checked against libc by the test)
void emu_glibc_rand(int *out, long long n) { fbenc_host::glibc_rand(out, n); }

}  // extern "C"
