// TEST INFRASTRUCTURE ONLY: runs the MANIAC ENCODE kernel source (fuif_b200/csrc/fb_maniac_enc.cu) under the CPU
// execution-model emulator, one warp per channel group, so that the CPU-only test tier can compare its per-group byte
// strings with the oracle encoder (which is byte-exact against the reference).  Built by tests/test_emu_maniac_enc.py.
#define FB_EMULATE 1
#include "../../fuif_b200/csrc/fb_maniac_enc.cu"

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
    std::vector<std::vector<long long>> scr((size_t)ngroups);
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
        scr[(size_t)g].resize(5 * 32);
        G.fleaves = fleaves[(size_t)g].data(); G.stack = stacks[(size_t)g].data(); G.scr = scr[(size_t)g].data();
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

}  // extern "C"
