// TEST INFRASTRUCTURE ONLY: runs the MANIAC decode kernel SOURCE of the product (fuif_b200/csrc/fb_maniac.cu, device part)
// under the CPU execution-model emulator (cuemu.h + maniac_emu_shim.h), so that the CPU-only test tier can compare the
// kernel's logic -- stream tickets, row wavefront, run-ahead walkers with forks, prologue warp, leaf cache, integer coder --
// with the oracle.  Built by tests/test_emu_maniac.py with  g++ -DFB_EMULATE.  The set-up below mirrors the host half of
// fb_maniac_decode() (descriptors, one stream per channel group, scratch, launch shape).
#define FB_EMULATE 1
#include "../../fuif_b200/csrc/fb_maniac.cu"

extern "C" {

// chdesc[nch][5] = w, h, hshift, vshift, q after meta_apply (what the library's host half takes from its channel list);  planes[nch] = w*h int16 each (outputs);  chout[nch][5] = minval, maxval, zero, q, holds samples
// ngroups > 0: one stream per channel group (group_off / group_first), else one stream for the whole file
// shape: 0 = one stream per block with 15 extra warps (single image), 1 = two streams per block with 7 extra warps each (batches)
// bytes_to_load: 0 = everything, else the responsive truncation point (-R k)
// returns the image status (0 = ok)
int emu_maniac_decode(const uint8_t *bytes, size_t nbytes, size_t body_pos, int max_properties, int n_orig, int nch, const int *chdesc,
                      int16_t **planes, int *chout, int ngroups, const long long *group_off, const int *group_first, int shape, int nblocks,
                      int cutoff, int alpha, int smem_kib, int debug, size_t bytes_to_load) {
    std::vector<uint8_t> file(nbytes + 16, 0);
    memcpy(file.data(), bytes, nbytes);
    std::vector<DChan> ch((size_t)nch);
    int maxw = 8;
    for (int i = 0; i < nch; i++) {
        DChan &d = ch[(size_t)i];
        memset(&d, 0, sizeof(d));
        d.w = chdesc[5 * i]; d.h = chdesc[5 * i + 1]; d.hshift = chdesc[5 * i + 2]; d.vshift = chdesc[5 * i + 3]; d.q = chdesc[5 * i + 4];
        d.group_off = -1;
        d.data = planes[i];
        if (!(d.w > 0 && d.h > 0)) { d.hdr_done = 1; d.rows_done = 0x7fffffff; }
        maxw = std::max(maxw, d.w);
    }
    DImage img;
    img.bytes = file.data(); img.nbytes = nbytes; img.bytes_to_load = bytes_to_load; img.ch = ch.data(); img.nch = nch;
    img.max_properties = max_properties; img.n_orig = n_orig; img.status = 0;
    std::vector<DStream> streams;
    if (ngroups > 0) {
        for (int g = 0; g < ngroups; g++) {
            DStream st;
            st.image = 0; st.first_channel = group_first[g]; st.end_channel = g + 1 < ngroups ? group_first[g + 1] : nch; st.max_groups = 1;
            st.offset = (unsigned long long)group_off[g];
            streams.push_back(st);
        }
    } else {
        DStream st;
        st.image = 0; st.first_channel = 0; st.end_channel = nch; st.max_groups = -1; st.offset = body_pos;
        streams.push_back(st);
    }
    std::vector<uint16_t> table(4096 * 2), meta(4096 * 2);
    build_table(meta.data(), 0xFFFFFFFFu / 19, 4096 - 2);
    build_table(table.data(), (uint32_t)alpha, (unsigned)(4096 - cutoff));
    const int wpb = shape == 1 ? 2 : (shape == 3 ? 8 : 1);     // shape 3: the throughput shape of big batches, 8 one-warp streams per block
    Params P;
    memset(&P, 0, sizeof(P));
    P.helpers = (shape == 2 || shape == 3) ? 0 : (shape ? 7 : 15);      // shapes 2, 3: no walkers at all (the one-warp path for every group)
    const int nslots = nblocks * wpb;
    std::vector<WarpScratch> ws((size_t)nslots);
    std::vector<std::vector<unsigned char>> arena((size_t)nslots);
    const size_t nodes_b = sizeof(TNode) * kMaxNodes, leaves_b = sizeof(uint16_t) * 32 * (kMaxNodes / 2), stack_b = sizeof(int) * 4 * (kMaxNodes / 2 + 2);
    const size_t refs_b = sizeof(int16_t) * (size_t)maxw * 12 + 256;
    for (int i = 0; i < nslots; i++) {
        arena[(size_t)i].assign(nodes_b + leaves_b + stack_b + refs_b, 0xEE);
        unsigned char *base = arena[(size_t)i].data();
        ws[(size_t)i].nodes = (TNode *)base;
        ws[(size_t)i].leaves = (uint16_t *)(base + nodes_b);
        ws[(size_t)i].stack = (int *)(base + nodes_b + leaves_b);
        ws[(size_t)i].refs = (int16_t *)(base + nodes_b + leaves_b + stack_b);
    }
    int ticket = 0;
    P.images = &img; P.streams = streams.data(); P.nstreams = (int)streams.size(); P.ticket = &ticket;
    P.table = table.data(); P.meta_table = meta.data(); P.scratch = ws.data(); P.maxw = maxw;
    P.debug = debug; P.walker_sleep = 100; P.walkers_used = 64; P.prefetch = 1;
    const size_t block_smem = (size_t)smem_kib * 1024;
    const size_t warp_smem = ((block_smem - 16384) / (size_t)wpb) & ~(size_t)15;
    P.warp_smem = (int)warp_smem;
    const size_t smem_bytes = 16384 + warp_smem * (size_t)wpb;
    cuemu::launch((unsigned)nblocks, (unsigned)(32 * wpb * (1 + P.helpers)), smem_bytes, nblocks > 1, [&]() { k_maniac_decode(P); });
    for (int i = 0; i < nch; i++) {
        const DChan &d = ch[(size_t)i];
        chout[5 * i] = d.minval; chout[5 * i + 1] = d.maxval; chout[5 * i + 2] = d.zero; chout[5 * i + 3] = d.q; chout[5 * i + 4] = d.state;
    }
    return img.status;
}

}  // extern "C"
