// Shared declarations of the fuif_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/fuif_b200.h"

// ---- host-side objects behind the opaque C-ABI handles -------------------------------------------------------

struct fb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    long long launches = 0;
    int sm_count = 148;
    // MANIAC decoder resources (fb_maniac.cu), created lazily
    void *maniac_state = nullptr;
    // fused unsqueeze (fb_fused_squeeze.cuh) device counters, 8 ints laid out as fq::VerifyParams::counters says;
    // fq_mode (FB_OPT_SQUEEZE_MODE): 0 default (direct per-step kernels), 1 tiled per-step kernels only, 2 fused tile
    // kernels + forced serial fallback, 3 fused + forced repair of every tile of the last launch, 4 fused
    int *fq_counters = nullptr;
    int fq_mode = 0;
    // packed unsqueeze kernels (fb_pk_squeeze.cuh): scratch for est / act / bad, arrival counters (kept at zero by the
    // kernels), [0] repaired segments [1] range-flagged segments; pk_mode: 1 = use them where eligible (default), 0 = off
    unsigned char *pk_scratch = nullptr;
    size_t pk_scratch_bytes = 0;
    int *pk_counters = nullptr;
    int pk_counters_n = 0;
    int *pk_stats = nullptr;
    int pk_mode = 1;
    int sq_maxval = -1;         // maxval of the image whose Squeeze is being undone (packed kernels: 0 .. 1023 only)
    // Plane memory recycled on the host side: every plane is used on this context's ONE stream, so a freed plane can be handed out
    // again without a driver call (60 cudaMallocAsync + 60 cudaFreeAsync per 4096^2 undo_transforms cost more host time than the
    // whole chain takes on the GPU).  Keyed by the rounded byte size; emptied by fb_ctx_destroy or when it holds more than 16 GiB.
    std::unordered_map<size_t, std::vector<void *>> plane_pool;
    size_t plane_pool_bytes = 0;
    std::unordered_map<void *, size_t> plane_sizes;       // every live plane allocation of this context -> its rounded size
    // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: the opt-ins are remembered per context (= per
    // device), not per process, so that a context on a second GPU of the same process gets them too
    unsigned smem_optin = 0;
    // Entropy backend (FB_OPT_ENTROPY_BACKEND): 0 = k_maniac_decode on the GPU (default), 1 = host threads (fb_host_entropy.cpp),
    // planes uploaded afterwards.  host_threads 0 = four per hardware thread, at most one per stream.  host_stage: pinned staging the host backend decodes
    // into (grow-only, reused from call to call).  device < 0 marks the context of a host-only image (fb_host_decode): no CUDA at all.
    int entropy_backend = 0, host_threads = 0, host_threads_used = 0;
    void *host_stage = nullptr;
    size_t host_stage_bytes = 0;
    enum { kOptHsqTiled = 1, kOptPyramid = 2, kOptDirect = 4, kOptFq = 8, kOptPkH = 16 /* << variant, 5 bits */ };
    // FB_KERNEL_TIMING=1: a CUDA event after every launch, dumped by fb_ctx_synchronize (development aid)
    bool timing = false, timing_stderr = false;
    struct Mark { std::string name; cudaEvent_t ev; double bytes; };
    std::vector<Mark> marks;
    // bytes: algorithmic HBM bytes of the launch just enqueued (0 = not accounted)
    void mark(const char *name, double bytes = 0) {
        if (!timing) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, stream);
        marks.push_back(Mark{name, e, bytes});
    }
};

struct FbChan {
    fb_plane_desc d;          // mirrors Channel (reference image/image.h:54-91)
    int16_t *dev = nullptr;   // w*h samples in HBM, row-major, no padding; nullptr = not decoded (data.size()==0)
    int16_t *host = nullptr;  // host-only images (fb_host_decode): the samples in host memory (inside fb_image::host_block)
};

struct FbXform {
    int id;
    std::vector<int> p;
};

struct fb_image {
    fb_ctx *ctx = nullptr;
    fb_image_info info{};
    std::vector<FbChan> ch;
    std::vector<FbXform> tr;
    std::vector<int64_t> group_off;
    std::vector<int32_t> group_first;
    // host-only image (made by fb_host_decode, no GPU involved): planes live in host_block until fb_image_upload moves them
    bool on_host = false, owns_ctx = false;
    std::vector<int16_t> host_block;
};

#define FB_CUDA(ctx, call)                                                                          \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                       \
            return FB_ERR_CUDA;                                                                     \
        }                                                                                           \
    } while (0)

// plane memory (stream-ordered pool)
int fb_plane_alloc(fb_ctx *ctx, size_t nsamples, int16_t **out);
void fb_plane_free(fb_ctx *ctx, int16_t *p);
void fb_plane_pool_release(fb_ctx *ctx);

// ---- transform launchers (fb_transforms.cu).  All enqueue on ctx->stream and bump ctx->launches. -------------

// inverse Squeeze steps (reference transform/squeeze.h:81-132, 173-224). res may be nullptr (all-zero residual).
int fb_launch_inv_hsqueeze(fb_ctx *ctx, const int16_t *avg, const int16_t *res, int16_t *out, int wa, int wr, int h);
int fb_launch_inv_vsqueeze(fb_ctx *ctx, const int16_t *avg, const int16_t *res, int16_t *out, int w, int ha, int hr);
// one squeeze step over up to four planes in one launch (tiled kernels with warm-up + in-block verification)
int fb_launch_inv_squeeze_batch(fb_ctx *ctx, int horizontal, int n, const int16_t *const *avg, const int16_t *const *res, int16_t *const *out,
                                const int *wa, const int *wr, const int *ha, const int *hr);
// a planned unsqueeze step on one plane (all buffers already allocated); ops are listed in execution order
struct FbSqOp {
    int step, horizontal;
    const int16_t *avg, *res;
    int16_t *out;
    int wa, wr, ha, hr;
};
// what follows the last step of the plan and may be fused into its final launch
struct FbSqEpilogue {
    int kind;                   // 0 none, 1 clamp every final plane, 2 inverse YCoCg on ycc[0..2] (+ clamp of every final plane)
    int maxval, lo, hi, do_clamp;
    const int16_t *ycc[3];
    int16_t *rout;              // kind 2: a spare W x H plane for R when the epilogue cannot work in place
};
// *epilogue_done = 0: not applied; 1: applied in place; 2: applied with R written to ep->rout (G, B in ycc[1], ycc[2])
int fb_run_inv_squeeze_plan(fb_ctx *ctx, const std::vector<FbSqOp> &ops, const FbSqEpilogue *ep, int *epilogue_done);
// forward Squeeze steps (squeeze.h:135-170, 227-263)
int fb_launch_fwd_hsqueeze(fb_ctx *ctx, const int16_t *in, int16_t *avg, int16_t *res, int w, int h);
int fb_launch_fwd_vsqueeze(fb_ctx *ctx, const int16_t *in, int16_t *avg, int16_t *res, int w, int h);
// YCoCg (ycocg.h:33-63 / 65-95), in place on three planes of n samples each
int fb_launch_ycocg(fb_ctx *ctx, int16_t *c0, int16_t *c1, int16_t *c2, size_t n, int maxval, int inverse, int clamp_lo, int clamp_hi, int do_clamp);
// YCbCr (ycbcr.h:33-63 / 65-95)
int fb_launch_ycbcr(fb_ctx *ctx, int16_t *c0, int16_t *c1, int16_t *c2, size_t n, int minval, int maxval, int inverse);
// v *= q / v /= q with int16 wrap (quantize.h:32-49 / 56-71)
int fb_launch_quantize(fb_ctx *ctx, int16_t *p, size_t n, int q, int inverse);
// final clamp of Image::undo_transforms (image.cpp:107-113)
int fb_launch_clamp(fb_ctx *ctx, int16_t *p, size_t n, int lo, int hi);
// 8x8 inverse DCT of one component: 64 coefficient planes (bw x bh each, planes[k] for coefficient k in the
// reference's block order, nullptr = absent) -> (8bw x 8bh) samples (dct.h:249-296)
int fb_launch_inv_dct(fb_ctx *ctx, const int16_t *const *planes64_dev, int16_t *out, int bw, int bh, float dc_offset);
// dequantise (quantize.h:32-49) + inverse DCT (+ inverse YCbCr, ycbcr.h:49-58, + clamp) of ncomp <= 3 components in one launch;
// planes[c][k] = coefficient k (block order) of component c (nullptr = zero), q[c][k] its quantisation factor, out[c] = 8bw x 8bh samples
int fb_launch_idct_fused(fb_ctx *ctx, const int16_t *const (*planes)[64], const int (*q)[64], int16_t *const *out, int ncomp, int bw, int bh,
                         float dc_offset, int ycbcr, int minval, int maxval);
// forward: samples (w x h, edge replicated) -> 64 planes (dct.h:298-336)
int fb_launch_fwd_dct(fb_ctx *ctx, const int16_t *in, int w, int h, int16_t *const *planes64_dev, int bw, int bh, float dc_offset);
// inv_palette (palette.h:32-68): out_planes[0] holds the indices and receives row 0 of the palette, planes 1..nb-1 the other rows
int fb_launch_palette_inv(fb_ctx *ctx, int16_t *const *out_planes, int nb, const int16_t *palette, int ncolors, size_t n);
// fwd_palette (palette.h:92-143) in two steps: distinct colours (hash set on the device, sorted on the host), then indices
int fb_palette_collect(fb_ctx *ctx, int16_t *const *planes, int nb, size_t n, int limit, std::vector<unsigned long long> &sorted, int *too_many);
int fb_launch_palette_index(fb_ctx *ctx, int16_t *const *planes, int nb, size_t n, const unsigned long long *sorted_dev, int count);
// inv_match (2dmatch.h:97-177, exact matches): roots of all samples by pointer jumping, then one gather per channel
int fb_match_resolve(fb_ctx *ctx, const int16_t *m, int n, int w, int maxcode, int **parent_out, int *bad);
int fb_match_soft(fb_ctx *ctx, const int16_t *m, const int16_t *orig, int16_t *out, int n, int w, int maxcode, int zero, int *bad);
int fb_launch_match_gather(fb_ctx *ctx, const int16_t *src, int16_t *dst, const int *parent, int n, int zero);
// Approximate (approximate.h:32-113): inverse ch = ch*q + chr (chr may be nullptr), forward ch, chr = floor-div / remainder
int fb_launch_approximate(fb_ctx *ctx, int16_t *ch, int16_t *chr, size_t n, int q, int inverse);
// chroma upscaling of one plane (ow x oh -> ow*srh x oh*srv), subsample.h:73-128
int fb_launch_inv_subsample(fb_ctx *ctx, const int16_t *in, int16_t *out, int ow, int oh, int srh, int srv);
// per-plane min / max (Channel::actual_minmax, image.cpp:82-92): out2_dev[0]=min out2_dev[1]=max (int32)
int fb_launch_minmax(fb_ctx *ctx, const int16_t *p, size_t n, int *out2_dev);
// planar int16 -> interleaved 8/16-bit big-endian samples (write_PAM_file layout)
int fb_launch_interleave(fb_ctx *ctx, const int16_t *const *planes_dev, int nch, size_t npix, int bytes_per_sample, void *dst);

// ---- MANIAC decode (fb_maniac.cu) ---------------------------------------------------------------------------
struct FbManiacJob {
    const uint8_t *bytes_host;    // whole file on the host, or (bytes_dev != nullptr) just its first header_len bytes
    const uint8_t *bytes_dev;     // non-null: the file already lives in HBM
    size_t nbytes;
    fb_image *img;            // channel list after meta_apply; planes get allocated + filled, ranges/q filled in
    size_t body_pos;          // byte offset of the first channel group
    size_t bytes_to_load;     // 0 = everything
    int max_properties;
    int cutoff, alpha;
    const int64_t *group_index;   // optional
    const int32_t *group_first;   // optional (required with bytes_dev)
    int n_groups;
};
int fb_maniac_decode(fb_ctx *ctx, std::vector<FbManiacJob> &jobs);
void fb_maniac_release(fb_ctx *ctx);
