// Host side of the fuif_b200 C ABI: contexts, the device-resident Image, the container parser and the
// transform-chain driver (which kernels run in which order on which planes).  No sample is touched on the
// host: every plane lives in HBM and every transform is a kernel from fb_transforms.cu.
#include "fb_common.cuh"

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <exception>
#include <new>

// Nothing may propagate through the C ABI: an exception inside an entry point (a failed host allocation while parsing an
// untrusted file, ...) becomes an error code.
template <class F>
static int abi_guard(fb_ctx *ctx, F &&f) {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        if (ctx) ctx->err = "out of host memory";
        return FB_ERR_NOMEM;
    } catch (const std::exception &e) {
        if (ctx) ctx->err = std::string("internal error: ") + e.what();
        return FB_ERR_INVALID;
    }
}

// ---------------------------------------------------------------------------------------------------------
// context + plane memory
// ---------------------------------------------------------------------------------------------------------

extern "C" int fb_ctx_create(int device, void *stream, fb_ctx **out) {
    if (!out) return FB_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        // fail loudly: there is no CPU fallback in this library
        fprintf(stderr, "fuif_b200: no usable CUDA device %d (found %d)\n", device, ndev);
        return FB_ERR_CUDA;
    }
    fb_ctx *ctx = new (std::nothrow) fb_ctx();
    if (!ctx) return FB_ERR_NOMEM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return FB_ERR_CUDA; }
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return FB_ERR_CUDA; }
        ctx->own_stream = true;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    ctx->timing_stderr = getenv("FB_KERNEL_TIMING") != nullptr;
    ctx->timing = ctx->timing_stderr;
    // FUIF_B200_ENTROPY=host|gpu presets FB_OPT_ENTROPY_BACKEND (fb_ctx_set_option still overrides it)
    if (const char *e = getenv("FUIF_B200_ENTROPY")) ctx->entropy_backend = !strcmp(e, "host") ? FB_ENTROPY_HOST : FB_ENTROPY_GPU;
    // keep freed plane memory in the stream-ordered pool instead of returning it to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = ctx;
    return FB_OK;
}

extern "C" void fb_ctx_destroy(fb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    fb_maniac_release(ctx);
    fb_plane_pool_release(ctx);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->fq_counters) cudaFree(ctx->fq_counters);
    if (ctx->pk_scratch) cudaFree(ctx->pk_scratch);
    if (ctx->pk_counters) cudaFree(ctx->pk_counters);
    if (ctx->pk_stats) cudaFree(ctx->pk_stats);
    if (ctx->host_stage) cudaFreeHost(ctx->host_stage);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *fb_last_error(fb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int fb_ctx_synchronize(fb_ctx *ctx) {
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->timing_stderr && ctx->marks.size() > 1) {
        for (size_t i = 1; i < ctx->marks.size(); i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->marks[i - 1].ev, ctx->marks[i].ev);
            fprintf(stderr, "[timing] %-32s %8.1f us\n", ctx->marks[i].name.c_str(), ms * 1000.f);
        }
        for (auto &mk : ctx->marks) cudaEventDestroy(mk.ev);
        ctx->marks.clear();
    }
    return FB_OK;
}

extern "C" long long fb_ctx_launch_count(fb_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int fb_ctx_set_option(fb_ctx *ctx, int option, int value) {
    if (!ctx) return FB_ERR_INVALID;
    if (option == FB_OPT_SQUEEZE_MODE && value >= 0 && value <= 4) { ctx->fq_mode = value; return FB_OK; }
    if (option == FB_OPT_SQUEEZE_PACKED && value >= 0 && value <= 1) { ctx->pk_mode = value; return FB_OK; }
    if (option == FB_OPT_ENTROPY_BACKEND && (value == FB_ENTROPY_GPU || value == FB_ENTROPY_HOST)) { ctx->entropy_backend = value; return FB_OK; }
    if (option == FB_OPT_HOST_THREADS && value >= 0 && value <= 4096) { ctx->host_threads = value; return FB_OK; }
    if (option == FB_OPT_KERNEL_TIMING) { ctx->timing = value != 0 || ctx->timing_stderr; return FB_OK; }
    return FB_ERR_INVALID;
}

extern "C" long long fb_ctx_counter(fb_ctx *ctx, int which) {
    if (ctx && which == FB_COUNTER_HOST_THREADS) return ctx->host_threads_used;
    if (ctx && (which == FB_COUNTER_PK_REPAIRED || which == FB_COUNTER_PK_RANGE_FLAGGED)) {
        if (!ctx->pk_stats) return 0;
        int v[2] = {0, 0};
        cudaSetDevice(ctx->device);
        if (cudaMemcpyAsync(v, ctx->pk_stats, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
        return v[which - FB_COUNTER_PK_REPAIRED];
    }
    if (!ctx || which < 0 || which > 1) return -1;
    if (!ctx->fq_counters) return 0;
    int v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaSetDevice(ctx->device);
    if (cudaMemcpyAsync(v, ctx->fq_counters, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    return v[4 + which];
}

extern "C" long long fb_ctx_timing_report(fb_ctx *ctx, char *buf, size_t cap) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -1;
    std::string out;
    for (size_t i = 1; i < ctx->marks.size(); i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->marks[i - 1].ev, ctx->marks[i].ev);
        char line[256];
        snprintf(line, sizeof(line), "%s\t%.3f\t%.0f\n", ctx->marks[i].name.c_str(), ms * 1000.0, ctx->marks[i].bytes);
        out += line;
    }
    for (auto &mk : ctx->marks) cudaEventDestroy(mk.ev);
    ctx->marks.clear();
    if (buf && cap) {
        const size_t n = std::min(cap - 1, out.size());
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return (long long)out.size();
}

int fb_plane_alloc(fb_ctx *ctx, size_t nsamples, int16_t **out) {
    *out = nullptr;
    size_t bytes = std::max<size_t>(nsamples * sizeof(int16_t), 16);
    bytes = (bytes + 255) & ~(size_t)255;
    auto it = ctx->plane_pool.find(bytes);
    if (it != ctx->plane_pool.end() && !it->second.empty()) {
        *out = (int16_t *)it->second.back();
        it->second.pop_back();
        ctx->plane_pool_bytes -= bytes;
        return FB_OK;
    }
    void *p = nullptr;
    FB_CUDA(ctx, cudaMallocAsync(&p, bytes, ctx->stream));
    ctx->plane_sizes[p] = bytes;
    *out = (int16_t *)p;
    return FB_OK;
}

void fb_plane_free(fb_ctx *ctx, int16_t *p) {
    if (!p) return;
    auto it = ctx->plane_sizes.find(p);
    if (it == ctx->plane_sizes.end() || ctx->plane_pool_bytes > ((size_t)16 << 30)) {
        if (it != ctx->plane_sizes.end()) ctx->plane_sizes.erase(it);
        cudaFreeAsync(p, ctx->stream);
        return;
    }
    ctx->plane_pool[it->second].push_back(p);
    ctx->plane_pool_bytes += it->second;
}
void fb_plane_pool_release(fb_ctx *ctx) {
    for (auto &kv : ctx->plane_pool)
        for (void *q : kv.second) { ctx->plane_sizes.erase(q); cudaFreeAsync(q, ctx->stream); }
    ctx->plane_pool.clear();
    ctx->plane_pool_bytes = 0;
}

// ---------------------------------------------------------------------------------------------------------
// Image helpers
// ---------------------------------------------------------------------------------------------------------

static void chan_defaults(fb_plane_desc &d) {       // Channel::Channel(), reference image/image.h:68
    memset(&d, 0, sizeof(d));
    d.q = 1;
    d.component = -1;
}

static inline int s16(int x) { return (int)(int16_t)x; }

static void chan_setzero(fb_plane_desc &d) {        // Channel::setzero, image/image.h:70-74
    if (d.minval > 0) d.zero = d.minval;
    else if (d.maxval < 0) d.zero = d.maxval;
    else d.zero = 0;
}

static size_t chan_samples(const fb_plane_desc &d) { return (d.w > 0 && d.h > 0) ? (size_t)d.w * d.h : 0; }

// Channel::resize() on an undecoded plane: w*h samples of `zero` (image/image.h:75-77)
static int chan_materialize(fb_ctx *ctx, FbChan &c) {
    if (c.dev) return FB_OK;
    size_t n = chan_samples(c.d);
    int rc = fb_plane_alloc(ctx, n, &c.dev);
    if (rc) return rc;
    if (n) {
        if (c.d.zero == 0) FB_CUDA(ctx, cudaMemsetAsync(c.dev, 0, n * sizeof(int16_t), ctx->stream));
        else {
            // memset16 through the clamp kernel: fill with zero then clamp to [zero, zero]
            FB_CUDA(ctx, cudaMemsetAsync(c.dev, 0, n * sizeof(int16_t), ctx->stream));
            rc = fb_launch_clamp(ctx, c.dev, n, c.d.zero, c.d.zero);
            if (rc) return rc;
        }
    }
    c.d.decoded = 1;
    return FB_OK;
}

extern "C" void fb_image_destroy(fb_image *img) {
    if (!img) return;
    if (img->ctx->device >= 0) {
        cudaSetDevice(img->ctx->device);
        for (auto &c : img->ch) fb_plane_free(img->ctx, c.dev);
    }
    if (img->owns_ctx) delete img->ctx;     // the CUDA-free context of a host-only image (fb_host_decode)
    delete img;
}

// entry points that compute on the planes need them in HBM
#define FB_NEEDS_DEVICE(img)                                                                                              \
    do {                                                                                                                  \
        if ((img) && (img)->on_host) { (img)->ctx->err = "this image lives in host memory (fb_host_decode): fb_image_upload() it first"; return FB_ERR_INVALID; } \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// Squeeze bookkeeping (reference transform/squeeze.h:266-408)
// ---------------------------------------------------------------------------------------------------------

// default_squeeze_parameters, squeeze.h:266-321 (MAX_FIRST_PREVIEW_SIZE = 8, config.h:41)
static void default_squeeze_parameters(std::vector<int> &p, const fb_image *img) {
    p.clear();
    const int nb = img->info.nb_channels, m = img->info.nb_meta_channels;
    int w = img->ch[m].d.w, h = img->ch[m].d.h;
    const bool wide = w > h;
    if (nb > 2 && img->ch[m + 1].d.w == w && img->ch[m + 1].d.h == h) {
        p.insert(p.end(), {3, m + 1, m + 2});       // horizontal chroma squeeze, residuals appended at the end
        p.insert(p.end(), {2, m + 1, m + 2});       // vertical chroma squeeze
    }
    if (!wide && h > 8) { p.insert(p.end(), {0, m, m + nb - 1}); h = (h + 1) / 2; }
    while (w > 8 || h > 8) {
        if (w > 8) { p.insert(p.end(), {1, m, m + nb - 1}); w = (w + 1) / 2; }
        if (h > 8) { p.insert(p.end(), {0, m, m + nb - 1}); h = (h + 1) / 2; }
    }
}

// meta_squeeze, squeeze.h:323-360: channel-list surgery only
static int meta_squeeze(fb_image *img, std::vector<int> &p) {
    if (p.empty()) default_squeeze_parameters(p, img);
    for (size_t i = 0; i + 2 < p.size(); i += 3) {
        const bool horizontal = p[i] & 1, in_place = !(p[i] & 2);
        const int beginc = p[i + 1], endc = p[i + 2];
        const int offset = in_place ? endc + 1 : img->info.nb_meta_channels + img->info.nb_channels;
        if (beginc < 0 || endc < beginc || endc >= (int)img->ch.size() || offset > (int)img->ch.size()) return FB_ERR_INVALID;
        for (int c = beginc; c <= endc; c++) {
            FbChan dummy;
            chan_defaults(dummy.d);
            fb_plane_desc &s = img->ch[c].d;
            dummy.d.hcshift = s.hcshift; dummy.d.vcshift = s.vcshift; dummy.d.component = s.component;
            if (horizontal) {
                int w = s.w;
                s.w = (w + 1) / 2; s.hshift++; s.hcshift++;
                dummy.d.w = w - (w + 1) / 2; dummy.d.h = s.h;
            } else {
                int h = s.h;
                s.h = (h + 1) / 2; s.vshift++; s.vcshift++;
                dummy.d.h = h - (h + 1) / 2; dummy.d.w = s.w;
            }
            dummy.d.hshift = s.hshift; dummy.d.vshift = s.vshift;
            img->ch.insert(img->ch.begin() + offset + c - beginc, dummy);
        }
    }
    return FB_OK;
}

// squeeze(..., inverse=true), squeeze.h:367-388 with inv_hsqueeze/inv_vsqueeze (:81-132, :173-224) as kernels
// ep_kind: what undo_transforms will do right after this Squeeze (0 nothing fusable, 1 final clamp, 2 inverse YCoCg,
// do_clamp: followed by the final clamp); *ep_done = 1 if the unsqueeze kernels already did it.
static int inv_squeeze(fb_image *img, const std::vector<int> &params, int ep_kind, int do_clamp, int *ep_done) {
    fb_ctx *ctx = img->ctx;
    std::vector<int> p = params;
    if (p.empty()) default_squeeze_parameters(p, img);
    // Pass 1: channel-list surgery and allocation, exactly in the reference's order; the kernels are only planned.
    std::vector<FbSqOp> ops;
    std::vector<int16_t *> to_free;
    int step = 0;
    for (int i = (int)p.size() - 3; i >= 0; i -= 3, step++) {
        const bool horizontal = p[i] & 1, in_place = !(p[i] & 2);
        const int beginc = p[i + 1], endc = p[i + 2];
        const int offset = in_place ? endc + 1 : img->info.nb_meta_channels + img->info.nb_channels;
        if (beginc < 0 || endc < beginc || offset + endc - beginc >= (int)img->ch.size()) {
            ctx->err = "Invalid parameters for squeeze transform";
            for (auto q : to_free) fb_plane_free(ctx, q);
            return FB_ERR_INVALID;
        }
        for (int c = beginc; c <= endc; c++) {
            FbChan &a = img->ch[c];
            FbChan &r = img->ch[offset + c - beginc];
            // the averages must exist; a missing residual plane acts as zeros (squeeze.h:379-383)
            int rc = chan_materialize(ctx, a);
            if (rc) return rc;
            FbChan out;
            out.d = a.d;
            if (horizontal) { out.d.w = a.d.w + r.d.w; out.d.hshift--; out.d.hcshift--; }
            else { out.d.h = a.d.h + r.d.h; out.d.vshift--; out.d.vcshift--; }
            chan_setzero(out.d);
            out.d.decoded = 1;
            rc = fb_plane_alloc(ctx, chan_samples(out.d), &out.dev);
            if (rc) return rc;
            ops.push_back(FbSqOp{step, horizontal ? 1 : 0, a.dev, r.dev, out.dev, a.d.w, r.d.w, a.d.h, r.d.h});
            to_free.push_back(a.dev);
            a = out;
        }
        for (int c = 0; c <= endc - beginc; c++) to_free.push_back(img->ch[offset + c].dev);
        img->ch.erase(img->ch.begin() + offset, img->ch.begin() + offset + (endc - beginc + 1));
    }
    // Pass 2: run the plan (fused tile kernels; per-level kernels for shapes the planner refuses).
    FbSqEpilogue ep{};
    const int m = img->info.nb_meta_channels;
    if (ep_kind) {
        // every plane of the image must come out of this plan, otherwise the epilogue stays with the caller
        bool all = true;
        for (auto &c : img->ch) {
            bool produced = false;
            for (auto &o : ops) if (o.out == c.dev) produced = true;
            if (!produced) all = false;
        }
        if (ep_kind == 2) {
            bool okc = img->info.nb_channels >= 3 && (int)img->ch.size() >= m + 3;
            if (okc) for (int k = 1; k < 3; k++) if (img->ch[m + k].d.w != img->ch[m].d.w || img->ch[m + k].d.h != img->ch[m].d.h) okc = false;
            if (okc) {
                ep.kind = 2;
                for (int k = 0; k < 3; k++) ep.ycc[k] = img->ch[m + k].dev;
                ep.do_clamp = (do_clamp && all) ? 1 : 0;
                if (do_clamp && !all) ep.kind = 0;     // keep the order "YCoCg, then clamp of everything" simple: do not fuse
            }
        } else if (all) {
            ep.kind = 1;
            ep.do_clamp = 1;
        }
        ep.maxval = img->info.maxval; ep.lo = img->info.minval; ep.hi = img->info.maxval;
    }
    if (ep.kind == 2) {
        int rc0 = fb_plane_alloc(ctx, chan_samples(img->ch[m].d), &ep.rout);
        if (rc0) return rc0;
    }
    int done = 0;
    ctx->sq_maxval = img->info.maxval;
    int rc = fb_run_inv_squeeze_plan(ctx, ops, ep.kind ? &ep : nullptr, &done);
    if (ep.rout) {
        if (!rc && done == 2) { to_free.push_back(img->ch[m].dev); img->ch[m].dev = ep.rout; }
        else to_free.push_back(ep.rout);
    }
    if (ep_done) *ep_done = done ? 1 : 0;
    // Pass 3: the consumed planes go back to the stream-ordered pool (after the kernels in stream order).
    for (auto q : to_free) fb_plane_free(ctx, q);
    return rc;
}

// squeeze(..., inverse=false), squeeze.h:389-406 with fwd_hsqueeze/fwd_vsqueeze (:135-170, :227-263)
static int fwd_squeeze(fb_image *img, const std::vector<int> &params) {
    fb_ctx *ctx = img->ctx;
    std::vector<int> p = params;
    if (p.empty()) default_squeeze_parameters(p, img);
    for (size_t i = 0; i + 2 < p.size(); i += 3) {
        const bool horizontal = p[i] & 1, in_place = !(p[i] & 2);
        const int beginc = p[i + 1], endc = p[i + 2];
        const int offset = in_place ? endc + 1 : img->info.nb_meta_channels + img->info.nb_channels;
        if (beginc < 0 || endc < beginc || endc >= (int)img->ch.size() || offset > (int)img->ch.size()) return FB_ERR_INVALID;
        for (int c = beginc; c <= endc; c++) {
            FbChan &in = img->ch[c];
            int rc = chan_materialize(ctx, in);
            if (rc) return rc;
            FbChan avg, res;
            avg.d = in.d;
            chan_defaults(res.d);
            if (horizontal) {
                avg.d.w = (in.d.w + 1) / 2; avg.d.hshift++; avg.d.hcshift++;
                res.d.w = in.d.w - avg.d.w; res.d.h = in.d.h;
                res.d.hshift = in.d.hshift + 1; res.d.vshift = in.d.vshift;
            } else {
                avg.d.h = (in.d.h + 1) / 2; avg.d.vshift++; avg.d.vcshift++;
                res.d.w = in.d.w; res.d.h = in.d.h - avg.d.h;
                res.d.hshift = in.d.hshift; res.d.vshift = in.d.vshift + 1;
            }
            res.d.hcshift = in.d.hcshift; res.d.vcshift = in.d.vcshift;
            res.d.minval = s16(avg.d.minval - avg.d.maxval); res.d.maxval = s16(avg.d.maxval - avg.d.minval);
            res.d.q = 1; res.d.component = in.d.component;
            chan_setzero(res.d); chan_setzero(avg.d);
            avg.d.decoded = res.d.decoded = 1;
            if ((rc = fb_plane_alloc(ctx, chan_samples(avg.d), &avg.dev))) return rc;
            if ((rc = fb_plane_alloc(ctx, chan_samples(res.d), &res.dev))) return rc;
            if (horizontal) rc = fb_launch_fwd_hsqueeze(ctx, in.dev, avg.dev, res.dev, in.d.w, in.d.h);
            else rc = fb_launch_fwd_vsqueeze(ctx, in.dev, avg.dev, res.dev, in.d.w, in.d.h);
            if (rc) return rc;
            fb_plane_free(ctx, in.dev);
            in = avg;
            img->ch.insert(img->ch.begin() + offset + c - beginc, res);
        }
    }
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// DCT bookkeeping (reference transform/dct.h:120-336)
// ---------------------------------------------------------------------------------------------------------

// The reference's coefficient scan (dct.h:120-130, "we use a variant"): index -> position in the 8x8 block
// laid out in L-shaped shells max(r,c)=s that hold scan indices s*s .. s*s+2s; even shells run down column
// s then left along row s, odd shells run right along row s then up column s; shell 1 is irregular.
static const int *scan_of_block_index() {
    static int zz[64];
    static bool ready = false;
    if (!ready) {
        zz[0] = 0; zz[1] = 1; zz[8] = 2; zz[9] = 3;
        for (int s = 2; s < 8; s++) {
            int idx = s * s;
            if (s % 2 == 0) {
                for (int r = 0; r <= s; r++) zz[r * 8 + s] = idx++;
                for (int c = s - 1; c >= 0; c--) zz[s * 8 + c] = idx++;
            } else {
                for (int c = 0; c <= s; c++) zz[s * 8 + c] = idx++;
                for (int r = s - 1; r >= 0; r--) zz[r * 8 + s] = idx++;
            }
        }
        ready = true;
    }
    return zz;
}

static int dct_cshift(int k) { return k == 0 ? 3 : (k < 4 ? 2 : (k < 16 ? 1 : 0)); }       // dct_cshifts, dct.h:159-171

static void default_dct_parameters(std::vector<int> &p, const fb_image *img) {              // dct.h:209-213
    p.clear();
    p.push_back(0);
    p.push_back(img->info.nb_channels - 1);
}

// meta_DCT, dct.h:215-246.  default_DCT_scanscript (:173-207) puts coefficient k of component c at position k*nb+c.
static int meta_dct(fb_image *img, std::vector<int> &p) {
    if (p.size() < 2) default_dct_parameters(p, img);
    const int beginc = img->info.nb_meta_channels + p[0], endc = img->info.nb_meta_channels + p[1];
    const int nb = endc - beginc + 1;
    if (nb < 1 || beginc < 0 || endc >= (int)img->ch.size()) return FB_ERR_INVALID;
    for (int c = beginc; c <= endc; c++) {
        fb_plane_desc &d = img->ch[c].d;
        d.w = (d.w + 7) / 8; d.h = (d.h + 7) / 8;
        d.hshift += 3; d.vshift += 3; d.hcshift += 3; d.vcshift += 3;
    }
    for (int i = nb; i < 64 * nb; i++) {
        FbChan dummy;
        chan_defaults(dummy.d);
        const fb_plane_desc &s = img->ch[beginc + i % nb].d;
        const int coeff = i / nb;
        dummy.d.w = s.w; dummy.d.h = s.h; dummy.d.hshift = s.hshift; dummy.d.vshift = s.vshift;
        dummy.d.hcshift = dct_cshift(coeff) + s.hcshift - 3;
        dummy.d.vcshift = dct_cshift(coeff) + s.vcshift - 3;
        dummy.d.component = s.component;
        img->ch.push_back(dummy);
    }
    return FB_OK;
}

// inv_DCT, dct.h:249-296
static int inv_dct(fb_image *img, std::vector<int> &p) {
    fb_ctx *ctx = img->ctx;
    if (p.size() < 2) default_dct_parameters(p, img);
    const int beginc = img->info.nb_meta_channels + p[0], endc = img->info.nb_meta_channels + p[1];
    const int nb = endc - beginc + 1;
    const int offset = (int)img->ch.size() - 63 * nb;
    if (nb < 1 || beginc < 0 || offset <= endc) { ctx->err = "Invalid number of channels to apply inverse DCT."; return FB_ERR_INVALID; }
    const int *zz = scan_of_block_index();
    const float dc_offset = (float)((img->info.maxval + 1.0) * 4.0);
    for (int c = beginc; c <= endc; c++) {
        int bw = img->ch[c - beginc + offset].d.w, bh = img->ch[c - beginc + offset].d.h;
        if (img->ch[c].d.w < bw) bw = img->ch[c].d.w;
        if (img->ch[c].d.h < bh) bh = img->ch[c].d.h;
        const int16_t *planes[64];
        planes[0] = img->ch[c].dev;
        bool ok = img->ch[c].dev == nullptr || (img->ch[c].d.w == bw && img->ch[c].d.h == bh);
        for (int i = 1; i < 64; i++) {
            const FbChan &s = img->ch[offset - nb + zz[i] * nb + (c - beginc)];
            planes[i] = s.dev;
            if (s.dev && (s.d.w != bw || s.d.h != bh)) ok = false;
        }
        if (!ok) { ctx->err = "inverse DCT: coefficient planes of unequal size are not supported"; return FB_ERR_UNSUPPORTED; }
        FbChan out;
        chan_defaults(out.d);
        out.d.w = bw * 8; out.d.h = bh * 8;
        out.d.component = img->ch[c].d.component;
        out.d.hshift = img->ch[c].d.hshift - 3; out.d.vshift = img->ch[c].d.vshift - 3;
        out.d.hcshift = img->ch[c].d.hcshift - 3; out.d.vcshift = img->ch[c].d.hcshift - 3;      // sic, dct.h:280
        out.d.decoded = 1;
        int rc = fb_plane_alloc(ctx, chan_samples(out.d), &out.dev);
        if (rc) return rc;
        if ((rc = fb_launch_inv_dct(ctx, planes, out.dev, bw, bh, dc_offset))) return rc;
        fb_plane_free(ctx, img->ch[c].dev);
        img->ch[c] = out;
    }
    for (int c = offset; c < offset + nb * 63; c++) fb_plane_free(ctx, img->ch[c].dev);
    img->ch.erase(img->ch.begin() + offset, img->ch.begin() + offset + nb * 63);
    return FB_OK;
}

// Quantize -> DCT (-> YCbCr) undone together: the tail of a JPEG-transcode chain in one launch (fb_idct_fused.cuh) instead of
// one launch per coefficient plane, one per component and one for the colour inverse.  *fused = how many transforms were
// undone (0: the shape is not one the fused kernel takes, the caller goes on transform by transform).
static int try_fused_dct_tail(fb_image *img, int keep, int *fused, bool *clamped) {
    fb_ctx *ctx = img->ctx;
    *fused = 0;
    static const bool off = getenv("FB_DCT_FUSED") && atoi(getenv("FB_DCT_FUSED")) == 0;
    const int nt = (int)img->tr.size();
    if (off || nt < 2 || nt - 2 < keep || img->tr[nt - 1].id != FB_TRANSFORM_QUANTIZE || img->tr[nt - 2].id != FB_TRANSFORM_DCT) return FB_OK;
    if (img->info.nb_meta_channels != 0) return FB_OK;
    std::vector<int> p = img->tr[nt - 2].p;
    if (p.size() < 2) default_dct_parameters(p, img);
    const int beginc = p[0], endc = p[1], nb = endc - beginc + 1;
    const int offset = (int)img->ch.size() - 63 * nb;
    if (nb < 1 || nb > 3 || beginc < 0 || offset <= endc) return FB_OK;
    const int *zz = scan_of_block_index();
    const int16_t *planes[3][64];
    int q[3][64];
    int bw = 0, bh = 0;
    for (int c = 0; c < nb; c++)
        for (int i = 0; i < 64; i++) {
            const FbChan &s = i == 0 ? img->ch[beginc + c] : img->ch[offset - nb + zz[i] * nb + c];
            planes[c][i] = s.dev;
            q[c][i] = s.d.q;
            if (c == 0 && i == 0) { bw = s.d.w; bh = s.d.h; }
            if (s.d.w != bw || s.d.h != bh) return FB_OK;       // planes of unequal size: the generic path reports it
        }
    if (bw < 1 || bh < 1 || bh > 65535) return FB_OK;
    const bool with_ycbcr = nt >= 3 && nt - 3 >= keep && img->tr[nt - 3].id == FB_TRANSFORM_YCBCR && nb == 3 && beginc == 0;
    // channels outside the DCT set are only dequantised (inv_quantize touches every non-meta channel)
    for (int c = 0; c < offset; c++) {
        if (c >= beginc && c <= endc) continue;
        FbChan &ch = img->ch[c];
        if (!ch.dev || ch.d.q == 1) continue;
        const int qq = ch.d.q;
        int rc = fb_launch_quantize(ctx, ch.dev, chan_samples(ch.d), qq, 1);
        if (rc) return rc;
        ch.d.minval = s16(ch.d.minval * qq); ch.d.maxval = s16(ch.d.maxval * qq); ch.d.q = 1;
    }
    FbChan out[3];
    int16_t *outp[3] = {nullptr, nullptr, nullptr};
    for (int c = 0; c < nb; c++) {
        const FbChan &dc = img->ch[beginc + c];
        chan_defaults(out[c].d);
        out[c].d.w = bw * 8; out[c].d.h = bh * 8;
        out[c].d.component = dc.d.component;
        out[c].d.hshift = dc.d.hshift - 3; out[c].d.vshift = dc.d.vshift - 3;
        out[c].d.hcshift = dc.d.hcshift - 3; out[c].d.vcshift = dc.d.hcshift - 3;      // sic, dct.h:280
        out[c].d.decoded = 1;
        int rc = fb_plane_alloc(ctx, chan_samples(out[c].d), &out[c].dev);
        if (rc) return rc;
        outp[c] = out[c].dev;
    }
    const float dc_offset = (float)((img->info.maxval + 1.0) * 4.0);
    int rc = fb_launch_idct_fused(ctx, planes, q, outp, nb, bw, bh, dc_offset, with_ycbcr ? 1 : 0, img->info.minval, img->info.maxval);
    if (rc) return rc;
    for (int c = 0; c < nb; c++) { fb_plane_free(ctx, img->ch[beginc + c].dev); img->ch[beginc + c] = out[c]; }
    for (int c = offset; c < offset + nb * 63; c++) fb_plane_free(ctx, img->ch[c].dev);
    img->ch.erase(img->ch.begin() + offset, img->ch.begin() + offset + nb * 63);
    *fused = with_ycbcr ? 3 : 2;
    // after the YCbCr clamp the final clamp of undo_transforms is the identity for these planes
    if (with_ycbcr && nt == 3 && keep == 0 && (int)img->ch.size() == 3) *clamped = true;
    return FB_OK;
}

// fwd_DCT, dct.h:298-336 (explicit parameters only: the reference dereferences an empty vector otherwise)
static int fwd_dct(fb_image *img, std::vector<int> &p, int *applied) {
    fb_ctx *ctx = img->ctx;
    *applied = 0;
    if (p.size() < 2) { ctx->err = "forward DCT needs explicit parameters"; return FB_ERR_INVALID; }
    const int beginc = img->info.nb_meta_channels + p[0], endc = img->info.nb_meta_channels + p[1];
    const int nb = endc - beginc + 1;
    if (nb < 1 || beginc < 0 || endc >= (int)img->ch.size()) return FB_ERR_INVALID;
    std::vector<FbChan> src(img->ch.begin() + beginc, img->ch.begin() + endc + 1);     // the reference's "Image tmp = input"
    for (auto &s : src) { int rc = chan_materialize(ctx, s); if (rc) return rc; }
    for (int c = beginc; c <= endc; c++) img->ch[c].dev = src[c - beginc].dev;
    const int offset = (int)img->ch.size();
    int rc = meta_dct(img, p);
    if (rc) return rc;
    const int *zz = scan_of_block_index();
    const float dc_offset = (float)((img->info.maxval + 1.0) * 4.0);
    // channels beginc .. end get fresh sample buffers (dct.h:316-318); channels between endc and offset are
    // resized in place by the reference, i.e. they keep their samples
    for (int c = beginc; c < offset + 63 * nb; c++) {
        if (c > endc && c < offset) { if ((rc = chan_materialize(ctx, img->ch[c]))) return rc; continue; }
        img->ch[c].dev = nullptr;
        if ((rc = fb_plane_alloc(ctx, chan_samples(img->ch[c].d), &img->ch[c].dev))) return rc;
        img->ch[c].d.decoded = 1;
    }
    for (int c = beginc; c <= endc; c++) {
        int16_t *planes[64];
        planes[0] = img->ch[c].dev;
        for (int i = 1; i < 64; i++) planes[i] = img->ch[offset - nb + zz[i] * nb + (c - beginc)].dev;
        const FbChan &s = src[c - beginc];
        if ((rc = fb_launch_fwd_dct(ctx, s.dev, s.d.w, s.d.h, planes, img->ch[c].d.w, img->ch[c].d.h, dc_offset))) return rc;
    }
    for (auto &s : src) fb_plane_free(ctx, s.dev);
    *applied = 1;
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// colour transforms / quantisation drivers
// ---------------------------------------------------------------------------------------------------------

// YCoCg(input, inverse), ycocg.h:33-101.  fuse_clamp: fold the final clamp of undo_transforms into the kernel.
static int do_ycocg(fb_image *img, bool inverse, bool fuse_clamp, int *applied) {
    fb_ctx *ctx = img->ctx;
    *applied = 0;
    const int m = img->info.nb_meta_channels;
    if (img->info.nb_channels < 3 || (int)img->ch.size() < m + 3) {
        if (inverse) { ctx->err = "Invalid number of channels to apply inverse YCoCg."; return FB_ERR_INVALID; }
        return FB_OK;       // forward: "not applied" (ycocg.h:67-70)
    }
    const int w = img->ch[m].d.w, h = img->ch[m].d.h;
    for (int k = 1; k < 3; k++)
        if (img->ch[m + k].d.w < w || img->ch[m + k].d.h < h) { ctx->err = "Invalid channel dimensions to apply YCoCg"; return FB_ERR_INVALID; }
    for (int k = 1; k < 3; k++)
        if (img->ch[m + k].d.w != w) { ctx->err = "YCoCg on planes of different width is not supported"; return FB_ERR_UNSUPPORTED; }
    for (int k = 0; k < 3; k++) { int rc = chan_materialize(ctx, img->ch[m + k]); if (rc) return rc; }
    int rc = fb_launch_ycocg(ctx, img->ch[m].dev, img->ch[m + 1].dev, img->ch[m + 2].dev, (size_t)w * h, img->info.maxval, inverse ? 1 : 0,
                             img->info.minval, img->info.maxval, fuse_clamp ? 1 : 0);
    if (rc) return rc;
    *applied = 1;
    return FB_OK;
}

// YCbCr(input, inverse), ycbcr.h:33-101: always channels 0..2
static int do_ycbcr(fb_image *img, bool inverse, int *applied) {
    fb_ctx *ctx = img->ctx;
    *applied = 0;
    if (img->ch.size() < 3) { ctx->err = "Invalid number of channels to apply YCbCr."; return FB_ERR_INVALID; }
    const int w = img->ch[0].d.w, h = img->ch[0].d.h;
    for (int k = 1; k < 3; k++)
        if (img->ch[k].d.w < w || img->ch[k].d.h < h) { ctx->err = "Invalid channel dimensions to apply YCbCr"; return FB_ERR_INVALID; }
    for (int k = 1; k < 3; k++)
        if (img->ch[k].d.w != w) { ctx->err = "YCbCr on planes of different width is not supported"; return FB_ERR_UNSUPPORTED; }
    for (int k = 0; k < 3; k++) { int rc = chan_materialize(ctx, img->ch[k]); if (rc) return rc; }
    int rc = fb_launch_ycbcr(ctx, img->ch[0].dev, img->ch[1].dev, img->ch[2].dev, (size_t)w * h, img->info.minval, img->info.maxval, inverse ? 1 : 0);
    if (rc) return rc;
    *applied = 1;
    return FB_OK;
}

// inv_quantize / fwd_quantize, quantize.h:32-49 / 56-71
static int do_quantize(fb_image *img, bool inverse, const std::vector<int> &p) {
    fb_ctx *ctx = img->ctx;
    for (int c = img->info.nb_meta_channels; c < (int)img->ch.size(); c++) {
        FbChan &ch = img->ch[c];
        if (inverse) {
            if (!ch.dev) continue;
            const int q = ch.d.q;
            if (q == 1) continue;
            int rc = fb_launch_quantize(ctx, ch.dev, chan_samples(ch.d), q, 1);
            if (rc) return rc;
            ch.d.minval = s16(ch.d.minval * q); ch.d.maxval = s16(ch.d.maxval * q); ch.d.q = 1;
        } else {
            if (p.empty()) return FB_ERR_INVALID;
            const int q = c < (int)p.size() ? p[c] : p.back();
            if (q == 0) return FB_ERR_INVALID;
            int rc = chan_materialize(ctx, ch);
            if (rc) return rc;
            if ((rc = fb_launch_quantize(ctx, ch.dev, chan_samples(ch.d), q, 0))) return rc;
            ch.d.minval = s16(ch.d.minval / q); ch.d.maxval = s16(ch.d.maxval / q); ch.d.q = q;
        }
    }
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// ChromaSubsample (reference transform/subsample.h): inverse + meta; the reference has no forward (:130-133)
// ---------------------------------------------------------------------------------------------------------

// check_subsample_parameters, subsample.h:33-69: one abbreviated parameter = 4:2:0 / 4:2:2 / 4:4:0 / 4:1:1; a list that is not
// a multiple of four is "invalid" and cleared (the transform then does nothing)
static std::vector<int> subsample_parameters(const std::vector<int> &p) {
    if (p.size() == 1 && p[0] >= 0 && p[0] <= 3) {
        static const int ab[4][2] = {{2, 2}, {2, 1}, {1, 2}, {4, 1}};
        return {1, 2, ab[p[0]][0], ab[p[0]][1]};
    }
    if (p.size() % 4) return {};
    return p;
}

// inv_subsample, subsample.h:73-128.  The upscaled plane is a fresh Channel(w*srh, h*srv, minval, maxval): shifts, q and
// component go back to their defaults, and its size can exceed the image's (odd dimensions), exactly as in the reference.
static int inv_subsample(fb_image *img, const std::vector<int> &params) {
    fb_ctx *ctx = img->ctx;
    const std::vector<int> p = subsample_parameters(params);
    const int nch = (int)img->ch.size();
    for (size_t i = 0; i + 3 < p.size(); i += 4) {
        const int c1 = p[i], c2 = p[i + 1], srh = p[i + 2], srv = p[i + 3];
        if (c1 < 0 || c2 >= nch || srh < 1 || srv < 1 || srh > 64 || srv > 64) { ctx->err = "inv_subsample: bad parameters"; return FB_ERR_INVALID; }
        for (int c = c1; c <= c2; c++) {
            FbChan &in = img->ch[c];
            const FbChan &ref = img->ch[img->info.nb_meta_channels];
            if (in.d.w >= ref.d.w && in.d.h >= ref.d.h) continue;       // LQIP / 1:16 decodes: already as large as the first channel
            if ((long long)in.d.w * srh * (long long)in.d.h * srv > 0x7fffffffLL) { ctx->err = "inv_subsample: plane too large"; return FB_ERR_UNSUPPORTED; }
            int rc = chan_materialize(ctx, in);         // reads of an undecoded plane give `zero` (image.h:82)
            if (rc) return rc;
            FbChan out;
            chan_defaults(out.d);
            out.d.w = in.d.w * srh; out.d.h = in.d.h * srv; out.d.minval = in.d.minval; out.d.maxval = in.d.maxval;
            chan_setzero(out.d);
            out.d.decoded = 1;
            rc = fb_plane_alloc(ctx, chan_samples(out.d), &out.dev);
            if (rc) return rc;
            rc = fb_launch_inv_subsample(ctx, in.dev, out.dev, in.d.w, in.d.h, srh, srv);
            if (rc) { fb_plane_free(ctx, out.dev); return rc; }
            fb_plane_free(ctx, in.dev);         // stream-ordered: released after the kernel above
            in = out;
        }
    }
    return FB_OK;
}

// meta_subsample, subsample.h:135-157 (the reference asserts ratios of 1 or 2 here)
static int meta_subsample(fb_image *img, const std::vector<int> &params) {
    const std::vector<int> p = subsample_parameters(params);
    const int nch = (int)img->ch.size();
    for (size_t i = 0; i + 3 < p.size(); i += 4) {
        const int c1 = p[i], c2 = p[i + 1], srh = p[i + 2], srv = p[i + 3];
        if ((srh != 1 && srh != 2) || (srv != 1 && srv != 2) || c1 < 0 || c2 >= nch) { img->ctx->err = "meta_subsample: bad parameters"; return FB_ERR_INVALID; }
        for (int c = c1; c <= c2; c++) {
            fb_plane_desc &d = img->ch[c].d;
            d.w = (d.w + srh - 1) / srh; d.h = (d.h + srv - 1) / srv;
            d.hshift += srh == 1 ? 0 : 1; d.vshift += srv == 1 ? 0 : 1;
        }
    }
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Palette (reference transform/palette.h): parameters = first channel, last channel (relative to the meta channels), colours.
// Forward (up to four channels), inverse and decode-time meta step.
// ---------------------------------------------------------------------------------------------------------

static int pl_unpack(unsigned long long k, int c) { return (int)((k >> (16 * (3 - c))) & 0xffffu) - 32768; }     // pl::unpack_colour, fb_palette.cuh

// meta_palette, palette.h:70-89: channels first+1..last disappear, a palette meta-channel (colours x channels, hshift -1) leads the list
static int meta_palette(fb_image *img, const std::vector<int> &p) {
    if (p.size() != 3) { img->ctx->err = "Palette: incorrect parameters"; return FB_ERR_INVALID; }
    const int begin_c = img->info.nb_meta_channels + p[0], end_c = img->info.nb_meta_channels + p[1];
    if (p[0] < 0 || begin_c > end_c || end_c >= (int)img->ch.size() || p[2] < 0) { img->ctx->err = "Palette: incorrect parameters"; return FB_ERR_INVALID; }
    const int nb = end_c - begin_c + 1;
    img->info.nb_meta_channels++;
    img->info.nb_channels -= nb - 1;
    for (int c = begin_c + 1; c <= end_c; c++) if (img->ch[c].dev) fb_plane_free(img->ctx, img->ch[c].dev);
    img->ch.erase(img->ch.begin() + begin_c + 1, img->ch.begin() + end_c + 1);
    FbChan pch;
    chan_defaults(pch.d);
    pch.d.w = p[2]; pch.d.h = nb; pch.d.minval = 0; pch.d.maxval = 1; pch.d.hshift = -1;
    chan_setzero(pch.d);
    img->ch.insert(img->ch.begin(), pch);
    return FB_OK;
}

// fwd_palette, palette.h:92-143.  *applied = 0 (and nothing changes) when the channels use more than p[2] colours; otherwise
// p[2] becomes the number of colours found, as in the reference.
static int fwd_palette(fb_image *img, std::vector<int> &p, int *applied) {
    fb_ctx *ctx = img->ctx;
    *applied = 0;
    if (p.size() != 3) { ctx->err = "Palette: incorrect parameters"; return FB_ERR_INVALID; }
    const int begin_c = img->info.nb_meta_channels + p[0], end_c = img->info.nb_meta_channels + p[1];
    if (p[0] < 0 || begin_c > end_c || end_c >= (int)img->ch.size()) { ctx->err = "Palette: incorrect parameters"; return FB_ERR_INVALID; }
    const int nb = end_c - begin_c + 1;
    if (nb > 4) { ctx->err = "Palette over more than four channels"; return FB_ERR_UNSUPPORTED; }
    const int w = img->ch[begin_c].d.w, h = img->ch[begin_c].d.h;
    int16_t *planes[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int c = 0; c < nb; c++) {
        FbChan &ch = img->ch[begin_c + c];
        if (ch.d.w != w || ch.d.h != h) { ctx->err = "Palette over channels of different sizes"; return FB_ERR_UNSUPPORTED; }
        int rc = chan_materialize(ctx, ch);
        if (rc) return rc;
        planes[c] = ch.dev;
    }
    const size_t n = chan_samples(img->ch[begin_c].d);
    std::vector<unsigned long long> sorted;
    int too_many = 0;
    int rc = fb_palette_collect(ctx, planes, nb, n, p[2], sorted, &too_many);
    if (rc) return rc;
    if (too_many) return FB_OK;
    const int count = (int)sorted.size();
    p[2] = count;
    // the palette meta-channel: `count` columns, one row per channel, hshift -1
    FbChan pch;
    chan_defaults(pch.d);
    pch.d.w = count; pch.d.h = nb; pch.d.minval = 0; pch.d.maxval = 1; pch.d.hshift = -1;
    chan_setzero(pch.d);
    pch.d.decoded = 1;
    std::vector<int16_t> pal((size_t)count * nb + 1);
    for (int k = 0; k < count; k++) for (int c = 0; c < nb; c++) pal[(size_t)c * count + k] = (int16_t)pl_unpack(sorted[(size_t)k], c);
    if ((rc = fb_plane_alloc(ctx, chan_samples(pch.d), &pch.dev))) return rc;
    unsigned long long *sorted_dev = nullptr;
    FB_CUDA(ctx, cudaMallocAsync((void **)&sorted_dev, (sorted.size() + 1) * sizeof(unsigned long long), ctx->stream));
    if (count) {
        FB_CUDA(ctx, cudaMemcpyAsync(pch.dev, pal.data(), (size_t)count * nb * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
        FB_CUDA(ctx, cudaMemcpyAsync(sorted_dev, sorted.data(), sorted.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
    }
    rc = fb_launch_palette_index(ctx, planes, nb, n, sorted_dev, count);
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // pal / sorted are host vectors about to go out of scope
    cudaFreeAsync(sorted_dev, ctx->stream);
    if (rc) return rc;
    img->info.nb_meta_channels++;
    img->info.nb_channels -= nb - 1;
    for (int c = begin_c + 1; c <= end_c; c++) if (img->ch[c].dev) fb_plane_free(ctx, img->ch[c].dev);
    img->ch.erase(img->ch.begin() + begin_c + 1, img->ch.begin() + end_c + 1);
    img->ch.insert(img->ch.begin(), pch);
    *applied = 1;
    return FB_OK;
}

// inv_palette, palette.h:32-68
static int inv_palette(fb_image *img, const std::vector<int> &p) {
    fb_ctx *ctx = img->ctx;
    if (img->info.nb_meta_channels < 1) { ctx->err = "Palette transform without palette"; return FB_ERR_INVALID; }
    if (p.size() != 3) { ctx->err = "Palette: incorrect parameters"; return FB_ERR_INVALID; }
    const int nb = img->ch[0].d.h;
    const int c0 = img->info.nb_meta_channels + p[0];
    if (p[0] < 0 || c0 >= (int)img->ch.size()) { ctx->err = "Palette: incorrect parameters"; return FB_ERR_INVALID; }
    if (nb < 1 || nb > 8 || img->ch[0].d.w < 1) { ctx->err = "Palette: more than 8 channels or an empty palette"; return FB_ERR_UNSUPPORTED; }
    if (!img->ch[c0].dev && chan_samples(img->ch[c0].d)) {
        // the reference then reads AND writes the channel's `zero` once per pixel (image.h:84): not reproduced
        ctx->err = "Palette: the index channel was not decoded";
        return FB_ERR_UNSUPPORTED;
    }
    const int w = img->ch[c0].d.w, h = img->ch[c0].d.h;
    for (int i = 1; i < nb; i++) {
        FbChan d;
        chan_defaults(d.d);
        d.d.w = w; d.d.h = h; d.d.minval = 0; d.d.maxval = 1;
        chan_setzero(d.d);
        d.d.decoded = 1;
        int rc = fb_plane_alloc(ctx, chan_samples(d.d), &d.dev);
        if (rc) return rc;
        img->ch.insert(img->ch.begin() + c0 + 1, d);
        img->ch[c0 + i].d.component = p[0] + i;         // as written in the reference: the channel at c0+i, whichever it is by now
    }
    int rc = chan_materialize(ctx, img->ch[0]);         // an undecoded palette reads as `zero` everywhere
    if (rc) return rc;
    int16_t *outs[8];
    for (int c = 0; c < nb; c++) outs[c] = img->ch[c0 + c].dev;
    if ((rc = fb_launch_palette_inv(ctx, outs, nb, img->ch[0].dev, img->ch[0].d.w, chan_samples(img->ch[c0].d)))) return rc;
    img->info.nb_channels += nb - 1;
    img->info.nb_meta_channels--;
    fb_plane_free(ctx, img->ch[0].dev);
    img->ch.erase(img->ch.begin());
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Permute (reference transform/permute.h) with explicit parameters: pure channel-list bookkeeping, the planes stay where
// they are in HBM.  The reference's other mode (permutation stored in a meta-channel) cannot round-trip there -- fwd_permute
// leaves the parameters in the Transform, so the decoder's meta_permute takes the explicit branch -- and is not offered.
// ---------------------------------------------------------------------------------------------------------

// meta_permute, permute.h:57-83: channel i of the list goes to position parameters[i]
static int meta_permute(fb_image *img, const std::vector<int> &p) {
    const int m = img->info.nb_meta_channels, nb = (int)img->ch.size() - m, np = (int)p.size();
    if (np == 0) { img->ctx->err = "Permute through a meta-channel is not supported"; return FB_ERR_UNSUPPORTED; }
    if (np > nb) { img->ctx->err = "Permute: incorrect number of parameters"; return FB_ERR_INVALID; }
    for (int i = 0; i < np; i++) {
        if (p[i] < 0 || p[i] >= np) { img->ctx->err = "Permute: invalid permutation"; return FB_ERR_INVALID; }
        for (int j = 0; j < i; j++) if (p[i] == p[j]) { img->ctx->err = "Permute: invalid permutation"; return FB_ERR_INVALID; }
    }
    const std::vector<FbChan> old(img->ch.begin() + m, img->ch.begin() + m + np);
    for (int i = 0; i < np; i++) img->ch[m + p[i]] = old[i];
    return FB_OK;
}
// fwd_permute, permute.h:85-124: a leading -1 selects the explicit mode and is dropped from the stored parameters
static int fwd_permute(fb_image *img, std::vector<int> &p) {
    if (p.size() < 3) { img->ctx->err = "Permute: not enough parameters"; return FB_ERR_INVALID; }
    if (p[0] != -1) { img->ctx->err = "Permute through a meta-channel is not supported"; return FB_ERR_UNSUPPORTED; }
    p.erase(p.begin());
    return meta_permute(img, p);
}
// inv_permute, permute.h:31-55: position i gets back the channel that sits at parameters[i]
static int inv_permute(fb_image *img, const std::vector<int> &p) {
    const int m = img->info.nb_meta_channels, np = (int)p.size();
    if (np == 0) { img->ctx->err = "Permute through a meta-channel is not supported"; return FB_ERR_UNSUPPORTED; }
    if (np > (int)img->ch.size() - m) { img->ctx->err = "Permute: incorrect number of parameters"; return FB_ERR_INVALID; }
    std::vector<char> seen((size_t)np, 0);          // a repeated index would alias one plane twice (and free it twice later)
    for (int i = 0; i < np; i++) {
        if (p[i] < 0 || p[i] >= np || seen[p[i]]) { img->ctx->err = "Permute: invalid permutation"; return FB_ERR_INVALID; }
        seen[p[i]] = 1;
    }
    const std::vector<FbChan> old(img->ch.begin() + m, img->ch.begin() + m + np);
    for (int i = 0; i < np; i++) img->ch[m + i] = old[p[i]];
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// 2DMatch (reference transform/2dmatch.h): inverse (exact and soft matches) + decode-time meta step, single-frame images.  The
// forward (a search heuristic, :196-385) is not offered.
// ---------------------------------------------------------------------------------------------------------

static std::vector<int> match_parameters(const fb_image *img, const std::vector<int> &p) {     // default_match_parameters, :89-95
    if (p.empty()) return {0, img->info.nb_channels - 1, 0, 1000000};
    return p;
}
// meta_match, 2dmatch.h:179-194
static int meta_match(fb_image *img, const std::vector<int> &p0) {
    const std::vector<int> p = match_parameters(img, p0);
    if (p.size() < 3) { img->ctx->err = "2DMatch: incorrect parameters"; return FB_ERR_INVALID; }
    const int begin_c = img->info.nb_meta_channels + p[0], end_c = img->info.nb_meta_channels + p[1];
    if (p[0] < 0 || begin_c > end_c || end_c >= (int)img->ch.size()) { img->ctx->err = "2DMatch: incorrect parameters"; return FB_ERR_INVALID; }
    img->info.nb_meta_channels++;
    FbChan mch;
    chan_defaults(mch.d);
    mch.d.w = img->ch[begin_c].d.w; mch.d.h = img->ch[begin_c].d.h; mch.d.minval = 0; mch.d.maxval = 1;
    chan_setzero(mch.d);
    img->ch.insert(img->ch.begin(), mch);
    return FB_OK;
}
// inv_match, 2dmatch.h:97-177
static int inv_match(fb_image *img, const std::vector<int> &p0) {
    fb_ctx *ctx = img->ctx;
    if (img->info.nb_meta_channels < 1) { ctx->err = "2DMatch transform without match channel"; return FB_ERR_INVALID; }
    const std::vector<int> p = match_parameters(img, p0);
    if (p.size() < 3) { ctx->err = "2DMatch: incorrect parameters"; return FB_ERR_INVALID; }
    const int c0 = img->info.nb_meta_channels + p[0], cn = img->info.nb_meta_channels + p[1], nch = (int)img->ch.size();
    if (p[0] < 0 || p[1] < p[0] || c0 >= nch || cn >= nch) { ctx->err = "2DMatch: incorrect parameters"; return FB_ERR_INVALID; }
    const bool softmatch = p[2] != 0;
    FbChan &m = img->ch[0];
    if (m.d.q != 1) { ctx->err = "2DMatch against previous frames (animations) is not supported"; return FB_ERR_UNSUPPORTED; }
    const int w = img->ch[c0].d.w, h = img->ch[c0].d.h;
    if (m.dev && chan_samples(m.d)) {           // an undecoded match channel reads as zero everywhere: nothing is matched
        if (m.d.w != w || m.d.h != h || (long long)w * h > 0x7fffffffLL) { ctx->err = "2DMatch: match channel of a different size"; return FB_ERR_UNSUPPORTED; }
        for (int c = c0; c <= cn; c++)
            if (!img->ch[c].dev || img->ch[c].d.w != w || img->ch[c].d.h != h) { ctx->err = "2DMatch over undecoded channels or channels of different sizes"; return FB_ERR_UNSUPPORTED; }
        const int n = w * h;
        int *parent = nullptr, bad = 0;
        int rc = FB_OK;
        if (softmatch) {        // sums along the chains, one channel at a time
            for (int c = c0; c <= cn; c++) {
                FbChan &ch = img->ch[c];
                int16_t *out = nullptr;
                if ((rc = fb_plane_alloc(ctx, (size_t)n, &out))) return rc;
                rc = fb_match_soft(ctx, m.dev, ch.dev, out, n, w, m.d.maxval, ch.d.zero, &bad);
                if (rc || bad) { fb_plane_free(ctx, out); if (rc) return rc; ctx->err = "2DMatch: match code out of range"; return FB_ERR_INVALID; }
                fb_plane_free(ctx, ch.dev);
                ch.dev = out;
            }
        } else {
        rc = fb_match_resolve(ctx, m.dev, n, w, m.d.maxval, &parent, &bad);
        if (rc) return rc;
        if (bad) { cudaFreeAsync(parent, ctx->stream); ctx->err = "2DMatch: match code out of range"; return FB_ERR_INVALID; }
        for (int c = c0; c <= cn && !rc; c++) {
            FbChan &ch = img->ch[c];
            int16_t *out = nullptr;
            if ((rc = fb_plane_alloc(ctx, (size_t)n, &out))) break;
            rc = fb_launch_match_gather(ctx, ch.dev, out, parent, n, ch.d.zero);
            fb_plane_free(ctx, ch.dev);
            ch.dev = out;
        }
        if (parent) cudaFreeAsync(parent, ctx->stream);
        if (rc) return rc;
        }
    }
    img->info.nb_meta_channels--;
    if (m.dev) fb_plane_free(ctx, m.dev);
    img->ch.erase(img->ch.begin());
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Approximate (reference transform/approximate.h): parameters = first channel, last channel, divisor - 1 per channel (the
// last one repeats; 0 = leave the channel alone).  Channel numbers are absolute (meta channels included), as in the reference.
// ---------------------------------------------------------------------------------------------------------

static int approx_param(const std::vector<int> &p, int c, int beginc) {      // approximate.h:37
    const size_t k = (size_t)(c + 2 - beginc);
    return k < p.size() ? p[k] : p.back();
}

// meta_approximate, approximate.h:64-80: a copy of every approximated channel (geometry, range, q) goes to the end of the list
static int meta_approximate(fb_image *img, const std::vector<int> &p) {
    if (p.size() < 3) { img->ctx->err = "Approximate: incorrect number of parameters"; return FB_ERR_INVALID; }
    const int nb = p[1] - p[0] + 1;
    if (nb < 1 || p[0] < 0 || p[1] >= (int)img->ch.size()) { img->ctx->err = "Approximate: incorrect parameters"; return FB_ERR_INVALID; }
    for (int c = p[0]; c <= p[1]; c++) {
        if (!approx_param(p, c, p[0])) continue;
        FbChan r;
        r.d = img->ch[c].d;
        r.d.decoded = 0;
        r.dev = nullptr;
        img->ch.push_back(r);
    }
    return FB_OK;
}

// fwd_approximate, approximate.h:83-113
static int fwd_approximate(fb_image *img, const std::vector<int> &p) {
    fb_ctx *ctx = img->ctx;
    const int offset = (int)img->ch.size();
    int rc = meta_approximate(img, p);
    if (rc) return rc;
    int i = 0;
    for (int c = p[0]; c <= p[1]; c++) {
        const int q = approx_param(p, c, p[0]) + 1;
        if (q == 1) continue;
        if (q < 1) { ctx->err = "Approximate: negative divisor"; return FB_ERR_INVALID; }
        FbChan &ch = img->ch[c], &chr = img->ch[offset + i];
        i++;
        if ((rc = chan_materialize(ctx, ch))) return rc;
        if ((rc = fb_plane_alloc(ctx, chan_samples(ch.d), &chr.dev))) return rc;
        chr.d.decoded = 1;
        if ((rc = fb_launch_approximate(ctx, ch.dev, chr.dev, chan_samples(ch.d), q, 0))) return rc;
        ch.d.minval = s16(ch.d.minval / q); ch.d.maxval = s16(ch.d.maxval / q);
        chr.d.minval = 0; chr.d.maxval = s16(q - 1);
        chr.d.q = ch.d.q;       // the quantisation factor travels with the remainder in case the quotient becomes all zero
    }
    return FB_OK;
}

// inv_approximate, approximate.h:32-62
static int inv_approximate(fb_image *img, const std::vector<int> &p) {
    fb_ctx *ctx = img->ctx;
    if (p.size() < 3) { ctx->err = "Approximate: incorrect number of parameters"; return FB_ERR_INVALID; }
    const int beginc = p[0], endc = p[1], nch = (int)img->ch.size();
    int offset = nch - (endc - beginc + 1);
    for (int c = beginc; c <= endc; c++) if (beginc <= endc && !approx_param(p, c, beginc)) offset++;
    if (beginc < 0 || endc >= nch || endc < beginc || offset <= endc || offset > nch) { ctx->err = "Approximate: incorrect parameters"; return FB_ERR_INVALID; }
    int i = 0;
    for (int c = beginc; c <= endc; c++) {
        const int q = approx_param(p, c, beginc) + 1;
        if (q == 1) continue;
        if (offset + i >= nch) { ctx->err = "Approximate: remainder channel missing"; return FB_ERR_INVALID; }
        FbChan &ch = img->ch[c];
        const FbChan &chr = img->ch[offset + i];
        i++;
        if (chr.dev) ch.d.q = chr.d.q;
        const size_t n = chan_samples(ch.d);
        if (ch.dev) {
            if (chr.dev && chan_samples(chr.d) != n) { ctx->err = "Approximate: remainder channel of a different size"; return FB_ERR_INVALID; }
            int rc = fb_launch_approximate(ctx, ch.dev, chr.dev, n, q, 1);
            if (rc) return rc;
        } else {
            // an undecoded channel has no samples: every ch.value(y, x) of the reference's loop is the channel's `zero`
            // (image.h:84 returns a reference to it), which is therefore multiplied w*h times
            int z = ch.d.zero;
            for (size_t k = 0; k < n && z != 0; k++) z = s16(z * q);
            ch.d.zero = z;
        }
    }
    for (int c = offset; c < nch; c++) if (img->ch[c].dev) fb_plane_free(ctx, img->ch[c].dev);
    img->ch.erase(img->ch.begin() + offset, img->ch.end());
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Transform dispatch (reference transform/transform.cpp:48-81)
// ---------------------------------------------------------------------------------------------------------

static int transform_meta_apply(fb_image *img, FbXform &t) {
    switch (t.id) {
    case FB_TRANSFORM_YCBCR: case FB_TRANSFORM_YCOCG: case FB_TRANSFORM_QUANTIZE: return FB_OK;
    case FB_TRANSFORM_SQUEEZE: return meta_squeeze(img, t.p);
    case FB_TRANSFORM_DCT: return meta_dct(img, t.p);
    case FB_TRANSFORM_SUBSAMPLE: return meta_subsample(img, t.p);
    case FB_TRANSFORM_APPROXIMATE: return meta_approximate(img, t.p);
    case FB_TRANSFORM_PALETTE: return meta_palette(img, t.p);
    case FB_TRANSFORM_PERMUTE: return meta_permute(img, t.p);
    case FB_TRANSFORM_2DMATCH: return meta_match(img, t.p);
    default:
        img->ctx->err = "transform " + std::to_string(t.id) + " is outside the hot path (SURVEY.md 8: out of scope)";
        return FB_ERR_UNSUPPORTED;
    }
}

static int fb_image_undo_transforms_impl(fb_image *img, int keep) {
    if (!img || keep < 0) return FB_ERR_INVALID;
    fb_ctx *ctx = img->ctx;
    cudaSetDevice(ctx->device);
    bool clamped = false;
    ctx->mark("undo_transforms begin");
    while ((int)img->tr.size() > keep) {
        FbXform &t = img->tr.back();
        int rc = FB_OK, applied = 1;
        const bool last = img->tr.size() == 1 && keep == 0;
        switch (t.id) {
        case FB_TRANSFORM_YCBCR: rc = do_ycbcr(img, true, &applied); break;
        case FB_TRANSFORM_YCOCG: {
            // the final clamp (image.cpp:107-113) is folded into the YCoCg kernel when YCoCg is the last
            // transform and covers every plane of the image
            const bool fuse = last && (int)img->ch.size() == img->info.nb_meta_channels + 3 && img->info.nb_meta_channels == 0;
            rc = do_ycocg(img, true, fuse, &applied);
            if (!rc && fuse) clamped = true;
            break;
        }
        case FB_TRANSFORM_QUANTIZE: {
            int fused = 0;
            rc = try_fused_dct_tail(img, keep, &fused, &clamped);
            if (!rc && fused) {
                for (int k = 0; k < fused - 1; k++) img->tr.pop_back();        // the last one is popped below
                break;
            }
            if (!rc) rc = do_quantize(img, true, t.p);
            break;
        }
        case FB_TRANSFORM_SQUEEZE: {
            // look ahead: an inverse YCoCg and / or the final clamp right after the Squeeze ride on its last launch
            const int nt = (int)img->tr.size();
            int ep_kind = 0, do_clamp = 0, ep_done = 0;
            if (nt >= 2 && nt - 2 >= keep && img->tr[nt - 2].id == FB_TRANSFORM_YCOCG && img->info.nb_meta_channels == 0) {
                ep_kind = 2;
                do_clamp = (nt == 2 && keep == 0) ? 1 : 0;
            } else if (last) {
                ep_kind = 1;
                do_clamp = 1;
            }
            const std::vector<int> params = t.p;
            rc = inv_squeeze(img, params, ep_kind, do_clamp, &ep_done);
            if (!rc && ep_done) {
                if (ep_kind == 2) img->tr.pop_back();      // the Squeeze; the YCoCg entry is popped below
                if (do_clamp) clamped = true;
            }
            break;
        }
        case FB_TRANSFORM_DCT: rc = inv_dct(img, t.p); break;
        case FB_TRANSFORM_SUBSAMPLE: rc = inv_subsample(img, t.p); break;
        case FB_TRANSFORM_APPROXIMATE: rc = inv_approximate(img, t.p); break;
        case FB_TRANSFORM_PALETTE: rc = inv_palette(img, t.p); break;
        case FB_TRANSFORM_PERMUTE: rc = inv_permute(img, t.p); break;
        case FB_TRANSFORM_2DMATCH: rc = inv_match(img, t.p); break;
        default:
            ctx->err = "cannot undo transform " + std::to_string(t.id) + " (outside the hot path)";
            rc = FB_ERR_UNSUPPORTED;
        }
        if (rc) { img->info.error = 1; return rc; }
        img->tr.pop_back();
    }
    if (!keep && !clamped) {
        for (auto &c : img->ch) {
            if (!c.dev) continue;
            int rc = fb_launch_clamp(ctx, c.dev, chan_samples(c.d), img->info.minval, img->info.maxval);
            if (rc) return rc;
        }
    }
    img->info.nb_planes = (int)img->ch.size();
    img->info.nb_transforms = (int)img->tr.size();
    return FB_OK;
}

extern "C" int fb_image_undo_transforms(fb_image *img, int keep) {
    FB_NEEDS_DEVICE(img);
    return abi_guard(img ? img->ctx : nullptr, [&]() { return fb_image_undo_transforms_impl(img, keep); });
}

static int fb_image_do_transform_impl(fb_image *img, int32_t id, const int32_t *params, int nparams, int *applied_out) {
    if (!img || nparams < 0 || (nparams && !params)) return FB_ERR_INVALID;
    cudaSetDevice(img->ctx->device);
    FbXform t;
    t.id = id;
    t.p.assign(params, params + nparams);
    int applied = 0, rc = FB_OK;
    switch (id) {
    case FB_TRANSFORM_YCBCR: rc = do_ycbcr(img, false, &applied); if (rc == FB_ERR_INVALID) { rc = FB_OK; applied = 0; } break;
    case FB_TRANSFORM_YCOCG: rc = do_ycocg(img, false, false, &applied); if (rc == FB_ERR_INVALID) { rc = FB_OK; applied = 0; } break;
    case FB_TRANSFORM_QUANTIZE: rc = do_quantize(img, false, t.p); applied = rc == FB_OK; break;
    case FB_TRANSFORM_SQUEEZE: rc = fwd_squeeze(img, t.p); applied = rc == FB_OK; break;
    case FB_TRANSFORM_DCT: rc = fwd_dct(img, t.p, &applied); break;
    case FB_TRANSFORM_APPROXIMATE: rc = fwd_approximate(img, t.p); applied = rc == FB_OK; break;
    case FB_TRANSFORM_PALETTE: rc = fwd_palette(img, t.p, &applied); break;
    case FB_TRANSFORM_PERMUTE: rc = fwd_permute(img, t.p); applied = rc == FB_OK; break;
    case FB_TRANSFORM_SUBSAMPLE: applied = 0; break;       // fwd_subsample is a stub in the reference: "return false"
    default:
        img->ctx->err = "transform " + std::to_string(id) + " is outside the hot path";
        rc = FB_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
    if (applied) {
        if (id == FB_TRANSFORM_DCT) t.p.assign(params, params + nparams);
        img->tr.push_back(t);
    }
    img->info.nb_planes = (int)img->ch.size();
    img->info.nb_transforms = (int)img->tr.size();
    if (applied_out) *applied_out = applied;
    return FB_OK;
}

extern "C" int fb_image_do_transform(fb_image *img, int32_t id, const int32_t *params, int nparams, int *applied_out) {
    FB_NEEDS_DEVICE(img);
    return abi_guard(img ? img->ctx : nullptr, [&]() { return fb_image_do_transform_impl(img, id, params, nparams, applied_out); });
}

extern "C" int fb_image_recompute_minmax(fb_image *img) {
    if (!img) return FB_ERR_INVALID;
    FB_NEEDS_DEVICE(img);
    fb_ctx *ctx = img->ctx;
    cudaSetDevice(ctx->device);
    const size_t n = img->ch.size();
    if (!n) return FB_OK;
    int *dev = nullptr;
    FB_CUDA(ctx, cudaMallocAsync((void **)&dev, n * 2 * sizeof(int), ctx->stream));
    for (size_t i = 0; i < n; i++) {
        int rc = fb_launch_minmax(ctx, img->ch[i].dev, img->ch[i].dev ? chan_samples(img->ch[i].d) : 0, dev + 2 * i);
        if (rc) return rc;
    }
    std::vector<int> host(n * 2);
    FB_CUDA(ctx, cudaMemcpyAsync(host.data(), dev, n * 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(dev, ctx->stream);
    for (size_t i = 0; i < n; i++) { img->ch[i].d.minval = s16(host[2 * i]); img->ch[i].d.maxval = s16(host[2 * i + 1]); }
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// create / query / download
// ---------------------------------------------------------------------------------------------------------

extern "C" int fb_image_create(fb_ctx *ctx, const fb_image_info *info, const fb_plane_desc *desc, const int16_t *const *planes,
                               const int32_t *tids, const int32_t *tnp, const int32_t *tparams, fb_image **out) {
    if (!ctx || !info || !out || (info->nb_planes && !desc)) return FB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    fb_image *img = new (std::nothrow) fb_image();
    if (!img) return FB_ERR_NOMEM;
    img->ctx = ctx;
    img->info = *info;
    img->ch.resize(info->nb_planes);
    for (int i = 0; i < info->nb_planes; i++) {
        img->ch[i].d = desc[i];
        const size_t n = chan_samples(desc[i]);
        if (desc[i].decoded && planes && planes[i]) {
            int rc = fb_plane_alloc(ctx, n, &img->ch[i].dev);
            if (rc) { fb_image_destroy(img); return rc; }
            if (n && cudaMemcpyAsync(img->ch[i].dev, planes[i], n * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
                ctx->err = "H2D plane copy failed";
                fb_image_destroy(img);
                return FB_ERR_CUDA;
            }
        } else {
            img->ch[i].d.decoded = 0;
        }
    }
    int k = 0;
    for (int i = 0; i < info->nb_transforms; i++) {
        FbXform t;
        t.id = tids[i];
        t.p.assign(tparams + k, tparams + k + tnp[i]);
        k += tnp[i];
        img->tr.push_back(t);
    }
    *out = img;
    return FB_OK;
}

extern "C" int fb_image_get_info(fb_image *img, fb_image_info *info) {
    if (!img || !info) return FB_ERR_INVALID;
    img->info.nb_planes = (int)img->ch.size();
    img->info.nb_transforms = (int)img->tr.size();
    *info = img->info;
    return FB_OK;
}

extern "C" int fb_image_get_plane(fb_image *img, int i, fb_plane_desc *desc) {
    if (!img || !desc || i < 0 || i >= (int)img->ch.size()) return FB_ERR_INVALID;
    *desc = img->ch[i].d;
    desc->decoded = (img->ch[i].dev || img->ch[i].host) ? 1 : 0;
    return FB_OK;
}

extern "C" int fb_image_get_transform(fb_image *img, int i, int32_t *id, int32_t *params, int cap) {
    if (!img || i < 0 || i >= (int)img->tr.size()) return -1;
    if (id) *id = img->tr[i].id;
    for (int k = 0; k < cap && k < (int)img->tr[i].p.size(); k++) params[k] = img->tr[i].p[k];
    return (int)img->tr[i].p.size();
}

extern "C" void *fb_image_plane_device_ptr(fb_image *img, int i) {
    if (!img || i < 0 || i >= (int)img->ch.size()) return nullptr;
    return img->ch[i].dev;
}

extern "C" int fb_image_download_plane(fb_image *img, int i, int16_t *dst) {
    if (img && dst && img->on_host && i >= 0 && i < (int)img->ch.size() && img->ch[i].host) {
        memcpy(dst, img->ch[i].host, chan_samples(img->ch[i].d) * sizeof(int16_t));
        return FB_OK;
    }
    if (!img || !dst || i < 0 || i >= (int)img->ch.size() || !img->ch[i].dev) return FB_ERR_INVALID;
    fb_ctx *ctx = img->ctx;
    cudaSetDevice(ctx->device);
    const size_t n = chan_samples(img->ch[i].d);
    if (n) FB_CUDA(ctx, cudaMemcpyAsync(dst, img->ch[i].dev, n * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FB_OK;
}

extern "C" int fb_image_download_interleaved(fb_image *img, int nch, int bps, void *dst) {
    FB_NEEDS_DEVICE(img);
    if (!img || !dst || nch < 1 || nch > 8 || nch > (int)img->ch.size() || (bps != 1 && bps != 2)) return FB_ERR_INVALID;
    fb_ctx *ctx = img->ctx;
    cudaSetDevice(ctx->device);
    const int w = img->ch[0].d.w, h = img->ch[0].d.h;
    const int16_t *planes[8];
    for (int c = 0; c < nch; c++) {
        if (!img->ch[c].dev || img->ch[c].d.w != w || img->ch[c].d.h != h) { ctx->err = "interleave: planes missing or of different size"; return FB_ERR_INVALID; }
        planes[c] = img->ch[c].dev;
    }
    const size_t npix = (size_t)w * h, bytes = npix * nch * bps;
    void *dev = nullptr;
    FB_CUDA(ctx, cudaMallocAsync(&dev, std::max<size_t>(bytes, 16), ctx->stream));
    int rc = fb_launch_interleave(ctx, planes, nch, npix, bps, dev);
    if (rc) return rc;
    FB_CUDA(ctx, cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(dev, ctx->stream);
    return FB_OK;
}

extern "C" int fb_image_group_index(fb_image *img, int64_t *offsets, int32_t *first_channel, int cap) {
    if (!img) return 0;
    const int n = (int)img->group_off.size();
    for (int i = 0; i < n && i < cap; i++) {
        if (offsets) offsets[i] = img->group_off[i];
        if (first_channel) first_channel[i] = img->group_first[i];
    }
    return n;
}

// ---------------------------------------------------------------------------------------------------------
// Container parser: fuif_decode up to the first channel group (reference encoding/encoding.cpp:599-706)
// ---------------------------------------------------------------------------------------------------------

namespace {
struct ByteReader {         // FileIO semantics (fileio.h:33-81): EOF only after a failed read
    const uint8_t *p; size_t n, pos = 0; bool eof = false;
    int get() { if (pos >= n) { eof = true; return -1; } return p[pos++]; }
};
int read_varint(ByteReader &io) {        // read_big_endian_varint, encoding.cpp:45-59
    int result = 0, bytes_read = 0;
    while (bytes_read++ < 10) {
        int number = io.get();
        if (number < 0) break;
        if (number < 128) return result + number;
        number -= 128;
        result += number;
        result = (int)((unsigned)result << 7);
    }
    return -1;
}
bool transform_has_parameters(int id) {  // Transform::has_parameters, transform/transform.h:85-102
    return id == 3 || id == 4 || id == 6 || id == 7 || id == 8 || id == 9 || id == 10;
}
// bounds on what an (untrusted) header may ask for; far above every configuration this library is meant for
constexpr int kMaxHeaderChannels = 4096;
constexpr long long kMaxHeaderPixels = 1ll << 31;
constexpr int kMaxHeaderTransforms = 4096;
struct Header {
    int w, h, bit_depth, nb_channels, colormodel, max_properties;
    int responsive_offsets[5];
    size_t after_offsets;
};
int parse_header(ByteReader &io, Header &hd) {
    if (io.n < 4) return FB_ERR_INVALID;
    bool multi = false;
    if (!memcmp(io.p, "FUAF", 4)) multi = true;
    else if (memcmp(io.p, "FUIF", 4)) return FB_ERR_INVALID;
    io.pos = 4;
    hd.nb_channels = read_varint(io) - '0';
    hd.bit_depth = read_varint(io) - '&';
    hd.w = read_varint(io) + 1;
    hd.h = read_varint(io) + 1;
    if (multi) {            // animation fields (encoding.cpp:614-623): frames remain a vertical filmstrip
        int nb_frames = read_varint(io) + 2;
        (void)read_varint(io);
        int numerator = read_varint(io);
        if (numerator) for (int i = 1; i < nb_frames && !io.eof; i++) (void)read_varint(io);
        (void)read_varint(io);
    }
    hd.colormodel = read_varint(io);
    hd.max_properties = read_varint(io);
    if (io.eof || hd.nb_channels < 0 || hd.bit_depth < 1 || hd.bit_depth > 16 || hd.w < 1 || hd.h < 1 || hd.max_properties < 0) return FB_ERR_INVALID;
    return FB_OK;
}
}  // namespace

extern "C" int fb_peek_header(const uint8_t *bytes, size_t nbytes, fb_image_info *info) {
    if (!bytes || !info) return FB_ERR_INVALID;
    ByteReader io{bytes, nbytes};
    Header hd;
    int rc = parse_header(io, hd);
    if (rc) return rc;
    memset(info, 0, sizeof(*info));
    info->w = hd.w; info->h = hd.h; info->minval = 0; info->maxval = (1 << hd.bit_depth) - 1;
    info->nb_channels = info->real_nb_channels = hd.nb_channels;
    info->colormodel = hd.colormodel;
    return FB_OK;
}

// Parses header + transform list, builds the (empty) channel list via meta_apply and hands the rest to the
// GPU MANIAC decoder.
static int parse_container(fb_ctx *ctx, const uint8_t *bytes_in, size_t nbytes, const fb_decode_options *opts, const int64_t *gidx,
                           const int32_t *gfirst, int ngroups, fb_image **out, FbManiacJob &job, std::vector<uint8_t> &header_copy) {
    // host or device bytes?  For a device buffer only the header (first 4 KiB) is read back.
    const uint8_t *bytes = bytes_in;
    const uint8_t *bytes_dev = nullptr;
    cudaPointerAttributes attr;
    if (ctx->device < 0) {
        // host-only decode (fb_host_decode): the bytes are host memory by contract, no CUDA call is made
    } else if (cudaPointerGetAttributes(&attr, bytes_in) == cudaSuccess && attr.type == cudaMemoryTypeDevice) {
        header_copy.resize(std::min<size_t>(nbytes, 4096));
        FB_CUDA(ctx, cudaMemcpyAsync(header_copy.data(), bytes_in, header_copy.size(), cudaMemcpyDeviceToHost, ctx->stream));
        FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        bytes = header_copy.data();
        bytes_dev = bytes_in;
    } else {
        cudaGetLastError();
    }
    ByteReader io{bytes, bytes_dev ? header_copy.size() : nbytes};
    Header hd;
    int rc = parse_header(io, hd);
    if (rc) { ctx->err = "not a FUIF file or corrupt header"; return rc; }
    // an untrusted header must not size allocations: the reference's Image would try to allocate w*h*nb_channels samples here;
    // this library bounds what it accepts (documented in include/fuif_b200.h)
    if (hd.nb_channels > kMaxHeaderChannels || hd.w < 1 || hd.h < 1 || (long long)hd.w * hd.h > kMaxHeaderPixels) {
        ctx->err = "header asks for more channels / pixels than this library accepts";
        return FB_ERR_UNSUPPORTED;
    }
    fb_image *img = new (std::nothrow) fb_image();
    if (!img) return FB_ERR_NOMEM;
    img->ctx = ctx;
    // Image(w, h, (1<<bit_depth)-1, nb_channels, colormodel), encoding.cpp:637 / image.h:114-120
    img->info.w = hd.w; img->info.h = hd.h; img->info.minval = 0; img->info.maxval = (1 << hd.bit_depth) - 1;
    img->info.nb_channels = img->info.real_nb_channels = hd.nb_channels;
    img->info.nb_meta_channels = 0; img->info.colormodel = hd.colormodel;
    img->ch.resize(hd.nb_channels);
    for (int i = 0; i < hd.nb_channels; i++) {
        chan_defaults(img->ch[i].d);
        img->ch[i].d.w = hd.w; img->ch[i].d.h = hd.h; img->ch[i].d.maxval = img->info.maxval; img->ch[i].d.component = i;
    }
    *out = img;
    job.bytes_host = bytes; job.bytes_dev = bytes_dev; job.nbytes = nbytes; job.img = img; job.max_properties = hd.max_properties;
    job.group_first = gfirst;
    job.cutoff = opts ? opts->maniac_cutoff : 6; job.alpha = opts ? opts->maniac_alpha : 0x0d000000;
    job.group_index = gidx; job.n_groups = ngroups; job.bytes_to_load = 0; job.body_pos = io.pos;
    if (hd.nb_channels < 1) return FB_OK;

    int rel = 0;
    for (int s = 0; s < 5; s++) { hd.responsive_offsets[s] = read_varint(io) + rel; rel = hd.responsive_offsets[s]; }
    rel = (int)io.pos;
    for (int s = 0; s < 5; s++) hd.responsive_offsets[s] += rel;

    const int nb_transforms = read_varint(io);
    if (nb_transforms < 0 || nb_transforms > kMaxHeaderTransforms) { ctx->err = "corrupt transform list"; return FB_ERR_INVALID; }
    for (int i = 0; i < nb_transforms; i++) {
        const int idp = read_varint(io);
        if (idp < 0 || io.eof) { ctx->err = "truncated transform list (or header larger than 4 KiB in a device buffer)"; return FB_ERR_INVALID; }
        FbXform t;
        t.id = idp & 0xf;
        if (transform_has_parameters(t.id)) {
            const int np = idp >> 4;
            if ((size_t)np > io.n - io.pos) { ctx->err = "transform with more parameters than the file has bytes"; return FB_ERR_INVALID; }
            for (int j = 0; j < np; j++) {
                const int v = read_varint(io);
                if (io.eof) { ctx->err = "truncated transform parameters (or header larger than 4 KiB in a device buffer)"; return FB_ERR_INVALID; }
                t.p.push_back(v);
            }
        }
        if ((rc = transform_meta_apply(img, t))) return rc;
        img->tr.push_back(t);
    }
    const int preview = opts ? opts->preview : -1;
    if (preview > 4) return FB_ERR_INVALID;
    if (preview >= 0) job.bytes_to_load = (size_t)hd.responsive_offsets[preview];
    job.body_pos = io.pos;
    img->info.nb_planes = (int)img->ch.size();
    img->info.nb_transforms = (int)img->tr.size();
    return FB_OK;
}

extern "C" int fb_decode_batch(fb_ctx *ctx, int n_images, const uint8_t *const *bytes, const size_t *nbytes, const fb_decode_options *opts,
                               const int64_t *const *group_index, const int32_t *const *group_first, const int *n_groups, fb_image **out) {
    if (!ctx || n_images < 0 || !bytes || !nbytes || !out) return FB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < n_images; i++) out[i] = nullptr;
    int rc = abi_guard(ctx, [&]() {
        std::vector<FbManiacJob> jobs(n_images);
        std::vector<std::vector<uint8_t>> headers(n_images);
        int r = FB_OK;
        for (int i = 0; i < n_images && !r; i++)
            r = parse_container(ctx, bytes[i], nbytes[i], opts, group_index ? group_index[i] : nullptr, group_first ? group_first[i] : nullptr,
                                n_groups ? n_groups[i] : 0, &out[i], jobs[i], headers[i]);
        if (!r) r = fb_maniac_decode(ctx, jobs);
        return r;
    });
    if (rc) {
        for (int i = 0; i < n_images; i++) { fb_image_destroy(out[i]); out[i] = nullptr; }
        return rc;
    }
    return FB_OK;
}

static thread_local std::string g_host_err;
extern "C" const char *fb_host_last_error(void) { return g_host_err.c_str(); }

extern "C" int fb_host_decode(const uint8_t *bytes, size_t nbytes, const fb_decode_options *opts, const int64_t *group_index,
                              const int32_t *group_first, int n_groups, int threads, fb_image **out) {
    if (!bytes || !out || threads < 0) return FB_ERR_INVALID;
    *out = nullptr;
    fb_ctx *hctx = new (std::nothrow) fb_ctx();      // a context without a device: no CUDA call is made on this path
    if (!hctx) return FB_ERR_NOMEM;
    hctx->device = -1;
    hctx->entropy_backend = FB_ENTROPY_HOST;
    hctx->host_threads = threads;
    fb_image *img = nullptr;
    int rc = abi_guard(hctx, [&]() {
        std::vector<FbManiacJob> jobs(1);
        std::vector<uint8_t> header;
        int r = parse_container(hctx, bytes, nbytes, opts, group_index, group_first, n_groups, &img, jobs[0], header);
        if (!r) r = fb_maniac_decode(hctx, jobs);
        return r;
    });
    if (rc) {
        g_host_err = hctx->err;
        if (img) { img->owns_ctx = false; fb_image_destroy(img); }
        delete hctx;
        return rc;
    }
    img->on_host = true;
    img->owns_ctx = true;
    *out = img;
    return FB_OK;
}

extern "C" int fb_image_upload(fb_ctx *ctx, fb_image *img) {
    if (!ctx || !img || ctx->device < 0) return FB_ERR_INVALID;
    if (!img->on_host) return img->ctx == ctx ? FB_OK : FB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    for (auto &c : img->ch) {
        if (!c.host) continue;
        const size_t n = chan_samples(c.d);
        int rc = fb_plane_alloc(ctx, n, &c.dev);
        if (rc) return rc;
        if (n) FB_CUDA(ctx, cudaMemcpyAsync(c.dev, c.host, n * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &c : img->ch) c.host = nullptr;
    std::vector<int16_t>().swap(img->host_block);
    if (img->owns_ctx) delete img->ctx;
    img->ctx = ctx;
    img->owns_ctx = false;
    img->on_host = false;
    return FB_OK;
}

extern "C" int fb_decode(fb_ctx *ctx, const uint8_t *bytes, size_t nbytes, const fb_decode_options *opts, const int64_t *group_index,
                         const int32_t *group_first, int n_groups, fb_image **out) {
    const int64_t *gi[1] = {group_index};
    const int32_t *gf[1] = {group_first};
    return fb_decode_batch(ctx, 1, &bytes, &nbytes, opts, group_index ? gi : nullptr, group_first ? gf : nullptr, &n_groups, out);
}

extern "C" int fb_decode_to_pixels(fb_ctx *ctx, const uint8_t *bytes, size_t nbytes, const fb_decode_options *opts, const int64_t *group_index,
                                   const int32_t *group_first, int n_groups, int bps, void *dst, size_t dst_bytes) {
    fb_image *img = nullptr;
    int rc = fb_decode(ctx, bytes, nbytes, opts, group_index, group_first, n_groups, &img);
    if (rc) return rc;
    rc = fb_image_undo_transforms(img, 0);
    if (!rc) {
        const int nch = img->info.nb_channels;
        if ((int)img->ch.size() < nch || nch < 1) rc = FB_ERR_INVALID;
        else if ((size_t)img->ch[0].d.w * img->ch[0].d.h * nch * bps > dst_bytes) { ctx->err = "destination buffer too small"; rc = FB_ERR_INVALID; }
        else rc = fb_image_download_interleaved(img, nch, bps, dst);
    }
    fb_image_destroy(img);
    return rc;
}
