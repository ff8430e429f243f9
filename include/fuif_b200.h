/* fuif_b200 -- C ABI of the B200-native FUIF hot path.
 *
 * The reference (cloudinary/fuif) has no plugin / FFI layer: its boundary is the C++ API in
 * encoding/encoding.h and image/image.h, called only from fuif.cpp and fuifplay.cpp.  This header is the
 * extern "C" layer a maintainer binds underneath that API: plain pointers and sizes,
 * int status codes (0 = ok), no C++ or torch types.  Each entry point names the reference interface it
 * replaces (file:line under the reference tree).
 *
 * Model: an fb_image is the device-resident mirror of the reference's `Image` (image/image.h:98-129): a list
 * of int16 planes in HBM (one per `Channel`, image/image.h:54-91, row-major, no padding) plus the transform
 * stack.  A context owns one CUDA device + stream; one context per host thread / GPU.
 */
#ifndef FUIF_B200_H
#define FUIF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FB_API __attribute__((visibility("default")))
#else
#define FB_API
#endif

#define FB_OK 0
#define FB_ERR_INVALID 1      /* bad argument / malformed stream ("return false" in the reference)            */
#define FB_ERR_CUDA 2         /* CUDA runtime error; fb_last_error() has the text                            */
#define FB_ERR_UNSUPPORTED 3  /* not offered: forward 2dmatch, permute via a meta-channel, palettes over > 4 channels */
#define FB_ERR_NOMEM 4

/* transform ids, reference transform/transform.h:30-70 */
#define FB_TRANSFORM_YCBCR 0
#define FB_TRANSFORM_YCOCG 1
#define FB_TRANSFORM_SUBSAMPLE 3
#define FB_TRANSFORM_DCT 4
#define FB_TRANSFORM_QUANTIZE 5
#define FB_TRANSFORM_PALETTE 6
#define FB_TRANSFORM_SQUEEZE 7
#define FB_TRANSFORM_2DMATCH 8
#define FB_TRANSFORM_PERMUTE 9
#define FB_TRANSFORM_APPROXIMATE 10

typedef struct fb_ctx fb_ctx;
typedef struct fb_image fb_image;

/* Mirrors class Channel (reference image/image.h:54-91) without the sample buffer. */
typedef struct fb_plane_desc {
    int32_t w, h;
    int32_t minval, maxval;
    int32_t zero;
    int32_t q;
    int32_t hshift, vshift;
    int32_t hcshift, vcshift;
    int32_t component;
    int32_t decoded;          /* 1 if the plane holds samples (data.size() != 0 in the reference) */
} fb_plane_desc;

/* Mirrors the scalar members of class Image (reference image/image.h:98-129). */
typedef struct fb_image_info {
    int32_t w, h;
    int32_t minval, maxval;
    int32_t nb_channels, real_nb_channels, nb_meta_channels;
    int32_t colormodel;
    int32_t nb_planes;        /* channel.size()   */
    int32_t nb_transforms;    /* transform.size() */
    int32_t error;            /* Image::error     */
} fb_image_info;

/* Mirrors struct fuif_options (reference encoding/encoding.h:32-59), decode-side members only. */
typedef struct fb_decode_options {
    int32_t preview;          /* -1 all, 0 LQIP, 1..4 = 1/16 .. 1/2 (encoding.h:34)                  */
    int32_t maniac_cutoff;    /* 6          (encoding.h:40; not in the bitstream)         */
    int32_t maniac_alpha;     /* 0x0d000000 (encoding.h:41)                                          */
    int32_t reserved;
} fb_decode_options;

/* ---- context ------------------------------------------------------------------------------------------ */

/* device: CUDA ordinal.  stream: a cudaStream_t to enqueue all work on (0 = a private stream is created). */
FB_API int fb_ctx_create(int device, void *stream, fb_ctx **out);
FB_API void fb_ctx_destroy(fb_ctx *ctx);
FB_API const char *fb_last_error(fb_ctx *ctx);
/* Blocks until everything enqueued on the context's stream has finished. */
FB_API int fb_ctx_synchronize(fb_ctx *ctx);
/* Number of kernels this library has launched on the context so far (bench.py's gpu_launches). */
FB_API long long fb_ctx_launch_count(fb_ctx *ctx);

/* Diagnostics / testing knobs (no reference counterpart).
 * FB_OPT_SQUEEZE_MODE: which kernels undo a Squeeze -- 0 = one launch per squeeze step on the direct (register
 *   resident, 128-bit access) kernels with the inverse YCoCg / clamp riding on the last step (default), 1 = the tiled
 *   per-step kernels only, 4 = multi-level fused tile kernels, 2 = fused + force the exact serial recompute,
 *   3 = fused + force the repair of every tile of the last launch.
 * FB_OPT_KERNEL_TIMING: 1 = record a CUDA event after every launch (fb_ctx_timing_report). */
#define FB_OPT_SQUEEZE_MODE 1
#define FB_OPT_KERNEL_TIMING 2
/* FB_OPT_SQUEEZE_PACKED: 1 (default) = squeeze steps of images with maxval <= 1023 whose planes qualify (width a multiple
 *   of 8, at least 64 x 32) run on the packed int16x2 kernels (TMA-fed horizontal step with the inverse YCoCg / clamp
 *   epilogue, coalesced vertical step); 0 = never.  Results are bit-exact either way. */
#define FB_OPT_SQUEEZE_PACKED 3
/* FB_OPT_ENTROPY_BACKEND: where fuif_decode_channel (encoding.cpp:259-429) runs.  FB_ENTROPY_GPU (default): k_maniac_decode,
 *   one stream per SM-resident warp team.  FB_ENTROPY_HOST: CPU threads (one channel group per thread with the group index,
 *   one image per thread without), planes copied to HBM afterwards; the transform chain stays on the GPU either way.  The
 *   serial coder of ONE group runs ~6x faster on a CPU core than on a GPU warp, so this is the backend for a single large
 *   image with a group index; batches that fill the GPU (hundreds of groups) are faster on the GPU backend.  Planes and
 *   errors are identical.  FB_OPT_HOST_THREADS: threads of the host backend, 0 (default) = four per hardware thread, at most one per
 *   stream (the large groups of a file come last in the claim order; oversubscribing starts them at once). */
#define FB_OPT_ENTROPY_BACKEND 4
#define FB_OPT_HOST_THREADS 5
#define FB_ENTROPY_GPU 0
#define FB_ENTROPY_HOST 1
FB_API int fb_ctx_set_option(fb_ctx *ctx, int option, int value);
/* The fused Squeeze inverse (mode >= 2) starts tiles speculatively and verifies them (results are bit-exact either way).
 * which = 0: Squeeze inverses so far that failed verification in an early launch and were recomputed serially;
 * which = 1: tiles of last launches so far that failed verification and were recomputed from exact states.
 * Synchronises the stream. */
#define FB_COUNTER_SERIAL_FALLBACKS 0
#define FB_COUNTER_REPAIRED_TILES 1
/* packed kernels: segments recomputed by the exact routine so far / segments that saw a value outside the packed range */
#define FB_COUNTER_PK_REPAIRED 2
#define FB_COUNTER_PK_RANGE_FLAGGED 3
/* threads the host entropy backend used in its last call */
#define FB_COUNTER_HOST_THREADS 4
FB_API long long fb_ctx_counter(fb_ctx *ctx, int which);
/* Device self-test of the packed 16x2 primitives against their exact 32-bit forms on pseudo-random inputs inside the
 * admitted range: which = 0 the unsqueeze pair (+ the range accumulator), 1 the inverse YCoCg.  scale = typical distance
 * between neighbouring values.  *mismatches = 0 on a correct device / build.  (No reference counterpart: a test hook.) */
FB_API int fb_selftest_packed(fb_ctx *ctx, int which, unsigned seed, int scale, int maxval, long long *mismatches);
/* With FB_OPT_KERNEL_TIMING: synchronises, then writes one line "name<TAB>microseconds<TAB>algorithmic bytes" per
 * launch recorded since the last report (NUL-terminated, truncated to cap) and returns the untruncated length. */
FB_API long long fb_ctx_timing_report(fb_ctx *ctx, char *buf, size_t cap);

/* ---- fuif_decode --------------------------------------------------------------------------------------- */

/* Replaces fuif_decode<BlobReader>() / fuif_decode_file() (reference encoding/encoding.cpp:599-720, 745-753):
 * parses the container on the host, runs the MANIAC range decoder + context model on the GPU
 * (fuif_decode_channel, encoding.cpp:259-429) and leaves the TRANSFORMED planes in HBM together with the
 * transform stack, exactly the state the reference's Image is in after fuif_decode().
 * bytes: the .fuif file, in HOST memory or already in DEVICE memory (detected; a device buffer is not copied,
 * only its first 4 KiB are read back for the header).  group_index (may be NULL): n_groups byte offsets of the
 * channel groups' headers (fb_image_group_index() of an earlier decode, or written by an encoder) -- with it
 * the groups decode concurrently, without it they decode back to back as the format dictates.
 * group_first (may be NULL for host bytes): first channel of each group, as fb_image_group_index() returns it.
 * End-of-stream follows FileIO (reference fileio.h:33-81).
 * Limits (the reference has none; exceeding one returns FB_ERR_UNSUPPORTED with a message, never a crash): at most 4096
 * channels and 2^31 pixels per channel in the header, 4096 transforms, MANIAC trees of at most 65535 nodes,
 * max_properties <= 18.  A damaged stream, or a group index that does not belong to the file, decodes to garbage planes or
 * FB_ERR_INVALID; the call always returns (a group header with an inverted value range, and a group that reaches beyond the planes
 * the index gives its stream, are FB_ERR_INVALID: the reference asserts / overruns there). */
FB_API int fb_decode(fb_ctx *ctx, const uint8_t *bytes, size_t nbytes, const fb_decode_options *opts,
              const int64_t *group_index, const int32_t *group_first, int n_groups, fb_image **out);

/* The host-threads entropy backend on its own, without any GPU (no fb_ctx): container parse + fuif_decode_channel on `threads`
 * CPU threads (0 = four per hardware thread, at most one per stream).  The image it returns lives in HOST memory: fb_image_get_info / get_plane /
 * get_transform / download_plane / group_index / destroy work on it, everything that computes returns FB_ERR_INVALID until
 * fb_image_upload() has moved it to a context's GPU.  fb_decode() with FB_OPT_ENTROPY_BACKEND = FB_ENTROPY_HOST is this plus the
 * upload (through pinned staging).  On failure *out is NULL and fb_host_last_error() (thread-local) has the message. */
FB_API int fb_host_decode(const uint8_t *bytes, size_t nbytes, const fb_decode_options *opts, const int64_t *group_index,
                          const int32_t *group_first, int n_groups, int threads, fb_image **out);
FB_API const char *fb_host_last_error(void);
/* Moves a host-only image into ctx's HBM (one H2D copy per plane); the image belongs to ctx afterwards. */
FB_API int fb_image_upload(fb_ctx *ctx, fb_image *img);

/* Same for a batch: every (image, group) is an independent stream of one kernel launch. */
FB_API int fb_decode_batch(fb_ctx *ctx, int n_images, const uint8_t *const *bytes, const size_t *nbytes,
                    const fb_decode_options *opts, const int64_t *const *group_index,
                    const int32_t *const *group_first, const int *n_groups, fb_image **out);

/* Byte offsets of the channel-group headers found while decoding (one per group, in stream order) and the
 * first channel of each group.  Returns the number of groups; copies at most cap entries. */
FB_API int fb_image_group_index(fb_image *img, int64_t *offsets, int32_t *first_channel, int cap);

/* ---- Image ----------------------------------------------------------------------------------------------- */

/* Builds a device image from host planes (what read_PAM_file + do_transform leave behind, or the output of the
 * reference's own fuif_decode): H2D copy of every plane.  planes[i] may be NULL for desc[i].decoded == 0.
 * transforms: ids / parameter counts / flattened parameters of Image::transform (image/image.h:101). */
FB_API int fb_image_create(fb_ctx *ctx, const fb_image_info *info, const fb_plane_desc *desc, const int16_t *const *planes,
                    const int32_t *transform_ids, const int32_t *transform_nparams, const int32_t *transform_params,
                    fb_image **out);
FB_API void fb_image_destroy(fb_image *img);

FB_API int fb_image_get_info(fb_image *img, fb_image_info *info);
FB_API int fb_image_get_plane(fb_image *img, int i, fb_plane_desc *desc);
/* Transform i of the stack: returns its parameter count; copies at most cap parameters. */
FB_API int fb_image_get_transform(fb_image *img, int i, int32_t *id, int32_t *params, int cap);
/* Device pointer of plane i (int16, w*h samples, row-major) or NULL if not decoded. Valid until the image is
 * transformed or destroyed. */
FB_API void *fb_image_plane_device_ptr(fb_image *img, int i);
/* D2H copy of plane i into dst (w*h int16). Synchronises the stream. */
FB_API int fb_image_download_plane(fb_image *img, int i, int16_t *dst);
/* D2H copy of the first n_channels planes interleaved as 8- or 16-bit samples (the layout write_PAM_file
 * emits, reference export/write_pam.h:29-168; 16-bit samples big-endian). dst holds w*h*n_channels samples. */
FB_API int fb_image_download_interleaved(fb_image *img, int n_channels, int bytes_per_sample, void *dst);

/* Replaces Image::undo_transforms(keep) (reference image/image.cpp:94-115): pops and inverts transforms until
 * `keep` are left, then (keep == 0) clamps every sample to [minval, maxval].  Runs entirely on the GPU:
 * Squeeze (transform/squeeze.h:363-388), Quantize (quantize.h:32-49), DCT (dct.h:249-296),
 * YCbCr (ycbcr.h:33-63), YCoCg (ycocg.h:33-63), ChromaSubsample (subsample.h:73-128), Approximate (approximate.h:32-62),
 * Palette (palette.h:32-68), Permute with explicit parameters (permute.h:31-55), 2DMatch (2dmatch.h:97-177). */
FB_API int fb_image_undo_transforms(fb_image *img, int keep);

/* Replaces Image::do_transform (reference image/image.cpp:117-122; forward direction of the same transforms).
 * *applied = 1 if the transform was applied and pushed on the stack, 0 if it did not apply (e.g. a Palette over channels
 * that use more colours than its third parameter allows, palette.h:109).  Palette records the number of colours it found
 * in that parameter, as the reference does. */
FB_API int fb_image_do_transform(fb_image *img, int32_t id, const int32_t *params, int nparams, int *applied);

/* Replaces Image::recompute_minmax (image/image.h:127; Channel::actual_minmax, image.cpp:82-92). */
FB_API int fb_image_recompute_minmax(fb_image *img);

/* ---- fuif_encode --------------------------------------------------------------------------------------- */

/* Mirrors struct fuif_options (reference encoding/encoding.h:32-59), encode-side members. */
typedef struct fb_encode_options {
    float nb_repeats;         /* 0.5: share of each plane's rows the tree is learned from (encoding.h:36)  */
    int32_t max_properties;   /* 12; at most 18 here                                  (encoding.h:38)      */
    int32_t maniac_cutoff;    /* 6                                                                         */
    int32_t maniac_alpha;     /* 0x0d000000                                                                */
    int32_t compress;         /* 1; 0 = plain binary coding of every sample           (encoding.h:41)      */
    int32_t max_group;        /* -1 = as many same-size channels per group as the format allows            */
    int32_t n_predictors;     /* per-channel predictor ids, the last one repeating    (encoding.h:44)      */
    const int32_t *predictor;
} fb_encode_options;

/* Replaces fuif_prepare_encode() + fuif_encode<BlobIO>() / fuif_encode_file() (reference encoding/encoding.cpp:737-743,
 * 455-573) on an image whose forward transforms have been applied (fb_image_do_transform): tightens the plane ranges,
 * then learns and writes every channel group on the GPU (fuif_encode_channels, encoding.cpp:74-207: MANIAC tree
 * learning, pruning, tree + sample coding; one warp per group, the groups concurrently) and assembles the container
 * on the host.  The file is the reference encoder's byte for byte, including the row order its learning pass draws
 * from libc rand() in a fresh process.  *bytes is malloc'ed: release it with fb_free().
 * group_index / group_first (may be NULL; at most cap entries are written, *n_groups gets the count): the sidecar
 * index fb_decode() takes -- the encoder knows it for free.  opts == NULL: the reference defaults. */
FB_API int fb_encode(fb_ctx *ctx, fb_image *img, const fb_encode_options *opts, uint8_t **bytes, size_t *nbytes,
              int64_t *group_index, int32_t *group_first, int cap, int *n_groups);
FB_API void fb_free(void *p);

/* ---- one-call convenience: file bytes in host memory -> pixels in host memory ------------------------------ */

/* fuif_decode + undo_transforms + interleave, the path `fuif -d in.fuif out.ppm` takes (reference
 * fuif.cpp:206-239).  dst must hold w*h*nb_channels samples of bytes_per_sample bytes each; query the sizes
 * with fb_peek_header() first. */
FB_API int fb_decode_to_pixels(fb_ctx *ctx, const uint8_t *bytes, size_t nbytes, const fb_decode_options *opts,
                        const int64_t *group_index, const int32_t *group_first, int n_groups, int bytes_per_sample,
                        void *dst, size_t dst_bytes);

/* Header-only parse (reference encoding.cpp:599-637, "identify"): fills w, h, maxval, nb_channels. */
FB_API int fb_peek_header(const uint8_t *bytes, size_t nbytes, fb_image_info *info);

#ifdef __cplusplus
}
#endif
#endif
