/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the FUIF hot path (see fuif_oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it. */
#ifndef FUIF_ORACLE_H
#define FUIF_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* transform ids: reference transform/transform.h:30-70 */
enum { FO_YCBCR = 0, FO_YCOCG = 1, FO_SUBSAMPLE = 3, FO_DCT = 4, FO_QUANTIZE = 5, FO_PALETTE = 6, FO_SQUEEZE = 7 };

typedef struct {            /* mirrors class Channel, reference image/image.h:54-91 */
    int16_t *data;
    size_t n;               /* data.size(): 0 = not decoded */
    int w, h;
    int minval, maxval, zero;
    int q;
    int hshift, vshift, hcshift, vcshift;
    int component;
} fo_channel;

typedef struct {            /* mirrors class Transform, reference transform/transform.h:77-106 */
    int id;
    int np;
    int *p;
} fo_transform;

typedef struct {            /* mirrors class Image, reference image/image.h:98-129 */
    fo_channel *ch;
    int nch, cap;
    fo_transform *tr;
    int ntr;
    int w, h, minval, maxval;
    int nb_channels, real_nb_channels, nb_meta_channels, colormodel;
    int error;
} fo_image;

fo_image *fo_image_new(int w, int h, int maxval, int nb_channels, int colormodel);
void fo_image_free(fo_image *img);
fo_image *fo_image_clone(const fo_image *img);

/* accessors for ctypes */
int fo_nplanes(const fo_image *img);
int fo_ntransforms(const fo_image *img);
/* out[12] = w h minval maxval zero q hshift vshift hcshift vcshift component nsamples */
void fo_plane_info(const fo_image *img, int i, long long *out);
int16_t *fo_plane_data(const fo_image *img, int i);
/* out[8] = w h minval maxval nb_channels real_nb_channels nb_meta_channels colormodel */
void fo_image_info(const fo_image *img, int *out);
int fo_transform_info(const fo_image *img, int i, int *id, int *params, int maxparams);
/* overwrite plane i (n must equal w*h of the plane, or 0 to mark it undecoded) */
int fo_plane_set(fo_image *img, int i, const int16_t *data, size_t n);
int fo_plane_set_range(fo_image *img, int i, int minval, int maxval, int q);
int fo_plane_reshape(fo_image *img, int i, int w, int h, int hshift, int vshift, int hcshift, int vcshift, int component);
int fo_push_transform(fo_image *img, int id, const int *params, int np);

/* fuif_decode over a memory blob (reference encoding/encoding.cpp:599-720) with the end-of-stream rules of
 * FileIO (fileio.h:33-81), the IO class fuif_decode_file uses.
 * preview = -1 for everything, 0..4 = responsive truncation point.
 * group_offsets (may be NULL) receives, for every channel group decoded, the byte offset of its header and
 * the first channel index: pairs (offset, beginc); *ngroups in: capacity (pairs), out: count.
 * Returns NULL on error. */
fo_image *fo_decode(const uint8_t *bytes, size_t n, int preview, int maniac_cutoff, int maniac_alpha,
                    long long *group_offsets, int *ngroups);

/* Image::undo_transforms(keep) (reference image/image.cpp:94-115). returns 0 on success */
int fo_undo_transforms(fo_image *img, int keep);
/* Image::do_transform (reference image/image.cpp:117-122). returns 1 if applied, 0 if not */
int fo_do_transform(fo_image *img, int id, const int *params, int np);
/* Image::recompute_minmax (image/image.h:127) */
void fo_recompute_minmax(fo_image *img);

#ifdef __cplusplus
}
#endif

/* ---- encoder (test infrastructure for the encode-side rows; reference encoding/encoding.cpp:455-573, 74-207) ----
 * fuif_prepare_encode + fuif_encode of `img` (transformed planes + transform list) with the reference's two-pass MANIAC
 * tree learning.  Mutates channel ranges / zero like the reference does.  Returns a malloc'ed buffer (fo_free). */
typedef struct {
    float nb_repeats;           /* 0.5 */
    int max_properties;         /* 12 */
    int maniac_cutoff;          /* 6 */
    int maniac_alpha;           /* 0x0d000000 */
    int compress;               /* 1 */
    int max_group;              /* -1 */
    int npred;
    const int *predictor;
} fo_enc_options;
uint8_t *fo_encode(fo_image *img, const fo_enc_options *opt, size_t *nbytes);
void fo_free(void *p);

#endif
