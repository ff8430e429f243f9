/* TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT PATH.
 *
 * Plain-C, single-threaded restatement of the FUIF decode hot path and of the forward / inverse
 * transform chain of the reference (cloudinary/fuif @ 49ff10b5).  It exists so that the CUDA path can
 * be checked bit-for-bit on machines where /root/reference is absent (the GPU box).  It is itself
 * pinned against the real reference: tests/test_oracle_vs_ref.py compares it with oracle/_ref/ref_driver
 * (the unmodified reference compiled by oracle/Makefile) and with the golden fixtures under tests/golden/
 * that were produced by that binary (tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it restates.  All arithmetic follows the C++ semantics of
 * the reference exactly, in particular pixel_type == int16_t (image/image.h:35): every value assigned to
 * a pixel_type local wraps to 16 bits at that point (S16 below).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may use this file.
 */
#include "fuif_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define S16(x) ((int)(int16_t)(x))
#define MAX_BIT_DEPTH 15          /* config.h:5 */
#define LARGEST_VAL 0x7FFF        /* image/image.h:36 */
#define SMALLEST_VAL (-0x7FFF)    /* image/image.h:37 (0x8001 as int16) */
#define NB_NONREF 13              /* context_predict.h:210 */

/* ------------------------------------------------------------------------------------------------ */
/* Image / Channel containers                                                                        */
/* ------------------------------------------------------------------------------------------------ */

static void ch_init(fo_channel *c) {            /* Channel::Channel(), image/image.h:68 */
    memset(c, 0, sizeof(*c));
    c->q = 1;
    c->component = -1;
}

static void ch_setzero(fo_channel *c) {         /* Channel::setzero, image/image.h:70-74 */
    if (c->minval > 0) c->zero = c->minval;
    else if (c->maxval < 0) c->zero = c->maxval;
    else c->zero = 0;
}

static void ch_resize(fo_channel *c) {          /* Channel::resize, image/image.h:75-77: data.resize(w*h, zero) */
    size_t want = (size_t)(c->w > 0 && c->h > 0 ? (size_t)c->w * c->h : 0);
    int16_t *d = (int16_t *)malloc(want * sizeof(int16_t) + 2);
    size_t keep = c->n < want ? c->n : want;
    if (keep) memcpy(d, c->data, keep * sizeof(int16_t));
    for (size_t i = keep; i < want; i++) d[i] = (int16_t)c->zero;
    free(c->data);
    c->data = d;
    c->n = want;
}

static void ch_fill(fo_channel *c, int v) {
    size_t want = (size_t)(c->w > 0 && c->h > 0 ? (size_t)c->w * c->h : 0);
    free(c->data);
    c->data = (int16_t *)malloc(want * sizeof(int16_t) + 2);
    for (size_t i = 0; i < want; i++) c->data[i] = (int16_t)v;
    c->n = want;
}

static void ch_make(fo_channel *c, int w, int h, int minval, int maxval, int q, int hs, int vs, int hcs, int vcs) {
    /* Channel(iw,ih,min,max,q,hsh,vsh,hcsh,vcsh), image/image.h:66-67: data(iw*ih, 0) */
    ch_init(c);
    c->w = w; c->h = h; c->minval = S16(minval); c->maxval = S16(maxval); c->q = q;
    c->hshift = hs; c->vshift = vs; c->hcshift = hcs; c->vcshift = vcs;
    ch_setzero(c);
    c->n = 0;
    {   size_t want = (size_t)(w > 0 && h > 0 ? (size_t)w * h : 0);
        c->data = (int16_t *)calloc(want + 1, sizeof(int16_t));
        c->n = want; }
}

/* Channel::value(r,c) read accessor, image/image.h:82: out-of-range reads return 'zero' */
static inline int ch_get(const fo_channel *c, int r, int col) {
    size_t idx = (size_t)((long long)r * c->w + col);
    if (idx >= c->n) return c->zero;
    return c->data[idx];
}
/* Channel::value(r,c) write accessor, image/image.h:84: out-of-range writes land in 'zero' (dropped here) */
static inline void ch_set(fo_channel *c, int r, int col, int v) {
    size_t idx = (size_t)((long long)r * c->w + col);
    if (idx >= c->n) { c->zero = S16(v); return; }
    c->data[idx] = (int16_t)v;
}

static void img_reserve(fo_image *img, int n) {
    if (n <= img->cap) return;
    int cap = img->cap ? img->cap : 16;
    while (cap < n) cap *= 2;
    img->ch = (fo_channel *)realloc(img->ch, (size_t)cap * sizeof(fo_channel));
    img->cap = cap;
}
static void img_insert(fo_image *img, int pos, const fo_channel *c) {   /* takes ownership of c->data */
    img_reserve(img, img->nch + 1);
    memmove(&img->ch[pos + 1], &img->ch[pos], (size_t)(img->nch - pos) * sizeof(fo_channel));
    img->ch[pos] = *c;
    img->nch++;
}
static void img_erase(fo_image *img, int from, int to) {                /* erase [from,to) */
    for (int i = from; i < to; i++) free(img->ch[i].data);
    memmove(&img->ch[from], &img->ch[to], (size_t)(img->nch - to) * sizeof(fo_channel));
    img->nch -= (to - from);
}

fo_image *fo_image_new(int w, int h, int maxval, int nb_channels, int colormodel) {
    /* Image(iw,ih,maxval,nb_chans,cm), image/image.h:114-120 */
    fo_image *img = (fo_image *)calloc(1, sizeof(fo_image));
    img->w = w; img->h = h; img->minval = 0; img->maxval = maxval;
    img->nb_channels = nb_channels; img->real_nb_channels = nb_channels; img->colormodel = colormodel;
    img_reserve(img, nb_channels > 0 ? nb_channels : 1);
    for (int i = 0; i < nb_channels; i++) {
        ch_make(&img->ch[i], w, h, 0, maxval, 1, 0, 0, 0, 0);
        img->ch[i].component = i;
    }
    img->nch = nb_channels > 0 ? nb_channels : 0;
    return img;
}

void fo_image_free(fo_image *img) {
    if (!img) return;
    for (int i = 0; i < img->nch; i++) free(img->ch[i].data);
    free(img->ch);
    for (int i = 0; i < img->ntr; i++) free(img->tr[i].p);
    free(img->tr);
    free(img);
}

fo_image *fo_image_clone(const fo_image *src) {
    fo_image *img = (fo_image *)calloc(1, sizeof(fo_image));
    *img = *src;
    img->ch = NULL; img->cap = 0; img->tr = NULL;
    img_reserve(img, src->nch > 0 ? src->nch : 1);
    for (int i = 0; i < src->nch; i++) {
        img->ch[i] = src->ch[i];
        img->ch[i].data = (int16_t *)malloc(src->ch[i].n * sizeof(int16_t) + 2);
        if (src->ch[i].n) memcpy(img->ch[i].data, src->ch[i].data, src->ch[i].n * sizeof(int16_t));
    }
    img->tr = (fo_transform *)calloc((size_t)(src->ntr > 0 ? src->ntr : 1), sizeof(fo_transform));
    for (int i = 0; i < src->ntr; i++) {
        img->tr[i] = src->tr[i];
        img->tr[i].p = (int *)malloc(sizeof(int) * (size_t)(src->tr[i].np > 0 ? src->tr[i].np : 1));
        memcpy(img->tr[i].p, src->tr[i].p, sizeof(int) * (size_t)src->tr[i].np);
    }
    return img;
}

static void img_push_transform(fo_image *img, int id, const int *p, int np) {
    img->tr = (fo_transform *)realloc(img->tr, sizeof(fo_transform) * (size_t)(img->ntr + 1));
    img->tr[img->ntr].id = id;
    img->tr[img->ntr].np = np;
    img->tr[img->ntr].p = (int *)malloc(sizeof(int) * (size_t)(np > 0 ? np : 1));
    if (np) memcpy(img->tr[img->ntr].p, p, sizeof(int) * (size_t)np);
    img->ntr++;
}

int fo_nplanes(const fo_image *img) { return img->nch; }
int fo_ntransforms(const fo_image *img) { return img->ntr; }
void fo_plane_info(const fo_image *img, int i, long long *o) {
    const fo_channel *c = &img->ch[i];
    o[0] = c->w; o[1] = c->h; o[2] = c->minval; o[3] = c->maxval; o[4] = c->zero; o[5] = c->q;
    o[6] = c->hshift; o[7] = c->vshift; o[8] = c->hcshift; o[9] = c->vcshift; o[10] = c->component; o[11] = (long long)c->n;
}
int16_t *fo_plane_data(const fo_image *img, int i) { return img->ch[i].data; }
void fo_image_info(const fo_image *img, int *o) {
    o[0] = img->w; o[1] = img->h; o[2] = img->minval; o[3] = img->maxval; o[4] = img->nb_channels;
    o[5] = img->real_nb_channels; o[6] = img->nb_meta_channels; o[7] = img->colormodel;
}
int fo_transform_info(const fo_image *img, int i, int *id, int *params, int maxparams) {
    *id = img->tr[i].id;
    int n = img->tr[i].np < maxparams ? img->tr[i].np : maxparams;
    for (int k = 0; k < n; k++) params[k] = img->tr[i].p[k];
    return img->tr[i].np;
}
int fo_plane_set(fo_image *img, int i, const int16_t *data, size_t n) {
    fo_channel *c = &img->ch[i];
    if (n != 0 && n != (size_t)c->w * c->h) return -1;
    free(c->data);
    c->data = (int16_t *)malloc(n * sizeof(int16_t) + 2);
    if (n) memcpy(c->data, data, n * sizeof(int16_t));
    c->n = n;
    return 0;
}
int fo_plane_set_range(fo_image *img, int i, int minval, int maxval, int q) {
    img->ch[i].minval = S16(minval); img->ch[i].maxval = S16(maxval); img->ch[i].q = q;
    ch_setzero(&img->ch[i]);
    return 0;
}

/* test plumbing: give plane i new dimensions / shifts (zero-filled), and push a transform without applying it -- the state
 * an Image is in when its channels were produced elsewhere (e.g. chroma planes of a JPEG) */
int fo_plane_reshape(fo_image *img, int i, int w, int h, int hshift, int vshift, int hcshift, int vcshift, int component) {
    if (i < 0 || i >= img->nch) return -1;
    fo_channel *c = &img->ch[i];
    c->w = w; c->h = h; c->hshift = hshift; c->vshift = vshift; c->hcshift = hcshift; c->vcshift = vcshift; c->component = component;
    ch_fill(c, 0);
    return 0;
}
int fo_push_transform(fo_image *img, int id, const int *params, int np) {
    img_push_transform(img, id, params, np);
    return 0;
}

void fo_recompute_minmax(fo_image *img) {       /* Channel::actual_minmax, image/image.cpp:82-92 */
    for (int i = 0; i < img->nch; i++) {
        int mn = LARGEST_VAL, mx = SMALLEST_VAL;
        for (size_t k = 0; k < img->ch[i].n; k++) {
            int v = img->ch[i].data[k];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        img->ch[i].minval = mn; img->ch[i].maxval = mx;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* Squeeze (transform/squeeze.h)                                                                     */
/* ------------------------------------------------------------------------------------------------ */

/* smooth_tendency, transform/squeeze.h:61-77 */
static inline int smooth_tendency(int B, int a, int n) {
    int diff = 0;
    if (B >= a && a >= n) {
        diff = S16((4 * B - 3 * n - a + 6) / 12);
        if (diff - (diff & 1) > 2 * (B - a)) diff = S16(2 * (B - a) + 1);
        if (diff + (diff & 1) > 2 * (a - n)) diff = S16(2 * (a - n));
    } else if (B <= a && a <= n) {
        diff = S16((4 * B - 3 * n - a - 6) / 12);
        if (diff + (diff & 1) < 2 * (B - a)) diff = S16(2 * (B - a) - 1);
        if (diff - (diff & 1) < 2 * (a - n)) diff = S16(2 * (a - n));
    }
    return diff;
}

/* the A/B reconstruction shared by inv_hsqueeze / inv_vsqueeze, squeeze.h:93-94 */
#define UNSQ_AB(avg, diff, A, B) do { \
        A = S16((((avg) << 1) + (diff) + ((diff) > 0 ? -((diff) & 1) : ((diff) & 1))) >> 1); \
        B = S16((A) - (diff)); } while (0)

/* inv_hsqueeze, transform/squeeze.h:81-132 */
static void inv_hsqueeze(fo_image *img, int c, int rc) {
    fo_channel *chin = &img->ch[c];
    fo_channel *res = &img->ch[rc];
    fo_channel out;
    ch_make(&out, chin->w + res->w, chin->h, chin->minval, chin->maxval, chin->q, chin->hshift - 1, chin->vshift, chin->hcshift - 1, chin->vcshift);
    out.component = chin->component;
    for (int y = 0; y < chin->h; y++) {
        int avg = chin->data[(size_t)y * chin->w + 0];
        int next_avg = (1 < chin->w ? chin->data[(size_t)y * chin->w + 1] : avg);
        int tendency = smooth_tendency(avg, avg, next_avg);
        int diff = S16(ch_get(res, y, 0) + tendency);
        int A, B;
        UNSQ_AB(avg, diff, A, B);
        ch_set(&out, y, 0, A);
        ch_set(&out, y, 1, B);
        for (int x = 1; x < res->w; x++) {
            int dmt = ch_get(res, y, x);
            avg = chin->data[(size_t)y * chin->w + x];
            next_avg = (x + 1 < chin->w ? chin->data[(size_t)y * chin->w + x + 1] : avg);
            int left = out.data[(size_t)y * out.w + (x << 1) - 1];
            tendency = smooth_tendency(left, avg, next_avg);
            diff = S16(dmt + tendency);
            UNSQ_AB(avg, diff, A, B);
            ch_set(&out, y, x << 1, A);
            ch_set(&out, y, (x << 1) + 1, B);
        }
        if (out.w & 1) ch_set(&out, y, out.w - 1, ch_get(chin, y, chin->w - 1));
    }
    free(chin->data);
    *chin = out;
}

/* inv_vsqueeze, transform/squeeze.h:173-224 */
static void inv_vsqueeze(fo_image *img, int c, int rc) {
    fo_channel *chin = &img->ch[c];
    fo_channel *res = &img->ch[rc];
    fo_channel out;
    ch_make(&out, chin->w, chin->h + res->h, chin->minval, chin->maxval, chin->q, chin->hshift, chin->vshift - 1, chin->hcshift, chin->vcshift - 1);
    out.component = chin->component;
    for (int x = 0; x < chin->w; x++) {
        int dmt = ch_get(res, 0, x);
        int avg = chin->data[x];
        int next_avg = avg;
        if (1 < chin->h) next_avg = chin->data[(size_t)chin->w + x];
        int tendency = smooth_tendency(avg, avg, next_avg);
        int diff = S16(dmt + tendency);
        int A, B;
        UNSQ_AB(avg, diff, A, B);
        ch_set(&out, 0, x, A);
        ch_set(&out, 1, x, B);
    }
    for (int y = 1; y < res->h; y++) {
        for (int x = 0; x < chin->w; x++) {
            int dmt = ch_get(res, y, x);
            int avg = chin->data[(size_t)y * chin->w + x];
            int next_avg = avg;
            if (y + 1 < chin->h) next_avg = chin->data[(size_t)(y + 1) * chin->w + x];
            int top = out.data[(size_t)((y << 1) - 1) * out.w + x];
            int tendency = smooth_tendency(top, avg, next_avg);
            int diff = S16(dmt + tendency);
            int A, B;
            UNSQ_AB(avg, diff, A, B);
            ch_set(&out, y << 1, x, A);
            ch_set(&out, (y << 1) + 1, x, B);
        }
    }
    if (out.h & 1) {
        int y = chin->h - 1;
        for (int x = 0; x < chin->w; x++) ch_set(&out, y << 1, x, ch_get(chin, y, x));
    }
    free(chin->data);
    *chin = out;
}

/* fwd_hsqueeze, transform/squeeze.h:135-170 */
static void fwd_hsqueeze(fo_image *img, int c, int rc) {
    fo_channel *chin = &img->ch[c];
    fo_channel out, res;
    ch_make(&out, (chin->w + 1) / 2, chin->h, chin->minval, chin->maxval, chin->q, chin->hshift + 1, chin->vshift, chin->hcshift + 1, chin->vcshift);
    ch_make(&res, chin->w - out.w, out.h, out.minval - out.maxval, out.maxval - out.minval, 1, chin->hshift + 1, chin->vshift, chin->hcshift, chin->vcshift);
    out.component = chin->component;
    res.component = chin->component;
    for (int y = 0; y < out.h; y++) {
        for (int x = 0; x < res.w; x++) {
            int A = ch_get(chin, y, x * 2), B = ch_get(chin, y, x * 2 + 1);
            int avg = S16((A + B + (A > B)) >> 1);
            ch_set(&out, y, x, avg);
            int diff = S16(A - B);
            int next_avg = avg;
            if (x + 1 < res.w) next_avg = S16((ch_get(chin, y, x * 2 + 2) + ch_get(chin, y, x * 2 + 3) + (ch_get(chin, y, x * 2 + 2) > ch_get(chin, y, x * 2 + 3))) >> 1);
            else if (chin->w & 1) next_avg = ch_get(chin, y, x * 2 + 2);
            int left = (x > 0 ? ch_get(chin, y, x * 2 - 1) : avg);
            int tendency = smooth_tendency(left, avg, next_avg);
            ch_set(&res, y, x, S16(diff - tendency));
        }
        if (chin->w & 1) {
            int x = out.w - 1;
            ch_set(&out, y, x, ch_get(chin, y, x * 2));
        }
    }
    free(chin->data);
    *chin = out;
    img_insert(img, rc, &res);
}

/* fwd_vsqueeze, transform/squeeze.h:227-263 */
static void fwd_vsqueeze(fo_image *img, int c, int rc) {
    fo_channel *chin = &img->ch[c];
    fo_channel out, res;
    ch_make(&out, chin->w, (chin->h + 1) / 2, chin->minval, chin->maxval, chin->q, chin->hshift, chin->vshift + 1, chin->hcshift, chin->vcshift + 1);
    ch_make(&res, chin->w, chin->h - out.h, out.minval - out.maxval, out.maxval - out.minval, 1, chin->hshift, chin->vshift + 1, chin->hcshift, chin->vcshift);
    out.component = chin->component;
    res.component = chin->component;
    for (int y = 0; y < res.h; y++) {
        for (int x = 0; x < out.w; x++) {
            int A = ch_get(chin, y * 2, x), B = ch_get(chin, y * 2 + 1, x);
            int avg = S16((A + B + (A > B)) >> 1);
            ch_set(&out, y, x, avg);
            int diff = S16(A - B);
            int next_avg = avg;
            if (y + 1 < res.h) next_avg = S16((ch_get(chin, y * 2 + 2, x) + ch_get(chin, y * 2 + 3, x) + (ch_get(chin, y * 2 + 2, x) > ch_get(chin, y * 2 + 3, x))) >> 1);
            else if (chin->h & 1) next_avg = ch_get(chin, y * 2 + 2, x);
            int top = (y > 0 ? ch_get(chin, y * 2 - 1, x) : avg);
            int tendency = smooth_tendency(top, avg, next_avg);
            ch_set(&res, y, x, S16(diff - tendency));
        }
    }
    if (chin->h & 1) {
        int y = out.h - 1;
        for (int x = 0; x < out.w; x++) ch_set(&out, y, x, ch_get(chin, y * 2, x));
    }
    free(chin->data);
    *chin = out;
    img_insert(img, rc, &res);
}

/* default_squeeze_parameters, transform/squeeze.h:266-321 (MAX_FIRST_PREVIEW_SIZE 8, config.h:41) */
static int default_squeeze_parameters(const fo_image *img, int *p) {
    int n = 0;
    int nb = img->nb_channels, m = img->nb_meta_channels;
    int w = img->ch[m].w, h = img->ch[m].h;
    int wide = (w > h);
    if (nb > 2 && img->ch[m + 1].w == w && img->ch[m + 1].h == h) {
        p[n++] = 3; p[n++] = m + 1; p[n++] = m + 2;
        p[n++] = 2; p[n++] = m + 1; p[n++] = m + 2;
    }
    if (!wide) {
        if (h > 8) { p[n++] = 0; p[n++] = m; p[n++] = m + nb - 1; h = (h + 1) / 2; }
    }
    while (w > 8 || h > 8) {
        if (w > 8) { p[n++] = 1; p[n++] = m; p[n++] = m + nb - 1; w = (w + 1) / 2; }
        if (h > 8) { p[n++] = 0; p[n++] = m; p[n++] = m + nb - 1; h = (h + 1) / 2; }
    }
    return n;
}

/* meta_squeeze, transform/squeeze.h:323-360 */
static void meta_squeeze(fo_image *img, const int *p, int np) {
    for (int i = 0; i + 2 < np; i += 3) {
        int horizontal = p[i] & 1, in_place = !(p[i] & 2);
        int beginc = p[i + 1], endc = p[i + 2];
        int offset = in_place ? endc + 1 : img->nb_meta_channels + img->nb_channels;
        for (int c = beginc; c <= endc; c++) {
            fo_channel d;
            ch_init(&d);
            d.hcshift = img->ch[c].hcshift; d.vcshift = img->ch[c].vcshift; d.component = img->ch[c].component;
            if (horizontal) {
                int w = img->ch[c].w;
                img->ch[c].w = (w + 1) / 2; img->ch[c].hshift++; img->ch[c].hcshift++;
                d.w = w - (w + 1) / 2; d.h = img->ch[c].h;
            } else {
                int h = img->ch[c].h;
                img->ch[c].h = (h + 1) / 2; img->ch[c].vshift++; img->ch[c].vcshift++;
                d.h = h - (h + 1) / 2; d.w = img->ch[c].w;
            }
            d.hshift = img->ch[c].hshift; d.vshift = img->ch[c].vshift;
            img_insert(img, offset + c - beginc, &d);
        }
    }
}

/* squeeze(), transform/squeeze.h:363-408 */
static int squeeze(fo_image *img, int inverse, const int *params, int np) {
    int adj[256];
    int n = np;
    if (np > 256) return 0;
    memcpy(adj, params, sizeof(int) * (size_t)np);
    if (!n) n = default_squeeze_parameters(img, adj);
    if (inverse) {
        for (int i = n - 3; i >= 0; i -= 3) {
            int horizontal = adj[i] & 1, in_place = !(adj[i] & 2);
            int beginc = adj[i + 1], endc = adj[i + 2];
            int offset = in_place ? endc + 1 : img->nb_meta_channels + img->nb_channels;
            if (offset + endc - beginc >= img->nch) return 0;
            for (int c = beginc; c <= endc; c++) {
                if (img->ch[offset + c - beginc].n == 0) ch_resize(&img->ch[offset + c - beginc]);
                if (horizontal) inv_hsqueeze(img, c, offset + c - beginc);
                else inv_vsqueeze(img, c, offset + c - beginc);
            }
            img_erase(img, offset, offset + (endc - beginc + 1));
        }
    } else {
        for (int i = 0; i + 2 < n; i += 3) {
            int horizontal = adj[i] & 1, in_place = !(adj[i] & 2);
            int beginc = adj[i + 1], endc = adj[i + 2];
            int offset = in_place ? endc + 1 : img->nb_meta_channels + img->nb_channels;
            for (int c = beginc; c <= endc; c++) {
                if (horizontal) fwd_hsqueeze(img, c, offset + c - beginc);
                else fwd_vsqueeze(img, c, offset + c - beginc);
            }
        }
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* Colour transforms and quantisation                                                                */
/* ------------------------------------------------------------------------------------------------ */

#define CLAMPI(x, l, u) ((x) < (l) ? (l) : ((x) > (u) ? (u) : (x)))

/* inv_YCoCg / fwd_YCoCg, transform/ycocg.h:33-63 / 65-95 */
static int ycocg(fo_image *img, int inverse) {
    int m = img->nb_meta_channels;
    if (img->nb_channels < 3) return 0;
    int w = img->ch[m].w, h = img->ch[m].h;
    if (img->ch[m + 1].w < w || img->ch[m + 1].h < h || img->ch[m + 2].w < w || img->ch[m + 2].h < h) return 0;
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
        if (inverse) {
            int Y = CLAMPI(ch_get(&img->ch[m], y, x), 0, img->maxval);
            int Co = ch_get(&img->ch[m + 1], y, x);
            int Cg = ch_get(&img->ch[m + 2], y, x);
            int G = CLAMPI(Y - ((-Cg) >> 1), 0, img->maxval);
            int B = CLAMPI(Y + ((1 - Cg) >> 1) - (Co >> 1), 0, img->maxval);
            int R = CLAMPI(Co + B, 0, img->maxval);
            ch_set(&img->ch[m], y, x, R); ch_set(&img->ch[m + 1], y, x, G); ch_set(&img->ch[m + 2], y, x, B);
        } else {
            int R = ch_get(&img->ch[m], y, x), G = ch_get(&img->ch[m + 1], y, x), B = ch_get(&img->ch[m + 2], y, x);
            int Y = (((R + B) >> 1) + G) >> 1;
            int Co = R - B;
            int Cg = G - ((R + B) >> 1);
            ch_set(&img->ch[m], y, x, Y); ch_set(&img->ch[m + 1], y, x, Co); ch_set(&img->ch[m + 2], y, x, Cg);
        }
    }
    return 1;
}

#define CLAMPD(x, l, u) ((x) < (l) ? (double)(l) : ((x) > (u) ? (double)(u) : (x)))

/* inv_YCbCr / fwd_YCbCr, transform/ycbcr.h:33-63 / 65-95 : float loads, double math, truncating store */
static int ycbcr(fo_image *img, int inverse) {
    if (img->nch < 3) return 0;
    int w = img->ch[0].w, h = img->ch[0].h;
    if (img->ch[1].w < w || img->ch[1].h < h || img->ch[2].w < w || img->ch[2].h < h) return 0;
    float half = (float)((img->maxval + 1) / 2);
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
        if (inverse) {
            float yy = (float)ch_get(&img->ch[0], y, x);
            float cb = (float)ch_get(&img->ch[1], y, x) - half;
            float cr = (float)ch_get(&img->ch[2], y, x) - half;
            double r = yy + 1.402 * cr + 0.5;
            double g = yy - 0.344136 * cb - 0.714136 * cr + 0.5;
            double b = yy + 1.772 * cb + 0.5;
            ch_set(&img->ch[0], y, x, (int16_t)CLAMPD(r, img->minval, img->maxval));
            ch_set(&img->ch[1], y, x, (int16_t)CLAMPD(g, img->minval, img->maxval));
            ch_set(&img->ch[2], y, x, (int16_t)CLAMPD(b, img->minval, img->maxval));
        } else {
            float r = (float)ch_get(&img->ch[0], y, x), g = (float)ch_get(&img->ch[1], y, x), b = (float)ch_get(&img->ch[2], y, x);
            double yy = 0.299 * r + 0.587 * g + 0.114 * b;
            double cb = half - 0.168736 * r - 0.331264 * g + 0.5 * b;
            double cr = half + 0.5 * r - 0.418688 * g - 0.081312 * b;
            ch_set(&img->ch[0], y, x, (int16_t)CLAMPD(yy, img->minval, img->maxval));
            ch_set(&img->ch[1], y, x, (int16_t)CLAMPD(cb, img->minval, img->maxval));
            ch_set(&img->ch[2], y, x, (int16_t)CLAMPD(cr, img->minval, img->maxval));
        }
    }
    return 1;
}

/* inv_quantize / fwd_quantize, transform/quantize.h:32-49 / 56-71 */
static int quantize(fo_image *img, int inverse, const int *p, int np) {
    for (int c = img->nb_meta_channels; c < img->nch; c++) {
        fo_channel *ch = &img->ch[c];
        if (inverse) {
            if (ch->n == 0) continue;
            int q = ch->q;
            if (q == 1) continue;
            for (int y = 0; y < ch->h; y++) for (int x = 0; x < ch->w; x++) ch_set(ch, y, x, S16(ch_get(ch, y, x) * q));
            ch->minval = S16(ch->minval * q); ch->maxval = S16(ch->maxval * q); ch->q = 1;
        } else {
            if (np <= 0) return 0;
            int q = (c < np ? p[c] : p[np - 1]);
            for (int y = 0; y < ch->h; y++) for (int x = 0; x < ch->w; x++) ch_set(ch, y, x, S16(ch_get(ch, y, x) / q));
            ch->minval = S16(ch->minval / q); ch->maxval = S16(ch->maxval / q); ch->q = q;
        }
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* DCT (transform/dct.h)                                                                             */
/* ------------------------------------------------------------------------------------------------ */

static double kDCT[64];       /* kDCTMatrix, dct.h:60-77: 0.5*alpha(u)*cos((2x+1)u*pi/16) rounded to 10 decimals */
static int zigzag[64];        /* the reference's own scan variant, dct.h:120-130 */
static int dct_tables_ready = 0;

static void dct_tables(void) {
    if (dct_tables_ready) return;
    /* the seven magnitudes of the 8-point DCT-II basis, to the 10 decimals the reference uses */
    static const double mag[8] = {0.3535533906, 0.4903926402, 0.4619397663, 0.4157348062,
                                  0.3535533906, 0.2777851165, 0.1913417162, 0.0975451610};
    for (int u = 0; u < 8; u++) for (int x = 0; x < 8; x++) {
        /* cos((2x+1)u*pi/16) = +-cos(k*pi/16) with k = (2x+1)u mod 32 folded into 0..8 */
        int k = ((2 * x + 1) * u) % 32;
        int sign = 1;
        if (k > 16) k = 32 - k;
        if (k > 8) { k = 16 - k; sign = -1; }
        double v = (u == 0) ? mag[0] : (k == 8 ? 0.0 : mag[k]);
        kDCT[8 * u + x] = sign * v;
    }
    /* scan order: L-shaped shells max(r,c)=s holding indices s*s .. s*s+2s; even shells run down column s
       then left along row s, odd shells run right along row s then up column s; shell 1 is irregular. */
    for (int s = 0; s < 8; s++) {
        int idx = s * s;
        if (s == 0) { zigzag[0] = 0; continue; }
        if (s == 1) { zigzag[0 * 8 + 1] = 1; zigzag[1 * 8 + 0] = 2; zigzag[1 * 8 + 1] = 3; continue; }
        if (s % 2 == 0) {
            for (int r = 0; r <= s; r++) zigzag[r * 8 + s] = idx++;
            for (int c = s - 1; c >= 0; c--) zigzag[s * 8 + c] = idx++;
        } else {
            for (int c = 0; c <= s; c++) zigzag[s * 8 + c] = idx++;
            for (int r = s - 1; r >= 0; r--) zigzag[r * 8 + s] = idx++;
        }
    }
    dct_tables_ready = 1;
}

static int dct_cshift(int k) { return k == 0 ? 3 : (k < 4 ? 2 : (k < 16 ? 1 : 0)); }   /* dct_cshifts, dct.h:159-171 */

/* DCT1d / IDCT1d / TransformBlock, dct.h:79-107: strict "out = 0.0; out += k*in" order, no FMA */
static void dct1d(const double *in, int stride, double *out, int inverse) {
    for (int x = 0; x < 8; ++x) {
        double acc = 0.0;
        for (int u = 0; u < 8; ++u) {
            double k = inverse ? kDCT[8 * u + x] : kDCT[8 * x + u];
            double prod = k * in[u * stride];
            acc = acc + prod;
        }
        out[x * stride] = acc;
    }
}
static void transform_block(double block[64], int inverse) {
    double tmp[64];
    for (int x = 0; x < 8; ++x) dct1d(&block[x], 8, &tmp[x], inverse);
    for (int y = 0; y < 8; ++y) dct1d(&tmp[8 * y], 1, &block[8 * y], inverse);
}

/* default_DCT_scanscript, dct.h:173-207: position p = coeff*nb + comp */
/* meta_DCT, dct.h:215-246 */
static int meta_dct(fo_image *img, const int *p) {
    int beginc = img->nb_meta_channels + p[0], endc = img->nb_meta_channels + p[1];
    int nb = endc - beginc + 1;
    if (nb < 1 || endc >= img->nch) return 0;
    for (int c = beginc; c <= endc; c++) {
        img->ch[c].w = (img->ch[c].w + 7) / 8; img->ch[c].h = (img->ch[c].h + 7) / 8;
        img->ch[c].hshift += 3; img->ch[c].vshift += 3; img->ch[c].hcshift += 3; img->ch[c].vcshift += 3;
    }
    for (int i = nb; i < 64 * nb; i++) {
        fo_channel d;
        ch_init(&d);
        int comp = i % nb, coeff = i / nb;
        int c = beginc + comp;
        d.w = img->ch[c].w; d.h = img->ch[c].h; d.hshift = img->ch[c].hshift; d.vshift = img->ch[c].vshift;
        d.hcshift = dct_cshift(coeff) + img->ch[c].hcshift - 3;
        d.vcshift = dct_cshift(coeff) + img->ch[c].vcshift - 3;
        d.component = img->ch[c].component;
        img_insert(img, img->nch, &d);
    }
    return 1;
}

/* inv_DCT, dct.h:249-296 */
static int inv_dct(fo_image *img, const int *p) {
    dct_tables();
    int beginc = img->nb_meta_channels + p[0], endc = img->nb_meta_channels + p[1];
    int nb = endc - beginc + 1;
    int offset = img->nch - 63 * nb;
    if (offset <= endc) return 0;
    for (int c = beginc; c <= endc; c++) {
        int bw = img->ch[c - beginc + offset].w, bh = img->ch[c - beginc + offset].h;
        if (img->ch[c].w < bw) bw = img->ch[c].w;
        if (img->ch[c].h < bh) bh = img->ch[c].h;
        fo_channel out;
        ch_make(&out, bw * 8, bh * 8, 0, 0, 1, 0, 0, 0, 0);
        out.component = img->ch[c].component;
        out.hshift = img->ch[c].hshift - 3; out.vshift = img->ch[c].vshift - 3;
        out.hcshift = img->ch[c].hcshift - 3; out.vcshift = img->ch[c].hcshift - 3;   /* sic, dct.h:280 */
        float DCoffset = (float)((img->maxval + 1.0) * 4.0);
        for (int by = 0; by < bh; by++) for (int bx = 0; bx < bw; bx++) {
            double block[64];
            block[0] = (float)ch_get(&img->ch[c], by, bx) + DCoffset;
            for (int i = 1; i < 64; i++) block[i] = ch_get(&img->ch[offset - nb + zigzag[i] * nb + (c - beginc)], by, bx);
            transform_block(block, 1);
            for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) ch_set(&out, by * 8 + y, bx * 8 + x, (int16_t)round(block[y * 8 + x]));
        }
        free(img->ch[c].data);
        img->ch[c] = out;
    }
    img_erase(img, offset, offset + nb * 63);
    return 1;
}

/* fwd_DCT, dct.h:298-336 (requires explicit parameters) */
static int fwd_dct(fo_image *img, const int *p, int np) {
    dct_tables();
    if (np < 2) return 0;
    fo_image *tmp = fo_image_clone(img);
    int beginc = img->nb_meta_channels + p[0], endc = img->nb_meta_channels + p[1];
    int nb = endc - beginc + 1;
    int offset = img->nch;
    if (!meta_dct(img, p)) { fo_image_free(tmp); return 0; }
    float DCoffset = (float)((img->maxval + 1.0) * 4.0);
    for (int c = beginc; c < offset + 63 * nb; c++) ch_resize(&img->ch[c]);
    for (int c = beginc; c <= endc; c++) {
        int bw = img->ch[c].w, bh = img->ch[c].h;
        const fo_channel *src = &tmp->ch[c];
        for (int by = 0; by < bh; by++) for (int bx = 0; bx < bw; bx++) {
            double block[64];
            for (int i = 0; i < 64; i++) {
                int r = by * 8 + (i >> 3), col = bx * 8 + (i & 7);
                r = r < 0 ? 0 : (r >= src->h ? src->h - 1 : r);
                col = col < 0 ? 0 : (col >= src->w ? src->w - 1 : col);
                block[i] = ch_get(src, r, col);
            }
            transform_block(block, 0);
            ch_set(&img->ch[c], by, bx, (int16_t)(round(block[0]) - DCoffset));
            for (int i = 1; i < 64; i++) ch_set(&img->ch[offset - nb + zigzag[i] * nb + (c - beginc)], by, bx, (int16_t)round(block[i]));
        }
    }
    fo_image_free(tmp);
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* ChromaSubsample, transform/subsample.h (inverse + meta; the reference has no forward, :130-133)    */
/* ------------------------------------------------------------------------------------------------ */

/* check_subsample_parameters, subsample.h:33-69: one abbreviated parameter expands to (first, last, ratio_h, ratio_v) */
static int subsample_parameters(const int *p, int np, int *out) {
    int n = 0;
    if (np == 1 && p[0] >= 0 && p[0] <= 3) {
        static const int abbrev[4][2] = {{2, 2}, {2, 1}, {1, 2}, {4, 1}};   /* 4:2:0, 4:2:2, 4:4:0, 4:1:1 */
        out[0] = 1; out[1] = 2; out[2] = abbrev[p[0]][0]; out[3] = abbrev[p[0]][1];
        return 4;
    }
    for (int i = 0; i < np && i < 64; i++) out[n++] = p[i];
    if (n % 4) return 0;            /* "invalid parameters": cleared */
    return n;
}

/* inv_subsample, subsample.h:73-128 */
static int inv_subsample(fo_image *img, const int *p0, int np0) {
    int p[64];
    const int np = subsample_parameters(p0, np0, p);
    for (int i = 0; i < np; i += 4) {
        const int c1 = p[i], c2 = p[i + 1], srh = p[i + 2], srv = p[i + 3];
        for (int c = c1; c <= c2; c++) {
            fo_channel *in = &img->ch[c];
            const int ow = in->w, oh = in->h;
            if (ow >= img->ch[img->nb_meta_channels].w && oh >= img->ch[img->nb_meta_channels].h) continue;   /* LQIP / 1:16 decodes */
            fo_channel out;
            ch_make(&out, ow * srh, oh * srv, in->minval, in->maxval, 1, 0, 0, 0, 0);
            if (srv <= 2 && srh <= 2) {
                if (srh == 2) {
                    for (int y = 0; y < oh; y++) for (int x = 0; x < ow; x++) {
                        ch_set(&out, y * srv, x * srh, (3 * ch_get(in, y, x) + ch_get(in, y, x ? x - 1 : 0) + 1) >> 2);
                        ch_set(&out, y * srv, x * srh + 1, (3 * ch_get(in, y, x) + ch_get(in, y, x + 1 < ow ? x + 1 : x) + 2) >> 2);
                    }
                } else {
                    for (int y = 0; y < oh; y++) for (int x = 0; x < ow; x++) ch_set(&out, y * srv, x, ch_get(in, y, x));
                }
                if (srv == 2) {
                    fo_channel orig = out;
                    orig.data = (int16_t *)malloc((out.n + 1) * sizeof(int16_t));
                    memcpy(orig.data, out.data, out.n * sizeof(int16_t));
                    for (int y = 0; y < oh; y++) for (int x = 0; x < ow * srh; x++) {
                        ch_set(&out, y * srv, x, (3 * ch_get(&orig, y * srv, x) + ch_get(&orig, y ? (y - 1) * srv : 0, x) + 1) >> 2);
                        ch_set(&out, y * srv + 1, x, (3 * ch_get(&orig, y * srv, x) + ch_get(&orig, y + 1 < oh ? (y + 1) * srv : y * srv, x) + 2) >> 2);
                    }
                    free(orig.data);
                }
            } else {
                for (int y = 0; y < oh * srv; y++) for (int x = 0; x < ow * srh; x++) ch_set(&out, y, x, ch_get(in, y / srv, x / srh));
            }
            free(in->data);
            *in = out;
        }
    }
    return 1;
}

/* meta_subsample, subsample.h:135-157 (the reference asserts ratios of 1 or 2 here) */
static int meta_subsample(fo_image *img, const int *p0, int np0) {
    int p[64];
    const int np = subsample_parameters(p0, np0, p);
    for (int i = 0; i < np; i += 4) {
        const int c1 = p[i], c2 = p[i + 1], srh = p[i + 2], srv = p[i + 3];
        if ((srh != 1 && srh != 2) || (srv != 1 && srv != 2)) return 0;
        if (c1 < 0 || c2 >= img->nch) return 0;
        for (int c = c1; c <= c2; c++) {
            img->ch[c].w = (img->ch[c].w + srh - 1) / srh;
            img->ch[c].h = (img->ch[c].h + srv - 1) / srv;
            img->ch[c].hshift += srh == 1 ? 0 : 1;
            img->ch[c].vshift += srv == 1 ? 0 : 1;
        }
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* Palette, transform/palette.h: channels begin..end become one index channel + a palette meta-channel */
/* ------------------------------------------------------------------------------------------------ */

/* meta_palette, palette.h:70-89 */
static int meta_palette(fo_image *img, const int *p, int np) {
    if (np != 3) return 0;
    const int begin_c = img->nb_meta_channels + p[0], end_c = img->nb_meta_channels + p[1];
    if (p[0] < 0 || begin_c > end_c || end_c >= img->nch) return 0;
    const int nb = end_c - begin_c + 1, nb_colors = p[2];
    img->nb_meta_channels++;
    img->nb_channels -= nb - 1;
    img_erase(img, begin_c + 1, end_c + 1);
    fo_channel pch;
    ch_make(&pch, nb_colors, nb, 0, 1, 1, 0, 0, 0, 0);
    pch.hshift = -1;
    img_insert(img, 0, &pch);
    return 1;
}

/* fwd_palette, palette.h:92-143: the colours in use, in lexicographic order (std::set of vectors); false when there are too many */
static int palette_cmp_nb;
static int palette_cmp(const void *a, const void *b) {
    const int16_t *x = (const int16_t *)a, *y = (const int16_t *)b;
    for (int i = 0; i < palette_cmp_nb; i++) if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    return 0;
}
static int fwd_palette(fo_image *img, int *p, int np) {
    if (np != 3) return 0;
    const int begin_c = img->nb_meta_channels + p[0], end_c = img->nb_meta_channels + p[1];
    if (p[0] < 0 || begin_c > end_c || end_c >= img->nch) return 0;
    const int nb = end_c - begin_c + 1;
    const int w = img->ch[begin_c].w, h = img->ch[begin_c].h;
    const size_t n = (size_t)w * h;
    int16_t *cols = (int16_t *)malloc((n + 1) * (size_t)nb * sizeof(int16_t));
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++)
        for (int c = 0; c < nb; c++) cols[((size_t)y * w + x) * nb + c] = (int16_t)ch_get(&img->ch[begin_c + c], y, x);
    palette_cmp_nb = nb;
    qsort(cols, n, (size_t)nb * sizeof(int16_t), palette_cmp);
    size_t count = 0;
    for (size_t k = 0; k < n; k++)
        if (!count || palette_cmp(cols + (count - 1) * nb, cols + k * nb)) { memmove(cols + count * nb, cols + k * nb, (size_t)nb * sizeof(int16_t)); count++; }
    if ((long long)count > (long long)p[2]) { free(cols); return 0; }      /* too many colours */
    p[2] = (int)count;
    fo_channel pch;
    ch_make(&pch, (int)count, nb, 0, 1, 1, 0, 0, 0, 0);
    pch.hshift = -1;
    for (size_t k = 0; k < count; k++) for (int c = 0; c < nb; c++) ch_set(&pch, c, (int)k, cols[k * nb + c]);
    int16_t key[64];
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
        for (int c = 0; c < nb && c < 64; c++) key[c] = (int16_t)ch_get(&img->ch[begin_c + c], y, x);
        size_t lo = 0, hi = count;         /* position of the colour in the sorted palette */
        while (lo < hi) { size_t mid = (lo + hi) / 2; if (palette_cmp(cols + mid * nb, key) < 0) lo = mid + 1; else hi = mid; }
        ch_set(&img->ch[begin_c], y, x, (int)lo);
    }
    free(cols);
    img->nb_meta_channels++;
    img->nb_channels -= nb - 1;
    img_erase(img, begin_c + 1, end_c + 1);
    img_insert(img, 0, &pch);
    return 1;
}

/* inv_palette, palette.h:32-68 */
static int inv_palette(fo_image *img, const int *p, int np) {
    if (img->nb_meta_channels < 1 || np != 3) return 0;
    const int nb = img->ch[0].h;
    const int c0 = img->nb_meta_channels + p[0];
    if (p[0] < 0 || c0 >= img->nch) return 0;
    const int w = img->ch[c0].w, h = img->ch[c0].h;
    for (int i = 1; i < nb; i++) {
        fo_channel d;
        ch_make(&d, w, h, 0, 1, 1, 0, 0, 0, 0);
        img_insert(img, c0 + 1, &d);
        img->ch[c0 + i].component = p[0] + i;       /* as written in the reference: the channel at c0+i, whichever it is by now */
    }
    const fo_channel *pal = &img->ch[0];
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
        int index = ch_get(&img->ch[c0], y, x);
        index = CLAMPI(index, 0, pal->w - 1);
        for (int c = 0; c < nb; c++) ch_set(&img->ch[c0 + c], y, x, ch_get(pal, c, index));
    }
    img->nb_channels += nb - 1;
    img->nb_meta_channels--;
    img_erase(img, 0, 1);
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* Permute, transform/permute.h, with explicit parameters.  (Its meta-channel mode cannot round-trip in */
/* the reference: fwd_permute leaves the parameters in the Transform, so the decoder's meta_permute      */
/* takes the explicit branch while the encoder added a meta-channel.  It is not restated.)               */
/* ------------------------------------------------------------------------------------------------ */

/* meta_permute, permute.h:57-83, explicit branch: channel i of the list goes to position parameters[i] */
static int meta_permute(fo_image *img, const int *p, int np) {
    const int nb = img->nch - img->nb_meta_channels, m = img->nb_meta_channels;
    if (np <= 0 || np > nb) return 0;
    for (int i = 0; i < np; i++) {
        if (p[i] < 0 || p[i] >= np) return 0;
        for (int j = 0; j < i; j++) if (p[i] == p[j]) return 0;
    }
    fo_channel *old = (fo_channel *)malloc(sizeof(fo_channel) * (size_t)np);
    memcpy(old, &img->ch[m], sizeof(fo_channel) * (size_t)np);
    for (int i = 0; i < np; i++) img->ch[m + p[i]] = old[i];
    free(old);
    return 1;
}
/* fwd_permute, permute.h:85-124: a leading -1 selects the explicit mode and is dropped from the stored parameters */
static int fwd_permute(fo_image *img, fo_transform *t) {
    if (t->np < 3 || t->p[0] != -1) return 0;
    memmove(t->p, t->p + 1, sizeof(int) * (size_t)(t->np - 1));
    t->np--;
    if (!meta_permute(img, t->p, t->np)) { img->error = 1; }       /* the reference flags the image and still reports success */
    return 1;
}
/* inv_permute, permute.h:31-55, explicit branch: position i gets back the channel that sits at parameters[i] */
static int inv_permute(fo_image *img, const int *p, int np) {
    const int m = img->nb_meta_channels;
    if (np <= 0 || np > img->nch - m) return 0;
    fo_channel *old = (fo_channel *)malloc(sizeof(fo_channel) * (size_t)np);
    for (int i = 0; i < np; i++) { if (p[i] < 0 || p[i] >= np) { free(old); return 0; } }
    memcpy(old, &img->ch[m], sizeof(fo_channel) * (size_t)np);
    for (int i = 0; i < np; i++) img->ch[m + i] = old[p[i]];
    free(old);
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* 2DMatch, transform/2dmatch.h: inverse and meta (the forward is a search heuristic, not restated).  */
/* Single-frame images only: the "corresponding pixel of the previous frame" mode needs nb_frames > 1. */
/* ------------------------------------------------------------------------------------------------ */

/* compute_offset, 2dmatch.h:52-77: offset code -> (dx, dy) of an earlier sample, spiralling outwards in "onion layers" */
static void match_offset(int code, int *xoffset, int *yoffset) {
    int layer = 0, size = 4;
    while (code > size) { code -= size; layer++; size += 4; }
    if (layer & 1) {
        if (code <= layer) { *xoffset = 1 + layer; *yoffset = -code; }
        else if (code <= 3 + 3 * layer) { *xoffset = 2 + 2 * layer - code; *yoffset = -1 - layer; }
        else { *xoffset = -1 - layer; *yoffset = -4 - 4 * layer + code; }
    } else {
        if (code <= 1 + layer) { *xoffset = -1 - layer; *yoffset = 1 - code; }
        else if (code <= 4 + 3 * layer) { *xoffset = -3 - 2 * layer + code; *yoffset = -1 - layer; }
        else { *xoffset = 1 + layer; *yoffset = -5 - 4 * layer + code; }
    }
}
static int match_parameters(const fo_image *img, const int *p, int np, int *out) {     /* default_match_parameters, :89-95 */
    if (!np) { out[0] = 0; out[1] = img->nb_channels - 1; out[2] = 0; out[3] = 1000000; return 4; }
    for (int i = 0; i < np && i < 8; i++) out[i] = p[i];
    return np < 8 ? np : 8;
}
/* meta_match, 2dmatch.h:179-194: a match channel of the size of the first matched channel leads the list */
static int meta_match(fo_image *img, const int *p0, int np0) {
    int p[8];
    const int np = match_parameters(img, p0, np0, p);
    if (np < 3) return 0;
    const int begin_c = img->nb_meta_channels + p[0], end_c = img->nb_meta_channels + p[1];
    if (p[0] < 0 || begin_c > end_c || end_c >= img->nch) return 0;
    img->nb_meta_channels++;
    fo_channel mch;
    ch_make(&mch, img->ch[begin_c].w, img->ch[begin_c].h, 0, 1, 1, 0, 0, 0, 0);
    img_insert(img, 0, &mch);
    return 1;
}
/* inv_match, 2dmatch.h:97-177, for m.q == 1: in scanline order every sample with a non-zero code takes (or, for soft
 * matches, adds) the already reconstructed sample at the coded offset.  Channel::value is flat-indexed: an offset that
 * leaves the row lands in the neighbouring row, one that leaves the plane reads `zero` (image.h:82). */
static int inv_match(fo_image *img, const int *p0, int np0) {
    if (img->nb_meta_channels < 1) return 0;
    int p[8];
    const int np = match_parameters(img, p0, np0, p);
    if (np < 3) return 0;
    const fo_channel *m = &img->ch[0];
    const int c0 = img->nb_meta_channels + p[0], cn = img->nb_meta_channels + p[1];
    if (p[0] < 0 || p[1] < 0 || c0 >= img->nch || cn >= img->nch) return 0;
    const int softmatch = p[2];
    const int w = img->ch[c0].w, h = img->ch[c0].h;
    if (m->q != 1) return 0;            /* previous-frame mode (animations) or "unexpected quantization factor" */
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
        const int z = ch_get(m, y, x);
        if (!z) continue;
        if (z < 0 || z > m->maxval) return 0;       /* the reference indexes its offsets table out of bounds here */
        int dx, dy;
        match_offset(z, &dx, &dy);
        for (int c = c0; c <= cn; c++) {
            fo_channel *ch = &img->ch[c];
            const int src = ch_get(ch, y + dy, x + dx);
            ch_set(ch, y, x, softmatch ? S16(ch_get(ch, y, x) + src) : src);
        }
    }
    img->nb_meta_channels--;
    img_erase(img, 0, 1);
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* Approximate, transform/approximate.h: channel = quotient, extra channel at the end = remainder    */
/* ------------------------------------------------------------------------------------------------ */

static int approx_q(const int *p, int np, int c, int beginc) {      /* approximate.h:37: the last divisor repeats */
    return c + 2 - beginc < np ? p[c + 2 - beginc] : p[np - 1];
}

/* meta_approximate, approximate.h:64-80: one copy of every approximated channel with a non-zero parameter goes to the end */
static int meta_approximate(fo_image *img, const int *p, int np) {
    if (np < 3) return 0;
    const int nb = p[1] - p[0] + 1;
    if (nb < 1 || p[0] < 0 || p[1] >= img->nch) return 0;
    for (int c = p[0]; c <= p[1]; c++) {
        if (!approx_q(p, np, c, p[0])) continue;
        fo_channel d = img->ch[c];
        d.data = (int16_t *)malloc(d.n * sizeof(int16_t) + 2);
        if (d.n) memcpy(d.data, img->ch[c].data, d.n * sizeof(int16_t));
        img_insert(img, img->nch, &d);
    }
    return 1;
}

/* fwd_approximate, approximate.h:83-113: floor division by q+1, remainder in [0, q] */
static int fwd_approximate(fo_image *img, const int *p, int np) {
    const int offset = img->nch;
    if (!meta_approximate(img, p, np)) { img->error = 1; return 1; }    /* the reference sets image.error and still returns true */
    const int beginc = p[0], endc = p[1];
    int i = 0;
    for (int c = beginc; c <= endc; c++) {
        const int q = approx_q(p, np, c, beginc) + 1;
        if (q == 1) continue;
        fo_channel *ch = &img->ch[c], *chr = &img->ch[offset + i];
        i++;
        for (size_t k = 0; k < ch->n; k++) {
            const int v = ch->data[k];
            int quotient = S16(v / q), r = S16(v % q);
            if (r < 0) { quotient = S16(quotient - 1); r = S16(r + q); }
            ch->data[k] = (int16_t)quotient;
            chr->data[k] = (int16_t)r;
        }
        ch->minval = S16(ch->minval / q);
        ch->maxval = S16(ch->maxval / q);
        chr->minval = 0;
        chr->maxval = S16(q - 1);
        chr->q = ch->q;
    }
    return 1;
}

/* inv_approximate, approximate.h:32-62 */
static int inv_approximate(fo_image *img, const int *p, int np) {
    if (np < 3) return 0;               /* the reference would read past its parameter vector */
    const int beginc = p[0], endc = p[1];
    int offset = img->nch - (endc - beginc + 1);
    for (int c = beginc; c <= endc; c++) if (!approx_q(p, np, c, beginc)) offset++;
    if (beginc < 0 || endc >= img->nch || offset < 0 || offset > img->nch) return 0;
    int i = 0;
    for (int c = beginc; c <= endc; c++) {
        const int q = approx_q(p, np, c, beginc) + 1;
        if (q == 1) continue;
        fo_channel *ch = &img->ch[c];
        const fo_channel *chr = &img->ch[offset + i];
        i++;
        if (chr->n) ch->q = chr->q;
        for (int y = 0; y < ch->h; y++) for (int x = 0; x < ch->w; x++) {
            int v = S16(ch_get(ch, y, x) * q);
            v = S16(v + (chr->n ? ch_get(chr, y, x) : 0));
            ch_set(ch, y, x, v);
        }
    }
    img_erase(img, offset, img->nch);
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* Transform dispatch (transform/transform.cpp:48-81) and Image::undo_transforms / do_transform       */
/* ------------------------------------------------------------------------------------------------ */

static int transform_apply(fo_image *img, fo_transform *t, int inverse) {
    switch (t->id) {
    case FO_YCBCR: return ycbcr(img, inverse);
    case FO_YCOCG: return ycocg(img, inverse);
    case FO_QUANTIZE: return quantize(img, inverse, t->p, t->np);
    case FO_SQUEEZE: return squeeze(img, inverse, t->p, t->np);
    case FO_DCT:
        if (t->np < 2) {    /* default_DCT_parameters, dct.h:209-213 (fills the Transform's own parameters) */
            if (!inverse) return 0;     /* the reference crashes here */
            t->p = (int *)realloc(t->p, 2 * sizeof(int)); t->np = 2; t->p[0] = 0; t->p[1] = img->nb_channels - 1;
        }
        return inverse ? inv_dct(img, t->p) : fwd_dct(img, t->p, t->np);
    case FO_SUBSAMPLE:
        if (!inverse) return 0;     /* fwd_subsample is a stub in the reference (subsample.h:130-133) */
        for (int i = 0; i + 3 < t->np || (t->np == 1 && i == 0); i += 4) {     /* channels must exist */
            if (t->np == 1) { if (img->nch < 3) return 0; break; }
            if (t->p[i] < 0 || t->p[i + 1] >= img->nch) return 0;
        }
        return inv_subsample(img, t->p, t->np);
    case FO_PALETTE: return inverse ? inv_palette(img, t->p, t->np) : fwd_palette(img, t->p, t->np);
    case 9: return inverse ? inv_permute(img, t->p, t->np) : fwd_permute(img, t);
    case 8: return inverse ? inv_match(img, t->p, t->np) : 0;       /* fwd_match: a search heuristic, not restated */
    case 10: return inverse ? inv_approximate(img, t->p, t->np) : fwd_approximate(img, t->p, t->np);
    default: return 0;
    }
}

static int transform_meta_apply(fo_image *img, fo_transform *t) {
    switch (t->id) {
    case FO_YCBCR: case FO_YCOCG: case FO_QUANTIZE: return 1;
    case FO_SQUEEZE:
        if (!t->np) {       /* meta_squeeze fills the parameters in place, squeeze.h:324 */
            int adj[256];
            int n = default_squeeze_parameters(img, adj);
            t->p = (int *)realloc(t->p, sizeof(int) * (size_t)(n > 0 ? n : 1)); memcpy(t->p, adj, sizeof(int) * (size_t)n); t->np = n;
        }
        meta_squeeze(img, t->p, t->np);
        return 1;
    case FO_DCT:
        if (t->np < 2) { t->p = (int *)realloc(t->p, 2 * sizeof(int)); t->np = 2; t->p[0] = 0; t->p[1] = img->nb_channels - 1; }
        return meta_dct(img, t->p);
    case FO_SUBSAMPLE: return meta_subsample(img, t->p, t->np);
    case 10: return meta_approximate(img, t->p, t->np);
    case FO_PALETTE: return meta_palette(img, t->p, t->np);
    case 9: return meta_permute(img, t->p, t->np);
    case 8: return meta_match(img, t->p, t->np);
    default: return 0;
    }
}

int fo_undo_transforms(fo_image *img, int keep) {   /* image/image.cpp:94-115 */
    while (img->ntr > keep) {
        fo_transform *t = &img->tr[img->ntr - 1];
        if (!transform_apply(img, t, 1)) { img->error = 1; return -1; }
        free(t->p);
        img->ntr--;
    }
    if (!keep) {
        for (int i = 0; i < img->nch; i++)
            for (size_t j = 0; j < img->ch[i].n; j++) {
                int v = img->ch[i].data[j];
                img->ch[i].data[j] = (int16_t)CLAMPI(v, img->minval, img->maxval);
            }
    }
    return 0;
}

int fo_do_transform(fo_image *img, int id, const int *params, int np) {   /* image/image.cpp:117-122 */
    fo_transform t;
    t.id = id; t.np = np;
    t.p = (int *)malloc(sizeof(int) * (size_t)(np > 0 ? np : 1));
    if (np) memcpy(t.p, params, sizeof(int) * (size_t)np);
    int did = transform_apply(img, &t, 0);
    if (did) img_push_transform(img, id, t.p, t.np);
    free(t.p);
    return did;
}

/* ------------------------------------------------------------------------------------------------ */
/* Byte reader and varints (encoding.cpp:45-59).  End-of-stream follows FileIO (fileio.h:33-81), the IO  */
/* class behind fuif_decode_file: isEOF() is feof(), i.e. it turns true only after a read has FAILED,   */
/* not when the last byte has been consumed (BlobReader, fileio.h:83-140, differs in exactly that).      */
/* ------------------------------------------------------------------------------------------------ */

typedef struct { const uint8_t *data; size_t size, pos; int eof; } blob;
static inline int io_getc(blob *io) { if (io->pos >= io->size) { io->eof = 1; return -1; } return io->data[io->pos++]; }
static inline int io_eof(const blob *io) { return io->eof; }

static int read_varint(blob *io) {
    int result = 0, bytes_read = 0;
    while (bytes_read++ < 10) {
        int number = io_getc(io);
        if (number < 0) break;
        if (number < 128) return result + number;
        number -= 128;
        result += number;
        result = (int)((unsigned)result << 7);
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------------ */
/* Range decoder (maniac/rac.h:35-114)                                                                */
/* ------------------------------------------------------------------------------------------------ */

typedef struct { blob *io; uint64_t range, low; } rac_in;

static inline uint64_t rac_byte(rac_in *r) {        /* read_catch_eof, rac.h:64-69: EOS (-1) read as all-ones garbage */
    int c = io_getc(r->io);
    return (uint64_t)(int64_t)c;
}
static void rac_init(rac_in *r, blob *io) {         /* RacInput ctor, rac.h:97-104 */
    r->io = io; r->range = 1u << 24; r->low = 0;
    uint64_t k = 1u << 24;
    while (k > 1) { r->low <<= 8; r->low |= rac_byte(r); k >>= 8; }
}
static inline void rac_input(rac_in *r) {           /* rac.h:70-81 */
    if (r->range <= (1u << 16)) { r->low <<= 8; r->range <<= 8; r->low |= rac_byte(r); }
    if (r->range <= (1u << 16)) { r->low <<= 8; r->range <<= 8; r->low |= rac_byte(r); }
}
static inline int rac_get(rac_in *r, uint64_t chance) {   /* rac.h:82-95 */
    if (r->low >= r->range - chance) { r->low -= r->range - chance; r->range = chance; rac_input(r); return 1; }
    r->range -= chance; rac_input(r); return 0;
}
static inline int rac_read_12bit(rac_in *r, int b12) { return rac_get(r, (r->range * (uint64_t)b12 + 0x800) >> 12); }   /* rac.h:42-52,107 */
static inline int rac_read_bit(rac_in *r) { return rac_get(r, r->range >> 1); }                                         /* rac.h:111 */

/* ------------------------------------------------------------------------------------------------ */
/* Adaptive chances (maniac/chance.h, chance.cpp) and the integer reader (maniac/symbol.h)            */
/* ------------------------------------------------------------------------------------------------ */

typedef struct { uint16_t next[4096][2]; } chance_table;

static void build_table(chance_table *t, uint32_t factor, unsigned max_p) {   /* chance.cpp:31-65 */
    const int64_t one = 1LL << 32;
    const int size = 4096;
    int64_t p;
    unsigned last_p8, p8, i;
    memset(t->next, 0, sizeof(t->next));
    last_p8 = 0;
    p = one / 2;
    for (i = 0; i < (unsigned)size / 2; i++) {
        p8 = (unsigned)((size * p + one / 2) >> 32);
        if (p8 <= last_p8) p8 = last_p8 + 1;
        if (last_p8 && last_p8 < (unsigned)size && p8 <= max_p) t->next[last_p8][1] = (uint16_t)p8;
        p += ((one - p) * factor + one / 2) >> 32;
        last_p8 = p8;
    }
    for (i = size - max_p; i <= max_p; i++) {
        if (t->next[i][1]) continue;
        p = (i * one + size / 2) / size;
        p += ((one - p) * factor + one / 2) >> 32;
        p8 = (unsigned)((size * p + one / 2) >> 32);
        if (p8 <= i) p8 = i + 1;
        if (p8 > max_p) p8 = max_p;
        t->next[i][1] = (uint16_t)p8;
    }
    for (i = 1; i < (unsigned)size; i++) t->next[i][0] = (uint16_t)(size - t->next[size - i][1]);
}

/* SymbolChance<BitChance,15>, symbol.h:72-139: [0]=zero [1]=sign [2..15]=exp[14] [16..30]=mant[15] */
typedef struct { uint16_t c[32]; } symchance;
#define SC_ZERO 0
#define SC_SIGN 1
#define SC_EXP 2
#define SC_MANT 16

static void symchance_init(symchance *s, int zero_chance) {    /* symbol.h:115-138 */
    s->c[SC_ZERO] = (uint16_t)zero_chance;
    s->c[SC_SIGN] = 0x800;
    uint64_t rp = 0x1000 - (uint64_t)zero_chance;
    for (int i = 0; i < MAX_BIT_DEPTH - 1; i++) {
        if (rp < 0x100) rp = 0x100;
        if (rp > 0xf00) rp = 0xf00;
        s->c[SC_EXP + i] = (uint16_t)(0x1000 - rp);
        rp = (rp * rp + 0x800) >> 12;
    }
    for (int i = 0; i < MAX_BIT_DEPTH; i++) s->c[SC_MANT + i] = 1024;
    s->c[31] = 0;
}

static inline int ilog2u(uint32_t l) { return l == 0 ? 0 : 31 - __builtin_clz(l); }   /* maniac/util.h:33-36 */

static long long st_bits = 0, st_steps = 0, st_syms = 0, st_same = 0, st_zero = 0, st_unk1 = 0; static int st_last = -1;   /* FO_STATS instrumentation only */
static inline int sym_read(rac_in *rac, const chance_table *t, symchance *s, int idx) {   /* compound.h:90-95 */
    st_bits++;
    int bit = rac_read_12bit(rac, s->c[idx]);
    s->c[idx] = t->next[s->c[idx]][bit];
    return bit;
}

/* reader<15>(coder,min,max), symbol.h:154-185 */
static int read_int(rac_in *rac, const chance_table *t, symchance *s, int min, int max) {
    if (min == max) return min;
    int sign;
    if (sym_read(rac, t, s, SC_ZERO)) return 0;
    if (min < 0) { if (max > 0) sign = sym_read(rac, t, s, SC_SIGN); else sign = 0; } else sign = 1;
    const int amax = (sign ? max : -min);
    const int emax = ilog2u((uint32_t)amax);
    int e = 0;
    for (; e < emax; e++) if (sym_read(rac, t, s, SC_EXP + e)) break;
    int have = (1 << e);
    for (int pos = e; pos > 0;) {
        pos--;
        int minabs1 = have | (1 << pos);
        if (minabs1 > amax) continue;
        if (sym_read(rac, t, s, SC_MANT + pos)) have = minabs1;
    }
    return sign ? have : -have;
}
static int read_int2(rac_in *rac, const chance_table *t, symchance *s, int min, int max) {   /* symbol.h:232-236 */
    if (min > 0) return read_int(rac, t, s, 0, max - min) + min;
    else if (max < 0) return read_int(rac, t, s, min - max, 0) + max;
    return read_int(rac, t, s, min, max);
}
/* UniformSymbolCoder::read_int(min,len), symbol.h:44-56 */
static int uniform_read(rac_in *rac, int min, int len) {
    while (len != 0) {
        int med = len / 2;
        if (rac_read_bit(rac)) { min = min + med + 1; len = len - (med + 1); }
        else len = med;
    }
    return min;
}

/* ------------------------------------------------------------------------------------------------ */
/* MANIAC tree (maniac/compound.h)                                                                   */
/* ------------------------------------------------------------------------------------------------ */

typedef struct { int16_t property; uint16_t childID; int32_t splitval; } tnode;    /* compound.h:41-51 */
typedef struct { tnode *n; int size, cap; } tree;

static int tree_push(tree *t) {
    if (t->size == t->cap) { t->cap = t->cap ? t->cap * 2 : 64; t->n = (tnode *)realloc(t->n, sizeof(tnode) * (size_t)t->cap); }
    t->n[t->size].property = -1; t->n[t->size].childID = 0; t->n[t->size].splitval = 0;
    return t->size++;
}

/* MetaPropertySymbolCoder::read_tree / read_subtree, compound.h:277-320 (recursion unrolled onto a stack) */
static int read_tree(rac_in *rac, const int (*range)[2], int nprops, tree *t) {
    static chance_table meta_table;
    static int meta_ready = 0;
    if (!meta_ready) { build_table(&meta_table, 0xFFFFFFFFu / 19, 4096 - 2); meta_ready = 1; }   /* cut=2, alpha=0xFFFFFFFF/19 */
    symchance coder[3];
    for (int i = 0; i < 3; i++) symchance_init(&coder[i], 1024);        /* SimpleSymbolCoder ctx(ZERO_CHANCE), symbol.h:219 */
    int (*sub)[2] = (int (*)[2])malloc(sizeof(int[2]) * (size_t)(nprops > 0 ? nprops : 1));
    memcpy(sub, range, sizeof(int[2]) * (size_t)nprops);
    typedef struct { int pos, p, oldmin, oldmax, splitval, stage; } frame;
    int fcap = 64, fsz = 0;
    frame *st = (frame *)malloc(sizeof(frame) * (size_t)fcap);
    t->size = 0;
    tree_push(t);
    st[fsz++] = (frame){0, 0, 0, 0, 0, 0};
    int ok = 1;
    while (fsz > 0 && ok) {
        frame *f = &st[fsz - 1];
        if (f->stage == 0) {
            int p = read_int2(rac, &meta_table, &coder[0], 0, nprops) - 1;
            t->n[f->pos].property = (int16_t)p;
            if (p == -1) { fsz--; continue; }
            f->p = p; f->oldmin = sub[p][0]; f->oldmax = sub[p][1];
            if (f->oldmin >= f->oldmax) { ok = 0; break; }        /* "Invalid tree", compound.h:285-288 */
            f->splitval = read_int2(rac, &meta_table, &coder[2], f->oldmin, f->oldmax - 1);
            t->n[f->pos].splitval = f->splitval;
            if (t->size + 2 > 65535) { ok = 0; break; }
            int child = t->size;
            t->n[f->pos].childID = (uint16_t)child;
            tree_push(t); tree_push(t);
            sub[p][0] = f->splitval + 1;
            f->stage = 1;
            if (fsz == fcap) { fcap *= 2; st = (frame *)realloc(st, sizeof(frame) * (size_t)fcap); f = &st[fsz - 1]; }
            st[fsz++] = (frame){child, 0, 0, 0, 0, 0};
        } else if (f->stage == 1) {
            sub[f->p][0] = f->oldmin;
            sub[f->p][1] = f->splitval;
            f->stage = 2;
            int child = t->n[f->pos].childID + 1;
            if (fsz == fcap) { fcap *= 2; st = (frame *)realloc(st, sizeof(frame) * (size_t)fcap); }
            st[fsz++] = (frame){child, 0, 0, 0, 0, 0};
        } else {
            sub[f->p][1] = f->oldmax;
            fsz--;
        }
    }
    free(st); free(sub);
    return ok;
}

/* ------------------------------------------------------------------------------------------------ */
/* Context model (encoding/context_predict.h)                                                        */
/* ------------------------------------------------------------------------------------------------ */

static inline int slog(int x16) {       /* slog(pixel_type), context_predict.h:54-61 */
    int x = S16(x16);
    if (x == 0) return 0;
    if (x > 0) return S16(32 - __builtin_clz((unsigned)x));
    return S16(-(32 - __builtin_clz((unsigned)(-x))));
}
static inline int fooabs(int x16) { int x = S16(x16); return S16(x < 0 ? -x : x); }   /* context_predict.h:63-65 */

/* init_properties, context_predict.h:67-120.  returns number of properties */
static int init_properties(int (*pr)[2], const fo_image *img, int beginc, int endc, int max_properties) {
    int n = 0, offset = 0;
    for (int j = beginc - 1; j >= 0 && offset < max_properties; j--) {
        if (img->ch[j].minval == img->ch[j].maxval) continue;
        if (img->ch[j].hshift < 0) continue;
        int minval = img->ch[j].minval; if (minval > 0) minval = 0;
        int maxval = img->ch[j].maxval; if (maxval < 0) maxval = 0;
        pr[n][0] = 0; pr[n][1] = fooabs(maxval > -minval ? maxval : minval); n++; offset++;
        pr[n][0] = slog(minval); pr[n][1] = slog(maxval); n++; offset++;
    }
    int minval = LARGEST_VAL, maxval = SMALLEST_VAL, maxh = 0, maxw = 0;
    for (int j = beginc; j <= endc; j++) {
        if (img->ch[j].minval < minval) minval = img->ch[j].minval;
        if (img->ch[j].maxval > maxval) maxval = img->ch[j].maxval;
        if (img->ch[j].h > maxh) maxh = img->ch[j].h;
        if (img->ch[j].w > maxw) maxw = img->ch[j].w;
    }
    if (minval > 0) minval = 0;
    if (maxval < 0) maxval = 0;
    int amax = fooabs(minval) > fooabs(maxval) ? fooabs(minval) : fooabs(maxval);
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

static inline int median3(int a, int b, int c) {    /* util.h:9-23 */
    if (a < b) { if (b < c) return b; return a < c ? c : a; }
    if (a < c) return a;
    return b < c ? c : b;
}

/* predict_and_compute_properties, context_predict.h:125-168 (the _no_edge_case variant :171-206 computes the
   same values where it is used: y>1, 1<x<w-1, predictor 0) */
static int predict_props(int *p, const fo_channel *ch, int x, int y, int predictor, int offset) {
    const int16_t *d = ch->data;
    const int w = ch->w;
    int left = (x ? d[(size_t)y * w + x - 1] : ch->zero);
    int top = (y ? d[(size_t)(y - 1) * w + x] : ch->zero);
    int topleft = (x && y ? d[(size_t)(y - 1) * w + x - 1] : left);
    int topright = (x + 1 < w && y ? d[(size_t)(y - 1) * w + x + 1] : top);
    int leftleft = (x > 1 ? d[(size_t)y * w + x - 2] : left);
    int toptop = (y > 1 ? d[(size_t)(y - 2) * w + x] : top);
    p[offset++] = fooabs(top);
    p[offset++] = fooabs(left);
    p[offset++] = slog(top);
    p[offset++] = slog(left);
    p[offset++] = y;
    p[offset++] = x;
    p[offset++] = left + top - topleft;
    p[offset++] = topleft + topright - top;
    p[offset++] = slog(left - topleft);
    p[offset++] = slog(topleft - top);
    p[offset++] = slog(top - topright);
    p[offset++] = slog(top - toptop);
    p[offset++] = slog(left - leftleft);
    switch (predictor) {
    case 0: return ch->zero;
    case 1: return S16((left + top) / 2);
    case 2: return median3(S16(left + top - topleft), left, top);
    case 3: return left;
    case 4: return top;
    case 5: return S16((left + topleft + top + topright) / 4);
    case 6: return S16(CLAMPI(left + top - topleft, ch->minval, ch->maxval));
    default: return median3(S16(left + top - topleft), left, top);
    }
}

/* precompute_references, context_predict.h:233-289.  refs is [w][nref] int */
static void precompute_references(const fo_channel *ch, int y, const fo_image *img, int i, int max_properties, int *refs, int nref) {
    int offset = 0;
    int oy = y << ch->vshift;
    for (int j = i - 1; j >= 0 && offset < max_properties; j--) {
        const fo_channel *cj = &img->ch[j];
        if (cj->minval == cj->maxval) continue;
        if (cj->hshift < 0) continue;
        int ry = oy >> cj->vshift;
        if (ry >= cj->h) ry = cj->h - 1;
        const int16_t *row = cj->data + (size_t)ry * cj->w;
        if (ch->hshift == cj->hshift && ch->w <= cj->w) {
            for (int x = 0; x < ch->w; x++) { int v = row[x]; refs[x * nref + offset] = fooabs(v); refs[x * nref + offset + 1] = slog(v); }
        } else if (ch->hshift < cj->hshift) {
            int stepsize = (1 << cj->hshift) >> ch->hshift;
            int x = 0, rx = 0, v;
            for (; rx < cj->w - 1; rx++) {
                v = row[rx];
                for (int s = 0; s < stepsize; s++, x++) if (x < ch->w) { refs[x * nref + offset] = fooabs(v); refs[x * nref + offset + 1] = slog(v); }
            }
            v = row[rx];
            while (x < ch->w) { refs[x * nref + offset] = fooabs(v); refs[x * nref + offset + 1] = slog(v); x++; }
        } else {
            for (int x = 0; x < ch->w; x++) {
                int ox = x << ch->hshift;
                int rx = ox >> cj->hshift;
                if (rx >= cj->w) rx = cj->w - 1;
                int v = row[rx];
                refs[x * nref + offset] = fooabs(v); refs[x * nref + offset + 1] = slog(v);
            }
        }
        offset += 2;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* fuif_decode_channel (encoding/encoding.cpp:259-429) and fuif_decode (:599-720)                     */
/* ------------------------------------------------------------------------------------------------ */

static int check_bit_depth(int minv, int maxv, int predictor) {   /* encoding.cpp:61-72 */
    int maxav = S16(abs(maxv));
    if (-minv > maxav) maxav = S16(-minv);
    if (predictor > 0 && maxv - minv > maxav) maxav = S16(maxv - minv);
    if (predictor > 0 && abs(minv - maxv) > maxav) maxav = S16(abs(minv - maxv));
    return ilog2u((uint32_t)maxav) + 1 <= MAX_BIT_DEPTH;
}

#define STOP(io, btl) (io_eof(io) || ((btl) && (io)->pos >= (btl)))

static int corrupt_or_truncated(blob *io, fo_channel *ch, size_t btl) {   /* encoding.cpp:209-219 */
    if (STOP(io, btl)) { ch_fill(ch, 0); return 1; }
    return 0;
}

typedef struct { chance_table table; int cutoff, alpha; } dec_ctx;

static int decode_channel_group(blob *io, dec_ctx *dc, int max_properties, int *pbeginc, fo_image *img, size_t btl) {
    int beginc = *pbeginc;
    if (STOP(io, btl)) return 1;
    int firstbyte = read_varint(io);
    if (STOP(io, btl)) return 1;
    int endc = beginc + (firstbyte >> 4);
    int compress = firstbyte & 1;
    int predictor = (firstbyte & 14) >> 1;
    int global_minv = S16(1 - read_varint(io));
    if (STOP(io, btl)) return 1;
    if (global_minv == 1) global_minv = S16(read_varint(io));
    if (STOP(io, btl)) return 1;
    int global_maxv = S16(global_minv + read_varint(io));
    if (STOP(io, btl)) return 1;
    if (endc >= img->nch || endc < beginc) return 0;

    int firstrealc = beginc;
    for (int i = beginc; i <= endc; i++) {
        fo_channel *ch = &img->ch[i];
        if (ch->w * ch->h <= 0) continue;
        ch->minval = global_minv; ch->maxval = global_maxv;
        if (endc > beginc && global_minv < global_maxv) {
            ch->minval = S16(ch->minval + read_varint(io));
            ch->maxval = S16(ch->minval + read_varint(io));
        }
        if (ch->minval == ch->maxval) { ch_fill(ch, ch->minval); firstrealc++; }
        if (ch->minval == 0 && ch->maxval == 0) continue;
        ch->q = read_varint(io);
        if (STOP(io, btl)) return corrupt_or_truncated(io, ch, btl);
        if (compress && !check_bit_depth(ch->minval, ch->maxval, predictor)) return 0;
    }
    if (firstrealc > endc) { *pbeginc = endc; return 1; }

    int pr[64][2];
    int nprops = init_properties(pr, img, beginc, endc, max_properties);
    int nref = nprops - NB_NONREF;

    int predictability = 2048;
    if (predictor == 0 && compress) {
        int rounded = read_varint(io);
        if (rounded < 1 || rounded > 127) return corrupt_or_truncated(io, &img->ch[firstrealc], btl);
        predictability = rounded * 32;
    }

    rac_in rac;
    rac_init(&rac, io);

    if (!compress) {
        for (int i = beginc; i <= endc; i++) {
            fo_channel *ch = &img->ch[i];
            if (ch->minval == ch->maxval) continue;
            ch_setzero(ch);
            ch_resize(ch);
            for (int y = 0; y < ch->h; y++) {
                if (STOP(io, btl)) break;
                for (int x = 0; x < ch->w; x++) ch->data[(size_t)y * ch->w + x] = (int16_t)uniform_read(&rac, ch->minval, ch->maxval - ch->minval);
            }
            if (STOP(io, btl)) break;
        }
        *pbeginc = endc;
        return 1;
    }

    tree t = {0};
    if (!read_tree(&rac, (const int (*)[2])pr, nprops, &t)) { free(t.n); return corrupt_or_truncated(io, &img->ch[beginc], btl); }

    /* FinalPropertySymbolCoder ctor, compound.h:213-225 */
    int nleaves = (t.size + 1) / 2;
    symchance *leaf = (symchance *)malloc(sizeof(symchance) * (size_t)nleaves);
    for (int i = 0; i < nleaves; i++) symchance_init(&leaf[i], predictability);
    for (int i = 0, leafID = 0; i < t.size; i++) if (t.n[i].property == -1) t.n[i].childID = (uint16_t)leafID++;

    int props[64];
    memset(props, 0, sizeof(props));

    for (int i = beginc; i <= endc; i++) {
        fo_channel *ch = &img->ch[i];
        if (ch->minval == ch->maxval) continue;
        ch_setzero(ch);
        ch_resize(ch);
        if (t.size == 1 && predictor == 0 && ch->zero == 0) {       /* fast track, encoding.cpp:371-383 */
            for (int y = 0; y < ch->h; y++) {
                if (STOP(io, btl)) { beginc = i; break; }
                for (int x = 0; x < ch->w; x++)
                    ch->data[(size_t)y * ch->w + x] = (int16_t)read_int(&rac, &dc->table, &leaf[0], ch->minval, ch->maxval);
            }
        } else {
            int *refs = (int *)calloc((size_t)(nref > 0 ? nref : 1) * (size_t)(ch->w > 0 ? ch->w : 1), sizeof(int));
            for (int y = 0; y < ch->h; y++) {
                if (STOP(io, btl)) { beginc = i; break; }
                precompute_references(ch, y, img, beginc, max_properties, refs, nref);
                for (int x = 0; x < ch->w; x++) {
                    for (int k = 0; k < nref; k++) props[k] = refs[x * nref + k];
                    int guess = predict_props(props, ch, x, y, predictor, nref);
                    int mn = ch->minval - guess, mx = ch->maxval - guess;
                    int diff;
                    if (mn == mx) diff = mn;
                    else {
                        int pos = 0;                                   /* find_leaf, compound.h:142-153 */
                        st_syms++;
                        while (t.n[pos].property != -1) {
                            st_steps++;
                            if (props[t.n[pos].property] > t.n[pos].splitval) pos = t.n[pos].childID;
                            else pos = t.n[pos].childID + 1;
                        }
                        if (pos == st_last) st_same++;
                        st_last = pos;
                        { int q = 0, d = 0; while (t.n[q].property != -1) { int pr_ = t.n[q].property - nref; if (pr_ == 1 || pr_ == 3 || pr_ == 6 || pr_ == 8 || pr_ == 12) break; d++; q = props[t.n[q].property] > t.n[q].splitval ? t.n[q].childID : t.n[q].childID + 1; } st_unk1 += d; }
                        diff = read_int(&rac, &dc->table, &leaf[t.n[pos].childID], mn, mx);
                        if (diff == 0) st_zero++;
                        if (getenv("FO_TRACE") && atoi(getenv("FO_TRACE")) == i && y < 2 && x < 6)
                            fprintf(stderr, "[oracle]   y %d x %d leaf %d mn %d mx %d diff %d guess %d pos %zu\n", y, x, t.n[pos].childID, mn, mx, diff, guess, io->pos);
                    }
                    ch->data[(size_t)y * ch->w + x] = (int16_t)(S16(diff) + guess);
                }
            }
            free(refs);
        }
        if (STOP(io, btl)) break;
    }
    if (getenv("FO_STATS")) {
        fprintf(stderr, "group %d-%d %dx%d pred %d nodes %d props %d syms %lld steps/sym %.2f bits/sym %.2f same %.2f zero %.2f knownprefix %.2f\n", *pbeginc, endc, img->ch[endc].w, img->ch[endc].h,
                predictor, t.size, nprops, st_syms, st_syms ? (double)st_steps / st_syms : 0.0, st_syms ? (double)st_bits / st_syms : 0.0, st_syms ? (double)st_same / st_syms : 0.0, st_syms ? (double)st_zero / st_syms : 0.0, st_syms ? (double)st_unk1 / st_syms : 0.0);
        st_bits = st_steps = st_syms = st_same = st_zero = st_unk1 = 0;
    }
    free(leaf); free(t.n);
    *pbeginc = endc;
    return 1;
}

fo_image *fo_decode(const uint8_t *bytes, size_t n, int preview, int maniac_cutoff, int maniac_alpha, long long *group_offsets, int *ngroups) {
    blob io = {bytes, n, 0, 0};
    int gcap = ngroups ? *ngroups : 0, gcount = 0;
    if (ngroups) *ngroups = 0;
    if (n < 4) return NULL;
    int multi = 0;
    if (!memcmp(bytes, "FUAF", 4)) multi = 1;
    else if (memcmp(bytes, "FUIF", 4)) return NULL;
    io.pos = 4;
    int nb_channels = read_varint(&io) - '0';
    int bit_depth = read_varint(&io) - '&';
    int w = read_varint(&io) + 1;
    int h = read_varint(&io) + 1;
    if (multi) {        /* animation header, encoding.cpp:614-623 (frames stay a vertical filmstrip) */
        int nb_frames = read_varint(&io) + 2;
        (void)read_varint(&io);
        int numerator = read_varint(&io);
        if (numerator) for (int i = 1; i < nb_frames; i++) (void)read_varint(&io);
        (void)read_varint(&io);
    }
    int colormodel = read_varint(&io);
    int max_properties = read_varint(&io);
    if (nb_channels < 1 || bit_depth < 1 || bit_depth > 16 || w < 1 || h < 1) return NULL;
    fo_image *img = fo_image_new(w, h, (1 << bit_depth) - 1, nb_channels, colormodel);

    int responsive_offsets[5], rel = 0;
    for (int s = 0; s < 5; s++) { responsive_offsets[s] = read_varint(&io) + rel; rel = responsive_offsets[s]; }
    rel = (int)io.pos;
    for (int s = 0; s < 5; s++) responsive_offsets[s] += rel;

    int nb_transforms = read_varint(&io);
    for (int i = 0; i < nb_transforms; i++) {
        int idp = read_varint(&io);
        int id = idp & 0xf;
        int has_params = (id == FO_SUBSAMPLE || id == FO_PALETTE || id == FO_SQUEEZE || id == FO_DCT || id == 8 || id == 9 || id == 10);
        int np = has_params ? (idp >> 4) : 0;
        int params[1024];
        if (np > 1024 || idp < 0) { fo_image_free(img); return NULL; }
        for (int j = 0; j < np; j++) params[j] = read_varint(&io);
        img_push_transform(img, id, params, np);
        if (!transform_meta_apply(img, &img->tr[img->ntr - 1])) { fo_image_free(img); return NULL; }
    }

    size_t btl = 0;
    if (preview >= 0) btl = (size_t)responsive_offsets[preview];

    dec_ctx *dc = (dec_ctx *)malloc(sizeof(dec_ctx));
    build_table(&dc->table, (uint32_t)maniac_alpha, (unsigned)(4096 - maniac_cutoff));
    int nch = img->nch;
    for (int i = 0; i < nch; i++) {
        if ((preview < 0 || io.pos < btl) && !io_eof(&io)) {
            if (!img->ch[i].w || !img->ch[i].h) continue;
            if (group_offsets && gcount < gcap) { group_offsets[2 * gcount] = (long long)io.pos; group_offsets[2 * gcount + 1] = i; gcount++; }
            if (!decode_channel_group(&io, dc, max_properties, &i, img, btl)) { free(dc); fo_image_free(img); return NULL; }
        } else break;
    }
    free(dc);
    if (ngroups) *ngroups = gcount;
    return img;
}

/* ================================================================================================ */
/* ENCODER (test infrastructure for the encode-side rows of SURVEY 8a: a20, a22, a24)                */
/*   fuif_encode                 encoding/encoding.cpp:455-573                                       */
/*   fuif_encode_channels        encoding/encoding.cpp:74-207 (learn pass + real pass)               */
/*   PropertySymbolCoder         maniac/compound_enc.h:243-518 (tree learning, simplify)             */
/*   CompoundSymbolBitCoder      maniac/compound_enc.h:76-136 (virtual chances, cost estimates)      */
/*   MetaPropertySymbolCoder     maniac/compound_enc.h:523-552 (write_tree)                          */
/*   writer<>, UniformSymbolCoder maniac/symbol_enc.h:28-109                                         */
/*   RacOutput                   maniac/rac_enc.h:28-100                                             */
/*   Log4kTable                  maniac/chance.cpp:67-91                                             */
/*   BlobIO                      fileio.h:148-272 (including its bytes_used = seek_pos + 1 quirk)    */
/* ================================================================================================ */

#define CONTEXT_TREE_SPLIT_THRESHOLD (5461 * 8 * 2)     /* config.h */
#define CONTEXT_TREE_MIN_SUBTREE_SIZE 10                /* config.h */

/* ---- BlobIO (fileio.h:148-272) ---- */
typedef struct { uint8_t *data; size_t cap, used, pos; } oblob;
static void ob_grow(oblob *b, size_t need) {
    if (need < b->cap) return;
    size_t ns = need < 4096 ? 4096 : need;
    if (ns < b->cap * 3 / 2) ns = b->cap * 3 / 2;
    b->data = (uint8_t *)realloc(b->data, ns);
    memset(b->data + b->cap, 0, ns - b->cap);       /* the reference leaves this memory uninitialised */
    b->cap = ns;
}
static void ob_putc(oblob *b, int c) {              /* BlobIO::fputc: bytes_used runs ONE byte ahead of the last byte written */
    ob_grow(b, b->pos + 1);
    b->data[b->pos++] = (uint8_t)c;
    if (b->used < b->pos) b->used = b->pos + 1;
}
static void ob_varint(oblob *b, size_t number, int done) {      /* write_big_endian_varint, encoding.cpp:30-41 */
    if (number < 128) ob_putc(b, (int)(done ? number : number + 128));
    else { size_t lsb = number & 127; ob_varint(b, number >> 7, 0); ob_varint(b, lsb, done); }
}

/* ---- RacOutput24 (rac_enc.h:28-100); out == NULL is RacDummy ---- */
typedef struct { oblob *out; uint64_t range, low; int delayed_byte, delayed_count; } rac_out;
static void racout_init(rac_out *r, oblob *out) { r->out = out; r->range = 1u << 24; r->low = 0; r->delayed_byte = -1; r->delayed_count = 0; }
static void racout_output(rac_out *r) {
    while (r->range <= (1u << 16)) {
        int byte = (int)(r->low >> 16);
        if (r->delayed_byte < 0) r->delayed_byte = byte;
        else if (((r->low + r->range) >> 8) < (1u << 16)) {
            ob_putc(r->out, r->delayed_byte);
            while (r->delayed_count) { ob_putc(r->out, 0xFF); r->delayed_count--; }
            r->delayed_byte = byte;
        } else if ((r->low >> 8) >= (1u << 16)) {
            ob_putc(r->out, r->delayed_byte + 1);
            while (r->delayed_count) { ob_putc(r->out, 0); r->delayed_count--; }
            r->delayed_byte = byte & 0xFF;
        } else r->delayed_count++;
        r->low = (r->low & ((1u << 16) - 1)) << 8;
        r->range <<= 8;
    }
}
static void racout_put(rac_out *r, uint64_t chance, int bit) {
    if (!r->out) return;
    if (bit) { r->low += r->range - chance; r->range = chance; } else r->range -= chance;
    racout_output(r);
}
static void racout_write_12bit(rac_out *r, int b12, int bit) { if (r->out) racout_put(r, (r->range * (uint64_t)b12 + 0x800) >> 12, bit); }
static void racout_write_bit(rac_out *r, int bit) { if (r->out) racout_put(r, r->range >> 1, bit); }
static void racout_flush(rac_out *r) {
    if (!r->out) return;
    r->low += (1u << 16) - 1;
    for (int k = 0; k < 4; k++) { r->range = (1u << 16) - 1; racout_output(r); }
}

/* ---- Log4kTable (chance.cpp:67-91) ---- */
static uint16_t log4k_data[4097];
static int log4k_ready = 0;
static uint32_t log4kf(int x, uint32_t base) {
    int bits = 8 * (int)sizeof(int) - __builtin_clz((unsigned)x);
    uint64_t y = ((uint64_t)x) << (32 - bits);
    uint32_t res = base * (uint32_t)(13 - bits);
    uint32_t add = base;
    while ((add > 1) && ((y & 0x7FFFFFFF) != 0)) {
        y = (((uint64_t)y) * y + 0x40000000) >> 31;
        add >>= 1;
        if ((y >> 32) != 0) { res -= add; y >>= 1; }
    }
    return res;
}
static void log4k_init(void) {
    if (log4k_ready) return;
    log4k_data[0] = 0;
    for (int i = 1; i <= 4096; i++) log4k_data[i] = (uint16_t)((log4kf(i, (uint32_t)((65535UL << 16) / 12)) + (1 << 15)) >> 16);
    log4k_ready = 1;
}
static inline void chance_estim(uint16_t chance, int bit, uint64_t *total) { *total += log4k_data[bit ? chance : 4096 - chance]; }   /* chance.h:80-82 */

/* ---- generic writer<15>(coder,min,max,value), symbol_enc.h:58-109; put(ctx, bit, index into symchance.c) ---- */
typedef void (*bit_sink)(void *ctx, int bit, int idx);
static void write_int_generic(bit_sink put, void *ctx, int min, int max, int value) {
    if (min == max) return;
    if (value == 0) { put(ctx, 1, SC_ZERO); return; }
    put(ctx, 0, SC_ZERO);
    int sign = (value > 0 ? 1 : 0);
    if (max > 0 && min < 0) put(ctx, sign, SC_SIGN);
    const int a = abs(value);
    const int e = ilog2u((uint32_t)a);
    int amax = sign ? abs(max) : abs(min);
    int emax = ilog2u((uint32_t)amax);
    int i = 0;
    while (i < emax) {
        if ((1 << (i + 1)) > amax) break;
        put(ctx, i == e, SC_EXP + i);
        if (i == e) break;
        i++;
    }
    int have = (1 << e);
    for (int pos = e; pos > 0;) {
        int bit = 1;
        --pos;
        int minabs1 = have | (1 << pos);
        if (minabs1 > amax) bit = 0;
        else { bit = (a >> pos) & 1; put(ctx, bit, SC_MANT + pos); }
        have |= (bit << pos);
    }
}
static void write_int2_generic(bit_sink put, void *ctx, int min, int max, int value) {      /* symbol.h:223-227 */
    if (min > 0) write_int_generic(put, ctx, 0, max - min, value - min);
    else if (max < 0) write_int_generic(put, ctx, min - max, 0, value - max);
    else write_int_generic(put, ctx, min, max, value);
}

/* SimpleSymbolBitCoder::write (symbol_enc.h:167-172) and FinalCompoundSymbolBitCoder::write (compound_enc.h:62-67) */
typedef struct { rac_out *rac; const chance_table *t; symchance *s; } simple_ctx;
static void simple_put(void *vc, int bit, int idx) {
    simple_ctx *c = (simple_ctx *)vc;
    racout_write_12bit(c->rac, c->s->c[idx], bit);
    c->s->c[idx] = c->t->next[c->s->c[idx]][bit];
}
static void uniform_write(rac_out *rac, int min, int max, int val) {        /* UniformSymbolCoder::write_int, symbol_enc.h:28-47 */
    if (min != 0) { max -= min; val -= min; }
    if (max == 0) return;
    int med = max / 2;
    if (val > med) { racout_write_bit(rac, 1); uniform_write(rac, med + 1, max, val); }
    else { racout_write_bit(rac, 0); uniform_write(rac, 0, med, val); }
}

/* ---- leaf of the learning pass: CompoundSymbolChances, compound_enc.h:29-59 ---- */
typedef struct {
    symchance real;
    symchance *virt;            /* [nprop][2]: first (selected when property > splitval), second */
    uint64_t realSize;
    uint64_t *virtSize;
    int64_t *virtPropSum;
    int32_t count;
    int16_t best_property;
} lleaf;
static void lleaf_init(lleaf *l, int nprop, int zero_chance) {
    symchance_init(&l->real, zero_chance);
    l->virt = (symchance *)malloc(sizeof(symchance) * 2 * (size_t)(nprop > 0 ? nprop : 1));
    for (int i = 0; i < 2 * nprop; i++) symchance_init(&l->virt[i], zero_chance);
    l->virtSize = (uint64_t *)calloc((size_t)(nprop > 0 ? nprop : 1), sizeof(uint64_t));
    l->virtPropSum = (int64_t *)calloc((size_t)(nprop > 0 ? nprop : 1), sizeof(int64_t));
    l->realSize = 0; l->count = 0; l->best_property = -1;
}
static void lleaf_copy(lleaf *d, const lleaf *s, int nprop) {
    d->real = s->real;
    d->virt = (symchance *)malloc(sizeof(symchance) * 2 * (size_t)(nprop > 0 ? nprop : 1));
    memcpy(d->virt, s->virt, sizeof(symchance) * 2 * (size_t)nprop);
    d->virtSize = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(nprop > 0 ? nprop : 1));
    memcpy(d->virtSize, s->virtSize, sizeof(uint64_t) * (size_t)nprop);
    d->virtPropSum = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nprop > 0 ? nprop : 1));
    memcpy(d->virtPropSum, s->virtPropSum, sizeof(int64_t) * (size_t)nprop);
    d->realSize = s->realSize; d->count = s->count; d->best_property = s->best_property;
}
static void lleaf_free(lleaf *l) { free(l->virt); free(l->virtSize); free(l->virtPropSum); }
static void lleaf_reset_counters(lleaf *l, int nprop) {     /* resetCounters, compound_enc.h:40-46 */
    l->best_property = -1; l->realSize = 0; l->count = 0;
    for (int i = 0; i < nprop; i++) { l->virtPropSum[i] = 0; l->virtSize[i] = 0; }
}

/* ---- PropertySymbolCoder (learning), compound_enc.h:243-518 ---- */
typedef struct {
    const chance_table *table;
    int nprop;
    int (*range)[2];
    lleaf *leaf; int nleaf, capleaf;
    tree *t;
    unsigned char *selection;
    int split_threshold;
    int (*cur)[2];              /* scratch: current_ranges */
} learner;
static inline int div_down(int64_t sum, int32_t count) {   /* compound_enc.h:256-260 */
    if (sum >= 0) return (int)(sum / count);
    return (int)-((-sum + count - 1) / count);
}
static inline int compute_splitval(const lleaf *ch, int p, int (*crange)[2]) {   /* compound_enc.h:261-284 */
    if (crange[p][0] < 0 && crange[p][1] > 0) return 0;
    int splitval = div_down(ch->virtPropSum[p], ch->count);
    if (splitval >= crange[p][1]) splitval = crange[p][1] - 1;
    return splitval;
}
/* find_leaf, compound_enc.h:307-366; returns the index of the leaf the symbol is coded with */
static int learner_find_leaf(learner *L, const int *props) {
    tree *t = L->t;
    int pos = 0;
    for (int i = 0; i < L->nprop; i++) { L->cur[i][0] = L->range[i][0]; L->cur[i][1] = L->range[i][1]; }
    while (t->n[pos].property != -1) {
        const int p = t->n[pos].property;
        if (props[p] > t->n[pos].splitval) { L->cur[p][0] = t->n[pos].splitval + 1; pos = t->n[pos].childID; }
        else { L->cur[p][1] = t->n[pos].splitval; pos = t->n[pos].childID + 1; }
    }
    int li = t->n[pos].childID;
    lleaf *result = &L->leaf[li];
    /* set_selection_and_update_property_sums, compound_enc.h:368-378 */
    result->count++;
    for (int i = 0; i < L->nprop; i++) {
        result->virtPropSum[i] += props[i];
        int splitval = compute_splitval(result, i, L->cur);
        L->selection[i] = (unsigned char)(props[i] > splitval);
    }
    const int bp = result->best_property;
    if (bp != -1 && result->realSize > result->virtSize[bp] + (uint64_t)L->split_threshold && L->nleaf < 0xFFFF && t->size < 0xFFFF &&
        L->cur[bp][0] < L->cur[bp][1]) {
        const int p = bp;
        const int splitval = compute_splitval(result, p, L->cur);
        const int new_inner = t->size;
        const tnode copy = t->n[pos];
        int a = tree_push(t); t->n[a] = copy;
        int b = tree_push(t); t->n[b] = copy;
        t->n[pos].splitval = splitval;
        t->n[pos].property = (int16_t)p;
        const int new_leaf = L->nleaf;
        lleaf_reset_counters(result, L->nprop);
        if (L->nleaf == L->capleaf) {
            L->capleaf = L->capleaf ? L->capleaf * 2 : 16;
            L->leaf = (lleaf *)realloc(L->leaf, sizeof(lleaf) * (size_t)L->capleaf);
            result = &L->leaf[li];
        }
        lleaf_copy(&L->leaf[L->nleaf], result, L->nprop);
        L->nleaf++;
        const int old_leaf = t->n[pos].childID;
        t->n[pos].childID = (uint16_t)new_inner;
        t->n[new_inner].childID = (uint16_t)old_leaf;
        t->n[new_inner + 1].childID = (uint16_t)new_leaf;
        return props[p] > t->n[pos].splitval ? old_leaf : new_leaf;
    }
    return li;
}
/* CompoundSymbolBitCoder::write with RacDummy = updateChances, compound_enc.h:91-109 */
typedef struct { learner *L; lleaf *leaf; } learn_ctx;
static void learn_put(void *vc, int bit, int idx) {
    learn_ctx *c = (learn_ctx *)vc;
    lleaf *ch = c->leaf;
    const chance_table *t = c->L->table;
    chance_estim(ch->real.c[idx], bit, &ch->realSize);
    ch->real.c[idx] = t->next[ch->real.c[idx]][bit];
    int best_property = -1;
    uint64_t best_size = ch->realSize;
    for (int j = 0; j < c->L->nprop; j++) {
        symchance *v = &ch->virt[2 * j + (c->L->selection[j] ? 0 : 1)];
        chance_estim(v->c[idx], bit, &ch->virtSize[j]);
        v->c[idx] = t->next[v->c[idx]][bit];
        if (ch->virtSize[j] < best_size) { best_size = ch->virtSize[j]; best_property = j; }
    }
    ch->best_property = (int16_t)best_property;
}
/* simplify, compound_enc.h:430-496 */
static void kill_children(tree *t, int pos) {
    if (t->n[pos].property == -1) t->n[pos].property = 0; else kill_children(t, t->n[pos].childID);
    if (t->n[pos + 1].property == -1) t->n[pos + 1].property = 0; else kill_children(t, t->n[pos + 1].childID);
}
static long long simplify_subtree(learner *L, int pos, int min_size) {
    tree *t = L->t;
    if (t->n[pos].property == -1) {
        if (L->leaf[t->n[pos].childID].count == 0) return -100;
        return L->leaf[t->n[pos].childID].count;
    }
    long long subtree_size = 0;
    subtree_size += simplify_subtree(L, t->n[pos].childID, min_size);
    subtree_size += simplify_subtree(L, t->n[pos].childID + 1, min_size);
    if (subtree_size < min_size) { t->n[pos].property = -1; kill_children(t, t->n[pos].childID); }
    return subtree_size;
}

/* MetaPropertySymbolCoder::write_subtree, compound_enc.h:523-546 */
typedef struct { rac_out *rac; const chance_table *t; symchance coder[3]; int nprop; } meta_out;
static void write_subtree(meta_out *m, const tree *t, int pos, int (*sub)[2]) {
    const tnode *n = &t->n[pos];
    const int p = n->property;
    simple_ctx c0 = {m->rac, m->t, &m->coder[0]};
    write_int2_generic(simple_put, &c0, 0, m->nprop, p + 1);
    if (p != -1) {
        const int oldmin = sub[p][0], oldmax = sub[p][1];
        simple_ctx c2 = {m->rac, m->t, &m->coder[2]};
        write_int2_generic(simple_put, &c2, oldmin, oldmax - 1, n->splitval);
        sub[p][0] = n->splitval + 1;
        write_subtree(m, t, n->childID, sub);
        sub[p][0] = oldmin;
        sub[p][1] = n->splitval;
        write_subtree(m, t, n->childID + 1, sub);
        sub[p][1] = oldmax;
    }
}

/* fuif_encode_channels<learn, compress>, encoding.cpp:74-207.  out == NULL: DummyIO / RacDummy (the learn pass). */
static int encode_channels(oblob *out, tree *t, const fo_enc_options *opt, const chance_table *table, int predictor, int beginc, int endc,
                           fo_image *img, size_t *header_pos, int learn, int compress) {
    oblob dummy = {0};
    oblob *io = out ? out : &dummy;
    ob_varint(io, (size_t)(((endc - beginc) << 4) + (predictor << 1) + (compress ? 1 : 0)), 1);
    int global_minv = LARGEST_VAL, global_maxv = SMALLEST_VAL;
    for (int i = beginc; i <= endc; i++) {
        const fo_channel *ch = &img->ch[i];
        if (ch->w * ch->h <= 0) continue;
        if (ch->minval < global_minv) global_minv = ch->minval;
        if (ch->maxval > global_maxv) global_maxv = ch->maxval;
    }
    if (global_minv <= 0) ob_varint(io, (size_t)(1 - global_minv), 1);
    else { ob_varint(io, 0, 1); ob_varint(io, (size_t)global_minv, 1); }
    ob_varint(io, (size_t)(global_maxv - global_minv), 1);
    int firstrealc = beginc;
    for (int i = beginc; i <= endc; i++) {
        fo_channel *ch = &img->ch[i];
        if (ch->w * ch->h <= 0) continue;
        const int minv = ch->minval, maxv = ch->maxval;
        if (endc > beginc && global_minv < global_maxv) { ob_varint(io, (size_t)(minv - global_minv), 1); ob_varint(io, (size_t)(maxv - minv), 1); }
        if (minv == maxv) firstrealc++;
        if (!check_bit_depth(minv, maxv, predictor)) { free(dummy.data); return 0; }
        if (minv == 0 && maxv == 0) continue;
        ob_varint(io, (size_t)ch->q, 1);
        ch_setzero(ch);
    }
    *header_pos = out ? io->pos : (size_t)-1;
    if (firstrealc > endc) { free(dummy.data); return 1; }

    int pr[64][2];
    const int nprops = init_properties(pr, img, beginc, endc, opt->max_properties);
    const int nref = nprops - NB_NONREF;
    int predictability = 2048;
    {
        const fo_channel *ch = &img->ch[firstrealc];
        if (predictor == 0 && compress) {
            uint64_t zeroes = 0, pixels = (uint64_t)(ch->h * ch->w);
            for (size_t k = 0; k < (size_t)ch->w * ch->h; k++) if (ch->data[k] == 0) zeroes++;
            int rounded = (int)(zeroes * 128 / pixels);
            if (rounded < 1) rounded = 1;
            if (rounded > 127) rounded = 127;
            ob_varint(io, (size_t)rounded, 1);
            predictability = rounded * 32;
        }
    }
    rac_out rac;
    racout_init(&rac, out);
    if (!compress) {
        for (int i = beginc; i <= endc; i++) {
            const fo_channel *ch = &img->ch[i];
            for (size_t k = 0; k < (size_t)ch->w * ch->h; k++) uniform_write(&rac, ch->minval, ch->maxval, ch->data[k]);
        }
    } else {
        learner L;
        memset(&L, 0, sizeof(L));
        symchance *fleaf = NULL;
        if (!learn) {
            static chance_table meta_table;
            static int meta_ready = 0;
            if (!meta_ready) { build_table(&meta_table, 0xFFFFFFFFu / 19, 4096 - 2); meta_ready = 1; }
            meta_out m;
            m.rac = &rac; m.t = &meta_table; m.nprop = nprops;
            for (int k = 0; k < 3; k++) symchance_init(&m.coder[k], 1024);
            int sub[64][2];
            for (int k = 0; k < nprops; k++) { sub[k][0] = pr[k][0]; sub[k][1] = pr[k][1]; }
            write_subtree(&m, t, 0, sub);
            /* FinalPropertySymbolCoder ctor, compound.h:213-225 */
            const int nleaves = (t->size + 1) / 2;
            fleaf = (symchance *)malloc(sizeof(symchance) * (size_t)nleaves);
            for (int k = 0; k < nleaves; k++) symchance_init(&fleaf[k], predictability);
            for (int k = 0, leafID = 0; k < t->size; k++) if (t->n[k].property == -1) t->n[k].childID = (uint16_t)leafID++;
        } else {
            L.table = table; L.nprop = nprops; L.range = pr; L.t = t; L.split_threshold = CONTEXT_TREE_SPLIT_THRESHOLD;
            L.capleaf = 16; L.leaf = (lleaf *)malloc(sizeof(lleaf) * 16); L.nleaf = 1;
            lleaf_init(&L.leaf[0], nprops, predictability);
            L.selection = (unsigned char *)calloc((size_t)(nprops > 0 ? nprops : 1), 1);
            L.cur = (int (*)[2])malloc(sizeof(int[2]) * 64);
        }
        int props[64];
        memset(props, 0, sizeof(props));
        for (int i = beginc; i <= endc; i++) {
            const fo_channel *ch = &img->ch[i];
            const int minv = ch->minval, maxv = ch->maxval;
            if (minv == maxv) continue;
            int rowslearned = 0;
            int *refs = (int *)calloc((size_t)(nref > 0 ? nref : 1) * (size_t)(ch->w > 0 ? ch->w : 1), sizeof(int));
            for (int y = 0; y < ch->h; y++) {
                if (learn) { if ((float)++rowslearned > opt->nb_repeats * (float)ch->h) break; }
                if (learn) y = rand() % ch->h;
                precompute_references(ch, y, img, beginc, opt->max_properties, refs, nref);
                for (int x = 0; x < ch->w; x++) {
                    for (int k = 0; k < nref; k++) props[k] = refs[x * nref + k];
                    const int guess = predict_props(props, ch, x, y, predictor, nref);
                    const int diff = S16(ch->data[(size_t)y * ch->w + x] - guess);
                    const int mn = minv - guess, mx = maxv - guess;
                    if (learn) {                    /* PropertySymbolCoder::write_int looks for (and may split) the leaf even when min == max */
                        const int li = learner_find_leaf(&L, props);
                        learn_ctx c = {&L, &L.leaf[li]};
                        write_int_generic(learn_put, &c, mn, mx, diff);
                    } else if (mn != mx) {          /* FinalPropertySymbolCoder::write_int, compound_enc.h:217-223 */
                        int pos = 0;
                        while (t->n[pos].property != -1) pos = props[t->n[pos].property] > t->n[pos].splitval ? t->n[pos].childID : t->n[pos].childID + 1;
                        simple_ctx c = {&rac, table, &fleaf[t->n[pos].childID]};
                        write_int_generic(simple_put, &c, mn, mx, diff);
                    }
                }
                if (learn) y = 0;
            }
            free(refs);
        }
        if (learn) {
            simplify_subtree(&L, 0, CONTEXT_TREE_MIN_SUBTREE_SIZE);
            for (int k = 0; k < L.nleaf; k++) lleaf_free(&L.leaf[k]);
            free(L.leaf); free(L.selection); free(L.cur);
        }
        free(fleaf);
    }
    racout_flush(&rac);
    free(dummy.data);
    return 1;
}

/* Image::recompute_downscales, image/image.cpp:124-136 */
static void recompute_downscales(const fo_image *img, int *ds) {
    ds[0] = img->nb_meta_channels + img->nb_channels - 1;
    for (int s = 1; s < 6; s++) {
        ds[s] = img->nch - 1;
        for (int k = ds[s - 1]; k < img->nch; k++) {
            const int rs = 32 >> s;
            if ((1 << img->ch[k].hcshift) < rs || (1 << img->ch[k].vcshift) < rs) break;
            if ((1 << img->ch[k].hcshift) == rs && (1 << img->ch[k].vcshift) == rs) ds[s] = k;
        }
    }
}

/* fuif_prepare_encode + fuif_encode, encoding.cpp:737-743, 455-573.  Returns a malloc'ed buffer. */
uint8_t *fo_encode(fo_image *img, const fo_enc_options *opt, size_t *nbytes) {
    *nbytes = 0;
    if (img->error) return NULL;
    log4k_init();
    srand(1);       /* the reference process calls rand() (encoding.cpp:185) from its initial state */
    fo_recompute_minmax(img);
    int downscales[6];
    recompute_downscales(img, downscales);
    chance_table *table = (chance_table *)malloc(sizeof(chance_table));
    build_table(table, (uint32_t)opt->maniac_alpha, (unsigned)(4096 - opt->maniac_cutoff));

    oblob real = {0};       /* FileIO: plain appends (no bytes_used quirk): only `pos` bytes are output */
    const char *magic = "FUIF";
    for (int k = 0; k < 4; k++) ob_putc(&real, magic[k]);
    int nb_channels = img->real_nb_channels;
    ob_varint(&real, (size_t)(nb_channels + '0'), 1);
    int bit_depth = 1, maxval = 1;
    while (maxval < img->maxval) { bit_depth++; maxval = maxval * 2 + 1; }
    ob_varint(&real, (size_t)(bit_depth + '&'), 1);
    ob_varint(&real, (size_t)(img->w - 1), 1);
    ob_varint(&real, (size_t)(img->h - 1), 1);
    ob_varint(&real, (size_t)img->colormodel, 1);
    ob_varint(&real, (size_t)opt->max_properties, 1);
    oblob io = {0};
    int ok = 1;
    if (nb_channels >= 1) {
        ob_varint(&io, (size_t)img->ntr, 1);
        for (int i = 0; i < img->ntr; i++) {
            const int id = img->tr[i].id;
            const int has_params = (id == FO_SUBSAMPLE || id == FO_PALETTE || id == FO_SQUEEZE || id == FO_DCT || id == 8 || id == 9 || id == 10);
            const int np = has_params ? img->tr[i].np : 0;
            ob_varint(&io, (size_t)((np << 4) + id), 1);
            for (int j = 0; j < np; j++) ob_varint(&io, (size_t)img->tr[i].p[j], 1);
        }
        nb_channels = img->nch;
        long long responsive_offsets[5] = {-1, -1, -1, -1, -1};
        for (int i = 0; i < nb_channels && ok; i++) {
            if (!img->ch[i].w || !img->ch[i].h) continue;
            int predictor = 0;
            if (opt->npred > i) predictor = opt->predictor[i]; else if (opt->npred > 0) predictor = opt->predictor[opt->npred - 1];
            int j = i;
            tree t = {0};
            tree_push(&t);          /* Tree(): one leaf */
            size_t header_pos = 0;
            if (!opt->compress) {
                ok = encode_channels(&io, &t, opt, table, predictor, i, j, img, &header_pos, 0, 0);
                free(t.n);
                continue;
            }
            for (int s = 1; s < 5; s++) if (j > downscales[s] && j < downscales[s + 1]) j = downscales[s + 1];
            for (int k = i + 1; k <= j; k++) if (img->ch[i].w != img->ch[k].w || img->ch[i].h != img->ch[k].h) { j = k - 1; break; }
            if (opt->max_group > 0 && j > i + opt->max_group - 1) j = i + opt->max_group - 1;
            ok = encode_channels(NULL, &t, opt, table, predictor, i, j, img, &header_pos, 1, 1);
            if (!ok) { free(t.n); break; }
            const size_t before = io.pos;
            ok = encode_channels(&io, &t, opt, table, predictor, i, j, img, &header_pos, 0, 1);
            if (!ok) { free(t.n); break; }
            size_t after = io.pos;
            float bits = (float)(after - header_pos) * 8.0f, pixels = 0.0f, ubits = 0.0f;
            for (int k = i; k <= j; k++) {
                float chpixels = (float)(img->ch[k].w * img->ch[k].h);
                float ubpp = (float)(ilog2u((uint32_t)(img->ch[k].maxval - img->ch[k].minval)) + 1);
                pixels += chpixels;
                if (img->ch[k].maxval > img->ch[k].minval) ubits += chpixels * ubpp;
            }
            if (ubits > 0.0f) ubits += 16;
            (void)pixels;
            if (bits >= ubits) {
                io.pos = before;
                ok = encode_channels(&io, &t, opt, table, predictor, i, j, img, &header_pos, 0, 0);
                after = io.pos;
            }
            for (int s = 0; s < 5; s++) if (downscales[s] >= i && downscales[s] <= j) responsive_offsets[s] = (long long)after;
            free(t.n);
            i = j;
        }
        long long relative_offset = 0;
        for (int s = 0; s < 5; s++) {
            if (responsive_offsets[s] < 0) responsive_offsets[s] = (long long)io.pos;
            long long offset = responsive_offsets[s] - relative_offset;
            ob_varint(&real, (size_t)offset, 1);        /* TRUNCATION_OFFSET_RESOLUTION == 1 */
            relative_offset = responsive_offsets[s];
        }
    }
    free(table);
    if (!ok) { free(real.data); free(io.data); return NULL; }
    /* realio gets every byte of the blob up to bytes_used (one past the last byte written, see ob_putc) */
    const size_t n = real.pos + io.used;
    uint8_t *outb = (uint8_t *)malloc(n ? n : 1);
    memcpy(outb, real.data, real.pos);
    if (io.used) memcpy(outb + real.pos, io.data, io.used);
    free(real.data); free(io.data);
    *nbytes = n;
    return outb;
}
void fo_free(void *p) { free(p); }
