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
    default: return 0;       /* subsample / palette / 2dmatch / permute / approximate: out of scope */
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
