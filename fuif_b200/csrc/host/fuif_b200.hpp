// fuif_b200.hpp -- C++ host side of the drop-in boundary.
//
// Source-compatible mirror of the reference's codec API for the decode / encode / transform hot path:
//   Channel, Image            reference image/image.h:54-129
//   Transform                 reference transform/transform.h:77-106
//   fuif_options              reference encoding/encoding.h:32-59
//   fuif_decode<IO>, fuif_decode_file, Image::undo_transforms, Image::do_transform
//                             reference encoding/encoding.h:68-71, image/image.h:125-126
//   fuif_prepare_encode, fuif_encode<IO>, fuif_encode_file
//                             reference encoding/encoding.h:61-66
// Same names, same argument meaning, same "return false + message on stderr" error behaviour, so a caller written
// against the reference (fuif.cpp:206-239, fuifplay.cpp:86-88) compiles against this header unchanged.  Every method
// that touches samples forwards to the extern "C" library (include/fuif_b200.h); there is no CPU implementation here.
//
// Header-only; link with -lfuif_b200 (fuif_b200/libfuif_b200.so).
#pragma once

#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../../include/fuif_b200.h"

namespace fuif_b200 {

typedef int16_t pixel_type;     // reference image/image.h:35

#define TRANSFORM_YCbCr 0
#define TRANSFORM_YCoCg 1
#define TRANSFORM_ChromaSubsample 3
#define TRANSFORM_DCT 4
#define TRANSFORM_QUANTIZE 5
#define TRANSFORM_PALETTE 6
#define TRANSFORM_SQUEEZE 7
#define TRANSFORM_2DMATCH 8
#define TRANSFORM_PERMUTE 9
#define TRANSFORM_APPROXIMATE 10

inline fb_ctx *default_context() {
    static fb_ctx *ctx = nullptr;
    if (!ctx && fb_ctx_create(0, nullptr, &ctx) != FB_OK) {
        fprintf(stderr, "fuif_b200: no CUDA device -- this library has no CPU fallback\n");
        ctx = nullptr;
    }
    return ctx;
}

class Channel {                 // reference image/image.h:54-91
public:
    std::vector<pixel_type> data;
    int w = 0, h = 0;
    pixel_type minval = 0, maxval = 0;
    mutable pixel_type zero = 0;
    int q = 1;
    int hshift = 0, vshift = 0;
    int hcshift = 0, vcshift = 0;
    int component = -1;
    Channel() {}
    Channel(int iw, int ih, pixel_type iminval, pixel_type imaxval) : data((size_t)iw * ih, 0), w(iw), h(ih), minval(iminval), maxval(imaxval) {}
    pixel_type value(int r, int c) const { size_t i = (size_t)r * w + c; return i >= data.size() ? zero : data[i]; }
};

class Transform {               // reference transform/transform.h:77-106
public:
    const int ID;
    std::vector<int> parameters;
    Transform(int id) : ID(id) {}
    Transform &operator=(const Transform &o) { const_cast<int &>(ID) = o.ID; parameters = o.parameters; return *this; }
    Transform(const Transform &o) : ID(o.ID), parameters(o.parameters) {}
};

struct fuif_options {           // reference encoding/encoding.h:32-59 (without debug / heatmap / max_dist)
    int preview = -1;
    bool identify = false;
    float nb_repeats = 0.5f;
    int max_properties = 12;
    int maniac_cutoff = 6;
    int maniac_alpha = 0x0d000000;
    bool compress = true;
    int max_group = -1;
    std::vector<int> predictor;
};
static const fuif_options default_fuif_options{};

// IO concept of the reference (fileio.h:33-308): anything with get_c() / isEOF(); BlobReader is provided.
class BlobReader {              // reference fileio.h:83-140
    const uint8_t *data_;
    size_t size_, pos_ = 0;
public:
    const int EOS = -1;
    BlobReader(const uint8_t *d, size_t n) : data_(d), size_(n) {}
    bool isEOF() const { return pos_ >= size_; }
    long ftell() const { return (long)pos_; }
    int get_c() { return pos_ >= size_ ? EOS : data_[pos_++]; }
    const uint8_t *raw() const { return data_; }
    size_t size() const { return size_; }
    static const char *getName() { return "BlobReader"; }
};

class Image {                   // reference image/image.h:98-129
public:
    std::vector<Channel> channel;
    std::vector<Transform> transform;
    int w = 0, h = 0;
    int nb_frames = 1, den = 10, loops = 0;
    std::vector<int> num;
    int minval = 0, maxval = 255;
    int nb_channels = 0, real_nb_channels = 0, nb_meta_channels = 0;
    int colormodel = 0;
    bool error = true;

    Image() {}
    Image(int iw, int ih, int maxv, int nb_chans, int cm = 0)
        : channel(nb_chans, Channel(iw, ih, 0, (pixel_type)maxv)), w(iw), h(ih), minval(0), maxval(maxv), nb_channels(nb_chans),
          real_nb_channels(nb_chans), colormodel(cm), error(false) {
        for (int i = 0; i < nb_chans; i++) channel[i].component = i;
    }

    // Image::undo_transforms (reference image/image.cpp:94-115): upload, invert on the GPU, download.
    void undo_transforms(int keep = 0) {
        fb_image *dev = upload();
        if (!dev) { error = true; return; }
        if (fb_image_undo_transforms(dev, keep) != FB_OK) {
            fprintf(stderr, "Error while undoing transforms: %s\n", fb_last_error(default_context()));
            error = true;
        } else {
            download(dev);
        }
        fb_image_destroy(dev);
    }

    // Image::do_transform (reference image/image.cpp:117-122)
    bool do_transform(const Transform &tr) {
        fb_image *dev = upload();
        if (!dev) return false;
        int applied = 0;
        std::vector<int32_t> p(tr.parameters.begin(), tr.parameters.end());
        int rc = fb_image_do_transform(dev, tr.ID, p.data(), (int)p.size(), &applied);
        if (rc == FB_OK && applied) download(dev);
        fb_image_destroy(dev);
        return rc == FB_OK && applied;
    }

    // ---- device <-> host plumbing (not part of the reference API)
    fb_image *upload() const {
        fb_ctx *ctx = default_context();
        if (!ctx) return nullptr;
        fb_image_info info{};
        info.w = w; info.h = h; info.minval = minval; info.maxval = maxval; info.nb_channels = nb_channels; info.real_nb_channels = real_nb_channels;
        info.nb_meta_channels = nb_meta_channels; info.colormodel = colormodel; info.nb_planes = (int)channel.size(); info.nb_transforms = (int)transform.size();
        std::vector<fb_plane_desc> desc(channel.size());
        std::vector<const int16_t *> ptrs(channel.size());
        for (size_t i = 0; i < channel.size(); i++) {
            const Channel &c = channel[i];
            const bool has = c.data.size() == (size_t)c.w * c.h && c.w * c.h > 0;
            desc[i] = fb_plane_desc{c.w, c.h, c.minval, c.maxval, c.zero, c.q, c.hshift, c.vshift, c.hcshift, c.vcshift, c.component, has ? 1 : 0};
            ptrs[i] = has ? c.data.data() : nullptr;
        }
        std::vector<int32_t> ids, nps, flat;
        for (const Transform &t : transform) {
            ids.push_back(t.ID); nps.push_back((int32_t)t.parameters.size());
            flat.insert(flat.end(), t.parameters.begin(), t.parameters.end());
        }
        fb_image *dev = nullptr;
        if (fb_image_create(ctx, &info, desc.data(), ptrs.data(), ids.data(), nps.data(), flat.data(), &dev) != FB_OK) {
            fprintf(stderr, "fuif_b200: upload failed: %s\n", fb_last_error(ctx));
            return nullptr;
        }
        return dev;
    }

    bool download(fb_image *dev) {
        fb_image_info info;
        if (fb_image_get_info(dev, &info) != FB_OK) return false;
        w = info.w; h = info.h; minval = info.minval; maxval = info.maxval; nb_channels = info.nb_channels; real_nb_channels = info.real_nb_channels;
        nb_meta_channels = info.nb_meta_channels; colormodel = info.colormodel; error = info.error != 0;
        channel.assign(info.nb_planes, Channel());
        for (int i = 0; i < info.nb_planes; i++) {
            fb_plane_desc d;
            fb_image_get_plane(dev, i, &d);
            Channel &c = channel[i];
            c.w = d.w; c.h = d.h; c.minval = (pixel_type)d.minval; c.maxval = (pixel_type)d.maxval; c.zero = (pixel_type)d.zero; c.q = d.q;
            c.hshift = d.hshift; c.vshift = d.vshift; c.hcshift = d.hcshift; c.vcshift = d.vcshift; c.component = d.component;
            if (d.decoded) {
                c.data.resize((size_t)d.w * d.h);
                if (!c.data.empty() && fb_image_download_plane(dev, i, c.data.data()) != FB_OK) return false;
            }
        }
        transform.clear();
        for (int i = 0; i < info.nb_transforms; i++) {
            int32_t id = 0;
            std::vector<int32_t> p(4096);
            int n = fb_image_get_transform(dev, i, &id, p.data(), (int)p.size());
            Transform t(id);
            t.parameters.assign(p.begin(), p.begin() + (n < 0 ? 0 : n));
            transform.push_back(t);
        }
        return true;
    }
};

// fuif_decode<IO> (reference encoding/encoding.cpp:599-720): the IO object is drained into memory, the container is
// parsed on the host and the channel groups are decoded on the GPU.  Returns false on error, true (possibly with
// partially filled channels) on truncation, like the reference.
template <typename IO>
bool fuif_decode(IO &io, Image &image, fuif_options options = default_fuif_options) {
    std::vector<uint8_t> bytes;
    for (int c; (c = io.get_c()) != io.EOS;) bytes.push_back((uint8_t)c);
    fb_ctx *ctx = default_context();
    if (!ctx) return false;
    if (options.identify) {
        fb_image_info info;
        if (fb_peek_header(bytes.data(), bytes.size(), &info) != FB_OK) { fprintf(stderr, "not a FUIF file\n"); return false; }
        printf("%i-channel, %ix%i image, maxval %i\n", info.nb_channels, info.w, info.h, info.maxval);
        return true;
    }
    fb_decode_options o{options.preview, options.maniac_cutoff, options.maniac_alpha, 0};
    fb_image *dev = nullptr;
    if (fb_decode(ctx, bytes.data(), bytes.size(), &o, nullptr, nullptr, 0, &dev) != FB_OK) {
        fprintf(stderr, "Could not decode: %s\n", fb_last_error(ctx));
        return false;
    }
    bool ok = image.download(dev);
    fb_image_destroy(dev);
    return ok;
}

// fuif_decode_file (reference encoding/encoding.cpp:745-753); "-" is stdin
inline bool fuif_decode_file(const char *filename, Image &image, fuif_options options = default_fuif_options) {
    FILE *f = !strcmp(filename, "-") ? stdin : fopen(filename, "rb");
    if (!f) return false;
    std::vector<uint8_t> bytes;
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) bytes.insert(bytes.end(), buf, buf + n);
    if (f != stdin) fclose(f);
    BlobReader io(bytes.data(), bytes.size());
    return fuif_decode(io, image, options);
}

// fuif_prepare_encode (reference encoding/encoding.cpp:737-743): tight ranges.  fuif_encode() does this on the device copy
// anyway; calling it keeps code written against the reference unchanged and refreshes the host-side ranges.
inline void fuif_prepare_encode(Image &image, fuif_options &) {
    fb_image *dev = image.upload();
    if (!dev) { image.error = true; return; }
    if (fb_image_recompute_minmax(dev) == FB_OK) {
        for (size_t i = 0; i < image.channel.size(); i++) {
            fb_plane_desc d;
            if (fb_image_get_plane(dev, (int)i, &d) == FB_OK) { image.channel[i].minval = (pixel_type)d.minval; image.channel[i].maxval = (pixel_type)d.maxval; }
        }
    }
    fb_image_destroy(dev);
}

// fuif_encode<IO> (reference encoding/encoding.cpp:455-573): IO is anything with fputc(int) (FileIO / BlobIO of fileio.h).
// The channel groups are learned and coded on the GPU; the bytes are the reference encoder's.
template <typename IO>
bool fuif_encode(IO &realio, const Image &image, fuif_options &options) {
    if (image.error) return false;
    fb_ctx *ctx = default_context();
    if (!ctx) return false;
    fb_image *dev = image.upload();
    if (!dev) return false;
    std::vector<int32_t> pred(options.predictor.begin(), options.predictor.end());
    fb_encode_options o{options.nb_repeats, options.max_properties, options.maniac_cutoff, options.maniac_alpha, options.compress ? 1 : 0, options.max_group,
                        (int32_t)pred.size(), pred.empty() ? nullptr : pred.data()};
    uint8_t *bytes = nullptr;
    size_t n = 0;
    const int rc = fb_encode(ctx, dev, &o, &bytes, &n, nullptr, nullptr, 0, nullptr);
    if (rc == FB_OK)        // Channel::zero is `mutable` in the reference and set by the encoder (encoding.cpp:118)
        for (size_t i = 0; i < image.channel.size(); i++) {
            fb_plane_desc d;
            if (fb_image_get_plane(dev, (int)i, &d) == FB_OK) image.channel[i].zero = (pixel_type)d.zero;
        }
    fb_image_destroy(dev);
    if (rc != FB_OK) { fprintf(stderr, "Could not encode: %s\n", fb_last_error(ctx)); return false; }
    for (size_t i = 0; i < n; i++) realio.fputc(bytes[i]);
    fb_free(bytes);
    return true;
}

// FileIO's writing half (reference fileio.h:33-81)
class FileWriter {
    FILE *f;
public:
    explicit FileWriter(FILE *fil) : f(fil) {}
    int fputc(int c) { return ::fputc(c, f); }
};

// fuif_encode_file (reference encoding/encoding.cpp:722-735); "-" is stdout
inline bool fuif_encode_file(const char *filename, const Image &image, fuif_options &options) {
    FILE *f = !strcmp(filename, "-") ? stdout : fopen(filename, "wb");
    if (!f) return false;
    FileWriter io(f);
    const bool ok = fuif_encode(io, image, options);
    if (f != stdout) fclose(f);
    return ok;
}

}  // namespace fuif_b200
