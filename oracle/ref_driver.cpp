// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// Driver around the UNMODIFIED reference (cloudinary/fuif) C++ API.  It is compiled
// by oracle/Makefile against the reference sources where they lie under
// /root/reference (nothing is copied into this repository); the binary goes to
// oracle/_ref/ref_driver (git-ignored).  It replaces the reference CLI (fuif.cpp),
// which cannot be built here because libpng / libjpeg headers are absent
// (SURVEY.md 8c), and additionally dumps the int16 planes after every single
// inverse / forward transform so that each CUDA kernel can be checked in isolation.
//
// Reference API used (all from /root/reference):
//   read_PAM_file              import/read_pam.h:46
//   write_PAM_file             export/write_pam.h:29
//   Image::do_transform        image/image.cpp:117
//   Image::undo_transforms     image/image.cpp:94
//   fuif_prepare_encode        encoding/encoding.cpp:737
//   fuif_encode_file           encoding/encoding.cpp:727
//   fuif_decode_file           encoding/encoding.cpp:745
//
// Sub-commands:
//   encode <in.pam> <out.fuif> [-C 0|1|2] [-J] [-S 0|1] [-q luma,chroma] [-E n] [-I f] [-G n] [-P digits] [-A k,q] [-L colours] [-M perm] [-D dist [-W]]
//   decode <in.fuif> <out.pam> [-R k]
//   dump   <in.fuif> <prefix>  [-R k]     planes after decode and after each inverse transform
//   fwd    <in.pam>  <prefix>  [same options as encode]   planes after each forward transform
//   time   <in.fuif> [reps] [out.pam]     JSON timing of entropy stage / transform chain; optionally writes the pixels
//   subsample <in.pam> <prefix> <p0[,p1,...]> [-F]   chroma planes decimated by this driver (the reference has no forward
//                                         subsampling), TRANSFORM_ChromaSubsample pushed with these parameters; dumps
//                                         <prefix>.b.fbpd, then the reference's inv_subsample -> <prefix>.a.fbpd;
//                                         -F: also encodes the subsampled image to <prefix>.fuif first
//
// Plane dump format ("FBPD1"): text header, then raw little-endian int16 planes:
//   FBPD1
//   image <w> <h> <minval> <maxval> <nb_channels> <real_nb_channels> <nb_meta_channels> <colormodel> <nplanes> <ntransforms>
//   transform <id> <nparams> <p0> ...              (one line per transform on the stack)
//   plane <w> <h> <minval> <maxval> <zero> <q> <hshift> <vshift> <hcshift> <vcshift> <component> <nsamples>
//   END

#include "encoding/encoding.h"
#include "import/read_pam.h"
#include "export/write_pam.h"

#include <chrono>
#include <string>
#include <cstring>
#include <cstdlib>

static bool write_dump(const std::string &fn, const Image &img) {
    FILE *f = fopen(fn.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "FBPD1\n");
    fprintf(f, "image %d %d %d %d %d %d %d %d %d %d\n", img.w, img.h, img.minval, img.maxval, img.nb_channels,
            img.real_nb_channels, img.nb_meta_channels, img.colormodel, (int)img.channel.size(), (int)img.transform.size());
    for (const Transform &t : img.transform) {
        fprintf(f, "transform %d %d", t.ID, (int)t.parameters.size());
        for (int p : t.parameters) fprintf(f, " %d", p);
        fprintf(f, "\n");
    }
    for (const Channel &c : img.channel) {
        fprintf(f, "plane %d %d %d %d %d %d %d %d %d %d %d %zu\n", c.w, c.h, (int)c.minval, (int)c.maxval, (int)c.zero, c.q,
                c.hshift, c.vshift, c.hcshift, c.vcshift, c.component, c.data.size());
    }
    fprintf(f, "END\n");
    for (const Channel &c : img.channel) {
        if (c.data.size()) fwrite(c.data.data(), sizeof(pixel_type), c.data.size(), f);
    }
    fclose(f);
    return true;
}

struct EncOpts {
    int colorspace = 2;      // 0 none, 1 YCbCr, 2 YCoCg
    bool dct = false;
    bool squeeze = true;
    int qluma = 0, qchroma = 0;  // 0 = lossless
    std::vector<int> permutation;   // -M a,b,c: TRANSFORM_PERMUTE with explicit parameters (-1 is prepended: fwd_permute's "no meta-channel" mode, permute.h:90-93)
    int match_soft = 0;      // -W: soft matches (the matched samples keep a residual that is added back, 2dmatch.h:119-129)
    int match_dist = 0;      // -D n: TRANSFORM_2DMATCH over all channels, exact matches, search distance n (fuif.cpp:440-447), before the colour transform
    int palette_colors = 0;  // -L n: all-channel TRANSFORM_PALETTE with at most n colours after the colour transform (fuif.cpp:398-407)
    int approx_k = 0, approx_q = 0;     // -A k,q: TRANSFORM_APPROXIMATE on the last k channels with divisor q+1 (fuif.cpp:504-510)
    fuif_options options = default_fuif_options;
};

static bool parse_enc_opts(int argc, char **argv, int start, EncOpts &o) {
    for (int i = start; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&](void) -> const char * { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2);} return argv[++i]; };
        if (a == "-C") o.colorspace = atoi(need());
        else if (a == "-J") o.dct = true;
        else if (a == "-S") o.squeeze = atoi(need()) != 0;
        else if (a == "-q") { if (sscanf(need(), "%d,%d", &o.qluma, &o.qchroma) != 2) return false; }
        else if (a == "-E") o.options.max_properties = atoi(need());
        else if (a == "-I") o.options.nb_repeats = atof(need());
        else if (a == "-G") o.options.max_group = atoi(need());
        else if (a == "-U") o.options.compress = false;
        else if (a == "-D") o.match_dist = atoi(need());
        else if (a == "-W") o.match_soft = 1;
        else if (a == "-L") o.palette_colors = atoi(need());
        else if (a == "-M") { const char *v = need(); char *dup = strdup(v); for (char *tok = strtok(dup, ","); tok; tok = strtok(nullptr, ",")) o.permutation.push_back(atoi(tok)); free(dup); }
        else if (a == "-A") { if (sscanf(need(), "%d,%d", &o.approx_k, &o.approx_q) != 2) return false; }
        else if (a == "-P") { const char *s = need(); while (*s) { if (*s >= '0' && *s <= '9') o.options.predictor.push_back(*s - '0'); s++; } }
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return false; }
    }
    return true;
}

// Builds the forward chain through the reference API.  The order of the transforms follows what the
// reference CLI does for PNM input (fuif.cpp:393-504): colour transform, then DCT or Squeeze, then
// Quantize, and for DCT a Squeeze of the DC planes at the end (fuif.cpp:576-578).  The quantisation
// constants are this driver's own simple schedule (q = base >> number of halvings, floor 1), not the
// CLI's quality tables; any constants give a valid stream and that is all the parity tests need.
static bool build_chain(Image &img, EncOpts &o, const std::string *dump_prefix) {
    int step = 0;
    auto dump = [&](void) { if (dump_prefix) { write_dump(*dump_prefix + ".f" + std::to_string(step) + ".fbpd", img); } step++; };
    img.recompute_minmax();
    dump();
    if (o.match_dist != 0) {
        Transform match(TRANSFORM_2DMATCH);
        match.parameters.push_back(0);
        match.parameters.push_back(img.nb_channels - 1);
        match.parameters.push_back(o.match_soft);
        match.parameters.push_back(o.match_dist);
        if (img.do_transform(match)) dump();
    }
    if (o.colorspace == 2) { if (img.do_transform(Transform(TRANSFORM_YCoCg))) dump(); }
    else if (o.colorspace == 1) { if (img.do_transform(Transform(TRANSFORM_YCbCr))) dump(); }
    if (!o.permutation.empty()) {
        Transform reorder(TRANSFORM_PERMUTE);
        reorder.parameters.push_back(-1);
        for (int c : o.permutation) reorder.parameters.push_back(c);
        if (!img.do_transform(reorder)) return false;
        dump();
    }
    if (o.palette_colors > 0 && img.nb_channels > 1) {
        Transform maybe_palette(TRANSFORM_PALETTE);
        maybe_palette.parameters.push_back(0);
        maybe_palette.parameters.push_back(img.nb_channels - 1);
        maybe_palette.parameters.push_back(o.palette_colors);
        if (img.do_transform(maybe_palette)) dump();
    }
    bool has_dct = false;
    if (o.dct) {
        Transform dct(TRANSFORM_DCT);
        dct.parameters.push_back(0);                      // explicit parameters: the CLI's empty-parameter
        dct.parameters.push_back(img.nb_channels - 1);    // path crashes in fwd_DCT
        if (!img.do_transform(dct)) return false;
        has_dct = true;
        dump();
    } else if (o.squeeze && img.channel[0].w * img.channel[0].h > 20) {
        if (!img.do_transform(Transform(TRANSFORM_SQUEEZE))) return false;
        if (o.options.max_group < 0) o.options.max_group = 1;
        dump();
    }
    if (o.qluma > 0) {
        Transform quantize(TRANSFORM_QUANTIZE);
        for (int i = 0; i < img.nb_meta_channels; i++) quantize.parameters.push_back(1);
        for (size_t i = img.nb_meta_channels; i < img.channel.size(); i++) {
            const Channel &ch = img.channel[i];
            bool chroma = (o.colorspace != 0 && ch.component > 0 && ch.component < 3);
            int q;
            if (has_dct) {
                q = chroma ? o.qchroma : o.qluma;       // flat tables
            } else {
                int shift = ch.hcshift + ch.vcshift;
                if (shift > 15) shift = 15;
                q = (chroma ? o.qchroma : o.qluma) >> shift;
            }
            if (q < 1) q = 1;
            quantize.parameters.push_back(q);
        }
        if (!img.do_transform(quantize)) return false;
        dump();
    }
    if (o.approx_k > 0) {
        Transform approximate(TRANSFORM_APPROXIMATE);
        approximate.parameters.push_back((int)img.channel.size() - o.approx_k);
        approximate.parameters.push_back((int)img.channel.size() - 1);
        approximate.parameters.push_back(o.approx_q);
        if (!img.do_transform(approximate)) return false;
        dump();
    }
    if (has_dct && o.squeeze) {
        if (!img.do_transform(Transform(TRANSFORM_SQUEEZE))) return false;
        dump();
    }
    if (o.options.predictor.size() == 0) {
        for (int i = 0; i < img.nb_meta_channels; i++) o.options.predictor.push_back(3);
        for (int i = 0; i < img.nb_channels; i++) o.options.predictor.push_back(2);
        o.options.predictor.push_back(0);
    }
    fuif_prepare_encode(img, o.options);
    dump();
    return true;
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s encode|decode|dump|fwd|time|subsample ...\n", argv[0]);
        return 2;
    }
    std::string cmd = argv[1];
    if (cmd == "encode" || cmd == "fwd") {
        if (argc < 4) return 2;
        EncOpts o;
        if (!parse_enc_opts(argc, argv, 4, o)) return 2;
        Image img = read_PAM_file(argv[2]);
        if (!img.w) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
        std::string prefix = argv[3];
        if (!build_chain(img, o, cmd == "fwd" ? &prefix : nullptr)) { fprintf(stderr, "transform chain failed\n"); return 1; }
        if (cmd == "encode") {
            double t0 = now_s();
            if (!fuif_encode_file(argv[3], img, o.options)) return 1;
            fprintf(stderr, "{\"encode_s\": %.6f, \"channels\": %d}\n", now_s() - t0, (int)img.channel.size());
        }
        return 0;
    }
    if (cmd == "decode" || cmd == "dump") {
        if (argc < 4) return 2;
        fuif_options options = default_fuif_options;
        for (int i = 4; i + 1 < argc; i++) if (!strcmp(argv[i], "-R")) options.preview = atoi(argv[i + 1]);
        Image img;
        if (!fuif_decode_file(argv[2], img, options)) { fprintf(stderr, "decode failed\n"); return 1; }
        if (cmd == "decode") {
            img.undo_transforms();
            if (img.error) return 1;
            write_PAM_file(argv[3], img);   // returns 0 on success (reference quirk Q6)
            return 0;
        }
        std::string prefix = argv[3];
        int step = 0;
        write_dump(prefix + ".s" + std::to_string(step++) + ".fbpd", img);
        while (img.transform.size() > 0) {
            img.undo_transforms((int)img.transform.size() - 1);
            if (img.error) return 1;
            write_dump(prefix + ".s" + std::to_string(step++) + ".fbpd", img);
        }
        return 0;
    }
    if (cmd == "subsample") {
        if (argc < 5) return 2;
        Image img = read_PAM_file(argv[2]);
        if (!img.w) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
        std::string prefix = argv[3];
        std::vector<int> params;
        for (char *tok = strtok(argv[4], ","); tok; tok = strtok(nullptr, ",")) params.push_back(atoi(tok));
        std::vector<int> full = params;
        if (full.size() == 1) {     // the abbreviations of check_subsample_parameters (subsample.h:33-69), spelled out for the decimation below
            static const int ab[4][2] = {{2, 2}, {2, 1}, {1, 2}, {4, 1}};
            int k = full[0];
            if (k < 0 || k > 3) return 2;
            full = {1, 2, ab[k][0], ab[k][1]};
        }
        if (full.size() % 4) return 2;
        for (size_t i = 0; i < full.size(); i += 4)
            for (int c = full[i]; c <= full[i + 1]; c++) {
                if (c < 0 || c >= (int)img.channel.size()) return 2;
                const Channel &in = img.channel[c];
                const int srh = full[i + 2], srv = full[i + 3];
                Channel out((in.w + srh - 1) / srh, (in.h + srv - 1) / srv, in.minval, in.maxval, in.q, in.hshift + (srh == 1 ? 0 : 1), in.vshift + (srv == 1 ? 0 : 1),
                            in.hcshift, in.vcshift);
                out.component = in.component;
                for (int y = 0; y < out.h; y++)
                    for (int x = 0; x < out.w; x++) {       // box average over the cell (edge cells are smaller)
                        int sum = 0, n = 0;
                        for (int dy = 0; dy < srv; dy++) for (int dx = 0; dx < srh; dx++)
                            if (y * srv + dy < in.h && x * srh + dx < in.w) { sum += in.value(y * srv + dy, x * srh + dx); n++; }
                        out.value(y, x) = (pixel_type)((sum + n / 2) / n);
                    }
                img.channel[c] = out;
            }
        Transform t(TRANSFORM_ChromaSubsample);
        t.parameters = params;
        img.transform.push_back(t);
        bool want_file = false;
        for (int i = 5; i < argc; i++) if (!strcmp(argv[i], "-F")) want_file = true;
        if (want_file) {
            fuif_options options = default_fuif_options;
            fuif_prepare_encode(img, options);
            if (!fuif_encode_file((prefix + ".fuif").c_str(), img, options)) return 1;
        }
        write_dump(prefix + ".b.fbpd", img);
        img.undo_transforms((int)img.transform.size() - 1);
        if (img.error) return 1;
        write_dump(prefix + ".a.fbpd", img);
        return 0;
    }
    if (cmd == "time") {
        int reps = argc > 3 ? atoi(argv[3]) : 1;
        double best_entropy = 1e30, best_chain = 1e30;
        int w = 0, h = 0;
        for (int r = 0; r < reps; r++) {
            Image img;
            double t0 = now_s();
            if (!fuif_decode_file(argv[2], img, default_fuif_options)) return 1;
            double t1 = now_s();
            img.undo_transforms();
            double t2 = now_s();
            if (t1 - t0 < best_entropy) best_entropy = t1 - t0;
            if (t2 - t1 < best_chain) best_chain = t2 - t1;
            w = img.w; h = img.h;
            if (r == reps - 1 && argc > 4) write_PAM_file(argv[4], img);
        }
        printf("{\"w\": %d, \"h\": %d, \"entropy_s\": %.6f, \"chain_s\": %.6f, \"total_s\": %.6f}\n", w, h, best_entropy, best_chain,
               best_entropy + best_chain);
        return 0;
    }
    fprintf(stderr, "unknown command %s\n", cmd.c_str());
    return 2;
}
