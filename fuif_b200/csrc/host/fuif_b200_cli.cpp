// fuif_b200 command line: the decode half of the reference CLI (fuif.cpp:206-239) on top of the C++ mirror API.
//   fuif_b200 -d [-R k] <input.fuif> <output.ppm|output.pam|null:|null_none:>
//   fuif_b200 -i <input.fuif>
// Output is binary PNM/PAM with the sample layout of the reference's write_PAM_file (export/write_pam.h:29-168).
#include <stdlib.h>

#include "fuif_b200.hpp"

using namespace fuif_b200;

static bool write_pnm(const char *fn, const Image &img) {
    const int nch = img.nb_channels;
    if (nch < 1 || nch > 4 || (int)img.channel.size() < nch) { fprintf(stderr, "cannot save %d channels as PNM\n", nch); return false; }
    const int w = img.channel[0].w, h = img.channel[0].h;
    FILE *f = fopen(fn, "wb");
    if (!f) return false;
    if (nch == 1) fprintf(f, "P5\n%u %u\n%i\n", w, h, img.maxval);
    else if (nch == 3) fprintf(f, "P6\n%u %u\n%i\n", w, h, img.maxval);
    else fprintf(f, "P7\nWIDTH %u\nHEIGHT %u\nDEPTH %d\nMAXVAL %i\nTUPLTYPE %s\nENDHDR\n", w, h, nch, img.maxval, nch == 2 ? "GRAYSCALE_ALPHA" : "RGB_ALPHA");
    std::vector<unsigned char> row((size_t)w * nch * 2);
    for (int y = 0; y < h; y++) {
        size_t k = 0;
        for (int x = 0; x < w; x++)
            for (int c = 0; c < nch; c++) {
                int v = img.channel[c].data[(size_t)y * w + x];
                if (img.maxval > 255) row[k++] = (unsigned char)(v >> 8);
                row[k++] = (unsigned char)(v & 0xFF);
            }
        fwrite(row.data(), 1, k, f);
    }
    fclose(f);
    return true;
}

int main(int argc, char **argv) {
    fuif_options options;
    bool decode = false;
    int i = 1;
    for (; i < argc && argv[i][0] == '-' && argv[i][1]; i++) {
        if (!strcmp(argv[i], "-d")) decode = true;
        else if (!strcmp(argv[i], "-i")) { options.identify = true; decode = true; }
        else if (!strcmp(argv[i], "-R") && i + 1 < argc) options.preview = atoi(argv[++i]);
        else { fprintf(stderr, "unknown option %s\n", argv[i]); return 3; }
    }
    if (!decode || argc - i < (options.identify ? 1 : 2)) {
        fprintf(stderr, "Usage: %s -d [-R k] <input.fuif> <output.ppm|null:|null_none:>\n       %s -i <input.fuif>\n(encoding is not part of the GPU hot path)\n", argv[0], argv[0]);
        return 2;
    }
    if (options.preview < -1 || options.preview > 4) { fprintf(stderr, "Invalid value for -R option (range: -1..4)\n"); return 1; }
    Image decoded;
    if (!fuif_decode_file(argv[i], decoded, options)) { fprintf(stderr, "Could not decode %s\n", argv[i]); return 1; }
    if (options.identify) return 0;
    if (!strcmp(argv[i + 1], "null_none:")) return 0;
    decoded.undo_transforms();
    if (decoded.error) return 1;
    if (!strcmp(argv[i + 1], "null:")) return 0;
    return write_pnm(argv[i + 1], decoded) ? 0 : 1;
}
