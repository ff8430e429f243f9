// TSAN harness: indexed multi-threaded host decode of a file, repeated
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
typedef int (*dec_t)(const uint8_t *, size_t, const void *, const int64_t *, const int32_t *, int, int, void **);
typedef void (*des_t)(void *);
typedef int (*gi_t)(void *, int64_t *, int32_t *, int);
int main(int argc, char **argv) {
    void *h = dlopen(argv[1], RTLD_NOW);
    if (!h) { printf("%s\n", dlerror()); return 1; }
    dec_t dec = (dec_t)dlsym(h, "fb_host_decode");
    des_t des = (des_t)dlsym(h, "fb_image_destroy");
    gi_t gi = (gi_t)dlsym(h, "fb_image_group_index");
    FILE *f = fopen(argv[2], "rb");
    std::vector<uint8_t> d; uint8_t buf[65536]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + n);
    void *img = nullptr;
    if (dec(d.data(), d.size(), nullptr, nullptr, nullptr, 0, 1, &img)) return 2;
    int ng = gi(img, nullptr, nullptr, 0);
    std::vector<int64_t> offs(ng); std::vector<int32_t> first(ng);
    gi(img, offs.data(), first.data(), ng);
    des(img);
    for (int r = 0; r < atoi(argv[3]); r++) {
        if (dec(d.data(), d.size(), nullptr, offs.data(), first.data(), ng, atoi(argv[4]), &img)) return 3;
        des(img);
    }
    printf("ok, %d groups\n", ng);
    return 0;
}
