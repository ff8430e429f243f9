// Host-threads backend of the entropy stage: the same fuif_decode_channel as fb_maniac.cu, run on CPU
// threads, one channel group per thread when the caller supplies the groups' byte offsets, one image per thread otherwise.
// Selected with FB_OPT_ENTROPY_BACKEND = FB_ENTROPY_HOST; the default backend is the GPU kernel.  Not the test oracle and not
// linked to it: this file and fb_host_entropy.cpp are product code with their own implementation.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace fbh {

struct Chan {           // one plane of the channel list after meta_apply (reference image/image.h:54-91)
    int w, h, minval, maxval, zero, q, hshift, vshift;
    int16_t *data;      // w * h samples, row-major, host memory
    int state;          // 0 untouched, 1 holds samples
    int hdr_done;       // the group header of this plane has been parsed: its range is final (other threads may read it)
    int rows_done;      // rows published so far (row wavefront for planes that back-reference this one)
    long long group_off;    // byte offset of the group header if this plane starts a group, else -1
};

struct Image {
    const uint8_t *bytes;
    unsigned long long nbytes, bytes_to_load;
    Chan *ch;
    int nch, max_properties;
    int n_orig;         // planes the Image constructor made (zero-filled), the others start out as `zero` (encoding.cpp:637)
    int status;         // 0 ok, FB_ERR_INVALID = corrupt stream, FB_ERR_UNSUPPORTED
};

struct Stream {         // planes [first_channel, end_channel) starting at byte `offset`; max_groups < 0: until the file ends
    int image, first_channel, end_channel, max_groups;
    unsigned long long offset;
};

// Decodes every stream.  A stream only ever waits for planes of streams with a lower index (its back-references), and
// streams are claimed in index order, so any thread count >= 1 makes progress.  threads <= 0: four per hardware thread (at most one per stream).
// Large planes may add a look-ahead helper thread each while hardware threads are free (never when threads == 1).
// Returns the number of worker threads used.
int decode(Image *images, int nimages, const Stream *streams, int nstreams, int cutoff, uint32_t alpha, int threads);

}  // namespace fbh
