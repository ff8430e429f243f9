"""Child process of tests/test_gpu_batch_many.py: decodes a batch of many small files in ONE launch with the many-streams-per-SM
shape of the MANIAC kernel forced (FB_MANIAC_SPB is read once per process) and compares every image with the oracle."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fuif_b200 import api  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests.util import gpu_plane_image, load_golden  # noqa: E402


def main():
    names = sys.argv[1].split(",")
    copies = int(sys.argv[2])
    po.build()
    ctx = api.Context(0)
    blobs = {n: load_golden(n) for n in names}
    refs = {n: po.parse_fbpd(blobs[n]["s0"]) for n in names}
    order = [names[i % len(names)] for i in range(copies * len(names))]
    seq = [api.fuif_decode(blobs[n]["fuif"], ctx=ctx) for n in names]
    index = {n: im.group_index() for n, im in zip(names, seq)}
    imgs = api.fuif_decode_batch([blobs[n]["fuif"] for n in order], ctx=ctx, group_indexes=[index[n] for n in order])
    bad = 0
    for n, im in zip(order, imgs):
        try:
            po.compare_plane_images(gpu_plane_image(po, im), refs[n], n)
        except AssertionError as e:
            bad += 1
            if bad < 3:
                print(str(e)[:300], file=sys.stderr)
    nstreams = sum(len(index[n][0]) for n in order)
    print(json.dumps({"images": len(order), "streams": nstreams, "bad": bad}))


if __name__ == "__main__":
    main()
