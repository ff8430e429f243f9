"""Times fb_encode (GPU) against the oracle encoder (CPU restatement of the reference encoder) on synthetic images and checks
that the two files are identical.  Usage: python tools/encode_bench.py [size ...]   (default 256 512 1024)
One JSON line per size.  The oracle is the checker here, never the thing shipped."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fuif_b200 import api  # noqa: E402
from fuif_b200.synth import synth_image  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [256, 512, 1024]
    po.build()
    ctx = api.Context(0)
    for n in sizes:
        pix = synth_image(n, n, 3, 255, seed=7)
        oi = po.OracleImage.from_pixels(pix, 255)
        oi.recompute_minmax()
        img = api.Image.from_pixels(pix, 255, ctx)
        img.recompute_minmax()
        for tid in (1, 7):
            assert oi.do_transform(tid, [])
            assert img.do_transform(api.Transform(tid, []))
        t0 = time.perf_counter()
        ref = oi.encode(predictor=[2, 2, 2, 0], nb_repeats=0.5, max_properties=12, compress=True, max_group=1)
        t1 = time.perf_counter()
        ctx.enable_kernel_timing(True)
        ctx.timing_report()
        mine, index = api.fuif_encode(img, api.fuif_options(max_group=1, predictor=[2, 2, 2, 0]), want_index=True)
        t2 = time.perf_counter()
        kern = {name: us for name, us, _ in ctx.timing_report() if "encode" in name}
        ctx.enable_kernel_timing(False)
        print(json.dumps({"size": n, "identical": mine == ref, "bytes": len(mine), "groups": len(index[0]), "oracle_s": round(t1 - t0, 3),
                          "fb_encode_s": round(t2 - t1, 3), "k_maniac_encode_ms": round(kern.get("k_maniac_encode", 0) / 1e3, 2),
                          "mpx_s_gpu": round(n * n / 1e6 / (t2 - t1), 3), "mpx_s_oracle": round(n * n / 1e6 / (t1 - t0), 3)}), flush=True)


if __name__ == "__main__":
    main()
