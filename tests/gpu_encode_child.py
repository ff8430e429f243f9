"""Child process of tests/test_zz_gpu_encode.py: runs fb_encode() on cuda:0 for the named cases and prints one JSON line per
case.  A separate process, because this kernel had its first GPU run after the round's GPU budget was spent: whatever it does
on real hardware (a sticky CUDA error, a hang) must not take the parity session of the other tests with it."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fuif_b200 import api  # noqa: E402
from fuif_b200.synth import read_pnm, synth_image  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests.cases import APPROX_CASES, CASES, PALETTE_CASES  # noqa: E402
from tests.test_oracle_encoder import _options  # noqa: E402
from tests.util import gpu_plane_image, load_golden, ordered  # noqa: E402


def _first_diff(a, b):
    n = min(len(a), len(b))
    return next((i for i in range(n) if a[i] != b[i]), n)


def check(name, pix, maxval, transforms, o, ctx, golden_file=None):
    """transforms: [(id, params)] to apply forward on both sides; o: options dict of tests.test_oracle_encoder._options"""
    t0 = time.time()
    oi = po.OracleImage.from_pixels(pix, maxval)
    oi.recompute_minmax()
    img = api.Image.from_pixels(pix, maxval, ctx)
    img.recompute_minmax()
    for tid, params in transforms:
        assert oi.do_transform(tid, params)
        assert img.do_transform(api.Transform(tid, params))
    ref = oi.encode(predictor=o["predictor"], nb_repeats=o["nb_repeats"], max_properties=o["max_properties"], compress=o["compress"], max_group=o["max_group"])
    t1 = time.time()
    opt = api.fuif_options(nb_repeats=o["nb_repeats"], max_properties=o["max_properties"], compress=o["compress"], max_group=o["max_group"],
                           predictor=list(o["predictor"]))
    mine, index = api.fuif_encode(img, opt, want_index=True)
    t2 = time.time()
    res = {"case": name, "ok": False, "bytes": len(mine), "ref_bytes": len(ref), "oracle_s": round(t1 - t0, 3), "gpu_s": round(t2 - t1, 3)}
    if mine != ref:
        res["why"] = f"file differs from the oracle encoder's at byte {_first_diff(mine, ref)} ({len(mine)} vs {len(ref)} bytes)"
        return res
    if golden_file is not None and (len(golden_file) != len(mine) or golden_file[:-1] != mine[:-1]):
        res["why"] = f"file differs from the reference's at byte {_first_diff(mine, golden_file)}"
        return res
    # the returned sidecar index is the one a sequential decode discovers, and the indexed GPU decode of our own file gives
    # back the planes we encoded
    seq = api.fuif_decode(mine, ctx=ctx)
    if seq.group_index() != index:
        res["why"] = f"group index {index} vs {seq.group_index()}"
        return res
    par = api.fuif_decode(mine, ctx=ctx, group_index=index)
    try:
        po.compare_plane_images(gpu_plane_image(po, par), gpu_plane_image(po, img), name + " decode(encode(image))", check_meta=False)
    except AssertionError as e:
        res["why"] = str(e)[:300]
        return res
    res["ok"] = True
    return res


def main(names):
    ctx = api.Context(0)
    for name in names:
        try:
            if name.startswith("synth"):
                w, h = {"synth256": (256, 256), "synth512": (512, 512), "synth1000x333": (1000, 333)}[name]
                pix = synth_image(w, h, 3, 255, seed=11)
                o = {"nb_repeats": 0.5, "max_properties": 12, "compress": True, "max_group": 1, "predictor": [2, 2, 2, 0]}
                res = check(name, pix, 255, [(1, []), (7, [])], o, ctx)
            elif name.startswith("noise"):
                rng = np.random.default_rng(5)
                pix = rng.integers(0, 256, size=(61, 67, 3)).astype(np.int32)
                o = {"nb_repeats": 0.5, "max_properties": 12, "compress": True, "max_group": -1, "predictor": [2, 2, 2, 0]}
                res = check(name, pix, 255, [], o, ctx)
            else:
                case = next(c for c in CASES + APPROX_CASES + PALETTE_CASES if c[0] == name)
                _n, w, h, c, maxval, seed, opts = case
                blob = load_golden(name)
                final = po.parse_fbpd(ordered(blob, "f")[-1])
                with tempfile.NamedTemporaryFile(suffix=".pnm", delete=False) as f:
                    f.write(blob["pnm"])
                    path = f.name
                try:
                    pix, _ = read_pnm(path)
                finally:
                    os.remove(path)
                trs = [(tid, params if tid in (4, 5, 6, 10) else []) for tid, params in final.transforms]
                res = check(name, pix, maxval, trs, _options(opts, c, final.transforms), ctx, golden_file=bytes(blob["fuif"]))
        except Exception as e:      # noqa: BLE001 -- reported to the parent, which fails the case
            res = {"case": name, "ok": False, "why": f"{type(e).__name__}: {e}"[:400]}
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    po.build()
    main(sys.argv[1:])
