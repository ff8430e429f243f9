"""CPU: the plain-C restatement of the reference ENCODER (oracle/fuif_oracle.c: fo_encode = fuif_prepare_encode + fuif_encode
with the two-pass MANIAC tree learning) against the files the unmodified reference wrote for the golden cases
(tests/golden/*.npz, entry `fuif`).  Byte-exact, which pins the tree learning (virtual chances, cost estimates, libc rand()
row order, pruning), the tree serialisation, the integer writer, the range encoder and the container layout.  This is the
checker the encode-side rows of SURVEY 8a (a20, a22, a24) will be held against; nothing in the product uses it."""
import os
import tempfile

import numpy as np
import pytest

from fuif_b200.synth import read_pnm
from tests.cases import APPROX_CASES, CASES, PALETTE_CASES, PERMUTE_CASES
from tests.util import load_golden, ordered


def _options(opts, nb_channels, transforms):
    """What oracle/ref_driver.cpp (parse_enc_opts + build_chain) hands to fuif_encode for these command-line options."""
    o = {"nb_repeats": 0.5, "max_properties": 12, "compress": True, "max_group": -1, "predictor": []}
    it = iter(opts)
    dct = False
    for a in it:
        if a == "-E": o["max_properties"] = int(next(it))
        elif a == "-I": o["nb_repeats"] = float(next(it))
        elif a == "-G": o["max_group"] = int(next(it))
        elif a == "-U": o["compress"] = False
        elif a == "-P": o["predictor"] = [int(ch) for ch in next(it) if ch.isdigit()]
        elif a == "-J": dct = True
        elif a in ("-C", "-S", "-q", "-A", "-L", "-M"): next(it)
    if not dct and any(t == 7 for t, _ in transforms) and o["max_group"] < 0:
        o["max_group"] = 1                      # build_chain: one channel per group after a Squeeze
    if not o["predictor"]:                          # build_chain: 3 for meta channels, 2 for the image's channels, then 0
        nmeta = sum(1 for t, _ in transforms if t == 6)
        nch = nb_channels - sum(p[1] - p[0] for t, p in transforms if t == 6)
        o["predictor"] = [3] * nmeta + [2] * nch + [0]
    return o


@pytest.mark.parametrize("case", CASES + APPROX_CASES + PALETTE_CASES + PERMUTE_CASES, ids=[c[0] for c in CASES + APPROX_CASES + PALETTE_CASES + PERMUTE_CASES])
def test_encoder_matches_reference_file(oracle, case):
    po = oracle
    name, w, h, c, maxval, seed, opts = case
    blob = load_golden(name)
    final = po.parse_fbpd(ordered(blob, "f")[-1])
    with tempfile.NamedTemporaryFile(suffix=".pnm", delete=False) as f:
        f.write(blob["pnm"])
        path = f.name
    try:
        pix, mv = read_pnm(path)
    finally:
        os.remove(path)
    oi = po.OracleImage.from_pixels(pix, maxval)
    oi.recompute_minmax()
    for tid, params in final.transforms:
        assert oi.do_transform(tid, [-1] + list(params) if tid == 9 else (params if tid in (4, 5, 6, 10) else []))
    o = _options(opts, c, final.transforms)
    mine = oi.encode(predictor=o["predictor"], nb_repeats=o["nb_repeats"], max_properties=o["max_properties"], compress=o["compress"],
                     max_group=o["max_group"])
    ref = bytes(blob["fuif"])
    assert len(mine) == len(ref), (len(mine), len(ref))
    # BlobIO (fileio.h:245-251) counts one byte past the last one it wrote, so the reference emits one trailing byte of
    # uninitialised buffer memory; every byte before it must match
    assert mine[:-1] == ref[:-1]
    # and the oracle's own decoder reads the pixels back
    back = po.OracleImage.decode(mine)
    back.undo_transforms(0)
    if not any(a in ("-q", "-A") for a in opts) or (name.startswith("approx") and "-q" not in opts):
        assert np.array_equal(back.pixels(), pix)


# Direct comparison with the unmodified reference where it is available (the build container): option combinations the
# golden cases do not cover -- several channels per group (-G), more / fewer learning iterations (-I), fewer reference
# properties (-E), other predictors, a 10-bit image.
LIVE = [
    ("g3", 90, 70, 3, 255, 21, ["-G", "3"]),
    ("gall", 64, 40, 3, 255, 22, ["-S", "0", "-G", "8"]),
    ("i2", 72, 56, 3, 255, 23, ["-I", "2"]),
    ("i0", 48, 48, 3, 255, 24, ["-I", "0"]),
    ("e4", 100, 60, 3, 255, 25, ["-E", "4"]),
    ("p6", 66, 50, 3, 255, 26, ["-P", "6630"]),
    ("bits10", 80, 64, 3, 1023, 27, []),
    ("gray16", 50, 70, 1, 16383, 28, ["-G", "2"]),
]


@pytest.mark.parametrize("case", LIVE, ids=[c[0] for c in LIVE])
def test_encoder_matches_live_reference(oracle, case, tmp_path):
    po = oracle
    if not po.have_ref():
        pytest.skip("oracle/_ref/ref_driver not built (needs /root/reference)")
    from fuif_b200.synth import synth_image, write_pnm
    name, w, h, c, maxval, seed, opts = case
    pix = synth_image(w, h, c, maxval, seed)
    pnm, out = str(tmp_path / "in.pnm"), str(tmp_path / "x.fuif")
    write_pnm(pnm, pix, maxval)
    po.ref_run("encode", pnm, out, *opts)
    ref = open(out, "rb").read()
    final = po.OracleImage.decode(ref).to_plane_image()         # only its transform list is used
    oi = po.OracleImage.from_pixels(pix, maxval)
    oi.recompute_minmax()
    for tid, params in final.transforms:
        assert oi.do_transform(tid, [-1] + list(params) if tid == 9 else (params if tid in (4, 5, 6, 10) else []))
    o = _options(opts, c, final.transforms)
    mine = oi.encode(predictor=o["predictor"], nb_repeats=o["nb_repeats"], max_properties=o["max_properties"], compress=o["compress"],
                     max_group=o["max_group"])
    assert len(mine) == len(ref), (len(mine), len(ref))
    assert mine[:-1] == ref[:-1]
