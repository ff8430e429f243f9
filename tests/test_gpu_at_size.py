"""GPU: decode + inverse chain at sizes the small golden fixtures do not reach, against the UNMODIFIED reference run on the spot
(oracle/_ref/ref_driver encodes the synthetic image, decodes it and dumps its planes; it travels with the repository).
Covers what only the benchmark exercised before: trees of thousands of nodes (leaf cache), the non-walker decode path
(value range > 256: every 14-bit plane), multi-plane groups of the DCT chain, the packed unsqueeze kernels behind a real
decode, and indexed vs sequential decode.  Bit-exact: decoded planes, final pixels."""
import os
import subprocess

import numpy as np
import pytest

from fuif_b200.synth import synth_image, write_pnm
from tests.util import gpu_plane_image

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")

CASES = [
    # name, w, h, channels, maxval, seed, ref_driver encode options
    ("cfg1_512", 512, 512, 3, 255, 1234, []),
    ("hd1080", 1920, 1080, 3, 255, 100, []),
    ("raw14_1024", 1024, 1024, 4, 16383, 9, ["-q", "12,64"]),
    ("dct_1024", 1024, 1024, 3, 255, 7, ["-C", "1", "-J", "-q", "8,12", "-G", "1"]),
    ("dct_odd_grouped", 1000, 520, 3, 255, 11, ["-C", "1", "-J", "-q", "8,12"]),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_decode_and_chain_vs_reference_at_size(oracle, ctx, case, tmp_path):
    from fuif_b200 import api
    if not os.access(REF, os.X_OK):
        pytest.skip("oracle/_ref/ref_driver is not built (needs /root/reference at build time)")
    po = oracle
    name, w, h, c, maxval, seed, opts = case
    pnm, fuif, pre = str(tmp_path / "in.pnm"), str(tmp_path / "x.fuif"), str(tmp_path / "d")
    write_pnm(pnm, synth_image(w, h, c, maxval, seed), maxval)
    subprocess.run([REF, "encode", pnm, fuif, *opts], check=True, capture_output=True)
    subprocess.run([REF, "dump", fuif, pre], check=True, capture_output=True)
    dumps = sorted((f for f in os.listdir(tmp_path) if f.startswith("d.s") and f.endswith(".fbpd")), key=lambda f: int(f[3:-5]))
    first = po.parse_fbpd(open(tmp_path / dumps[0], "rb").read())
    last = po.parse_fbpd(open(tmp_path / dumps[-1], "rb").read())
    data = open(fuif, "rb").read()
    # sequential decode (what the bare format allows), then the same with the group index it yields
    seq = api.fuif_decode(data, ctx=ctx)
    po.compare_plane_images(gpu_plane_image(po, seq), first, name + " decoded planes")
    index = seq.group_index()
    par = api.fuif_decode(data, ctx=ctx, group_index=index)
    po.compare_plane_images(gpu_plane_image(po, par), first, name + " decoded planes (indexed)")
    par.undo_transforms(0)
    po.compare_plane_images(gpu_plane_image(po, par), last, name + " final planes", check_meta=False)
    seq.undo_transforms(0)
    po.compare_plane_images(gpu_plane_image(po, seq), last, name + " final planes (sequential)", check_meta=False)
    assert ctx.pk_range_flagged == 0 or maxval > 1023


@pytest.mark.parametrize("shape", [(256, 256, 3, 255), (203, 117, 3, 255), (64, 64, 3, 1023), (1024, 520, 3, 255), (320, 200, 1, 255)])
@pytest.mark.parametrize("keep", [0, 1])
def test_fused_dct_tail_vs_oracle(oracle, ctx, shape, keep):
    """Quantize -> DCT (-> YCbCr) undone in ONE call = one fused launch (fb_idct_fused.cuh); keep = 1 leaves the colour transform
    in place (two transforms fused).  Against the oracle's transform-by-transform result, bit for bit."""
    from fuif_b200 import api
    from tests.util import default_squeeze_parameters
    po = oracle
    w, h, c, maxval = shape
    pix = synth_image(w, h, c, maxval, seed=3 * w + h)
    oi = po.OracleImage.from_pixels(pix, maxval)
    gi = api.Image.from_pixels(pix, maxval, ctx)
    q = ([8, 12, 12] if c == 3 else [8]) * 64
    sq = default_squeeze_parameters((w + 7) // 8, (h + 7) // 8, c)
    chain = ([(0, [])] if c == 3 else []) + [(4, [0, c - 1]), (5, q), (7, sq)]
    for tid, params in chain:
        assert oi.do_transform(tid, params) and gi.do_transform(api.Transform(tid, params))
    k = keep if c == 3 else 0
    n0 = ctx.launches
    gi.undo_transforms(k)
    oi.undo_transforms(k)
    po.compare_plane_images(gpu_plane_image(po, gi), oi.to_plane_image(), f"fused dct tail keep={k} {shape}")
    assert ctx.launches - n0 < 40, "the fused kernel did not take the chain (one launch per coefficient plane again?)"
