"""GPU: files with a Palette transform (reference transform/palette.h) through the C ABI against golden vectors made by the
unmodified reference: decode with a palette meta-channel (meta step at decode time, MANIAC decode of a channel with hshift -1),
every inverse step, responsive decodes.  The gather kernel is also checked on the CPU (tests/test_emu_palette.py, emulator)."""
import os
import tempfile

import pytest

from fuif_b200.synth import read_pnm
from tests.cases import PALETTE_CASES
from tests.util import gpu_plane_image, load_golden, ordered

pytestmark = pytest.mark.gpu      # hardware runs on record: GPUTEST_r01.json (XPASS), round 2 calls (passed)


@pytest.mark.parametrize("case", PALETTE_CASES, ids=lambda c: c[0])
def test_decode_and_undo_vs_golden(oracle, ctx, case):
    from fuif_b200 import api
    po = oracle
    blob = load_golden(case[0])
    steps = [po.parse_fbpd(b) for b in ordered(blob, "s")]
    img = api.fuif_decode(blob["fuif"], ctx=ctx)
    po.compare_plane_images(gpu_plane_image(po, img), steps[0], case[0] + " s0")
    ntr = len(steps[0].transforms)
    for k, ref in enumerate(steps[1:]):
        img.undo_transforms(ntr - 1 - k)
        po.compare_plane_images(gpu_plane_image(po, img), ref, f"{case[0]} s{k + 1}")


@pytest.mark.parametrize("case", PALETTE_CASES, ids=lambda c: c[0])
def test_indexed_decode_matches_sequential(oracle, ctx, case):
    from fuif_b200 import api
    po = oracle
    blob = load_golden(case[0])
    seq = api.fuif_decode(blob["fuif"], ctx=ctx)
    index = seq.group_index()
    par = api.fuif_decode(blob["fuif"], ctx=ctx, group_index=index)
    po.compare_plane_images(gpu_plane_image(po, par), po.parse_fbpd(blob["s0"]), case[0] + " indexed vs golden")


@pytest.mark.parametrize("case", [c for c in PALETTE_CASES if c[0] in ("pal", "pal4")], ids=lambda c: c[0])
@pytest.mark.parametrize("preview", [0, 2, 4])
def test_responsive_decode(oracle, ctx, case, preview):
    from fuif_b200 import api
    po = oracle
    blob = load_golden(case[0])
    img = api.fuif_decode(blob["fuif"], api.fuif_options(preview=preview), ctx=ctx)
    img.undo_transforms(0)
    po.compare_plane_images(gpu_plane_image(po, img), po.parse_fbpd(blob[f"r{preview}"]), f"{case[0]} R{preview}", check_meta=False)


@pytest.mark.parametrize("case", PALETTE_CASES, ids=lambda c: c[0])
def test_forward_chain_vs_golden(oracle, ctx, case):
    """fwd_palette on the GPU (hash-set collect, host sort of the few keys, index kernel) inside the forward chain"""
    from fuif_b200 import api
    po = oracle
    name, w, h, c, maxval, seed, opts = case
    blob = load_golden(name)
    steps = [po.parse_fbpd(b) for b in ordered(blob, "f")]
    with tempfile.NamedTemporaryFile(suffix=".pnm", delete=False) as f:
        f.write(blob["pnm"])
        path = f.name
    try:
        pix, _ = read_pnm(path)
    finally:
        os.remove(path)
    img = api.Image.from_pixels(pix, maxval, ctx)
    img.recompute_minmax()
    k = 1
    for tid, params in steps[-1].transforms:
        limit = list(params)
        if tid == 6:
            limit[2] = 5000                 # what the caller allows; the transform records what it found
        assert img.do_transform(api.Transform(tid, limit if tid in (4, 5, 6, 10) else []))
        got = gpu_plane_image(po, img)
        if k == len(steps) - 1:
            img.recompute_minmax()
            got = gpu_plane_image(po, img)
        po.compare_plane_images(got, steps[k], f"{name} f{k}", check_meta=(k == len(steps) - 1))
        k += 1
    assert [list(t.parameters) for t in img.transform if t.ID == 6] == [list(p) for t, p in steps[-1].transforms if t == 6]


def test_palette_with_too_many_colours_does_not_apply(ctx):
    from fuif_b200 import api
    from fuif_b200.synth import synth_image
    img = api.Image.from_pixels(synth_image(64, 48, 3, 255, seed=3), 255, ctx)
    assert not img.do_transform(api.Transform(6, [0, 2, 16]))
    assert len(img.transform) == 0 and img.nb_planes() == 3
