"""GPU: the Approximate transform (reference transform/approximate.h) through the C ABI against golden vectors made by the
unmodified reference: decode of files whose chain ends in Approximate (meta step at decode time), every inverse step, every
forward step, responsive decodes.  The kernels are also checked on the CPU (tests/test_emu_approximate.py, emulator)."""
import os
import tempfile

import pytest

from fuif_b200.synth import read_pnm
from tests.cases import APPROX_CASES
from tests.util import gpu_plane_image, load_golden, ordered

pytestmark = pytest.mark.gpu      # hardware runs on record: GPUTEST_r01.json (XPASS), round 2 calls (passed)


@pytest.mark.parametrize("case", APPROX_CASES, ids=lambda c: c[0])
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


@pytest.mark.parametrize("case", APPROX_CASES, ids=lambda c: c[0])
def test_forward_chain_vs_golden(oracle, ctx, case):
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
        assert img.do_transform(api.Transform(tid, params if tid in (4, 5, 10) else []))
        got = gpu_plane_image(po, img)
        if k == len(steps) - 1:
            img.recompute_minmax()
            got = gpu_plane_image(po, img)
        po.compare_plane_images(got, steps[k], f"{name} f{k}", check_meta=(k == len(steps) - 1))
        k += 1


@pytest.mark.parametrize("case", [c for c in APPROX_CASES if c[0] in ("approx", "approx_q", "approx14")], ids=lambda c: c[0])
@pytest.mark.parametrize("preview", [0, 2, 4])
def test_responsive_decode(oracle, ctx, case, preview):
    """partial decodes leave remainder channels (they come last) undecoded: approximate.h:45-55"""
    from fuif_b200 import api
    po = oracle
    blob = load_golden(case[0])
    img = api.fuif_decode(blob["fuif"], api.fuif_options(preview=preview), ctx=ctx)
    img.undo_transforms(0)
    po.compare_plane_images(gpu_plane_image(po, img), po.parse_fbpd(blob[f"r{preview}"]), f"{case[0]} R{preview}", check_meta=False)
