"""GPU: files with a Palette transform (reference transform/palette.h) through the C ABI against golden vectors made by the
unmodified reference: decode with a palette meta-channel (meta step at decode time, MANIAC decode of a channel with hshift -1),
every inverse step, responsive decodes.  The gather kernel is also checked on the CPU (tests/test_emu_palette.py, emulator)."""
import pytest

from tests.cases import PALETTE_CASES
from tests.util import gpu_plane_image, load_golden, ordered

pytestmark = [pytest.mark.gpu,
              # Not strict: a pass is reported as XPASS.  Written after the round's GPU budget was spent; the marker goes away with the first recorded hardware run.
              pytest.mark.xfail(strict=False, reason="Palette has been verified under the CPU emulator only; this is its first run on hardware")]


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
