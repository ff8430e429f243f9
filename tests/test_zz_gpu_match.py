"""GPU: files with a 2DMatch transform (exact matches, reference transform/2dmatch.h) through the C ABI against golden vectors
made by the unmodified reference: decode with a match meta-channel, every inverse step (inv_match = pointer jumping on the
GPU), responsive decodes.  The kernels are also checked on the CPU (tests/test_emu_match.py, emulator)."""
import pytest

from tests.cases import MATCH_CASES
from tests.util import gpu_plane_image, load_golden, ordered

pytestmark = pytest.mark.gpu      # hardware runs on record: GPUTEST_r01.json (XPASS), round 2 calls (passed)


@pytest.mark.parametrize("case", MATCH_CASES, ids=lambda c: c[0])
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


@pytest.mark.parametrize("case", MATCH_CASES, ids=lambda c: c[0])
def test_indexed_decode_and_full_undo(oracle, ctx, case):
    from fuif_b200 import api
    po = oracle
    blob = load_golden(case[0])
    seq = api.fuif_decode(blob["fuif"], ctx=ctx)
    par = api.fuif_decode(blob["fuif"], ctx=ctx, group_index=seq.group_index())
    par.undo_transforms(0)
    po.compare_plane_images(gpu_plane_image(po, par), po.parse_fbpd(ordered(blob, "s")[-1]), case[0] + " indexed decode + undo", check_meta=False)


@pytest.mark.parametrize("preview", [2, 4])
def test_responsive_decode(oracle, ctx, preview):
    from fuif_b200 import api
    po = oracle
    blob = load_golden("match")
    img = api.fuif_decode(blob["fuif"], api.fuif_options(preview=preview), ctx=ctx)
    img.undo_transforms(0)
    po.compare_plane_images(gpu_plane_image(po, img), po.parse_fbpd(blob[f"r{preview}"]), f"match R{preview}", check_meta=False)
