"""GPU: ChromaSubsample (reference transform/subsample.h) through the C ABI against golden vectors made by the unmodified
reference: the in-memory inverse (fb_image_undo_transforms) and the decode of reference-encoded subsampled files
(meta_subsample at decode time).  The kernel source is also checked on the CPU (tests/test_oracle_subsample.py, emulator)."""
import pytest

from tests.cases import SUBSAMPLE_CASES
from tests.util import gpu_plane_image, load_golden, upload_plane_image

pytestmark = pytest.mark.gpu      # hardware runs on record: GPUTEST_r01.json (XPASS), round 2 calls (passed)


@pytest.mark.parametrize("case", SUBSAMPLE_CASES, ids=lambda c: c[0])
def test_inv_subsample_vs_reference(oracle, ctx, case):
    from fuif_b200 import api
    po = oracle
    name = case[0]
    blob = load_golden("sub_" + name)
    before, after = po.parse_fbpd(blob["b"]), po.parse_fbpd(blob["a"])
    img = upload_plane_image(api, before, ctx)
    img.undo_transforms(len(before.transforms) - 1)
    po.compare_plane_images(gpu_plane_image(po, img), after, name + " after inv_subsample")


@pytest.mark.parametrize("case", [c for c in SUBSAMPLE_CASES if c[7]], ids=lambda c: c[0])
def test_decode_subsampled_file_vs_reference(oracle, ctx, case):
    from fuif_b200 import api
    po = oracle
    name = case[0]
    blob = load_golden("sub_" + name)
    before, after = po.parse_fbpd(blob["b"]), po.parse_fbpd(blob["a"])
    img = api.fuif_decode(blob["fuif"], ctx=ctx)
    po.compare_plane_images(gpu_plane_image(po, img), before, name + " decode")
    img.undo_transforms(len(before.transforms) - 1)
    po.compare_plane_images(gpu_plane_image(po, img), after, name + " decode + inv_subsample", check_meta=False)
