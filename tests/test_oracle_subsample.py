"""CPU: the oracle's ChromaSubsample (inverse + meta, reference transform/subsample.h) against golden vectors made by the
unmodified reference (tests/golden/make_golden_subsample.py)."""
import pytest

from tests.cases import SUBSAMPLE_CASES
from tests.util import load_golden


@pytest.mark.parametrize("case", SUBSAMPLE_CASES, ids=lambda c: c[0])
def test_inv_subsample_vs_reference(oracle, case):
    po = oracle
    name = case[0]
    blob = load_golden("sub_" + name)
    before, after = po.parse_fbpd(blob["b"]), po.parse_fbpd(blob["a"])
    img = po.OracleImage.from_plane_image(before)
    po.compare_plane_images(img.to_plane_image(), before, name + " upload")
    img.undo_transforms(len(before.transforms) - 1)
    po.compare_plane_images(img.to_plane_image(), after, name + " after inv_subsample")


@pytest.mark.parametrize("case", [c for c in SUBSAMPLE_CASES if c[7]], ids=lambda c: c[0])
def test_decode_subsampled_file_vs_reference(oracle, case):
    """meta_subsample at decode time: the planes of the reference-encoded file come out with the subsampled geometry"""
    po = oracle
    name = case[0]
    blob = load_golden("sub_" + name)
    before, after = po.parse_fbpd(blob["b"]), po.parse_fbpd(blob["a"])
    img = po.OracleImage.decode(blob["fuif"])
    po.compare_plane_images(img.to_plane_image(), before, name + " decode")
    img.undo_transforms(len(before.transforms) - 1)
    po.compare_plane_images(img.to_plane_image(), after, name + " decode + inv_subsample", check_meta=False)
