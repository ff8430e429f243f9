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


@pytest.mark.parametrize("case", SUBSAMPLE_CASES, ids=lambda c: c[0])
def test_subsample_kernel_source_vs_reference(case):
    """the product's k_inv_subsample (fuif_b200/csrc/fb_subsample.cuh) executed by the CPU emulator, plane by plane, against
    the reference's dumps; the parameter expansion mirrors fb_image.cu's subsample_parameters"""
    import numpy as np
    from oracle import pyoracle as po
    from tests import emu_util
    name, params = case[0], case[6]
    blob = load_golden("sub_" + name)
    before, after = po.parse_fbpd(blob["b"]), po.parse_fbpd(blob["a"])
    if len(params) == 1:
        params = [1, 2] + {0: [2, 2], 1: [2, 1], 2: [1, 2], 3: [4, 1]}[params[0]]
    for i in range(0, len(params), 4):
        c1, c2, srh, srv = params[i:i + 4]
        for c in range(c1, c2 + 1):
            src = np.ascontiguousarray(before.planes[c].data.astype(np.int16))
            want = after.planes[c].data
            assert want.shape == (src.shape[0] * srv, src.shape[1] * srh)
            out = np.full(want.shape, 0x5A5A, dtype=np.int16)
            emu_util.lib().emu_inv_subsample(src.ctypes.data, out.ctypes.data, src.shape[1], src.shape[0], srh, srv)
            assert np.array_equal(out, want), f"{name} plane {c}"
