"""CPU: the product's palette gather kernel (fuif_b200/csrc/fb_palette.cuh) executed by the emulator against the reference's
dumps before / after inv_palette of every palette golden case, plus out-of-range indices (clamped, palette.h:58)."""
import ctypes as C

import numpy as np
import pytest

from tests import emu_util
from tests.cases import PALETTE_CASES
from tests.util import load_golden, ordered


def _run(index, palette):
    nb, ncolors = palette.shape
    planes = [np.ascontiguousarray(index.astype(np.int16))] + [np.full(index.shape, 0x5A5A, dtype=np.int16) for _ in range(nb - 1)]
    ptrs = (C.c_void_p * nb)(*[a.ctypes.data for a in planes])
    pal = np.ascontiguousarray(palette.astype(np.int16))
    emu_util.lib().emu_palette_inv(ptrs, nb, pal.ctypes.data, ncolors, planes[0].size)
    return planes


@pytest.mark.parametrize("case", PALETTE_CASES, ids=lambda c: c[0])
def test_palette_kernel_vs_reference(oracle, case):
    po = oracle
    blob = load_golden(case[0])
    steps = [po.parse_fbpd(b) for b in ordered(blob, "s")]
    k = next(i for i, s in enumerate(steps) if s.transforms and s.transforms[-1][0] == 6)
    before, after = steps[k], steps[k + 1]
    p = before.transforms[-1][1]
    assert before.nb_meta_channels == 1 and after.nb_meta_channels == 0
    palette = before.planes[0].data
    nb = palette.shape[0]
    got = _run(before.planes[1 + p[0]].data, palette)
    for c in range(nb):
        assert np.array_equal(got[c], after.planes[p[0] + c].data), f"{case[0]} channel {c}"


def test_palette_indices_are_clamped():
    rng = np.random.default_rng(3)
    palette = rng.integers(-3000, 3000, size=(4, 37)).astype(np.int16)
    index = rng.integers(-100, 140, size=(19, 23)).astype(np.int16)
    got = _run(index, palette)
    idx = np.clip(index.astype(np.int64), 0, 36)
    for c in range(4):
        assert np.array_equal(got[c], palette[c][idx])
