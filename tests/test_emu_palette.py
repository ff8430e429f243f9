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


def _fwd(chans, limit, table_cap=4096):
    nb = len(chans)
    planes = [np.ascontiguousarray(c.astype(np.int16)) for c in chans]
    ptrs = (C.c_void_p * nb)(*[a.ctypes.data for a in planes])
    pal = np.zeros(nb * max(1, limit) + 8, dtype=np.int16)
    count = emu_util.lib().emu_palette_fwd(ptrs, nb, planes[0].size, limit, pal.ctypes.data, table_cap)
    if count < 0:
        return None, None
    return pal[:nb * count].reshape(nb, count), planes[0]


@pytest.mark.parametrize("case", PALETTE_CASES, ids=lambda c: c[0])
def test_forward_palette_kernels_vs_reference(oracle, case):
    """collect (hash set) + sort + index against the palette and the index plane the reference's fwd_palette produced"""
    po = oracle
    blob = load_golden(case[0])
    steps = [po.parse_fbpd(b) for b in ordered(blob, "f")]
    k = next(i for i, s in enumerate(steps) if s.transforms and s.transforms[-1][0] == 6)
    before, after = steps[k - 1], steps[k]
    p = after.transforms[-1][1]
    chans = [before.planes[c].data for c in range(p[0], p[1] + 1)]
    pal, idx = _fwd(chans, 5000)
    assert pal.shape[1] == p[2]
    assert np.array_equal(pal, after.planes[0].data)
    assert np.array_equal(idx, after.planes[1 + p[0]].data)
    # exactly at the limit it still applies, one below it does not (palette.h:109)
    assert _fwd(chans, p[2])[0] is not None
    assert _fwd(chans, p[2] - 1)[0] is None


def test_forward_palette_order_collisions_and_full_table():
    """negative values sort before positive ones (signed lexicographic order of std::set<std::vector<pixel_type>>), a tiny
    table forces long probe chains, and a table smaller than the number of colours reports 'too many'"""
    rng = np.random.default_rng(8)
    colours = rng.integers(-32768, 32768, size=(50, 3)).astype(np.int16)
    colours[0] = [32767, 32767, 32767]
    colours[1] = [-32768, -32768, -32768]
    pick = rng.integers(0, 50, size=(31, 29))
    chans = [colours[pick, c] for c in range(3)]
    used = sorted({tuple(int(v) for v in colours[k]) for k in np.unique(pick)})
    for cap in (4096, 64):
        pal, idx = _fwd(chans, 64, table_cap=cap)
        assert [tuple(int(v) for v in pal[:, k]) for k in range(pal.shape[1])] == used
        assert all(tuple(int(chans[c][y, x]) for c in range(3)) == used[idx[y, x]] for y in range(31) for x in range(29))
    assert _fwd(chans, 64, table_cap=32)[0] is None
    # four channels whose colour packs to the all-ones key (the hash set's empty marker)
    ones = [np.full((3, 5), 32767, dtype=np.int16) for _ in range(4)]
    ones[3][1, 2] = -5
    pal, idx = _fwd(ones, 10)
    assert pal.shape == (4, 2) and [int(v) for v in pal[:, 1]] == [32767] * 4 and int(idx[1, 2]) == 0 and int(idx[0, 0]) == 1
