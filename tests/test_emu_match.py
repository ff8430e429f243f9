"""CPU: the product's 2DMatch inverse (fuif_b200/csrc/fb_match.cuh: parents, pointer jumping, gather) executed by the emulator
against the reference's dumps before / after inv_match, and against a sequential model of the reference's loop on random
codes -- long chains, offsets that leave the row or the plane, and (on narrow planes) offsets that point forward."""
import ctypes as C

import numpy as np
import pytest

from tests import emu_util
from tests.cases import MATCH_CASES
from tests.util import load_golden, ordered


def _run(m, chans, maxcode, zeros):
    h, w = m.shape
    planes = [np.ascontiguousarray(c.astype(np.int16)) for c in chans]
    ptrs = (C.c_void_p * len(planes))(*[a.ctypes.data for a in planes])
    mm = np.ascontiguousarray(m.astype(np.int16))
    z = (C.c_int * len(planes))(*zeros)
    bad = emu_util.lib().emu_match_inv(mm.ctypes.data, ptrs, len(planes), w, h, maxcode, z)
    return bad, planes


def _offset(code):
    """compute_offset, reference transform/2dmatch.h:52-77"""
    layer, size = 0, 4
    while code > size:
        code -= size; layer += 1; size += 4
    if layer & 1:
        if code <= layer: return 1 + layer, -code
        if code <= 3 + 3 * layer: return 2 + 2 * layer - code, -1 - layer
        return -1 - layer, -4 - 4 * layer + code
    if code <= 1 + layer: return -1 - layer, 1 - code
    if code <= 4 + 3 * layer: return -3 - 2 * layer + code, -1 - layer
    return 1 + layer, -5 - 4 * layer + code


def _sequential(m, chan, zero, soft=False):
    """the loops of inv_match (2dmatch.h:119-141) with Channel::value's flat indexing (image.h:82-85)"""
    h, w = m.shape
    flat = chan.astype(np.int16).reshape(-1).copy()
    n = flat.size
    for y in range(h):
        for x in range(w):
            z = int(m[y, x])
            if z:
                dx, dy = _offset(z)
                src = (y + dy) * w + (x + dx)
                v = int(flat[src]) if 0 <= src < n else zero
                flat[y * w + x] = np.int16(((int(flat[y * w + x]) + v + 32768) % 65536) - 32768) if soft else v
    return flat.reshape(h, w)


@pytest.mark.parametrize("case", MATCH_CASES, ids=lambda c: c[0])
def test_match_kernels_vs_reference(oracle, case):
    po = oracle
    blob = load_golden(case[0])
    steps = [po.parse_fbpd(b) for b in ordered(blob, "s")]
    k = next(i for i, s in enumerate(steps) if s.transforms and s.transforms[-1][0] == 8)
    before, after = steps[k], steps[k + 1]
    m = before.planes[0]
    assert int((m.data != 0).sum()) > 100
    chans = [p.data for p in before.planes[1:]]
    params = before.transforms[-1][1]
    if params and params[2]:            # soft matches: the summing kernels, one channel at a time
        got = []
        mm = np.ascontiguousarray(m.data.astype(np.int16))
        for pl_ in before.planes[1:]:
            g = np.ascontiguousarray(pl_.data.astype(np.int16))
            assert emu_util.lib().emu_match_soft(mm.ctypes.data, g.ctypes.data, g.shape[1], g.shape[0], m.maxval, pl_.zero) == 0
            got.append(g)
    else:
        bad, got = _run(m.data, chans, m.maxval, [p.zero for p in before.planes[1:]])
        assert bad == 0
    for c, g in enumerate(got):
        assert np.array_equal(g, after.planes[c].data), f"{case[0]} channel {c}"


@pytest.mark.parametrize("shape", [(23, 31), (40, 3), (64, 1), (5, 200)])
def test_match_kernels_vs_sequential_model(shape):
    h, w = shape
    rng = np.random.default_rng(h * 100 + w)
    maxcode = 400
    m = np.where(rng.random((h, w)) < 0.7, rng.integers(1, maxcode + 1, size=(h, w)), 0).astype(np.int16)
    m[:, : w // 2] = np.where(rng.random((h, w // 2)) < 0.9, 1, m[:, : w // 2])     # code 1 = "left neighbour": chains as long as half a row
    chan = rng.integers(-3000, 3000, size=(h, w)).astype(np.int16)
    bad, got = _run(m, [chan, -chan], maxcode, [7, -9])
    assert bad == 0
    assert np.array_equal(got[0], _sequential(m, chan, 7))
    assert np.array_equal(got[1], _sequential(m, -chan, -9))


def test_match_code_out_of_range_is_reported():
    m = np.zeros((4, 4), dtype=np.int16)
    m[2, 2] = 50
    assert _run(m, [np.zeros((4, 4), dtype=np.int16)], 10, [0])[0] == 1


@pytest.mark.parametrize("shape", [(23, 31), (40, 3), (64, 1), (5, 200)])
def test_soft_match_kernels_vs_sequential_model(shape):
    """soft matches add the source instead of copying it (2dmatch.h:119-129): sums along the chains with int16 wrap-around"""
    h, w = shape
    rng = np.random.default_rng(h * 7 + w)
    maxcode = 400
    m = np.where(rng.random((h, w)) < 0.7, rng.integers(1, maxcode + 1, size=(h, w)), 0).astype(np.int16)
    m[:, : w // 2] = np.where(rng.random((h, w // 2)) < 0.9, 1, m[:, : w // 2])
    for zero, scale in ((0, 3000), (-11, 32767)):       # the second one wraps many times along a chain
        chan = rng.integers(-scale, scale + 1, size=(h, w)).astype(np.int16)
        got = np.ascontiguousarray(chan.copy())
        mm = np.ascontiguousarray(m)
        assert emu_util.lib().emu_match_soft(mm.ctypes.data, got.ctypes.data, w, h, maxcode, zero) == 0
        assert np.array_equal(got, _sequential(m, chan, zero, soft=True))
