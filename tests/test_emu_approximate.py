"""CPU: the product's Approximate kernels (fuif_b200/csrc/fb_approx.cuh) executed by the emulator against the reference's
dumps: the forward step of every golden case (quotient + remainder planes) and its inverse, plus int16 wrap cases against
the oracle."""
import numpy as np
import pytest

from tests import emu_util
from tests.cases import APPROX_CASES
from tests.util import load_golden, ordered


def _steps(po, blob):
    """(planes before Approximate, planes after, parameters) from the forward dumps of a golden case"""
    dumps = [po.parse_fbpd(b) for b in ordered(blob, "f")]
    for k in range(1, len(dumps)):
        if len(dumps[k].transforms) == len(dumps[k - 1].transforms) + 1 and dumps[k].transforms[-1][0] == 10:
            return dumps[k - 1], dumps[k], dumps[k].transforms[-1][1]
    raise AssertionError("no Approximate step in the dumps")


@pytest.mark.parametrize("case", APPROX_CASES, ids=lambda c: c[0])
def test_approximate_kernels_vs_reference(oracle, case):
    po = oracle
    blob = load_golden(case[0])
    before, after, p = _steps(po, blob)
    beginc, endc = p[0], p[1]
    offset, i = len(before.planes), 0
    for c in range(beginc, endc + 1):
        q = (p[c + 2 - beginc] if c + 2 - beginc < len(p) else p[-1]) + 1
        if q == 1:
            assert np.array_equal(after.planes[c].data, before.planes[c].data)
            continue
        ch = np.ascontiguousarray(before.planes[c].data.astype(np.int16))
        chr_ = np.full(ch.shape, 0x5A5A, dtype=np.int16)
        emu_util.lib().emu_approximate(ch.ctypes.data, chr_.ctypes.data, ch.size, q, 0)
        assert np.array_equal(ch, after.planes[c].data), f"quotient of channel {c}"
        assert np.array_equal(chr_, after.planes[offset + i].data), f"remainder of channel {c}"
        emu_util.lib().emu_approximate(ch.ctypes.data, chr_.ctypes.data, ch.size, q, 1)
        assert np.array_equal(ch, before.planes[c].data), f"inverse of channel {c}"
        i += 1
    assert offset + i == len(after.planes)


@pytest.mark.parametrize("q", [2, 7, 100, 32767])
def test_approximate_inverse_wraps_like_int16(q):
    """ch * q and + remainder both narrow to int16 in the reference (pixel_type arithmetic, approximate.h:53-55)"""
    rng = np.random.default_rng(q)
    ch = rng.integers(-32768, 32768, size=4099).astype(np.int16)
    chr_ = rng.integers(-32768, 32768, size=4099).astype(np.int16)
    want = ((ch.astype(np.int64) * q).astype(np.int16).astype(np.int64) + chr_).astype(np.int16)
    got = ch.copy()
    emu_util.lib().emu_approximate(got.ctypes.data, chr_.ctypes.data, got.size, q, 1)
    assert np.array_equal(got, want)
    got = ch.copy()
    emu_util.lib().emu_approximate(got.ctypes.data, None, got.size, q, 1)
    assert np.array_equal(got, (ch.astype(np.int64) * q).astype(np.int16))


@pytest.mark.parametrize("q", [2, 3, 8, 1000])
def test_approximate_forward_is_floor_division(q):
    v = np.arange(-32768, 32768, dtype=np.int64)
    ch = v.astype(np.int16)
    chr_ = np.zeros_like(ch)
    emu_util.lib().emu_approximate(ch.ctypes.data, chr_.ctypes.data, ch.size, q, 0)
    assert np.array_equal(ch, np.floor_divide(v, q)) and np.array_equal(chr_, np.mod(v, q))
