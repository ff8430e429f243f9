"""CPU tier: the packed (int16x2) per-step unsqueeze kernels (fuif_b200/csrc/fb_pk_squeeze.cuh, planned by fb_pk_plan.h)
executed by the emulator in tests/emu against the oracle's undo_transforms -- bit-exact, with and without the fused
inverse-YCoCg / clamp epilogue; the packed pair and the packed colour inverse against their exact 32-bit forms on the
whole range the kernels admit; full-range garbage (every segment range-flagged and recomputed by the exact routine);
moderate garbage (packed arithmetic valid, speculation failures repaired)."""
import numpy as np
import pytest

from fuif_b200.synth import synth_image
from tests import emu_util
from tests.test_emu_direct_squeeze import aligned_plane
from tests.util import default_squeeze_parameters

YCOCG, SQUEEZE = 1, 7


def run_case(po, pix, maxval, params, garbage=None, colour=True, ep_clamp=1, sm_count=4):
    h, w, nch = pix.shape
    img = po.OracleImage.from_pixels(pix, maxval)
    if colour and nch >= 3:
        assert img.do_transform(YCOCG)
    assert img.do_transform(SQUEEZE, params)
    pi = img.to_plane_image()
    coeff = [p.data.copy() for p in pi.planes]
    if garbage is not None:
        rng = np.random.default_rng(garbage[0])
        coeff = [rng.integers(-garbage[1], garbage[1] + 1, size=c.shape, dtype=np.int64).astype(np.int16) for c in coeff]
        for i, c in enumerate(coeff):
            a = np.ascontiguousarray(c)
            po.lib().fo_plane_set(img.h, i, a.ctypes.data, a.size)
    ref = img.clone()
    ref.undo_transforms(0)
    expect = ref.to_plane_image()
    dims = [(p.w, p.h) for p in pi.planes]
    ops, nplanes, final = emu_util.plan_inverse_squeeze(dims, params, 0, nch)
    planes = [aligned_plane(c.shape, c) for c in coeff]
    shapes = {o[4]: ((o[7] if o[1] else o[7] + o[8]), (o[5] + o[6] if o[1] else o[5])) for o in ops}
    for i in range(len(dims), nplanes):
        planes.append(aligned_plane(shapes[i], 0x5A5A))
    use_ycocg = colour and nch >= 3
    final_ids = [f[0] for f in final]
    ops10 = []
    for o in ops:
        is_final = o[4] in final_ids
        clamp = 1 if (is_final and not (use_ycocg and o[4] in final_ids[:3])) else 0
        ops10.append(tuple(o) + (clamp,))
    ep = [0] * 9
    rplane = None
    if use_ycocg:
        planes.append(aligned_plane((h, w), 0x7777))
        rplane = len(planes) - 1
        ep = [1, final_ids[0], rplane, final_ids[1], final_ids[2], maxval, 0, maxval, ep_clamp]
    st = emu_util.run_pk(planes, ops10, ep, 0, maxval, sm_count)
    got = [planes[i].copy() for i in final_ids]
    if use_ycocg:
        if st[3]:
            got[0] = planes[rplane]
        else:       # the epilogue did not ride on the last step: apply it here the way the library's separate kernel does
            Y, Co, Cg = [g.astype(np.int32) for g in got[:3]]
            Y = np.clip(Y, 0, maxval)
            G = np.clip(Y - ((-Cg) >> 1), 0, maxval)
            B = np.clip(Y + ((1 - Cg) >> 1) - (Co >> 1), 0, maxval)
            R = np.clip(Co + B, 0, maxval)
            got[:3] = [R.astype(np.int16), G.astype(np.int16), B.astype(np.int16)]
    for k in range(len(final_ids)):
        want = expect.planes[k].data
        if not np.array_equal(got[k], want):
            bad = np.argwhere(got[k] != want)
            raise AssertionError(f"plane {k} {want.shape} differs at {len(bad)} samples, first {bad[0]}: {got[k][tuple(bad[0])]} vs {want[tuple(bad[0])]}; stats {st}")
    return st


def _i16(a):
    return np.ascontiguousarray(a.astype(np.int16))


def test_packed_pair_equals_exact_pair_inside_the_admitted_range():
    rng = np.random.default_rng(11)
    n = 400000
    L = emu_util.lib()
    # averages within +-2047, residuals within +-4095, previous B anywhere a checked step can leave it (+-8189)
    for scale in (1, 8, 64, 2047):
        av = _i16(rng.integers(-min(scale * 4, 2047), min(scale * 4, 2047) + 1, n))
        nx = _i16(np.clip(av + rng.integers(-scale, scale + 1, n), -2047, 2047))
        pv = _i16(np.clip(av + rng.integers(-2 * scale, 2 * scale + 1, n), -8189, 8189))
        rs = _i16(rng.integers(-min(4 * scale, 4095), min(4 * scale, 4095) + 1, n))
        assert L.emu_check_pk_pair(pv.ctypes.data, av.ctypes.data, nx.ctypes.data, rs.ctypes.data, n) == 0
    # corners of the range
    c = np.array([-8189, -2047, -1, 0, 1, 2047, 8189])
    P, A, N, R = np.meshgrid(c, c[1:-1], c[1:-1], np.array([-4095, -1, 0, 1, 4095]), indexing="ij")
    pv, av, nx, rs = [_i16(np.repeat(x.ravel(), 2)) for x in (P, A, N, R)]
    assert L.emu_check_pk_pair(pv.ctypes.data, av.ctypes.data, nx.ctypes.data, rs.ctypes.data, len(pv)) == 0


def test_packed_ycocg_equals_exact():
    rng = np.random.default_rng(12)
    n = 300000
    L = emu_util.lib()
    for maxval in (255, 1023):
        y = _i16(rng.integers(-32768, 32768, n))
        co = _i16(rng.integers(-8189, 8190, n))
        cg = _i16(rng.integers(-8189, 8190, n))
        assert L.emu_check_pk_ycocg(y.ctypes.data, co.ctypes.data, cg.ctypes.data, n, maxval) == 0


@pytest.mark.parametrize("w,h,nch,sm", [(256, 192, 3, 4), (512, 128, 3, 2), (128, 512, 1, 4), (320, 200, 3, 8), (1024, 64, 3, 1), (272, 130, 3, 4), (1920, 136, 3, 6)])
def test_pk_unsqueeze_matches_oracle(oracle, w, h, nch, sm):
    pix = synth_image(w, h, nch, 255, seed=w * 3 + h)
    st = run_case(oracle, pix, 255, default_squeeze_parameters(w, h, nch), sm_count=sm)
    assert st[1] > 0, st
    if nch >= 3 and w % 16 == 0:
        assert st[3] == 1, st
    assert st[5] == 0, st       # nothing outside the packed range in an 8-bit image


def test_pk_unsqueeze_10bit(oracle):
    pix = synth_image(256, 128, 3, 1023, seed=3)
    st = run_case(oracle, pix, 1023, default_squeeze_parameters(256, 128, 3))
    assert st[1] > 0 and st[5] == 0, st


def test_pk_unsqueeze_full_range_garbage(oracle):
    """Coefficients anywhere in int16: every segment is range-flagged and recomputed by the exact routine."""
    pix = synth_image(256, 128, 3, 255, seed=2)
    st = run_case(oracle, pix, 255, default_squeeze_parameters(256, 128, 3), garbage=(3, 32767))
    assert st[1] > 0 and st[5] > 0 and st[4] >= st[5], st


def test_pk_unsqueeze_moderate_garbage(oracle):
    """Noise inside the packed range: the arithmetic stays packed, chains do not re-join as fast, repairs happen."""
    pix = synth_image(512, 128, 3, 255, seed=2)
    st = run_case(oracle, pix, 255, default_squeeze_parameters(512, 128, 3), garbage=(5, 900), sm_count=16)
    assert st[1] > 0, st


def test_pk_unsqueeze_noise_image(oracle):
    pix = np.random.default_rng(5).integers(0, 256, size=(160, 640, 3)).astype(np.int32)
    run_case(oracle, pix, 255, default_squeeze_parameters(640, 160, 3), sm_count=8)


@pytest.mark.parametrize("garbage", [None, (9, 32767)])
def test_pk_colour_epilogue_without_final_clamp(oracle, garbage):
    pix = synth_image(256, 128, 3, 255, seed=4)
    st = run_case(oracle, pix, 255, default_squeeze_parameters(256, 128, 3), garbage=garbage, ep_clamp=0)
    assert st[3] == 1, st
