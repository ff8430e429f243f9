"""CPU tier: the direct per-step unsqueeze kernels (fuif_b200/csrc/fb_direct_squeeze.cuh, planned by fb_direct_plan.h)
executed by the emulator in tests/emu against the oracle's undo_transforms -- bit-exact, with and without the fused
inverse-YCoCg / clamp epilogue, on shapes that mix eligible and ineligible (odd, unaligned) steps."""
import numpy as np
import pytest

from fuif_b200.synth import synth_image
from tests import emu_util
from tests.util import default_squeeze_parameters

YCOCG, SQUEEZE = 1, 7


def aligned_plane(shape, fill=None):
    n = int(np.prod(shape))
    raw = np.empty(n * 2 + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 64
    a = raw[off:off + 2 * n].view(np.int16).reshape(shape)
    if fill is not None:
        a[...] = fill
    return a


def run_case(po, pix, maxval, params, garbage=None, colour=True, ep_clamp=1):
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
    last_op_of = {o[4]: k for k, o in enumerate(ops)}
    ops10 = []
    for k, o in enumerate(ops):
        is_final = o[4] in final_ids
        clamp = 1 if (is_final and not (use_ycocg and o[4] in final_ids[:3])) else 0
        ops10.append(tuple(o) + (clamp,))
    ep = [0] * 9
    rplane = None
    if use_ycocg:
        planes.append(aligned_plane((h, w), 0x7777))
        rplane = len(planes) - 1
        ep = [1, final_ids[0], rplane, final_ids[1], final_ids[2], maxval, 0, maxval, ep_clamp]
    st = emu_util.run_direct(planes, ops10, ep, 0, maxval)
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


@pytest.mark.parametrize("w,h,nch", [(256, 192, 3), (512, 128, 3), (128, 512, 1), (320, 200, 4), (131, 77, 3), (1024, 64, 3)])
def test_direct_unsqueeze_matches_oracle(oracle, w, h, nch):
    maxval = 255 if nch != 4 else 16383
    pix = synth_image(w, h, nch, maxval, seed=w * 3 + h)
    st = run_case(oracle, pix, maxval, default_squeeze_parameters(w, h, nch))
    if w % 32 == 0 and h % 32 == 0:
        assert st[1] > 0, st
        if nch >= 3:
            assert st[3] == 1, st


def test_direct_unsqueeze_full_range_garbage(oracle):
    pix = synth_image(256, 128, 3, 255, seed=2)
    st = run_case(oracle, pix, 255, default_squeeze_parameters(256, 128, 3), garbage=(3, 32767))
    assert st[1] > 0


def test_direct_unsqueeze_noise(oracle):
    pix = np.random.default_rng(5).integers(0, 256, size=(160, 640, 3)).astype(np.int32)
    run_case(oracle, pix, 255, default_squeeze_parameters(640, 160, 3))


@pytest.mark.parametrize("garbage", [None, (9, 32767)])
def test_colour_epilogue_without_final_clamp(oracle, garbage):
    """The library leaves the per-pixel final clamp out of the colour epilogue when the clamp range contains [0, maxval]
    (fb_transforms.cu): inverse YCoCg already leaves R, G, B there, so the pixels must not change -- also for garbage input."""
    pix = synth_image(256, 128, 3, 255, seed=4)
    st = run_case(oracle, pix, 255, default_squeeze_parameters(256, 128, 3), garbage=garbage, ep_clamp=0)
    assert st[3] == 1, st
