"""GPU: the packed int16x2 unsqueeze kernels (fb_pk_squeeze.cuh: TMA-fed horizontal step with the inverse YCoCg / clamp
epilogue, coalesced vertical step) through the C ABI -- against the oracle, against the 32-bit kernels, on smooth images
(no repairs expected), on noise inside the packed range (speculation misses repaired) and on full-range garbage (every
segment range-flagged and recomputed).  Bit-exact everywhere."""
import numpy as np
import pytest

from fuif_b200.synth import synth_image
from tests.util import default_squeeze_parameters, gpu_plane_image, upload_plane_image

pytestmark = pytest.mark.gpu


def _squeezed(po, w, h, c, maxval, seed):
    pix = synth_image(w, h, c, maxval, seed=seed)
    oi = po.OracleImage.from_pixels(pix, maxval)
    if c >= 3:
        assert oi.do_transform(1)
    assert oi.do_transform(7, default_squeeze_parameters(w, h, c))
    return pix, oi


@pytest.mark.parametrize("scale", [1, 8, 64, 512, 2047])
def test_packed_primitives_on_device(ctx, scale):
    """VIADD.16x2 / VIMNMX / VIADDMNMX formulations against the exact 32-bit pair and colour inverse, on the hardware."""
    assert ctx.selftest_packed(0, seed=scale, scale=scale) == 0
    for maxval in (255, 1023):
        assert ctx.selftest_packed(1, seed=scale, scale=scale, maxval=maxval) == 0


@pytest.mark.parametrize("shape", [(1024, 1024, 3, 255), (2048, 512, 3, 255), (1920, 1080, 3, 255), (4096, 256, 1, 255), (1024, 640, 3, 1023),
                                   (1032, 520, 4, 255), (512, 2048, 3, 255)])
def test_packed_unsqueeze_vs_oracle(oracle, shape):
    from fuif_b200 import api
    po = oracle
    w, h, c, maxval = shape
    pix, oi = _squeezed(po, w, h, c, maxval, 5 * w + h)
    pi = oi.to_plane_image()
    oi.undo_transforms(0)
    want = oi.to_plane_image()
    cx = api.Context(0)
    try:
        for packed in (True, False):
            cx.set_squeeze_packed(packed)
            n0 = cx.launches
            gi = upload_plane_image(api, pi, cx)
            gi.undo_transforms(0)
            po.compare_plane_images(gpu_plane_image(po, gi), want, f"packed={packed} {shape}")
            assert np.array_equal(gi.pixels(), pix)
            assert cx.launches > n0
        assert cx.pk_range_flagged == 0, "an 8/10-bit image left the packed range"
        assert cx.pk_repaired == 0, "speculative segment starts missed on a smooth image"
        # keep = 1: the colour transform stays, no epilogue on the last step
        if c >= 3:
            cx.set_squeeze_packed(True)
            pix2, oi2 = _squeezed(po, w, h, c, maxval, 5 * w + h)
            gi2 = upload_plane_image(api, oi2.to_plane_image(), cx)
            gi2.undo_transforms(1)
            oi2.undo_transforms(1)
            po.compare_plane_images(gpu_plane_image(po, gi2), oi2.to_plane_image(), f"keep=1 {shape}")
    finally:
        cx.close()


@pytest.mark.parametrize("amp,expect_flags", [(300, False), (900, True), (32767, True)])
def test_packed_unsqueeze_garbage(oracle, amp, expect_flags):
    """Noise in every coefficient plane: chains re-join slowly (repairs), values leave the packed range (range flags)."""
    from fuif_b200 import api
    po = oracle
    w, h = 1024, 512
    pix, oi = _squeezed(po, w, h, 3, 255, 77)
    pi = oi.to_plane_image()
    rng = np.random.default_rng(amp)
    L = po.lib()
    for i, p in enumerate(pi.planes):
        p.data = rng.integers(-amp, amp + 1, size=p.data.shape).astype(np.int16)
        a = np.ascontiguousarray(p.data)
        L.fo_plane_set(oi.h, i, a.ctypes.data, a.size)
    oi.undo_transforms(0)
    cx = api.Context(0)
    try:
        gi = upload_plane_image(api, pi, cx)
        gi.undo_transforms(0)
        po.compare_plane_images(gpu_plane_image(po, gi), oi.to_plane_image(), f"garbage {amp}")
        if expect_flags:
            assert cx.pk_range_flagged > 0 and cx.pk_repaired >= cx.pk_range_flagged
        else:
            assert cx.pk_range_flagged == 0
    finally:
        cx.close()


def test_packed_unsqueeze_repeated_runs_reuse_scratch(oracle):
    """Arrival counters are left at zero by the kernels: the second and third image on one context must come out the same."""
    from fuif_b200 import api
    po = oracle
    pix, oi = _squeezed(po, 1024, 768, 3, 255, 3)
    pi = oi.to_plane_image()
    cx = api.Context(0)
    try:
        for _ in range(3):
            gi = upload_plane_image(api, pi, cx)
            gi.undo_transforms(0)
            assert np.array_equal(gi.pixels(), pix)
    finally:
        cx.close()
