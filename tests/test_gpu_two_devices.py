"""GPU (needs two devices): one context per GPU in ONE process.  The large-shared-memory opt-ins of the transform kernels are
per-device attributes; they used to be remembered in process-wide statics, so a context on a second GPU never got them
(ADVICE r1).  Both contexts must undo the same chain bit-exactly."""
import numpy as np
import pytest

from fuif_b200.synth import synth_image
from tests.util import default_squeeze_parameters, gpu_plane_image, upload_plane_image

pytestmark = pytest.mark.gpu


def test_two_contexts_on_two_devices(oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from fuif_b200 import api
    po = oracle
    cxs = [api.Context(1), api.Context(0)]        # device 1 FIRST: its opt-ins must not be skipped because of an earlier device
    try:
        for shape in ((640, 480, 3, 255), (1000, 333, 3, 255), (257, 513, 4, 16383), (1024, 512, 3, 255)):
            w, h, c, maxval = shape
            pix = synth_image(w, h, c, maxval, seed=w + 7 * h)
            oi = po.OracleImage.from_pixels(pix, maxval)
            if c >= 3:
                assert oi.do_transform(1)
            assert oi.do_transform(7, default_squeeze_parameters(w, h, c))
            pi = oi.to_plane_image()
            oi.undo_transforms(0)
            want = oi.to_plane_image()
            for mode in (0, 1, 4):
                for cx in cxs:
                    cx.set_squeeze_mode(mode)
                    gi = upload_plane_image(api, pi, cx)
                    gi.undo_transforms(0)
                    po.compare_plane_images(gpu_plane_image(po, gi), want, f"device ctx mode {mode} {shape}")
                    assert np.array_equal(gi.pixels(), pix)
    finally:
        for cx in cxs:
            cx.close()
