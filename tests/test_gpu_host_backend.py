"""GPU: the host-threads entropy backend behind fb_decode (FB_OPT_ENTROPY_BACKEND = FB_ENTROPY_HOST): same golden tests as the GPU
backend -- planes after the decode, planes after every inverse transform -- plus the batch call, a file that already lives in
HBM, and fb_image_upload of an image decoded without any GPU."""
import numpy as np
import pytest

from tests.cases import CASES
from tests.util import gpu_plane_image, load_golden, ordered

pytestmark = pytest.mark.gpu


@pytest.fixture()
def hctx():
    from fuif_b200 import api
    c = api.Context(0)
    c.set_entropy_backend("host", 4)
    yield c
    c.close()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_host_backend_decode_and_undo_vs_golden(oracle, hctx, case):
    from fuif_b200 import api
    po = oracle
    blob = load_golden(case[0])
    steps = [po.parse_fbpd(b) for b in ordered(blob, "s")]
    seq = api.fuif_decode(blob["fuif"], ctx=hctx)
    assert hctx.host_threads_used == 1          # no index: one stream
    po.compare_plane_images(gpu_plane_image(po, seq), steps[0], case[0] + " host-backend s0")
    img = api.fuif_decode(blob["fuif"], ctx=hctx, group_index=seq.group_index())
    po.compare_plane_images(gpu_plane_image(po, img), steps[0], case[0] + " host-backend indexed s0")
    ntr = len(steps[0].transforms)
    for k, ref in enumerate(steps[1:]):
        img.undo_transforms(ntr - 1 - k)
        po.compare_plane_images(gpu_plane_image(po, img), ref, f"{case[0]} host-backend s{k + 1}")


def test_host_backend_batch_and_device_bytes(oracle, hctx):
    import torch
    from fuif_b200 import api
    po = oracle
    names = ["sq128", "rgba14", "dct", "odd", "gray"]
    datas = [bytes(load_golden(n)["fuif"]) for n in names]
    outs = api.fuif_decode_batch(datas, ctx=hctx)
    for n, o in zip(names, outs):
        po.compare_plane_images(gpu_plane_image(po, o), po.parse_fbpd(load_golden(n)["s0"]), n + " host-backend batch")
    dev = torch.frombuffer(bytearray(datas[0]), dtype=torch.uint8).cuda()
    img = api.fuif_decode((dev.data_ptr(), dev.numel()), ctx=hctx)
    po.compare_plane_images(gpu_plane_image(po, img), po.parse_fbpd(load_golden("sq128")["s0"]), "host-backend, file in HBM")


def test_host_image_upload_then_chain(oracle, ctx):
    from fuif_b200 import api
    po = oracle
    blob = load_golden("sq128")
    steps = [po.parse_fbpd(b) for b in ordered(blob, "s")]
    img = api.fuif_host_decode(blob["fuif"], threads=2)
    with pytest.raises(api.FuifError):
        img.undo_transforms(0)
    img.upload(ctx)
    po.compare_plane_images(gpu_plane_image(po, img), steps[0], "uploaded s0")
    img.undo_transforms(0)
    po.compare_plane_images(gpu_plane_image(po, img), steps[-1], "uploaded, undone")


def test_backends_agree_on_a_1080p_image(ctx, hctx):
    """A reference-sized case through both backends: identical pixels."""
    from fuif_b200 import api
    from fuif_b200.synth import synth_image
    pix = synth_image(1920, 1080, 3, 255, 77)
    img = api.Image.from_pixels(pix, 255, ctx)
    img.recompute_minmax()
    for tid in (1, 7):
        assert img.do_transform(api.Transform(tid, []))
    data, index = api.fuif_encode(img, api.fuif_options(max_group=1, predictor=[2, 2, 2, 0]), want_index=True)
    a = api.decode_to_pixels(data, ctx=ctx, group_index=index)
    b = api.decode_to_pixels(data, ctx=hctx, group_index=index)
    assert hctx.host_threads_used == 4
    assert np.array_equal(a, b) and np.array_equal(np.asarray(a).reshape(pix.shape).astype(np.int32), pix)

