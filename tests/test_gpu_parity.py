"""GPU: the CUDA path, called through the C ABI, against golden vectors of the unmodified reference and against the
C restatement on fresh seeded inputs.  Everything here is bit-exact (integer / byte work; the DCT and YCbCr kernels are
IEEE double without contraction, so they are bit-exact too)."""
import os
import tempfile

import numpy as np
import pytest

from tests.cases import CASES
from tests.util import default_squeeze_parameters, gpu_plane_image, load_golden, ordered, upload_plane_image
from fuif_b200.synth import read_pnm, synth_image

pytestmark = pytest.mark.gpu


def golden_pixels(blob):
    with tempfile.NamedTemporaryFile(suffix=".pnm", delete=False) as f:
        f.write(blob["pnm"])
        path = f.name
    try:
        return read_pnm(path)
    finally:
        os.remove(path)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_decode_and_undo_vs_golden(oracle, ctx, case):
    """fuif_decode on the GPU (MANIAC + context model), then every inverse transform one at a time."""
    from fuif_b200 import api
    po = oracle
    blob = load_golden(case[0])
    steps = [po.parse_fbpd(b) for b in ordered(blob, "s")]
    img = api.fuif_decode(blob["fuif"], ctx=ctx)
    po.compare_plane_images(gpu_plane_image(po, img), steps[0], case[0] + " s0")
    ntr = len(steps[0].transforms)
    for k, ref in enumerate(steps[1:]):
        img.undo_transforms(ntr - 1 - k)
        po.compare_plane_images(gpu_plane_image(po, img), ref, f"{case[0]} s{k + 1}")


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_indexed_decode_matches_sequential(oracle, ctx, case):
    """With the group index every channel group is its own stream; the planes must not change."""
    from fuif_b200 import api
    po = oracle
    blob = load_golden(case[0])
    seq = api.fuif_decode(blob["fuif"], ctx=ctx)
    offs, first = seq.group_index()
    _, ooffs = po.OracleImage.decode(blob["fuif"], want_offsets=True)
    assert list(zip(offs, first)) == [(int(a), int(b)) for a, b in ooffs]
    par = api.fuif_decode(blob["fuif"], ctx=ctx, group_index=offs)
    po.compare_plane_images(gpu_plane_image(po, par), gpu_plane_image(po, seq), case[0] + " indexed")
    par2 = api.fuif_decode(blob["fuif"], ctx=ctx, group_index=(offs, first))
    po.compare_plane_images(gpu_plane_image(po, par2), gpu_plane_image(po, seq), case[0] + " indexed+first")
    po.compare_plane_images(gpu_plane_image(po, par), po.parse_fbpd(blob["s0"]), case[0] + " indexed vs golden")


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_forward_chain_vs_golden(oracle, ctx, case):
    """Image::do_transform on the GPU: planes after every forward transform."""
    from fuif_b200 import api
    po = oracle
    name, w, h, c, maxval, seed, opts = case
    blob = load_golden(name)
    steps = [po.parse_fbpd(b) for b in ordered(blob, "f")]
    pix, mv = golden_pixels(blob)
    img = api.Image.from_pixels(pix, maxval, ctx)
    img.recompute_minmax()
    po.compare_plane_images(gpu_plane_image(po, img), steps[0], name + " f0")
    k = 1
    for tid, params in steps[-1].transforms:
        assert img.do_transform(api.Transform(tid, params if tid in (4, 5) else []))
        got = gpu_plane_image(po, img)
        if k == len(steps) - 1:
            img.recompute_minmax()      # the last dump is taken after fuif_prepare_encode (encoding.cpp:737-743)
            got = gpu_plane_image(po, img)
        po.compare_plane_images(got, steps[k], f"{name} f{k}", check_meta=(k == len(steps) - 1))
        k += 1


@pytest.mark.parametrize("case", [c for c in CASES if c[0] in ("odd", "sq128", "rgba14", "dct", "gray")], ids=lambda c: c[0])
@pytest.mark.parametrize("preview", [0, 1, 2, 3, 4])
def test_responsive_decode(oracle, ctx, case, preview):
    from fuif_b200 import api
    po = oracle
    blob = load_golden(case[0])
    img = api.fuif_decode(blob["fuif"], api.fuif_options(preview=preview), ctx=ctx)
    po.compare_plane_images(gpu_plane_image(po, img), po.parse_fbpd(blob[f"r{preview}s0"]), f"{case[0]} R{preview} s0", check_meta=False)
    img.undo_transforms(0)
    po.compare_plane_images(gpu_plane_image(po, img), po.parse_fbpd(blob[f"r{preview}"]), f"{case[0]} R{preview}", check_meta=False)


def test_decode_to_pixels_lossless_roundtrip(ctx):
    """Lossless files must reproduce the input image exactly (the reference's own self-check)."""
    from fuif_b200 import api
    for name in ("odd", "tall", "wide", "gray", "sq128", "nosq", "pred", "e0", "unc", "tiny", "one"):
        blob = load_golden(name)
        pix, maxval = golden_pixels(blob)
        out = api.decode_to_pixels(blob["fuif"], ctx=ctx)
        assert np.array_equal(out.astype(np.int32), pix), name


def test_batch_decode(oracle, ctx):
    """Several files in one launch, with and without index."""
    from fuif_b200 import api
    po = oracle
    names = ["odd", "rgba14", "dct", "sq128", "gray", "unc"]
    blobs = [load_golden(n) for n in names]
    imgs = api.fuif_decode_batch([b["fuif"] for b in blobs], ctx=ctx)
    idx = []
    for n, b, im in zip(names, blobs, imgs):
        po.compare_plane_images(gpu_plane_image(po, im), po.parse_fbpd(b["s0"]), n + " batch")
        idx.append(im.group_index()[0])
    imgs2 = api.fuif_decode_batch([b["fuif"] for b in blobs], ctx=ctx, group_indexes=idx)
    for n, b, im in zip(names, blobs, imgs2):
        po.compare_plane_images(gpu_plane_image(po, im), po.parse_fbpd(b["s0"]), n + " batch indexed")
        im.undo_transforms(0)
        steps = ordered(b, "s")
        po.compare_plane_images(gpu_plane_image(po, im), po.parse_fbpd(steps[-1]), n + " batch final")


@pytest.mark.parametrize("shape", [(512, 512, 3, 255), (1000, 333, 3, 255), (257, 513, 4, 16383), (1920, 1080, 3, 255), (64, 4096, 1, 255)])
def test_transform_chain_vs_oracle_random(oracle, ctx, shape):
    """Forward chain on the GPU == oracle, inverse chain on the GPU == oracle == original pixels (Squeeze + YCoCg)."""
    from fuif_b200 import api
    po = oracle
    w, h, c, maxval = shape
    pix = synth_image(w, h, c, maxval, seed=w + h)
    oi = po.OracleImage.from_pixels(pix, maxval)
    gi = api.Image.from_pixels(pix, maxval, ctx)
    if c >= 3:
        assert oi.do_transform(1) and gi.do_transform(api.Transform(1))
    sq = default_squeeze_parameters(w, h, c)
    assert oi.do_transform(7, sq) and gi.do_transform(api.Transform(7, sq))
    po.compare_plane_images(gpu_plane_image(po, gi), oi.to_plane_image(), f"fwd {shape}")
    gi.undo_transforms(0)
    oi.undo_transforms(0)
    po.compare_plane_images(gpu_plane_image(po, gi), oi.to_plane_image(), f"inv {shape}")
    assert np.array_equal(gi.pixels(), pix)


@pytest.mark.parametrize("shape", [(256, 256, 3, 255), (203, 117, 3, 255), (64, 64, 3, 1023)])
def test_dct_chain_vs_oracle_random(oracle, ctx, shape):
    """YCbCr + DCT + Quantize + Squeeze forward and back: double-precision kernels must match bit for bit."""
    from fuif_b200 import api
    po = oracle
    w, h, c, maxval = shape
    pix = synth_image(w, h, c, maxval, seed=3 * w + h)
    oi = po.OracleImage.from_pixels(pix, maxval)
    gi = api.Image.from_pixels(pix, maxval, ctx)
    q = [8, 12, 12] * 64
    sq = default_squeeze_parameters((w + 7) // 8, (h + 7) // 8, 3)
    for tid, params in ((0, []), (4, [0, 2]), (5, q), (7, sq)):
        assert oi.do_transform(tid, params) and gi.do_transform(api.Transform(tid, params))
        po.compare_plane_images(gpu_plane_image(po, gi), oi.to_plane_image(), f"fwd {tid} {shape}")
    for keep in (3, 2, 1, 0):
        gi.undo_transforms(keep)
        oi.undo_transforms(keep)
        po.compare_plane_images(gpu_plane_image(po, gi), oi.to_plane_image(), f"inv keep={keep} {shape}")


def test_extreme_values_wrap_like_int16(oracle, ctx):
    """pixel_type is int16 in the reference: feed planes at the edge of the range so that every wrap point is hit."""
    from fuif_b200 import api
    po = oracle
    rng = np.random.default_rng(5)
    w, h = 193, 77
    pix = rng.integers(0, 16384, size=(h, w, 3)).astype(np.int32)
    pix[::7, ::3] = 16383
    pix[3::5, 1::4] = 0
    oi = po.OracleImage.from_pixels(pix, 16383)
    gi = api.Image.from_pixels(pix, 16383, ctx)
    assert oi.do_transform(1) and gi.do_transform(api.Transform(1))
    sq = default_squeeze_parameters(w, h, 3)
    assert oi.do_transform(7, sq) and gi.do_transform(api.Transform(7, sq))
    pi = oi.to_plane_image()
    po.compare_plane_images(gpu_plane_image(po, gi), pi, "extreme fwd")
    # now corrupt the residuals with full-range noise and undo: garbage in, identical garbage out
    for p in pi.planes[3:]:
        p.data = rng.integers(-32768, 32768, size=p.data.shape).astype(np.int16)
    L = po.lib()
    for i, p in enumerate(pi.planes):
        a = np.ascontiguousarray(p.data)
        L.fo_plane_set(oi.h, i, a.ctypes.data, a.size)
    gi2 = upload_plane_image(api, pi, ctx)
    gi2.undo_transforms(0)
    oi.undo_transforms(0)
    po.compare_plane_images(gpu_plane_image(po, gi2), oi.to_plane_image(), "extreme inv")


def test_cli_decodes_like_fuif_d(ctx, tmp_path):
    """The C++ mirror API + CLI (`fuif_b200_cli -d in.fuif out.ppm`) against the reference's decoded pixels."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "fuif_b200", "fuif_b200_cli")
    assert os.path.exists(cli), "run __graft_entry__.build()"
    for name in ("sq128", "rgba14", "gray", "dct"):
        blob = load_golden(name)
        fin, fout = tmp_path / (name + ".fuif"), tmp_path / (name + ".pnm")
        fin.write_bytes(blob["fuif"])
        subprocess.run([cli, "-d", str(fin), str(fout)], check=True)
        got, _ = read_pnm(str(fout))
        final = ordered(blob, "s")[-1]
        from oracle import pyoracle as po
        ref = po.parse_fbpd(final)
        want = np.stack([p.data.astype(np.int32) for p in ref.planes[:ref.nb_channels]], axis=-1)
        assert np.array_equal(got, want), name


@pytest.mark.parametrize("mode", [0, 1, 4, 2, 3], ids=["direct", "tiled", "fused", "fused_forced_fallback", "fused_forced_repair"])
@pytest.mark.parametrize("shape", [(512, 384, 3, 255), (1000, 333, 3, 255), (257, 513, 4, 16383), (640, 480, 1, 255), (1024, 768, 4, 16383)])
def test_unsqueeze_modes_vs_oracle(oracle, shape, mode):
    """The fused tile kernels, the per-level kernels and the serial fallback kernel must all reproduce the oracle."""
    from fuif_b200 import api
    po = oracle
    c2 = api.Context(0)
    try:
        c2.set_squeeze_mode(mode)
        w, h, c, maxval = shape
        pix = synth_image(w, h, c, maxval, seed=2 * w + h)
        oi = po.OracleImage.from_pixels(pix, maxval)
        if c >= 3:
            assert oi.do_transform(1)
        sq = default_squeeze_parameters(w, h, c)
        assert oi.do_transform(7, sq)
        gi = upload_plane_image(api, oi.to_plane_image(), c2)
        gi.undo_transforms(0)
        oi.undo_transforms(0)
        po.compare_plane_images(gpu_plane_image(po, gi), oi.to_plane_image(), f"mode {mode} {shape}")
        assert np.array_equal(gi.pixels(), pix)
        if mode == 2:
            assert c2.fallbacks >= 1
        if mode == 3:
            assert c2.repaired_tiles >= 1 and c2.fallbacks == 0
        if mode == 4:
            assert c2.fallbacks == 0, "speculative tile starts failed verification on a smooth image"
        # keep = 1 (colour transform left in place): no epilogue
        if c >= 3:
            oi2 = po.OracleImage.from_pixels(pix, maxval)
            assert oi2.do_transform(1) and oi2.do_transform(7, sq)
            gi2 = upload_plane_image(api, oi2.to_plane_image(), c2)
            gi2.undo_transforms(1)
            oi2.undo_transforms(1)
            po.compare_plane_images(gpu_plane_image(po, gi2), oi2.to_plane_image(), f"mode {mode} keep=1 {shape}")
    finally:
        c2.close()


def test_fused_unsqueeze_full_range_garbage(oracle):
    """Full-range noise in every coefficient plane: wraps everywhere, speculation may fail -> verified / repaired."""
    from fuif_b200 import api
    po = oracle
    c2 = api.Context(0)
    try:
        rng = np.random.default_rng(11)
        w, h = 640, 400
        pix = synth_image(w, h, 3, 255, seed=1)
        oi = po.OracleImage.from_pixels(pix, 255)
        sq = default_squeeze_parameters(w, h, 3)
        assert oi.do_transform(1) and oi.do_transform(7, sq)
        pi = oi.to_plane_image()
        L = po.lib()
        for i, p in enumerate(pi.planes):
            p.data = rng.integers(-32768, 32768, size=p.data.shape).astype(np.int16)
            a = np.ascontiguousarray(p.data)
            L.fo_plane_set(oi.h, i, a.ctypes.data, a.size)
        oi.undo_transforms(0)
        for mode in (0, 4):
            c2.set_squeeze_mode(mode)
            gi = upload_plane_image(api, pi, c2)
            gi.undo_transforms(0)
            po.compare_plane_images(gpu_plane_image(po, gi), oi.to_plane_image(), f"garbage mode {mode}")
    finally:
        c2.close()
