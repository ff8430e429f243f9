"""CPU: the plain-C restatement (oracle/fuif_oracle.c) against golden vectors produced by the unmodified reference."""
import numpy as np
import pytest

from tests.cases import APPROX_CASES, CASES, MATCH_CASES, PALETTE_CASES, PERMUTE_CASES

ALL_CASES = CASES + APPROX_CASES + PALETTE_CASES + PERMUTE_CASES
from tests.util import load_golden, ordered
from fuif_b200.synth import read_pnm  # noqa: F401


@pytest.mark.parametrize("case", ALL_CASES + MATCH_CASES, ids=[c[0] for c in ALL_CASES + MATCH_CASES])
def test_decode_and_undo(oracle, case):
    po = oracle
    blob = load_golden(case[0])
    steps = [po.parse_fbpd(b) for b in ordered(blob, "s")]
    img = po.OracleImage.decode(blob["fuif"])
    po.compare_plane_images(img.to_plane_image(), steps[0], case[0] + " s0")
    ntr = len(steps[0].transforms)
    assert len(steps) == ntr + 1
    for k, ref in enumerate(steps[1:]):
        img.undo_transforms(ntr - 1 - k)
        po.compare_plane_images(img.to_plane_image(), ref, f"{case[0]} s{k + 1}")


@pytest.mark.parametrize("case", ALL_CASES, ids=[c[0] for c in ALL_CASES])
def test_forward_chain(oracle, case):
    po = oracle
    name, w, h, c, maxval, seed, opts = case
    blob = load_golden(name)
    steps = [po.parse_fbpd(b) for b in ordered(blob, "f")]
    import io, os, tempfile
    with tempfile.NamedTemporaryFile(suffix=".pnm", delete=False) as f:
        f.write(blob["pnm"])
        path = f.name
    try:
        pix, mv = read_pnm(path)
    finally:
        os.remove(path)
    assert mv == maxval
    oi = po.OracleImage.from_pixels(pix, maxval)
    oi.recompute_minmax()
    po.compare_plane_images(oi.to_plane_image(), steps[0], name + " f0")
    k = 1
    for tid, params in steps[-1].transforms:
        assert oi.do_transform(tid, [-1] + list(params) if tid == 9 else (params if tid in (4, 5, 6, 10) else []))
        po.compare_plane_images(oi.to_plane_image(), steps[k], f"{name} f{k}")
        k += 1


@pytest.mark.parametrize("case", [c for c in ALL_CASES + MATCH_CASES if c[0] in ("match", "odd", "sq128", "rgba14", "dct", "gray", "approx", "approx_q", "approx14", "pal", "pal4")], ids=lambda c: c[0])
@pytest.mark.parametrize("preview", [0, 1, 2, 3, 4])
def test_responsive_decode(oracle, case, preview):
    """-R k partial decodes (encoding.cpp:704-716, squeeze.h:379-383)."""
    po = oracle
    blob = load_golden(case[0])
    img = po.OracleImage.decode(blob["fuif"], preview=preview)
    po.compare_plane_images(img.to_plane_image(), po.parse_fbpd(blob[f"r{preview}s0"]), f"{case[0]} R{preview} s0", check_meta=False)
    img.undo_transforms(0)
    po.compare_plane_images(img.to_plane_image(), po.parse_fbpd(blob[f"r{preview}"]), f"{case[0]} R{preview}", check_meta=False)


def test_group_offsets_are_consistent(oracle):
    """The offsets the oracle reports are where each group's header starts: decoding is a pure function of them."""
    po = oracle
    blob = load_golden("sq128")
    img, offs = po.OracleImage.decode(blob["fuif"], want_offsets=True)
    assert len(offs) > 10
    assert all(b[0] > a[0] for a, b in zip(offs, offs[1:]))
    assert all(b[1] > a[1] for a, b in zip(offs, offs[1:]))
