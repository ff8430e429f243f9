"""CPU tier: the product's fused unsqueeze kernel SOURCE (fuif_b200/csrc/fb_fused_squeeze.cuh + planner), executed by
the execution-model emulator in tests/emu, against the oracle's undo_transforms -- bit-exact, including the
verification / serial-fallback path.  The same kernel is compared with the oracle on the GPU in test_gpu_parity.py."""
import numpy as np
import pytest

from fuif_b200.synth import synth_image
from tests import emu_util
from tests.util import default_squeeze_parameters

YCOCG, SQUEEZE = 1, 7


def forward_planes(po, pix, maxval, params, ycocg=True):
    img = po.OracleImage.from_pixels(pix, maxval)
    if ycocg:
        assert img.do_transform(YCOCG)
    assert img.do_transform(SQUEEZE, params)
    return img


def run_case(po, pix, maxval, params, opts, ycocg=True, garbage=None, keep_colour=False):
    """Unsqueezes the oracle's coefficient planes with the emulated kernel and compares with the oracle's own inverse."""
    h, w, nch = pix.shape
    img = forward_planes(po, pix, maxval, params, ycocg)
    pi = img.to_plane_image()
    coeff = [p.data.copy() for p in pi.planes]
    if garbage is not None:
        rng = np.random.default_rng(garbage[0])
        coeff = [rng.integers(-garbage[1], garbage[1] + 1, size=c.shape, dtype=np.int64).astype(np.int16) for c in coeff]
        for i, c in enumerate(coeff):
            img_set = np.ascontiguousarray(c)
            po.lib().fo_plane_set(img.h, i, img_set.ctypes.data, img_set.size)
    ref = img.clone()
    ref.undo_transforms(0 if not keep_colour else 1)
    expect = ref.to_plane_image()
    dims = [(p.w, p.h) for p in pi.planes]
    ops, nplanes, final = emu_util.plan_inverse_squeeze(dims, params, 0, nch)
    planes = [np.ascontiguousarray(c) for c in coeff]
    shapes = {}
    for o in ops:
        shapes[o[4]] = (o[7] if o[1] else o[7] + o[8], o[5] + o[6] if o[1] else o[5])
    for i in range(len(dims), nplanes):
        planes.append(np.full(shapes[i], 0x5A5A, dtype=np.int16))
    if keep_colour or not ycocg:
        ep = [1 if not keep_colour else 0, maxval, 0, maxval, 0 if keep_colour else 1, -1, -1, -1]
    else:
        ep = [2, maxval, 0, maxval, 1] + [f[0] for f in final[:3]]
    st = emu_util.run_plan(planes, ops, ep, opts)
    assert st[0] == 1, "planner refused the shape"
    if ep[0] and not st[5]:
        pytest.skip("epilogue not fused for this shape")
    for k, f in enumerate(final):
        got = planes[f[0]]
        want = expect.planes[k].data
        assert got.shape == want.shape
        if not np.array_equal(got, want):
            bad = np.argwhere(got != want)
            raise AssertionError(f"plane {k} {got.shape} differs at {len(bad)} samples, first {bad[0]}: {got[tuple(bad[0])]} vs {want[tuple(bad[0])]}; stats {st}")
    return st


def test_closed_form_pair_equals_reference_formulation():
    rng = np.random.default_rng(1)
    n = 2_000_000
    arrs = [rng.integers(-32768, 32768, size=n, dtype=np.int64).astype(np.int16) for _ in range(4)]
    # bias half of the samples towards monotone / near-equal triples, where the clamps of smooth_tendency act
    arrs[1][: n // 2] = (arrs[0][: n // 2].astype(np.int32) + rng.integers(-40, 41, size=n // 2)).clip(-32768, 32767).astype(np.int16)
    arrs[2][: n // 2] = (arrs[1][: n // 2].astype(np.int32) + rng.integers(-40, 41, size=n // 2)).clip(-32768, 32767).astype(np.int16)
    L = emu_util.lib()
    assert L.emu_check_pair(*[a.ctypes.data for a in arrs], n) == 0


@pytest.mark.parametrize("w,h,nch,tile,levels,coarse", [
    (96, 80, 3, (32, 32), 4, 16),
    (200, 120, 3, (64, 32), 4, 32),
    (131, 77, 3, (32, 48), 3, 16),
    (64, 48, 1, (16, 16), 4, 8),
    (37, 29, 3, (16, 16), 2, 8),
    (260, 40, 4, (64, 16), 4, 16),
    (5, 131, 3, (16, 32), 4, 16),
])
def test_fused_unsqueeze_matches_oracle(oracle, w, h, nch, tile, levels, coarse):
    maxval = 255 if nch != 4 else 16383
    pix = synth_image(w, h, nch, maxval, seed=w + h)
    params = default_squeeze_parameters(w, h, nch)
    st = run_case(oracle, pix, maxval, params, [tile[0], tile[1], levels, coarse, 64, 0], ycocg=nch >= 3)
    assert st[2] == 0, f"speculation failed on a smooth image: {st}"


def test_serial_fallback_is_exact_when_forced(oracle):
    pix = synth_image(96, 80, 3, 255, seed=3)
    params = default_squeeze_parameters(96, 80, 3)
    st = run_case(oracle, pix, 255, params, [32, 32, 4, 16, 64, 1])
    assert st[2] == 1


def test_full_range_garbage_is_exact(oracle):
    """Coefficients drawn from the whole int16 range: every wrap point is exercised and the warm-up cannot be
    expected to converge, so the verification has to catch it and the fallback has to repair it."""
    pix = synth_image(96, 80, 3, 255, seed=5)
    params = default_squeeze_parameters(96, 80, 3)
    st = run_case(oracle, pix, 255, params, [32, 32, 4, 16, 64, 0], garbage=(7, 32767))
    assert st[3] > 0


def test_noise_residuals_moderate_range(oracle):
    pix = synth_image(128, 96, 3, 255, seed=6)
    params = default_squeeze_parameters(128, 96, 3)
    run_case(oracle, pix, 255, params, [32, 32, 4, 16, 64, 0], garbage=(8, 300))


def test_keep_colour_transform_no_epilogue(oracle):
    pix = synth_image(100, 60, 3, 255, seed=9)
    params = default_squeeze_parameters(100, 60, 3)
    run_case(oracle, pix, 255, params, [32, 32, 4, 16, 64, 0], keep_colour=True)


@pytest.mark.parametrize("shape", [(200, 120, 3), (131, 77, 3), (260, 40, 4), (64, 48, 1)])
def test_repair_every_tile_of_last_launch(oracle, shape):
    """force = 2: every tile of the last launch is recomputed in exact mode (chains start from the recorded act values)."""
    w, h, nch = shape
    maxval = 255 if nch != 4 else 16383
    pix = synth_image(w, h, nch, maxval, seed=w)
    params = default_squeeze_parameters(w, h, nch)
    st = run_case(oracle, pix, maxval, params, [32, 32, 4, 16, 64, 2], ycocg=nch >= 3)
    assert st[8] > 0 and st[2] == 0, st


def test_short_warmup_failures_are_repaired_locally(oracle):
    """warm-up of 2 pairs in the last launch: many speculative starts fail; the tiles are repaired, no serial fallback."""
    pix = synth_image(256, 192, 3, 255, seed=21, noise=0.1)
    params = default_squeeze_parameters(256, 192, 3)
    st = run_case(oracle, pix, 255, params, [32, 32, 4, 32, 64, 0, 2, 12])
    assert st[4] > 0 and st[8] > 0 and st[2] == 0, st


def test_short_warmup_in_an_early_launch_takes_the_serial_fallback(oracle):
    pix = synth_image(256, 192, 3, 255, seed=22, noise=0.1)
    params = default_squeeze_parameters(256, 192, 3)
    st = run_case(oracle, pix, 255, params, [32, 32, 2, 16, 64, 0, 8, 1], keep_colour=True)
    assert st[4] > 0 and st[2] == 1, st
