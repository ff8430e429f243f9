"""The host-threads entropy backend (fb_host_decode) against the golden vectors of the unmodified
reference: no GPU involved, so these run in the CPU tier.  The same planes must come out with and without the group
index, for any thread count."""
import numpy as np
import pytest

from tests.cases import CASES
from tests.util import gpu_plane_image, load_golden


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_host_decode_vs_golden(oracle, case):
    from fuif_b200 import api
    po = oracle
    blob = load_golden(case[0])
    ref = po.parse_fbpd(blob["s0"])
    img = api.fuif_host_decode(blob["fuif"], threads=1)
    po.compare_plane_images(gpu_plane_image(po, img), ref, case[0] + " host s0")
    offs, first = img.group_index()
    _, ooffs = po.OracleImage.decode(blob["fuif"], want_offsets=True)
    assert list(zip(offs, first)) == [(int(a), int(b)) for a, b in ooffs]
    for threads, gi in ((4, (offs, first)), (3, offs), (0, (offs, first))):
        par = api.fuif_host_decode(blob["fuif"], group_index=gi, threads=threads)
        po.compare_plane_images(gpu_plane_image(po, par), ref, f"{case[0]} host indexed t{threads}")


ALL_WITH_FUIF = ["approx", "approx14", "approx_noop", "approx_nosq", "approx_q", "match", "match_gray", "match_nosq", "match_soft", "pal", "pal4",
                 "pal_c0", "pal_nosq", "perm", "perm2", "perm_nosq", "sub_420", "sub_420odd", "sub_422", "sub_440", "sub_one_chan", "sub_tiny"]


@pytest.mark.parametrize("name", ALL_WITH_FUIF)
def test_host_decode_vs_oracle_other_chains(oracle, name):
    """Files of the other transform chains (palette and match meta-channels, permutations, subsampled chroma): same planes as
    the C restatement of fuif_decode, sequentially and one thread per group."""
    from fuif_b200 import api
    po = oracle
    data = bytes(load_golden(name)["fuif"])
    ref = po.OracleImage.decode(data).to_plane_image()
    img = api.fuif_host_decode(data, threads=1)
    po.compare_plane_images(gpu_plane_image(po, img), ref, name + " host")
    par = api.fuif_host_decode(data, group_index=img.group_index(), threads=8)
    po.compare_plane_images(gpu_plane_image(po, par), ref, name + " host indexed")


@pytest.mark.parametrize("case", [c for c in CASES if c[0] in ("odd", "sq128", "rgba14", "dct", "gray")], ids=lambda c: c[0])
@pytest.mark.parametrize("preview", [0, 1, 2, 3, 4])
def test_host_responsive_decode(oracle, case, preview):
    from fuif_b200 import api
    po = oracle
    blob = load_golden(case[0])
    img = api.fuif_host_decode(blob["fuif"], api.fuif_options(preview=preview))
    po.compare_plane_images(gpu_plane_image(po, img), po.parse_fbpd(blob[f"r{preview}s0"]), f"{case[0]} host R{preview} s0", check_meta=False)


@pytest.mark.parametrize("name", ["sq128", "rgba14", "dct", "unc", "pred"])
def test_host_truncated_and_damaged_like_oracle(oracle, name):
    """Truncation is tolerated the way the reference tolerates it (zero fill, encoding.cpp:209-219); a damaged stream gives the
    oracle's planes or an error where the oracle fails -- and always returns."""
    from fuif_b200 import api
    po = oracle
    data = bytes(load_golden(name)["fuif"])
    rng = np.random.default_rng(len(data))
    variants = [data[:cut] for cut in (len(data) - 1, len(data) * 3 // 4, len(data) // 2, len(data) // 7, 64)]
    for k in range(8):
        d = bytearray(data)
        for _ in range(1 + 2 * k):
            d[int(rng.integers(40, len(d)))] ^= 1 << int(rng.integers(0, 8))
        variants.append(bytes(d))
    for k, d in enumerate(variants):
        try:
            ref = po.OracleImage.decode(d).to_plane_image()
        except RuntimeError:
            ref = None
        try:
            got = gpu_plane_image(po, api.fuif_host_decode(d, threads=2))
        except api.FuifError:
            got = None
        # where the oracle (like the reference) fails, the backend must fail too; the backend may ALSO refuse a header the reference
        # only survives by luck (an inverted value range: the reference runs into its asserts / unbounded recursion there)
        assert got is None or ref is not None, f"{name} variant {k}: the oracle fails, the host backend decodes"
        if ref is not None and got is not None:
            po.compare_plane_images(got, ref, f"{name} damaged {k}", check_meta=False)


def test_host_image_needs_upload_before_computing():
    from fuif_b200 import api
    img = api.fuif_host_decode(bytes(load_golden("sq128")["fuif"]))
    with pytest.raises(api.FuifError):
        img.undo_transforms(0)


import os  # noqa: E402
import subprocess  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")

AT_SIZE = [
    # name, w, h, channels, maxval, seed, ref_driver encode options: trees of thousands of nodes, 14-bit planes (full leaf layout),
    # multi-plane groups of the DCT chain, a non-zero predictor through the chunked row decoder
    ("rgb_512", 512, 512, 3, 255, 1234, []),
    ("raw14_640", 640, 480, 4, 16383, 9, ["-q", "12,64"]),
    ("dct_odd_grouped", 600, 328, 3, 255, 11, ["-C", "1", "-J", "-q", "8,12"]),
    ("pred6_gray", 700, 300, 1, 255, 5, ["-P", "6"]),
]


@pytest.mark.parametrize("case", AT_SIZE, ids=[c[0] for c in AT_SIZE])
def test_host_decode_vs_reference_at_size(oracle, case, tmp_path):
    """Against the UNMODIFIED reference run on the spot (ref_driver encodes a synthetic image and dumps its decoded planes)."""
    from fuif_b200 import api
    from fuif_b200.synth import synth_image, write_pnm
    if not os.access(REF, os.X_OK):
        pytest.skip("oracle/_ref/ref_driver is not built (needs /root/reference at build time)")
    po = oracle
    name, w, h, c, maxval, seed, opts = case
    pnm, fuif, pre = str(tmp_path / "in.pnm"), str(tmp_path / "x.fuif"), str(tmp_path / "d")
    write_pnm(pnm, synth_image(w, h, c, maxval, seed), maxval)
    subprocess.run([REF, "encode", pnm, fuif, *opts], check=True, capture_output=True)
    subprocess.run([REF, "dump", fuif, pre], check=True, capture_output=True)
    dumps = sorted((f for f in os.listdir(tmp_path) if f.startswith("d.s") and f.endswith(".fbpd")), key=lambda f: int(f[3:-5]))
    first = po.parse_fbpd(open(tmp_path / dumps[0], "rb").read())
    data = open(fuif, "rb").read()
    seq = api.fuif_host_decode(data, threads=1)
    po.compare_plane_images(gpu_plane_image(po, seq), first, name + " host decode")
    for rep in range(3):        # several runs with more threads than cores: the row wavefront between dependent groups under preemption
        par = api.fuif_host_decode(data, group_index=seq.group_index(), threads=12)
        po.compare_plane_images(gpu_plane_image(po, par), first, f"{name} host decode indexed, run {rep}")


@pytest.mark.timeout(120)
@pytest.mark.parametrize("name", ["sq128", "dct", "rgba14"])
def test_host_wrong_group_index_returns(oracle, name):
    """A sidecar that does not belong to the file (shifted, shuffled, first entries missing, offsets into the middle of a group)
    must not hang a stream on planes nobody decodes: the call returns -- the right planes where the index is recognisably
    unusable and the decoder falls back to one stream, garbage or an error otherwise."""
    from fuif_b200 import api
    po = oracle
    data = bytes(load_golden(name)["fuif"])
    ref = po.parse_fbpd(load_golden(name)["s0"])
    good = api.fuif_host_decode(data, threads=1)
    offs, first = good.group_index()
    offs, first = list(offs), list(first)
    n = len(offs)
    # the first planes belong to no stream: recognised, decoded as one stream
    img = api.fuif_host_decode(data, group_index=(offs[2:], first[2:]), threads=4)
    po.compare_plane_images(gpu_plane_image(po, img), ref, name + " index without its first entries")
    rng = np.random.default_rng(7)
    variants = [(offs[::-1], first), ([o + 1 for o in offs], first), (offs, [min(f + 1, first[-1]) for f in first]), (offs[: n // 2], first[: n // 2]),
                ([int(x) for x in rng.integers(0, len(data), n)], first), (offs, first[:1] + first[2:] + first[-1:])]
    for k, gi in enumerate(variants):
        try:
            api.fuif_host_decode(data, group_index=gi, threads=4)
        except api.FuifError:
            pass


def test_host_fuzz_returns():
    """900 damaged inputs (bit flips, overwritten bytes, random group offsets) in a child process: every call returns, the process
    survives.  Found with it: a group header with an inverted value range (endless loop of the uniform coder, reference: assert) and
    a group that reaches into the planes of another stream (lowers their `rows_done` after the owner released it: waiters spin
    forever) -- both are corrupt streams now, in the GPU kernel as well."""
    import sys
    child = os.path.join(ROOT, "tests", "host_fuzz_child.py")
    r = subprocess.run([sys.executable, child, "11", "900"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    assert "returned 900 times" in r.stdout
