"""CPU: the product's MANIAC decode kernel SOURCE (fuif_b200/csrc/fb_maniac.cu, device part) executed by the execution-model
emulator (tests/emu: one fibre per CUDA thread, warp collectives as rendezvous, polls yield) against the oracle, on
golden files written by the unmodified reference.  This checks the kernel's logic without a GPU -- stream tickets, the
row wavefront between channel groups, run-ahead walkers with property-12 forks, the prologue warp, the leaf cache, the
integer coder, the one-warp fallback path -- not its timing and not SIMT lockstep (see tests/emu/maniac_emu_shim.h)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.util import load_golden

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU_DIR = os.path.join(HERE, "emu")
LIB = os.path.join(EMU_DIR, "libfb_emu_maniac.so")
SRCS = [os.path.join(EMU_DIR, "emu_maniac.cpp"), os.path.join(EMU_DIR, "cuemu.h"), os.path.join(EMU_DIR, "maniac_emu_shim.h"),
        os.path.join(ROOT, "fuif_b200", "csrc", "fb_maniac.cu")]
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(s) for s in SRCS):
            subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-DFB_EMULATE", "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas",
                                   "-Wno-unused-variable", "-fno-strict-aliasing", "-I", EMU_DIR, "-I", os.path.join(ROOT, "fuif_b200", "csrc"), SRCS[0], "-o", LIB])
        L = C.CDLL(LIB)
        L.emu_maniac_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_size_t]
        _lib = L
    return _lib


def _varints(data, pos, n):
    out = []
    for _ in range(n):
        v = 0
        while True:
            b = data[pos]; pos += 1
            if b < 128:
                v += b
                break
            v = (v + b - 128) << 7
        out.append(v)
    return out, pos


def emu_decode(po, data, indexed=True, shape=0, nblocks=1, smem_kib=226, preview=-1):
    """Returns (status, [(plane array or None, minval, maxval, zero, q)]) for the channel list of the file."""
    full, offs = po.OracleImage.decode(data, want_offsets=True)
    ref = full if preview < 0 else po.OracleImage.decode(data, preview=preview)
    pi = ref.to_plane_image()
    (nbch, _bd, _w, _h, _cm, max_properties), pos = _varints(data, 4, 6)
    btl = 0
    if preview >= 0:        # responsive truncation points, encoding.cpp:641-650
        rel, pos = _varints(data, pos, 5)
        btl = sum(rel[:preview + 1]) + pos
        offs = [o for o in offs if o[0] < btl]      # what the library's host half does with a group index
    nbch -= ord('0')
    nch = len(pi.planes)
    desc = (C.c_int * (5 * nch))()
    planes = []
    ptrs = (C.c_void_p * nch)()
    for i, p in enumerate(pi.planes):
        desc[5 * i:5 * i + 5] = [p.w, p.h, p.hshift, p.vshift, p.q]     # (q of an all-zero plane is never in the stream: the channel list's q stays)
        a = np.full((max(p.h, 0), max(p.w, 0)), -12345, dtype=np.int16)
        planes.append(a)
        ptrs[i] = a.ctypes.data if a.size else None
    chout = (C.c_int * (5 * nch))()
    ng = len(offs) if indexed else 0
    goff = (C.c_longlong * max(1, ng))(*[o for o, _ in offs][:ng])
    gfirst = (C.c_int * max(1, ng))(*[f for _, f in offs][:ng])
    st = lib().emu_maniac_decode(data, len(data), offs[0][0], max_properties, nbch, nch, desc, ptrs, chout, ng, goff, gfirst, shape, nblocks, 6, 0x0d000000,
                                 smem_kib, 0, btl)
    out = []
    for i in range(nch):
        out.append((planes[i] if chout[5 * i + 4] else None, chout[5 * i], chout[5 * i + 1], chout[5 * i + 2], chout[5 * i + 3]))
    return st, out, pi


def _check(po, data, meta=True, **kw):
    st, got, pi = emu_decode(po, data, **kw)
    assert st == 0
    for i, (p, g) in enumerate(zip(pi.planes, got)):
        if p.data is None:
            assert g[0] is None, f"plane {i} decoded by the kernel but not by the oracle"
            continue
        assert g[0] is not None, f"plane {i} missing"
        if meta:
            assert (g[1], g[2], g[4]) == (p.minval, p.maxval, p.q), f"plane {i} range / q: {(g[1], g[2], g[4])} vs {(p.minval, p.maxval, p.q)}"
        assert np.array_equal(g[0], p.data), f"plane {i} ({p.w}x{p.h}) differs"


SMALL = ["odd", "tiny", "one", "gray", "nosq", "pred", "e0", "unc", "tall", "wide", "dct", "dctodd", "lossyq", "rgba14", "sq128"]


@pytest.mark.parametrize("name", SMALL)
def test_kernel_source_decodes_golden_indexed(oracle, name):
    _check(oracle, bytes(load_golden(name)["fuif"]), indexed=True)


@pytest.mark.parametrize("name,indexed", [("approx", True), ("approx_nosq", False), ("approx14", True), ("pal", True), ("pal", False), ("pal_nosq", True),
                                          ("pal4", False), ("pal_c0", True), ("match_gray", True), ("match_gray", False), ("match_nosq", True), ("perm", True)])
def test_kernel_source_decodes_meta_and_remainder_channels(oracle, name, indexed):
    """files whose channel list starts with a meta-channel -- a palette (hshift -1: never a back-reference,
    context_predict.h:73) or a 2DMatch code plane (hshift 0: it IS one) -- or ends with Approximate's remainder channels"""
    _check(oracle, bytes(load_golden(name)["fuif"]), indexed=indexed)


@pytest.mark.parametrize("name", ["odd", "gray", "unc", "tiny"])
def test_kernel_source_decodes_golden_sequential(oracle, name):
    _check(oracle, bytes(load_golden(name)["fuif"]), indexed=False)


@pytest.mark.parametrize("name", ["odd", "nosq", "dct", "sq128"])
def test_batch_launch_shape_and_two_blocks(oracle, name):
    """two streams per block with 5 walkers each, two co-resident blocks (streams wait for planes decoded by the other block)"""
    _check(oracle, bytes(load_golden(name)["fuif"]), indexed=True, shape=1, nblocks=2)


def test_leaf_cache_path(oracle):
    """a shared-memory budget too small for the leaves of the larger groups: direct-mapped write-back leaf cache"""
    _check(oracle, bytes(load_golden("sq128")["fuif"]), indexed=True, smem_kib=72)


@pytest.mark.parametrize("name", ["odd", "gray", "sq128", "dct"])
@pytest.mark.parametrize("preview", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("indexed", [True, False], ids=["indexed", "sequential"])
def test_responsive_truncation(oracle, name, preview, indexed):
    """-R k: the stream stops at a responsive offset (encoding.cpp:704-716): planes beyond it stay undecoded, a plane cut in
    the middle keeps its initial fill"""
    _check(oracle, bytes(load_golden(name)["fuif"]), meta=False, indexed=indexed, preview=preview)


@pytest.mark.parametrize("name", ["odd", "dct", "sq128", "rgba14", "pred", "unc"])
def test_throughput_shape_eight_one_warp_streams_per_block(oracle, name):
    """the launch shape of big batches (fb_maniac.cu: from 12 streams per SM on): 8 streams per block, one warp each, no walkers; two
    blocks, so streams also wait for planes decoded in the other block"""
    _check(oracle, bytes(load_golden(name)["fuif"]), indexed=True, shape=3, nblocks=2)


@pytest.mark.parametrize("shape", [0, 1, 3], ids=["one_stream_per_block", "two_streams_two_blocks", "eight_one_warp_streams"])
def test_kernel_source_survives_damaged_input(shape):
    """120 damaged files / bogus group offsets per launch shape through the emulated kernel in a child process: every launch ends
    (no spin-wait on rows nobody will publish, no endless coder loop).  The same inputs hang a GPU if they hang here."""
    import sys
    child = os.path.join(HERE, "emu_maniac_fuzz_child.py")
    lib()       # build the emulator library here, not under the child's timeout
    r = subprocess.run([sys.executable, child, str(21 + shape), "120", str(shape)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-800:]
    assert "returned 120 times" in r.stdout
