"""CPU: the MANIAC ENCODE kernel source (fuif_b200/csrc/fb_maniac_enc.cu: one warp per channel group, lane p = property p)
executed by the execution-model emulator against the oracle encoder (byte-exact against the reference's files,
tests/test_oracle_encoder.py): every group's byte string -- header, learned + pruned tree, entropy-coded samples, or the
plain form when the reference rolls back -- must be identical."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from fuif_b200.synth import read_pnm
from tests.cases import APPROX_CASES, CASES, PALETTE_CASES, PERMUTE_CASES
from tests.test_oracle_encoder import _options
from tests.util import load_golden, ordered

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU_DIR = os.path.join(HERE, "emu")
LIB = os.path.join(EMU_DIR, "libfb_emu_maniac_enc.so")
SRCS = [os.path.join(EMU_DIR, "emu_maniac_enc.cpp"), os.path.join(EMU_DIR, "cuemu.h"), os.path.join(EMU_DIR, "maniac_emu_shim.h"),
        os.path.join(ROOT, "fuif_b200", "csrc", "fb_maniac_enc.cu"), os.path.join(ROOT, "fuif_b200", "csrc", "fb_encode_host.h")]
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(s) for s in SRCS):
            subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-DFB_EMULATE", "-fno-strict-aliasing", "-Wall", "-Wno-unused-function",
                                   "-Wno-unknown-pragmas", "-Wno-unused-variable", "-I", EMU_DIR, "-I", os.path.join(ROOT, "fuif_b200", "csrc"), SRCS[0], "-o", LIB])
        L = C.CDLL(LIB)
        L.emu_maniac_encode.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_void_p), C.c_uint,
                                        C.POINTER(C.c_int), C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int]
        L.emu_fuif_encode.restype = C.c_longlong
        L.emu_fuif_encode.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.c_float, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_longlong,
                                      C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
        L.emu_glibc_rand.argtypes = [C.c_void_p, C.c_longlong]
        _lib = L
    return _lib


def build_table(factor, max_p):
    """build_table, reference maniac/chance.cpp:31-65"""
    one, size = 1 << 32, 4096
    t = np.zeros((size, 2), dtype=np.uint16)
    last_p8, p = 0, one // 2
    for _ in range(size // 2):
        p8 = (size * p + one // 2) >> 32
        if p8 <= last_p8:
            p8 = last_p8 + 1
        if last_p8 and last_p8 < size and p8 <= max_p:
            t[last_p8, 1] = p8
        p += ((one - p) * factor + one // 2) >> 32
        last_p8 = p8
    for i in range(size - max_p, max_p + 1):
        if t[i, 1]:
            continue
        p = (i * one + size // 2) // size
        p += ((one - p) * factor + one // 2) >> 32
        p8 = (size * p + one // 2) >> 32
        if p8 <= i:
            p8 = i + 1
        if p8 > max_p:
            p8 = max_p
        t[i, 1] = p8
    for i in range(1, size):
        t[i, 0] = size - int(t[size - i, 1])
    return np.ascontiguousarray(t)


def log4k_table():
    """Log4kTable, reference maniac/chance.cpp:67-91"""
    def log4kf(x, base):
        bits = x.bit_length()
        y = x << (32 - bits)
        res = (base * (13 - bits)) & 0xFFFFFFFF
        add = base
        while add > 1 and (y & 0x7FFFFFFF) != 0:
            y = (y * y + 0x40000000) >> 31
            add >>= 1
            if (y >> 32) != 0:
                res = (res - add) & 0xFFFFFFFF
                y >>= 1
        return res
    base = (65535 << 16) // 12
    out = np.zeros(4097, dtype=np.uint16)
    for i in range(1, 4097):
        out[i] = ((log4kf(i, base) + (1 << 15)) >> 16) & 0xFFFF
    return out


def libc_rand(n):
    """the first n values of libc rand() from its initial state (what the reference process sees)"""
    libc = C.CDLL(None)
    libc.srand(1)
    return np.array([libc.rand() for _ in range(n)], dtype=np.int32)


def _varint(data, pos):
    v = 0
    while True:
        b = data[pos]; pos += 1
        if b < 128:
            return v + b, pos
        v = (v + b - 128) << 7


def learn_rows(h, nb_repeats):
    """iterations of the learning loop of encoding.cpp:180-203 for a plane of h rows (= rand() calls)"""
    n, y, rows = 0, 0, 0
    while y < h:
        rows += 1
        if np.float32(rows) > np.float32(nb_repeats) * np.float32(h):
            break
        n += 1
        y = 1
    return n


@pytest.mark.parametrize("case", [c for c in CASES if c[0] in ("odd", "gray", "tiny", "one", "nosq", "pred", "e0", "unc", "tall", "wide", "lossyq", "dct", "sq128")],
                         ids=lambda c: c[0])
def test_encode_kernel_matches_oracle_encoder(oracle, case):
    po = oracle
    name, w, h, c, maxval, seed, opts = case
    blob = load_golden(name)
    final = po.parse_fbpd(ordered(blob, "f")[-1])
    with tempfile.NamedTemporaryFile(suffix=".pnm", delete=False) as f:
        f.write(blob["pnm"])
        path = f.name
    try:
        pix, _ = read_pnm(path)
    finally:
        os.remove(path)
    oi = po.OracleImage.from_pixels(pix, maxval)
    oi.recompute_minmax()
    for tid, params in final.transforms:
        assert oi.do_transform(tid, [-1] + list(params) if tid == 9 else (params if tid in (4, 5, 6, 10) else []))
    o = _options(opts, c, final.transforms)
    ref = oi.encode(predictor=o["predictor"], nb_repeats=o["nb_repeats"], max_properties=o["max_properties"], compress=o["compress"], max_group=o["max_group"])
    assert ref[:-1] == bytes(blob["fuif"])[:-1]
    # the groups of the reference file: offsets, first channels, channel counts (first header varint), predictors
    dec, offs = po.OracleImage.decode(ref, want_offsets=True)
    pi = oi.to_plane_image()                   # planes as the encoder sees them (ranges tight, zero set by the encode above)
    nch = len(pi.planes)
    groups, ends = [], [o_ for o_, _ in offs[1:]] + [len(ref) - 1]
    rand_off = 0
    for (off, first), end in zip(offs, ends):
        fb, _ = _varint(ref, off)
        beginc, endc, predictor = first, first + (fb >> 4), (fb & 14) >> 1
        pred_opt = o["predictor"][beginc] if beginc < len(o["predictor"]) else o["predictor"][-1]
        assert predictor == pred_opt
        groups.append((beginc, endc, predictor, rand_off, off, end))
        if o["compress"] and any(pi.planes[k].minval != pi.planes[k].maxval for k in range(beginc, endc + 1)):
            rand_off += sum(learn_rows(pi.planes[k].h, o["nb_repeats"]) for k in range(beginc, endc + 1) if pi.planes[k].minval != pi.planes[k].maxval)
    desc = (C.c_int * (8 * nch))()
    ptrs = (C.c_void_p * nch)()
    keep = []
    for i, p in enumerate(pi.planes):
        desc[8 * i:8 * i + 8] = [p.w, p.h, p.minval, p.maxval, p.zero, p.q, p.hshift, p.vshift]
        a = np.ascontiguousarray(p.data.astype(np.int16)) if p.data is not None else np.zeros(1, dtype=np.int16)
        keep.append(a)
        ptrs[i] = a.ctypes.data
    ng = len(groups)
    gdesc = (C.c_longlong * (4 * ng))(*[v for g in groups for v in g[:4]])
    cap = 1 << 20
    outs = [np.zeros(cap, dtype=np.uint8) for _ in range(ng)]
    optrs = (C.c_void_p * ng)(*[a.ctypes.data for a in outs])
    glen = (C.c_int * (3 * ng))()
    table, meta, l4k = build_table(0x0d000000, 4096 - 6), build_table(0xFFFFFFFF // 19, 4096 - 2), log4k_table()
    rnd = libc_rand(rand_off + 16)
    rc = lib().emu_maniac_encode(nch, desc, ptrs, ng, gdesc, optrs, cap, glen, o["max_properties"], o["nb_repeats"], 1 if o["compress"] else 0,
                                 table.ctypes.data, meta.ctypes.data, l4k.ctypes.data, rnd.ctypes.data, len(rnd), 4096)
    assert rc == 0, [glen[3 * g + 2] for g in range(ng)]
    for gi, (beginc, endc, predictor, _ro, off, end) in enumerate(groups):
        mine = bytes(outs[gi][:glen[3 * gi]])
        if gi + 1 < ng:
            assert mine == ref[off:end], f"group {gi} (channels {beginc}-{endc}): {len(mine)} bytes vs {end - off}"
        else:       # after the last group the reference's blob may keep the tail of a longer, rolled-back compressed attempt and,
            #           every other length, one byte of BlobIO's bytes_used = seek_pos + 1 (fileio.h:245-251)
            assert len(mine) <= len(ref) - off and mine == ref[off:off + len(mine)], f"last group {gi}: {len(mine)} bytes vs {len(ref) - off}"
    # and the oracle decoder reads the kernel's groups back: same planes as from the reference's file
    mine_file = ref[:offs[0][0]] + b"".join(bytes(outs[gi][:glen[3 * gi]]) for gi in range(ng))
    back = po.OracleImage.decode(mine_file + b"\0")
    po.compare_plane_images(back.to_plane_image(), dec.to_plane_image(), name + " decode of the kernel's output")


ENC_CASES = list(CASES) + list(APPROX_CASES) + list(PALETTE_CASES) + list(PERMUTE_CASES)


def test_rand_restatement_matches_libc():
    """the product tabulates libc rand() itself (glibc TYPE_3, seed 1): the learning pass visits the reference's rows"""
    n = 5000
    mine = np.zeros(n, dtype=np.int32)
    lib().emu_glibc_rand(mine.ctypes.data, n)
    assert np.array_equal(mine, libc_rand(n))


@pytest.mark.parametrize("case", ENC_CASES, ids=lambda c: c[0])
def test_encode_file_matches_oracle_encoder(oracle, case):
    """fb_encode() end to end with the kernel emulated: the product's host side (group planning, rand() table, chance and
    cost tables, container assembly with BlobIO's quirks) around the kernel must give the oracle encoder's file, every byte,
    and the group index it returns must be the one a decode finds."""
    po = oracle
    name, w, h, c, maxval, seed, opts = case
    blob = load_golden(name)
    final = po.parse_fbpd(ordered(blob, "f")[-1])
    with tempfile.NamedTemporaryFile(suffix=".pnm", delete=False) as f:
        f.write(blob["pnm"])
        path = f.name
    try:
        pix, _ = read_pnm(path)
    finally:
        os.remove(path)
    oi = po.OracleImage.from_pixels(pix, maxval)
    oi.recompute_minmax()
    for tid, params in final.transforms:
        assert oi.do_transform(tid, [-1] + list(params) if tid == 9 else (params if tid in (4, 5, 6, 10) else []))
    _check_file(po, oi, _options(opts, c, final.transforms))


def _check_file(po, oi, o):
    oi.recompute_minmax()                       # fuif_prepare_encode
    pi = oi.to_plane_image()                    # zero not yet set by an encode: the host side has to do it
    ref = oi.encode(predictor=o["predictor"], nb_repeats=o["nb_repeats"], max_properties=o["max_properties"], compress=o["compress"], max_group=o["max_group"])
    _dec, offs = po.OracleImage.decode(ref, want_offsets=True)
    nch = len(pi.planes)
    desc = (C.c_int * (10 * nch))()
    ptrs = (C.c_void_p * nch)()
    keep = []
    for i, p in enumerate(pi.planes):
        desc[10 * i:10 * i + 10] = [p.w, p.h, p.minval, p.maxval, p.zero, p.q, p.hshift, p.vshift, p.hcshift, p.vcshift]
        a = np.ascontiguousarray(p.data.astype(np.int16)) if p.data is not None else np.zeros(1, dtype=np.int16)
        keep.append(a)
        ptrs[i] = a.ctypes.data
    info = (C.c_int * 7)(pi.w, pi.h, pi.maxval, pi.colormodel, pi.real_nb_channels, pi.nb_channels, pi.nb_meta_channels)
    flat = []
    for tid, params in pi.transforms:
        flat += [tid, len(params)] + list(params)
    tdesc = (C.c_int * max(1, len(flat)))(*flat)
    pred = (C.c_int * max(1, len(o["predictor"])))(*o["predictor"])
    cap = len(ref) + 4096
    out = np.zeros(cap, dtype=np.uint8)
    goffs = (C.c_longlong * 512)()
    gfirst = (C.c_int * 512)()
    ng = C.c_int()
    n = lib().emu_fuif_encode(nch, desc, ptrs, info, len(pi.transforms), tdesc, o["nb_repeats"], o["max_properties"], 6, 0x0d000000, 1 if o["compress"] else 0,
                              o["max_group"], len(o["predictor"]), pred, out.ctypes.data, cap, goffs, gfirst, 512, C.byref(ng))
    assert n >= 0, f"status {-n}"
    mine = bytes(out[:n])
    assert len(mine) == len(ref), f"{len(mine)} bytes vs {len(ref)}"
    assert mine == ref, f"first difference at byte {next(i for i in range(len(ref)) if mine[i] != ref[i])}"
    assert [(int(goffs[g]), int(gfirst[g])) for g in range(ng.value)] == [(int(a), int(b)) for a, b in offs]
    return ref, offs


@pytest.mark.parametrize("kind", ["noise", "noise_sq", "flat", "noise16"])
def test_encode_file_entropy_extremes(oracle, kind):
    """noise: the compressed form of a group is no smaller than the plain one, so the reference rolls back and the tail of
    the abandoned attempt stays in its blob; flat: constant planes, groups without entropy-coded data"""
    po = oracle
    rng = np.random.default_rng(5)
    maxval = 65535 >> 2 if kind == "noise16" else 255
    w, h, c = (37, 29, 3) if kind != "noise16" else (24, 20, 1)
    pix = np.full((h, w, c), 77, dtype=np.int32) if kind == "flat" else rng.integers(0, maxval + 1, size=(h, w, c)).astype(np.int32)
    oi = po.OracleImage.from_pixels(pix, maxval)
    oi.recompute_minmax()
    o = {"nb_repeats": 0.5, "max_properties": 12, "compress": True, "max_group": -1, "predictor": [2] * c + [0]}
    if kind in ("noise_sq", "flat"):
        assert oi.do_transform(1, [])
        assert oi.do_transform(7, [])
        o["max_group"] = 1
    ref, offs = _check_file(po, oi, o)
    plain = [(_varint(ref, off)[0] & 1) == 0 for off, _ in offs]
    if kind.startswith("noise"):
        assert any(plain), "expected at least one rolled-back group"


@pytest.mark.parametrize("order", ["reverse", "shuffle"])
def test_encode_is_independent_of_lane_order(order):
    """The emulator gives the lanes of a warp their turns in lane order by default, which would hide a race between lanes inside
    one barrier interval (on hardware they run together).  Same files with the order reversed / rotated at random."""
    env = dict(os.environ, CUEMU_ORDER=order)
    lib()       # built by now, so the child does not race another build
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "file_matches and (odd or dct or pred or unc) or extremes"], env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-1000:]


@pytest.mark.parametrize("variant", ["repeats2", "repeats01", "props2", "props18", "group3", "pred3", "pred0_nosq", "gray16"])
def test_encode_file_option_variants(oracle, variant):
    """encode options the golden cases leave out: more / fewer learning iterations (rand() offsets of later groups shift),
    fewer / more back-reference properties (18 = the most one warp holds), several channels per group, the remaining predictors"""
    po = oracle
    rng = np.random.default_rng(21)
    w, h, c, maxval = 41, 30, 3, 255
    if variant == "gray16":
        c, maxval = 1, 4095
    yy, xx = np.mgrid[0:h, 0:w]
    pix = np.stack([np.clip((maxval * (0.5 + 0.4 * np.sin(xx / (5.0 + k)) * np.cos(yy / (7.0 + k))) + rng.normal(0, maxval * 0.02, (h, w))), 0, maxval)
                    for k in range(c)], axis=-1).astype(np.int32)
    oi = po.OracleImage.from_pixels(pix, maxval)
    oi.recompute_minmax()
    o = {"nb_repeats": 0.5, "max_properties": 12, "compress": True, "max_group": 1, "predictor": [2] * c + [0]}
    squeeze = True
    if variant == "repeats2": o["nb_repeats"] = 2.0
    elif variant == "repeats01": o["nb_repeats"] = 0.1
    elif variant == "props2": o["max_properties"] = 2
    elif variant == "props18": o["max_properties"] = 18
    elif variant == "group3": o["max_group"] = 3
    elif variant == "pred3": o["predictor"] = [3, 3, 3, 3]
    elif variant == "pred0_nosq": o["predictor"] = [0]; squeeze = False; o["max_group"] = -1
    if c >= 3:
        assert oi.do_transform(1, [])
    if squeeze:
        assert oi.do_transform(7, [])
    _check_file(po, oi, o)
