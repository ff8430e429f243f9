"""TEST INFRASTRUCTURE ONLY: builds tests/emu/libfb_emu.so (the product's fused-kernel SOURCE compiled for the CPU
execution-model emulator, tests/emu/cuemu.h) and mirrors, in Python, the channel-list surgery of the library's
inv_squeeze (fuif_b200/csrc/fb_image.cu; reference transform/squeeze.h:367-388) so that a plan can be run on numpy planes."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU_DIR = os.path.join(HERE, "emu")
LIB = os.path.join(EMU_DIR, "libfb_emu.so")
SRCS = [os.path.join(EMU_DIR, "emu_squeeze.cpp"), os.path.join(EMU_DIR, "cuemu.h"),
        os.path.join(ROOT, "fuif_b200", "csrc", "fb_fused_squeeze.cuh"), os.path.join(ROOT, "fuif_b200", "csrc", "fb_fused_plan.h"),
        os.path.join(ROOT, "fuif_b200", "csrc", "fb_port.h"), os.path.join(ROOT, "fuif_b200", "csrc", "fb_direct_squeeze.cuh"),
        os.path.join(ROOT, "fuif_b200", "csrc", "fb_direct_plan.h"), os.path.join(ROOT, "fuif_b200", "csrc", "fb_subsample.cuh"),
        os.path.join(ROOT, "fuif_b200", "csrc", "fb_approx.cuh"), os.path.join(ROOT, "fuif_b200", "csrc", "fb_palette.cuh"),
        os.path.join(ROOT, "fuif_b200", "csrc", "fb_match.cuh"), os.path.join(ROOT, "fuif_b200", "csrc", "fb_pk_squeeze.cuh"),
        os.path.join(ROOT, "fuif_b200", "csrc", "fb_pk_plan.h")]
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(s) for s in SRCS):
            subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-DFB_EMULATE", "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas",
                                   "-I", EMU_DIR, "-I", os.path.join(ROOT, "fuif_b200", "csrc"), SRCS[0], "-o", LIB])
        L = C.CDLL(LIB)
        L.emu_run_plan.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.emu_check_pair.argtypes = [C.c_void_p] * 4 + [C.c_int]
        L.emu_run_direct.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.emu_inv_subsample.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.emu_inv_subsample.restype = None
        L.emu_approximate.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int]
        L.emu_approximate.restype = None
        L.emu_palette_inv.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, C.c_longlong]
        L.emu_palette_inv.restype = None
        L.emu_match_inv.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.emu_match_soft.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.emu_palette_fwd.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_longlong, C.c_int, C.c_void_p, C.c_int]
        L.emu_run_pk.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.emu_check_pk_pair.argtypes = [C.c_void_p] * 4 + [C.c_int]
        L.emu_check_pk_ycocg.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int]
        _lib = L
    return _lib


def plan_inverse_squeeze(dims, params, nb_meta, nb_channels):
    """dims: [(w, h)] of the channel list after meta_squeeze.  Returns (ops, nplanes, final) where ops are
    (step, horizontal, avg, res, out, wa, wr, ha, hr) over plane ids (ids >= len(dims) are new planes) and final is the
    list of (plane id, w, h) of the channel list after the inverse."""
    chans = [(i, w, h) for i, (w, h) in enumerate(dims)]
    nplanes = len(dims)
    ops = []
    step = 0
    for i in range(len(params) - 3, -1, -3):
        horizontal, in_place = params[i] & 1, not (params[i] & 2)
        beginc, endc = params[i + 1], params[i + 2]
        offset = endc + 1 if in_place else nb_meta + nb_channels
        for c in range(beginc, endc + 1):
            a, r = chans[c], chans[offset + c - beginc]
            if horizontal:
                out = (nplanes, a[1] + r[1], a[2])
            else:
                out = (nplanes, a[1], a[2] + r[2])
            nplanes += 1
            ops.append((step, horizontal, a[0], r[0], out[0], a[1], r[1], a[2], r[2]))
            chans[c] = out
        del chans[offset:offset + endc - beginc + 1]
        step += 1
    return ops, nplanes, chans


def run_plan(planes, ops, ep, opts):
    """planes: list of contiguous int16 arrays (outputs preallocated).  Returns the stats list."""
    L = lib()
    ptrs = (C.c_void_p * len(planes))(*[p.ctypes.data for p in planes])
    od = (C.c_int * (9 * len(ops)))(*[int(v) for o in ops for v in o])
    e = (C.c_int * 8)(*ep)
    o = (C.c_int * 8)(*(list(opts) + [0] * (8 - len(opts))))
    st = (C.c_int * 9)()
    L.emu_run_plan(len(planes), ptrs, len(ops), od, e, o, st)
    return list(st)


def run_direct(planes, ops10, ep9, lo, hi):
    """ops10: (step, horizontal, avg, res, out, wa, wr, ha, hr, clamp).  Returns [launches, direct ops, serial ops, epilogue done]."""
    L = lib()
    ptrs = (C.c_void_p * len(planes))(*[p.ctypes.data for p in planes])
    od = (C.c_int * (10 * len(ops10)))(*[int(v) for o in ops10 for v in o])
    e = (C.c_int * 9)(*ep9)
    st = (C.c_int * 4)()
    L.emu_run_direct(len(planes), ptrs, len(ops10), od, e, lo, hi, st)
    return list(st)


def run_pk(planes, ops10, ep9, lo, hi, sm_count=4):
    """The packed per-step kernels (fb_pk_squeeze.cuh) under the emulator; same descriptors as run_direct.
    Returns [launches, packed ops, serial ops, epilogue done, repaired segments, range-flagged segments]."""
    L = lib()
    ptrs = (C.c_void_p * len(planes))(*[p.ctypes.data for p in planes])
    od = (C.c_int * (10 * len(ops10)))(*[int(v) for o in ops10 for v in o])
    e = (C.c_int * 9)(*ep9)
    st = (C.c_int * 6)()
    rc = L.emu_run_pk(len(planes), ptrs, len(ops10), od, e, lo, hi, sm_count, st)
    assert rc == 0, f"emu_run_pk rc={rc}"
    return list(st)
