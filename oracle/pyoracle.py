"""TEST INFRASTRUCTURE ONLY: ctypes bindings for oracle/libfuif_oracle.so (the plain-C restatement of the
reference hot path) and helpers around oracle/_ref/ref_driver (the unmodified reference, when present).

Nothing under fuif_b200/ imports this module; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfuif_oracle.so")
REF_DRIVER = os.path.join(HERE, "_ref", "ref_driver")

_lib = None


def build(force: bool = False) -> None:
    """Compiles the C restatement (and, when /root/reference exists, the reference driver)."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "fuif_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference") and (force or not os.path.exists(REF_DRIVER)
                                             or os.path.getmtime(REF_DRIVER) < os.path.getmtime(os.path.join(HERE, "ref_driver.cpp"))):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_DRIVER) and os.access(REF_DRIVER, os.X_OK)


class EncOptions(C.Structure):
    """fo_enc_options (oracle/fuif_oracle.h) = the encode half of fuif_options, reference encoding/encoding.h:32-59."""
    _fields_ = [("nb_repeats", C.c_float), ("max_properties", C.c_int), ("maniac_cutoff", C.c_int), ("maniac_alpha", C.c_int),
                ("compress", C.c_int), ("max_group", C.c_int), ("npred", C.c_int), ("predictor", C.POINTER(C.c_int))]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.fo_image_new.restype = C.c_void_p
        L.fo_image_new.argtypes = [C.c_int] * 5
        L.fo_image_free.argtypes = [C.c_void_p]
        L.fo_image_clone.restype = C.c_void_p
        L.fo_image_clone.argtypes = [C.c_void_p]
        L.fo_nplanes.argtypes = [C.c_void_p]
        L.fo_ntransforms.argtypes = [C.c_void_p]
        L.fo_plane_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
        L.fo_plane_data.restype = C.POINTER(C.c_int16)
        L.fo_plane_data.argtypes = [C.c_void_p, C.c_int]
        L.fo_image_info.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.fo_transform_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
        L.fo_plane_set.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.fo_plane_set_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.fo_plane_reshape.argtypes = [C.c_void_p] + [C.c_int] * 8
        L.fo_push_transform.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int]
        L.fo_decode.restype = C.c_void_p
        L.fo_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
        L.fo_undo_transforms.argtypes = [C.c_void_p, C.c_int]
        L.fo_do_transform.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int]
        L.fo_recompute_minmax.argtypes = [C.c_void_p]
        L.fo_encode.restype = C.c_void_p
        L.fo_encode.argtypes = [C.c_void_p, C.POINTER(EncOptions), C.POINTER(C.c_size_t)]
        L.fo_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


@dataclass
class Plane:
    """One Channel (reference image/image.h:54-91) as plain data."""
    w: int
    h: int
    minval: int = 0
    maxval: int = 0
    zero: int = 0
    q: int = 1
    hshift: int = 0
    vshift: int = 0
    hcshift: int = 0
    vcshift: int = 0
    component: int = -1
    data: np.ndarray | None = None   # int16 (h, w) or None when not decoded

    def meta(self):
        return (self.w, self.h, self.minval, self.maxval, self.q, self.hshift, self.vshift, self.hcshift, self.vcshift, self.component)


@dataclass
class PlaneImage:
    """One Image (reference image/image.h:98-129) as plain data."""
    w: int
    h: int
    minval: int
    maxval: int
    nb_channels: int
    real_nb_channels: int
    nb_meta_channels: int
    colormodel: int
    planes: list = field(default_factory=list)
    transforms: list = field(default_factory=list)   # [(id, [params])]


def read_fbpd(path: str) -> PlaneImage:
    """Reads a plane dump written by oracle/ref_driver.cpp."""
    with open(path, "rb") as f:
        return parse_fbpd(f.read())


def parse_fbpd(buf: bytes) -> PlaneImage:
    buf = bytes(buf)
    end = buf.index(b"END\n") + 4
    lines = buf[:end].decode().strip().split("\n")
    assert lines[0] == "FBPD1"
    it = list(map(int, lines[1].split()[1:]))
    img = PlaneImage(*it[:8])
    nplanes, ntr = it[8], it[9]
    k = 2
    for _ in range(ntr):
        v = list(map(int, lines[k].split()[1:]))
        img.transforms.append((v[0], v[2:2 + v[1]]))
        k += 1
    sizes = []
    for _ in range(nplanes):
        v = list(map(int, lines[k].split()[1:]))
        img.planes.append(Plane(*v[:11]))
        sizes.append(v[11])
        k += 1
    pos = end
    for p, n in zip(img.planes, sizes):
        if n and n == p.w * p.h:
            p.data = np.frombuffer(buf, dtype="<i2", count=n, offset=pos).reshape(p.h, p.w).copy()
        elif n:
            # the reference keeps the original full-size (all-zero) buffer for planes it never decoded
            p.data = None
        pos += 2 * n
    return img


class OracleImage:
    """Owns an fo_image* of the C restatement."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle returned NULL")
        self.h = handle

    def __del__(self):
        if getattr(self, "h", None):
            lib().fo_image_free(self.h)
            self.h = None

    @staticmethod
    def decode(data: bytes, preview: int = -1, cutoff: int = 6, alpha: int = 0x0d000000, want_offsets: bool = False):
        cap = C.c_int(4096)
        offs = (C.c_longlong * (2 * 4096))()
        hnd = lib().fo_decode(data, len(data), preview, cutoff, alpha, offs, C.byref(cap))
        img = OracleImage(hnd)
        if want_offsets:
            return img, [(offs[2 * i], offs[2 * i + 1]) for i in range(cap.value)]
        return img

    @staticmethod
    def from_pixels(pix: np.ndarray, maxval: int) -> "OracleImage":
        """(h, w, c) integer array -> Image as read_PAM_file would build it (reference import/read_pam.h:126)."""
        h, w, c = pix.shape
        img = OracleImage(lib().fo_image_new(w, h, maxval, c, 0))
        for i in range(c):
            a = np.ascontiguousarray(pix[:, :, i].astype(np.int16))
            lib().fo_plane_set(img.h, i, a.ctypes.data, a.size)
        return img

    @staticmethod
    def from_plane_image(pi: "PlaneImage") -> "OracleImage":
        """An Image whose planes and transform stack are given as they are (nothing is applied): the state after fuif_decode,
        or of an importer that delivers already-transformed planes.  The plane count must be nb_channels."""
        assert len(pi.planes) == pi.nb_channels and pi.nb_meta_channels == 0
        img = OracleImage(lib().fo_image_new(pi.w, pi.h, pi.maxval, pi.nb_channels, pi.colormodel))
        for i, p in enumerate(pi.planes):
            lib().fo_plane_reshape(img.h, i, p.w, p.h, p.hshift, p.vshift, p.hcshift, p.vcshift, p.component)
            if p.data is not None:
                a = np.ascontiguousarray(p.data.astype(np.int16))
                lib().fo_plane_set(img.h, i, a.ctypes.data, a.size)
            lib().fo_plane_set_range(img.h, i, p.minval, p.maxval, p.q)
        for tid, params in pi.transforms:
            arr = (C.c_int * max(1, len(params)))(*params)
            lib().fo_push_transform(img.h, tid, arr, len(params))
        return img

    def clone(self) -> "OracleImage":
        return OracleImage(lib().fo_image_clone(self.h))

    def undo_transforms(self, keep: int = 0) -> None:
        if lib().fo_undo_transforms(self.h, keep) != 0:
            raise RuntimeError("oracle undo_transforms failed")

    def do_transform(self, tid: int, params=()) -> bool:
        arr = (C.c_int * max(1, len(params)))(*params)
        return bool(lib().fo_do_transform(self.h, tid, arr, len(params)))

    def recompute_minmax(self) -> None:
        lib().fo_recompute_minmax(self.h)

    def encode(self, predictor=(), nb_repeats: float = 0.5, max_properties: int = 12, cutoff: int = 6, alpha: int = 0x0d000000,
               compress: bool = True, max_group: int = -1) -> bytes:
        """fuif_prepare_encode + fuif_encode of the (already transformed) image, reference encoding/encoding.cpp:737-743, 455-573."""
        pred = (C.c_int * max(1, len(predictor)))(*predictor)
        opt = EncOptions(nb_repeats, max_properties, cutoff, alpha, 1 if compress else 0, max_group, len(predictor), pred)
        n = C.c_size_t(0)
        ptr = lib().fo_encode(self.h, C.byref(opt), C.byref(n))
        if not ptr:
            raise RuntimeError("oracle fo_encode failed")
        try:
            return C.string_at(ptr, n.value)
        finally:
            lib().fo_free(ptr)

    def to_plane_image(self) -> PlaneImage:
        L = lib()
        info = (C.c_int * 8)()
        L.fo_image_info(self.h, info)
        pi = PlaneImage(*list(info))
        for i in range(L.fo_nplanes(self.h)):
            v = (C.c_longlong * 12)()
            L.fo_plane_info(self.h, i, v)
            p = Plane(*[int(x) for x in v[:11]])
            n = int(v[11])
            if n and n == p.w * p.h:
                ptr = L.fo_plane_data(self.h, i)
                p.data = np.ctypeslib.as_array(ptr, shape=(p.h, p.w)).copy()
            pi.planes.append(p)
        for i in range(L.fo_ntransforms(self.h)):
            tid = C.c_int()
            params = (C.c_int * 1024)()
            n = L.fo_transform_info(self.h, i, C.byref(tid), params, 1024)
            pi.transforms.append((tid.value, list(params[:n])))
        return pi

    def pixels(self) -> np.ndarray:
        """(h, w, c) int32 of the first nb_channels planes (after undo_transforms)."""
        pi = self.to_plane_image()
        return np.stack([p.data.astype(np.int32) for p in pi.planes[:pi.nb_channels]], axis=-1)


def ref_run(*args: str, check: bool = True) -> subprocess.CompletedProcess:
    return subprocess.run([REF_DRIVER, *args], capture_output=True, text=True, check=check)


def compare_plane_images(a: PlaneImage, b: PlaneImage, what: str = "", check_meta: bool = True) -> None:
    assert len(a.planes) == len(b.planes), f"{what}: plane count {len(a.planes)} vs {len(b.planes)}"
    for i, (p, q) in enumerate(zip(a.planes, b.planes)):
        if check_meta:
            assert p.meta() == q.meta(), f"{what}: plane {i} meta {p.meta()} vs {q.meta()}"
        else:
            assert (p.w, p.h) == (q.w, q.h), f"{what}: plane {i} dims"
        if p.data is None or q.data is None:
            pz = p.data is None or not p.data.any()
            qz = q.data is None or not q.data.any()
            assert pz and qz, f"{what}: plane {i} decoded on one side only"
            continue
        if not np.array_equal(p.data, q.data):
            bad = np.argwhere(p.data != q.data)
            raise AssertionError(f"{what}: plane {i} ({p.w}x{p.h}) differs at {len(bad)} samples, first {bad[0]}: "
                                 f"{p.data[tuple(bad[0])]} vs {q.data[tuple(bad[0])]}")
