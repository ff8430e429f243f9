"""Python host side of fuif_b200: a thin ctypes mirror of the reference's Image / Transform / fuif_decode API
(reference image/image.h:54-129, transform/transform.h:77-106, encoding/encoding.h:32-71) over the C ABI in
include/fuif_b200.h.  All sample work happens in the CUDA library; this module only moves descriptors.

There is no CPU fallback: importing the library fails loudly when libfuif_b200.so has not been built, and
creating a Context fails loudly when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfuif_b200.so")

# transform ids (reference transform/transform.h:30-70)
TRANSFORM_YCbCr = 0
TRANSFORM_YCoCg = 1
TRANSFORM_ChromaSubsample = 3
TRANSFORM_2DMATCH = 8
TRANSFORM_PERMUTE = 9
TRANSFORM_APPROXIMATE = 10
TRANSFORM_DCT = 4
TRANSFORM_QUANTIZE = 5
TRANSFORM_PALETTE = 6
TRANSFORM_SQUEEZE = 7

FB_OK = 0
FB_OPT_SQUEEZE_MODE = 1
FB_OPT_KERNEL_TIMING = 2
FB_OPT_SQUEEZE_PACKED = 3
FB_OPT_ENTROPY_BACKEND = 4
FB_OPT_HOST_THREADS = 5
FB_ENTROPY_GPU, FB_ENTROPY_HOST = 0, 1


class FuifError(RuntimeError):
    pass


class PlaneDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("w", "h", "minval", "maxval", "zero", "q", "hshift", "vshift", "hcshift", "vcshift", "component", "decoded")]


class ImageInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("w", "h", "minval", "maxval", "nb_channels", "real_nb_channels", "nb_meta_channels", "colormodel",
                                         "nb_planes", "nb_transforms", "error")]


class DecodeOptions(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("preview", "maniac_cutoff", "maniac_alpha", "reserved")]


class EncodeOptions(C.Structure):
    _fields_ = [("nb_repeats", C.c_float), ("max_properties", C.c_int32), ("maniac_cutoff", C.c_int32), ("maniac_alpha", C.c_int32),
                ("compress", C.c_int32), ("max_group", C.c_int32), ("n_predictors", C.c_int32), ("predictor", C.POINTER(C.c_int32))]


# every symbol include/fuif_b200.h declares (tests/test_abi.py checks that the library exports all of them)
ABI_SYMBOLS = [
    "fb_ctx_create", "fb_ctx_destroy", "fb_last_error", "fb_ctx_synchronize", "fb_ctx_launch_count",
    "fb_decode", "fb_decode_batch", "fb_image_group_index",
    "fb_image_create", "fb_image_destroy", "fb_image_get_info", "fb_image_get_plane", "fb_image_get_transform",
    "fb_image_plane_device_ptr", "fb_image_download_plane", "fb_image_download_interleaved",
    "fb_image_undo_transforms", "fb_image_do_transform", "fb_image_recompute_minmax",
    "fb_decode_to_pixels", "fb_peek_header", "fb_ctx_set_option", "fb_ctx_counter", "fb_ctx_timing_report",
    "fb_encode", "fb_free", "fb_selftest_packed", "fb_host_decode", "fb_host_last_error", "fb_image_upload",
]

_lib = None


def load_library():
    """Loads libfuif_b200.so (built in-tree by __graft_entry__.build() / fuif_b200/csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(fuif_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32p, i64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    L.fb_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.fb_ctx_destroy.argtypes = [vp]
    L.fb_ctx_destroy.restype = None
    L.fb_last_error.argtypes = [vp]
    L.fb_last_error.restype = C.c_char_p
    L.fb_ctx_synchronize.argtypes = [vp]
    L.fb_ctx_launch_count.argtypes = [vp]
    L.fb_ctx_launch_count.restype = C.c_longlong
    L.fb_ctx_set_option.argtypes = [vp, C.c_int, C.c_int]
    L.fb_ctx_counter.argtypes = [vp, C.c_int]
    L.fb_ctx_counter.restype = C.c_longlong
    L.fb_selftest_packed.argtypes = [vp, C.c_int, C.c_uint, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
    L.fb_ctx_timing_report.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.fb_ctx_timing_report.restype = C.c_longlong
    L.fb_decode.argtypes = [vp, vp, C.c_size_t, C.POINTER(DecodeOptions), i64p, i32p, C.c_int, C.POINTER(vp)]
    L.fb_decode_batch.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(DecodeOptions), C.POINTER(i64p),
                                  C.POINTER(i32p), C.POINTER(C.c_int), C.POINTER(vp)]
    L.fb_image_group_index.argtypes = [vp, i64p, i32p, C.c_int]
    L.fb_image_create.argtypes = [vp, C.POINTER(ImageInfo), C.POINTER(PlaneDesc), C.POINTER(vp), i32p, i32p, i32p, C.POINTER(vp)]
    L.fb_image_destroy.argtypes = [vp]
    L.fb_image_destroy.restype = None
    L.fb_image_get_info.argtypes = [vp, C.POINTER(ImageInfo)]
    L.fb_image_get_plane.argtypes = [vp, C.c_int, C.POINTER(PlaneDesc)]
    L.fb_image_get_transform.argtypes = [vp, C.c_int, i32p, i32p, C.c_int]
    L.fb_image_plane_device_ptr.argtypes = [vp, C.c_int]
    L.fb_image_plane_device_ptr.restype = vp
    L.fb_image_download_plane.argtypes = [vp, C.c_int, vp]
    L.fb_image_download_interleaved.argtypes = [vp, C.c_int, C.c_int, vp]
    L.fb_image_undo_transforms.argtypes = [vp, C.c_int]
    L.fb_image_do_transform.argtypes = [vp, C.c_int32, i32p, C.c_int, C.POINTER(C.c_int)]
    L.fb_image_recompute_minmax.argtypes = [vp]
    L.fb_decode_to_pixels.argtypes = [vp, vp, C.c_size_t, C.POINTER(DecodeOptions), i64p, i32p, C.c_int, C.c_int, vp, C.c_size_t]
    L.fb_peek_header.argtypes = [vp, C.c_size_t, C.POINTER(ImageInfo)]
    L.fb_encode.argtypes = [vp, vp, C.POINTER(EncodeOptions), C.POINTER(vp), C.POINTER(C.c_size_t), i64p, i32p, C.c_int, C.POINTER(C.c_int)]
    L.fb_free.argtypes = [vp]
    L.fb_free.restype = None
    L.fb_host_decode.argtypes = [vp, C.c_size_t, C.POINTER(DecodeOptions), i64p, i32p, C.c_int, C.c_int, C.POINTER(vp)]
    L.fb_host_last_error.argtypes = []
    L.fb_host_last_error.restype = C.c_char_p
    L.fb_image_upload.argtypes = [vp, vp]
    _lib = L
    return L


@dataclass
class fuif_options:
    """struct fuif_options (reference encoding/encoding.h:32-59); the encode-side members are read by fuif_encode only."""
    preview: int = -1
    maniac_cutoff: int = 6
    maniac_alpha: int = 0x0d000000
    nb_repeats: float = 0.5
    max_properties: int = 12
    compress: bool = True
    max_group: int = -1
    predictor: list = field(default_factory=list)

    def _c(self) -> DecodeOptions:
        return DecodeOptions(self.preview, self.maniac_cutoff, self.maniac_alpha, 0)


default_fuif_options = fuif_options()


@dataclass
class Transform:
    """class Transform (reference transform/transform.h:77-106)."""
    ID: int
    parameters: list

    def __init__(self, ID: int, parameters=()):
        self.ID = ID
        self.parameters = list(parameters)


@dataclass
class Channel:
    """class Channel (reference image/image.h:54-91) with its samples downloaded to the host."""
    w: int
    h: int
    minval: int
    maxval: int
    zero: int
    q: int
    hshift: int
    vshift: int
    hcshift: int
    vcshift: int
    component: int
    data: np.ndarray | None

    def meta(self):
        return (self.w, self.h, self.minval, self.maxval, self.q, self.hshift, self.vshift, self.hcshift, self.vcshift, self.component)


class Context:
    """One CUDA device + stream (fb_ctx)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.fb_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != FB_OK:
            raise FuifError(f"fb_ctx_create(device={device}) failed with {rc}: no usable CUDA device (fuif_b200 has no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.fb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int, what: str):
        if rc != FB_OK:
            raise FuifError(f"{what} failed ({rc}): {self.lib.fb_last_error(self.h).decode()}")

    def synchronize(self):
        self.check(self.lib.fb_ctx_synchronize(self.h), "fb_ctx_synchronize")

    @property
    def launches(self) -> int:
        return int(self.lib.fb_ctx_launch_count(self.h))

    def set_squeeze_mode(self, mode: int) -> None:
        """0 direct per-step kernels (default), 1 tiled per-step kernels only, 4 fused tile kernels, 2 fused + forced serial
        recompute, 3 fused + forced repair of every tile of the last launch (tests)."""
        self.check(self.lib.fb_ctx_set_option(self.h, FB_OPT_SQUEEZE_MODE, mode), "fb_ctx_set_option")

    @property
    def fallbacks(self) -> int:
        """Squeeze inverses that failed verification in an early launch (recomputed serially; still exact)."""
        return int(self.lib.fb_ctx_counter(self.h, 0))

    @property
    def repaired_tiles(self) -> int:
        """Tiles of last launches whose speculative start failed verification (recomputed from exact states)."""
        return int(self.lib.fb_ctx_counter(self.h, 1))

    def set_squeeze_packed(self, on: bool) -> None:
        """Packed int16x2 unsqueeze kernels (TMA-fed; default on for images with maxval <= 1023); off = the 32-bit kernels."""
        self.check(self.lib.fb_ctx_set_option(self.h, FB_OPT_SQUEEZE_PACKED, 1 if on else 0), "fb_ctx_set_option")

    @property
    def pk_repaired(self) -> int:
        """Segments the packed unsqueeze kernels recomputed with the exact routine (speculation miss or range flag)."""
        return int(self.lib.fb_ctx_counter(self.h, 2))

    @property
    def pk_range_flagged(self) -> int:
        """Segments in which the packed unsqueeze kernels met a value outside the packed range."""
        return int(self.lib.fb_ctx_counter(self.h, 3))

    def selftest_packed(self, which: int, seed: int = 1, scale: int = 64, maxval: int = 255) -> int:
        """Mismatches of the packed 16x2 primitives against their exact forms on the device (0 = correct)."""
        n = C.c_longlong(-1)
        self.check(self.lib.fb_selftest_packed(self.h, which, seed, scale, maxval, C.byref(n)), "fb_selftest_packed")
        return int(n.value)

    def set_entropy_backend(self, backend: str, threads: int = 0) -> None:
        """'gpu' (default): k_maniac_decode; 'host': fuif_decode_channel on CPU threads (0 = four per hardware thread, at most one per stream), planes uploaded
        afterwards -- the faster backend for ONE large image with a group index.  Identical planes and errors."""
        self.check(self.lib.fb_ctx_set_option(self.h, FB_OPT_ENTROPY_BACKEND, {"gpu": FB_ENTROPY_GPU, "host": FB_ENTROPY_HOST}[backend]), "fb_ctx_set_option")
        self.check(self.lib.fb_ctx_set_option(self.h, FB_OPT_HOST_THREADS, threads), "fb_ctx_set_option")

    @property
    def host_threads_used(self) -> int:
        """Threads the host entropy backend used in its last call."""
        return int(self.lib.fb_ctx_counter(self.h, 4))

    def enable_kernel_timing(self, on: bool = True) -> None:
        self.check(self.lib.fb_ctx_set_option(self.h, FB_OPT_KERNEL_TIMING, 1 if on else 0), "fb_ctx_set_option")

    def timing_report(self) -> list:
        """[(kernel / launcher name, microseconds, algorithmic bytes)] since the last report (synchronises)."""
        buf = C.create_string_buffer(1 << 20)
        self.lib.fb_ctx_timing_report(self.h, buf, len(buf))
        out = []
        for ln in buf.value.decode().splitlines():
            name, us, b = ln.split("\t")
            out.append((name, float(us), float(b)))
        return out


class HostContext:
    """What stands behind a host-only Image (fuif_host_decode): the library without a GPU."""

    def __init__(self):
        self.lib = load_library()
        self.h = True       # "open"; never handed to the library
        self.device = -1

    def check(self, rc: int, what: str):
        if rc != FB_OK:
            raise FuifError(f"{what} failed ({rc}): {self.lib.fb_host_last_error().decode()}")


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class Image:
    """class Image (reference image/image.h:98-129), resident in HBM (fb_image)."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._handle = handle

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        if self.__dict__.get("_handle"):
            if self.ctx.h:      # a closed context has already released the device
                self.ctx.lib.fb_image_destroy(self._handle)
            self._handle = None

    def upload(self, ctx: "Context") -> "Image":
        """Moves a host-only image (fuif_host_decode) into ctx's HBM."""
        ctx.check(ctx.lib.fb_image_upload(ctx.h, self._handle), "fb_image_upload")
        self.ctx = ctx
        return self

    # ---- construction ----------------------------------------------------------------------------------------
    @staticmethod
    def from_pixels(pix: np.ndarray, maxval: int, ctx: Context | None = None) -> "Image":
        """(h, w, c) integers -> Image as read_PAM_file builds it (reference import/read_pam.h:126-153)."""
        hh, ww, cc = pix.shape
        planes = [Channel(ww, hh, 0, maxval, 0, 1, 0, 0, 0, 0, i, np.ascontiguousarray(pix[:, :, i].astype(np.int16))) for i in range(cc)]
        return Image.from_planes(ww, hh, 0, maxval, cc, cc, 0, 0, planes, [], ctx)

    @staticmethod
    def from_planes(w, h, minval, maxval, nb_channels, real_nb_channels, nb_meta_channels, colormodel, planes, transforms, ctx=None) -> "Image":
        ctx = ctx or default_context()
        L = ctx.lib
        info = ImageInfo(w, h, minval, maxval, nb_channels, real_nb_channels, nb_meta_channels, colormodel, len(planes), len(transforms), 0)
        desc = (PlaneDesc * max(1, len(planes)))()
        ptrs = (C.c_void_p * max(1, len(planes)))()
        keep = []
        for i, p in enumerate(planes):
            has = p.data is not None
            desc[i] = PlaneDesc(p.w, p.h, p.minval, p.maxval, p.zero, p.q, p.hshift, p.vshift, p.hcshift, p.vcshift, p.component, 1 if has else 0)
            if has:
                a = np.ascontiguousarray(p.data, dtype=np.int16)
                assert a.size == p.w * p.h
                keep.append(a)
                ptrs[i] = a.ctypes.data
        ids = (C.c_int32 * max(1, len(transforms)))(*[t.ID for t in transforms])
        nps = (C.c_int32 * max(1, len(transforms)))(*[len(t.parameters) for t in transforms])
        flat = [x for t in transforms for x in t.parameters]
        par = (C.c_int32 * max(1, len(flat)))(*flat)
        h_ = C.c_void_p()
        ctx.check(L.fb_image_create(ctx.h, C.byref(info), desc, ptrs, ids, nps, par, C.byref(h_)), "fb_image_create")
        ctx.synchronize()
        return Image(ctx, h_)

    # ---- queries ---------------------------------------------------------------------------------------------
    def info(self) -> ImageInfo:
        inf = ImageInfo()
        self.ctx.check(self.ctx.lib.fb_image_get_info(self._handle, C.byref(inf)), "fb_image_get_info")
        return inf

    def __getattr__(self, name):
        if name in ("w", "h", "minval", "maxval", "nb_channels", "real_nb_channels", "nb_meta_channels", "colormodel", "error"):
            return getattr(self.info(), name)
        raise AttributeError(name)

    @property
    def transform(self) -> list:
        out = []
        n = self.info().nb_transforms
        for i in range(n):
            tid = C.c_int32()
            buf = (C.c_int32 * 4096)()
            k = self.ctx.lib.fb_image_get_transform(self._handle, i, C.byref(tid), buf, 4096)
            out.append(Transform(tid.value, list(buf[:k])))
        return out

    def plane_desc(self, i: int) -> PlaneDesc:
        d = PlaneDesc()
        self.ctx.check(self.ctx.lib.fb_image_get_plane(self._handle, i, C.byref(d)), "fb_image_get_plane")
        return d

    def nb_planes(self) -> int:
        return self.info().nb_planes

    def channel(self, i: int, download: bool = True) -> Channel:
        d = self.plane_desc(i)
        data = None
        if d.decoded and download:
            data = np.empty((d.h, d.w), dtype=np.int16)
            self.ctx.check(self.ctx.lib.fb_image_download_plane(self._handle, i, data.ctypes.data), "fb_image_download_plane")
        return Channel(d.w, d.h, d.minval, d.maxval, d.zero, d.q, d.hshift, d.vshift, d.hcshift, d.vcshift, d.component, data)

    def channels(self) -> list:
        return [self.channel(i) for i in range(self.nb_planes())]

    def device_ptr(self, i: int) -> int:
        return int(self.ctx.lib.fb_image_plane_device_ptr(self._handle, i) or 0)

    def group_index(self):
        n = self.ctx.lib.fb_image_group_index(self._handle, None, None, 0)
        offs = (C.c_int64 * max(1, n))()
        first = (C.c_int32 * max(1, n))()
        self.ctx.lib.fb_image_group_index(self._handle, offs, first, n)
        return list(offs[:n]), list(first[:n])

    # ---- the reference's Image methods -----------------------------------------------------------------------------
    def undo_transforms(self, keep: int = 0) -> None:
        """Image::undo_transforms (reference image/image.cpp:94-115)."""
        self.ctx.check(self.ctx.lib.fb_image_undo_transforms(self._handle, keep), "fb_image_undo_transforms")

    def do_transform(self, t: Transform) -> bool:
        """Image::do_transform (reference image/image.cpp:117-122)."""
        par = (C.c_int32 * max(1, len(t.parameters)))(*t.parameters)
        applied = C.c_int()
        self.ctx.check(self.ctx.lib.fb_image_do_transform(self._handle, t.ID, par, len(t.parameters), C.byref(applied)), "fb_image_do_transform")
        return bool(applied.value)

    def recompute_minmax(self) -> None:
        self.ctx.check(self.ctx.lib.fb_image_recompute_minmax(self._handle), "fb_image_recompute_minmax")

    def pixels(self) -> np.ndarray:
        """(h, w, nb_channels) int32, via the interleave kernel (layout of write_PAM_file, export/write_pam.h:29-168)."""
        inf = self.info()
        d = self.plane_desc(0)
        bps = 2 if inf.maxval > 255 else 1
        out = np.empty((d.h, d.w, inf.nb_channels), dtype=(">u2" if bps == 2 else np.uint8))
        self.ctx.check(self.ctx.lib.fb_image_download_interleaved(self._handle, inf.nb_channels, bps, out.ctypes.data), "fb_image_download_interleaved")
        return out.astype(np.int32)


def _index_args(group_index):
    """group_index: None, a list of byte offsets, or (offsets, first_channels) as Image.group_index() returns it."""
    if group_index is None or len(group_index) == 0:
        return None, None, 0
    first = None
    if isinstance(group_index, tuple):
        group_index, first = group_index
    arr = (C.c_int64 * len(group_index))(*group_index)
    farr = (C.c_int32 * len(first))(*first) if first is not None else None
    return arr, farr, len(group_index)


def fuif_decode(data: bytes, options: fuif_options = default_fuif_options, ctx: Context | None = None, group_index=None) -> Image:
    """fuif_decode (reference encoding/encoding.cpp:599-720): .fuif bytes -> Image holding the transformed planes."""
    ctx = ctx or default_context()
    arr, farr, n = _index_args(group_index)
    h = C.c_void_p()
    opts = options._c()
    if isinstance(data, tuple):         # (device pointer, nbytes): the file already lives in HBM
        ptr, size = data
    else:
        buf = np.frombuffer(data, dtype=np.uint8)
        ptr, size = buf.ctypes.data, buf.size
    ctx.check(ctx.lib.fb_decode(ctx.h, ptr, size, C.byref(opts), arr, farr, n, C.byref(h)), "fb_decode")
    return Image(ctx, h)


def fuif_host_decode(data: bytes, options: fuif_options = default_fuif_options, group_index=None, threads: int = 0) -> Image:
    """fuif_decode on CPU threads only (fb_host_decode): no GPU, no Context.  The Image lives in host memory: planes, ranges, the
    transform list and the group index can be read; Image.upload(ctx) moves it to a GPU for everything else."""
    hc = HostContext()
    arr, farr, n = _index_args(group_index)
    h = C.c_void_p()
    opts = options._c()
    buf = np.frombuffer(data, dtype=np.uint8)
    hc.check(hc.lib.fb_host_decode(buf.ctypes.data, buf.size, C.byref(opts), arr, farr, n, threads, C.byref(h)), "fb_host_decode")
    return Image(hc, h)


def fuif_encode(image: "Image", options: fuif_options = default_fuif_options, want_index: bool = False):
    """fuif_prepare_encode + fuif_encode (reference encoding/encoding.cpp:737-743, 455-573): the .fuif bytes of an image whose
    forward transforms have been applied; with want_index also (offsets, first channels) of the channel groups, the sidecar
    index fuif_decode() takes."""
    ctx = image.ctx
    pred = (C.c_int32 * max(1, len(options.predictor)))(*options.predictor)
    eo = EncodeOptions(options.nb_repeats, options.max_properties, options.maniac_cutoff, options.maniac_alpha, 1 if options.compress else 0,
                       options.max_group, len(options.predictor), pred)
    out = C.c_void_p()
    n = C.c_size_t()
    cap = 4096
    offs = (C.c_int64 * cap)()
    first = (C.c_int32 * cap)()
    ng = C.c_int()
    ctx.check(ctx.lib.fb_encode(ctx.h, image._handle, C.byref(eo), C.byref(out), C.byref(n), offs, first, cap, C.byref(ng)), "fb_encode")
    try:
        data = C.string_at(out, n.value)
    finally:
        ctx.lib.fb_free(out)
    if want_index:
        return data, (list(offs[:ng.value]), list(first[:ng.value]))
    return data


def fuif_encode_file(filename: str, image: "Image", options: fuif_options = default_fuif_options) -> None:
    """fuif_encode_file (reference encoding/encoding.cpp:722-735)."""
    with open(filename, "wb") as f:
        f.write(fuif_encode(image, options))


def fuif_decode_file(filename: str, options: fuif_options = default_fuif_options, ctx: Context | None = None, group_index=None) -> Image:
    """fuif_decode_file (reference encoding/encoding.cpp:745-753)."""
    with open(filename, "rb") as f:
        return fuif_decode(f.read(), options, ctx, group_index)


def fuif_decode_batch(datas, options: fuif_options = default_fuif_options, ctx: Context | None = None, group_indexes=None) -> list:
    """Decodes several files in one kernel launch: every (image, channel group) is an independent stream."""
    ctx = ctx or default_context()
    n = len(datas)
    bufs, plist, slist = [], [], []
    for d in datas:
        if isinstance(d, tuple):            # (device pointer, nbytes)
            plist.append(d[0]); slist.append(d[1])
        else:
            b = np.frombuffer(d, dtype=np.uint8)
            bufs.append(b); plist.append(b.ctypes.data); slist.append(b.size)
    ptrs = (C.c_void_p * n)(*plist)
    sizes = (C.c_size_t * n)(*slist)
    out = (C.c_void_p * n)()
    opts = options._c()
    gi_ptrs = gf_ptrs = None
    gi_n = (C.c_int * n)()
    keep = []
    if group_indexes is not None:
        gi_ptrs = (C.POINTER(C.c_int64) * n)()
        gf_ptrs = (C.POINTER(C.c_int32) * n)()
        have_first = False
        for i, gi in enumerate(group_indexes):
            if not gi:
                continue
            first = None
            if isinstance(gi, tuple):
                gi, first = gi
            a = (C.c_int64 * len(gi))(*gi)
            keep.append(a)
            gi_ptrs[i] = C.cast(a, C.POINTER(C.c_int64))
            gi_n[i] = len(gi)
            if first is not None:
                f = (C.c_int32 * len(first))(*first)
                keep.append(f)
                gf_ptrs[i] = C.cast(f, C.POINTER(C.c_int32))
                have_first = True
        if not have_first:
            gf_ptrs = None
    ctx.check(ctx.lib.fb_decode_batch(ctx.h, n, ptrs, sizes, C.byref(opts), gi_ptrs, gf_ptrs, gi_n, out), "fb_decode_batch")
    return [Image(ctx, C.c_void_p(out[i])) for i in range(n)]


def decode_to_pixels(data: bytes, options: fuif_options = default_fuif_options, ctx: Context | None = None, group_index=None,
                     out: np.ndarray | None = None) -> np.ndarray:
    """`fuif -d in.fuif out.ppm` in one call (reference fuif.cpp:206-239): host bytes in, host pixels out."""
    ctx = ctx or default_context()
    L = ctx.lib
    buf = np.frombuffer(data, dtype=np.uint8)
    inf = ImageInfo()
    if L.fb_peek_header(buf.ctypes.data, buf.size, C.byref(inf)) != FB_OK:
        raise FuifError("not a FUIF file")
    bps = 2 if inf.maxval > 255 else 1
    if out is None:
        out = np.empty((inf.h, inf.w, inf.nb_channels), dtype=(">u2" if bps == 2 else np.uint8))
    arr, farr, n = _index_args(group_index)
    opts = options._c()
    ctx.check(L.fb_decode_to_pixels(ctx.h, buf.ctypes.data, buf.size, C.byref(opts), arr, farr, n, bps, out.ctypes.data, out.nbytes), "fb_decode_to_pixels")
    return out
