"""Deterministic synthetic images for parity tests and benchmarks.

per channel k:  v = clip(0.5 + 0.25*sin(x/(37+11k)) + 0.25*cos(y/(53+7k))*sin((x+y)/(101+k)) + N(0, 0.02), 0, 1) * maxval
rounded with rint, NumPy default_rng(seed).  There is no image corpus in the container, so every
benchmark / test input is produced by this generator and written as binary PNM/PAM, the one input
format the reference reads without external libraries (reference import/read_pam.h:46-156).
"""
from __future__ import annotations

import numpy as np


def synth_image(w: int, h: int, nchan: int = 3, maxval: int = 255, seed: int = 7, noise: float = 0.02) -> np.ndarray:
    """Returns an (h, w, nchan) int32 array with values in [0, maxval]."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    out = np.empty((h, w, nchan), dtype=np.int32)
    for k in range(nchan):
        v = 0.5 + 0.25 * np.sin(x / (37 + 11 * k)) + 0.25 * np.cos(y / (53 + 7 * k)) * np.sin((x + y) / (101 + k))
        v = v + rng.normal(0.0, noise, size=(h, w))
        out[:, :, k] = np.rint(np.clip(v, 0.0, 1.0) * maxval).astype(np.int32)
    return out


def write_pnm(path: str, img: np.ndarray, maxval: int = 255) -> None:
    """Writes P5 (1 channel), P6 (3 channels) or P7 (2 or 4 channels); 16-bit samples are big-endian."""
    h, w, c = img.shape
    if maxval > 255:
        data = img.astype(">u2").tobytes()
    else:
        data = img.astype(np.uint8).tobytes()
    with open(path, "wb") as f:
        if c == 1:
            f.write(b"P5\n%d %d\n%d\n" % (w, h, maxval))
        elif c == 3:
            f.write(b"P6\n%d %d\n%d\n" % (w, h, maxval))
        else:
            tupl = {2: b"GRAYSCALE_ALPHA", 4: b"RGB_ALPHA"}[c]
            f.write(b"P7\nWIDTH %d\nHEIGHT %d\nDEPTH %d\nMAXVAL %d\nTUPLTYPE %s\nENDHDR\n" % (w, h, c, maxval, tupl))
        f.write(data)


def read_pnm(path: str) -> tuple[np.ndarray, int]:
    """Reads P5/P6/P7 written by write_pnm or by the reference's write_PAM_file. Returns ((h,w,c) int32, maxval)."""
    with open(path, "rb") as f:
        buf = f.read()
    pos = 0

    def token() -> bytes:
        nonlocal pos
        while buf[pos:pos + 1].isspace():
            pos += 1
        start = pos
        while not buf[pos:pos + 1].isspace():
            pos += 1
        return buf[start:pos]

    magic = token()
    if magic in (b"P5", b"P6"):
        w = int(token()); h = int(token()); maxval = int(token())
        pos += 1
        c = 1 if magic == b"P5" else 3
    elif magic == b"P7":
        w = h = c = maxval = 0
        while True:
            t = token()
            if t == b"ENDHDR":
                pos += 1
                break
            if t == b"WIDTH": w = int(token())
            elif t == b"HEIGHT": h = int(token())
            elif t == b"DEPTH": c = int(token())
            elif t == b"MAXVAL": maxval = int(token())
            elif t == b"TUPLTYPE": token()
    else:
        raise ValueError("unsupported PNM magic %r" % magic)
    n = w * h * c
    if maxval > 255:
        a = np.frombuffer(buf, dtype=">u2", count=n, offset=pos).astype(np.int32)
    else:
        a = np.frombuffer(buf, dtype=np.uint8, count=n, offset=pos).astype(np.int32)
    return a.reshape(h, w, c), maxval
