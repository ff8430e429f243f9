"""CPU: the C-ABI library loads and exports every symbol include/fuif_b200.h declares; host-only entry points work;
creating a context without a GPU fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fuif_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from fuif_b200 import api
    lib = api.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fuif_b200.h but not exported"
    assert sorted(api.ABI_SYMBOLS) == names


def test_peek_header_is_host_only():
    from fuif_b200 import api
    from tests.util import load_golden
    import numpy as np
    lib = api.load_library()
    blob = load_golden("rgba14")
    buf = np.frombuffer(blob["fuif"], dtype=np.uint8)
    inf = api.ImageInfo()
    assert lib.fb_peek_header(buf.ctypes.data, buf.size, C.byref(inf)) == 0
    assert (inf.w, inf.h, inf.nb_channels, inf.maxval) == (96, 80, 4, 16383)
    bad = np.frombuffer(b"NOPE" + bytes(20), dtype=np.uint8)
    assert lib.fb_peek_header(bad.ctypes.data, bad.size, C.byref(inf)) != 0


def test_no_cpu_fallback():
    import torch
    from fuif_b200 import api
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.FuifError):
        api.Context(0)
