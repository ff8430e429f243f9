"""Child process of tests/test_emu_maniac.py::test_kernel_source_survives_damaged_input: the MANIAC decode kernel SOURCE under the
execution-model emulator on damaged files and bogus group offsets.  A spin-wait that never ends (what would hang a GPU) shows up
here as this process timing out.  usage: seed iterations shape"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.pyoracle as po  # noqa: E402
from tests import test_emu_maniac as T  # noqa: E402
from tests.util import load_golden  # noqa: E402

names = ["odd", "gray", "e0", "unc", "pred", "tiny"]
rng = np.random.default_rng(int(sys.argv[1]))
shape = int(sys.argv[3])
info = {}
for n in names:
    data = bytes(load_golden(n)["fuif"])
    full, offs = po.OracleImage.decode(data, want_offsets=True)
    (nbch, _bd, _w, _h, _cm, maxp), _ = T._varints(data, 4, 6)
    info[n] = (data, offs, full.to_plane_image(), nbch - ord('0'), maxp)
hist = {}
for it in range(int(sys.argv[2])):
    data, offs, pi, nbch, maxp = info[names[it % len(names)]]
    d = bytearray(data)
    for _ in range(1 + int(rng.integers(0, 5))):
        p = int(rng.integers(min(40, len(d) - 1), len(d)))
        d[p] = [d[p] ^ (1 << int(rng.integers(0, 8))), int(rng.integers(0, 256)), 0xFF, 0][it % 4]
    nch = len(pi.planes)
    desc = (C.c_int * (5 * nch))()
    ptrs = (C.c_void_p * nch)()
    keep = []
    for i, p in enumerate(pi.planes):
        desc[5 * i:5 * i + 5] = [p.w, p.h, p.hshift, p.vshift, p.q]
        a = np.zeros((max(p.h, 0), max(p.w, 0)), dtype=np.int16)
        keep.append(a)
        ptrs[i] = a.ctypes.data if a.size else None
    chout = (C.c_int * (5 * nch))()
    o = [x for x, _ in offs]
    if it % 3 == 1:     # offsets that do not belong to the file
        o = [int(x) for x in rng.integers(offs[0][0], len(d), len(o))]
    ng = len(o) if it % 3 != 2 else 0
    goff = (C.c_longlong * max(1, ng))(*o[:ng])
    gfirst = (C.c_int * max(1, ng))(*[f for _, f in offs][:ng])
    st = T.lib().emu_maniac_decode(bytes(d), len(d), offs[0][0], maxp, nbch, nch, desc, ptrs, chout, ng, goff, gfirst, shape, 2 if shape else 1, 6, 0x0d000000,
                                   226, 0, 0)
    hist[st] = hist.get(st, 0) + 1
print(f"returned {sum(hist.values())} times: {hist}")
