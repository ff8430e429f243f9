"""Child process of tests/test_host_entropy.py::test_host_fuzz_returns: mutated files and bogus group indexes through the host
entropy backend, in-process -- a crash or a hang shows up as this process dying or timing out.  usage: seed iterations"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fuif_b200 import api  # noqa: E402
from tests.util import load_golden  # noqa: E402

names = ["sq128", "rgba14", "dct", "unc", "pred", "pal", "match", "e0", "odd"]
rng = np.random.default_rng(int(sys.argv[1]))
clean = {n: bytes(load_golden(n)["fuif"]) for n in names}
index = {}
for n in names:
    offs, first = api.fuif_host_decode(clean[n], threads=1).group_index()
    index[n] = (list(offs), list(first))
n_ok = n_err = 0
for it in range(int(sys.argv[2])):
    name = names[it % len(names)]
    data = bytearray(clean[name])
    mode = it % 4
    for _ in range(1 + int(rng.integers(0, 6))):
        pos = int(rng.integers(min(40, len(data) - 1), len(data)))
        data[pos] = [data[pos] ^ (1 << int(rng.integers(0, 8))), int(rng.integers(0, 256)), 0xFF, 0][mode]
    gi = None
    if it % 3 == 0:
        offs, first = index[name]
        if it % 2:
            offs = [int(x) for x in rng.integers(0, len(data), len(offs))]
        gi = (offs, first)
    try:
        api.fuif_host_decode(bytes(data), group_index=gi, threads=3)
        n_ok += 1
    except api.FuifError:
        n_err += 1
print(f"returned {n_ok + n_err} times: {n_ok} decodes, {n_err} errors")
