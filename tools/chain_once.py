"""Inverse transform chain alone (no entropy stage): builds the coefficient planes of a synthetic image with the
library's own forward transforms on the GPU, then times Image::undo_transforms with CUDA events (L2 flushed before
every iteration) and prints per-launch times from the library's timing report.

usage: chain_once.py W H C [iters] [--dct]      tunables come from the FB_FQ_* / FB_SQUEEZE_MODE environment."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fuif_b200 import api
from fuif_b200.synth import synth_image
from tests.util import default_squeeze_parameters

w, h, c = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 and not sys.argv[4].startswith("-") else 5
dct = "--dct" in sys.argv
maxval = 255 if c != 4 else 16383
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = api.Context(0, stream.cuda_stream)
pix = synth_image(w, h, c, maxval, seed=7)
img = api.Image.from_pixels(pix, maxval, ctx)
if dct:
    sq = default_squeeze_parameters((w + 7) // 8, (h + 7) // 8, 3)
    for tid, params in ((0, []), (4, [0, 2]), (5, [8, 12, 12] * 64), (7, sq)):
        assert img.do_transform(api.Transform(tid, params))
else:
    if c >= 3:
        assert img.do_transform(api.Transform(1))
    assert img.do_transform(api.Transform(7, default_squeeze_parameters(w, h, c)))
inf = img.info()
planes = img.channels()
trs = img.transform
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
times, reports = [], []
ctx.enable_kernel_timing(True)
for it in range(iters + 2):
    g = api.Image.from_planes(inf.w, inf.h, inf.minval, inf.maxval, inf.nb_channels, inf.real_nb_channels, inf.nb_meta_channels, inf.colormodel, planes, trs, ctx)
    flush.zero_()
    torch.cuda.synchronize()
    ctx.timing_report()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches
    e0.record(stream); g.undo_transforms(0); e1.record(stream)
    torch.cuda.synchronize()
    rep = ctx.timing_report()
    if it >= 2:
        times.append(e0.elapsed_time(e1))
        reports.append(rep)
    nl = ctx.launches - l0
exact = bool(np.array_equal(g.pixels(), pix)) if not dct else None
alg = 4.0 * w * h * c
ms = float(np.mean(times))
kern = {}
for rep in reports:
    for i, (name, us, b) in enumerate(rep[1:]):       # rep[0] is the interval since the report reset
        kern.setdefault(f"{i}:{name}", []).append((us, b))
peak = 6538.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = {"shape": [w, h, c], "dct": dct, "env": {k: v for k, v in os.environ.items() if k.startswith("FB_")}, "chain_ms_mean": ms, "chain_ms_min": float(np.min(times)),
       "launches": nl, "chain_GBps": alg / ms / 1e6, "chain_frac": alg / ms / 1e6 / peak, "exact": exact,
       "repaired_tiles": ctx.repaired_tiles, "serial_fallbacks": ctx.fallbacks,
       "kernels": {k: {"us": round(float(np.mean([u for u, _ in v])), 2), "GBps": round(v[0][1] / np.mean([u for u, _ in v]) / 1e3, 1) if v[0][1] else None} for k, v in kern.items()}}
print(json.dumps(out))
