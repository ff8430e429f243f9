"""Times Image::undo_transforms (inverse Squeeze + YCoCg + clamp) on a synthetic image whose coefficient planes are made by the
forward chain on the GPU (no entropy stage involved): per iteration upload the planes, flush L2, time the chain with CUDA
events on the library's stream, and list every launch with its own time (FB_OPT_KERNEL_TIMING).
usage: chain_synth.py W H C [iters] [packed 0/1] [maxval]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fuif_b200 import api
from fuif_b200.synth import synth_image
from tests.util import default_squeeze_parameters
w, h, c = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
packed = int(sys.argv[5]) if len(sys.argv) > 5 else 1
maxval = int(sys.argv[6]) if len(sys.argv) > 6 else 255
pix = synth_image(w, h, c, maxval, seed=7)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = api.Context(0, stream.cuda_stream)
ctx.set_squeeze_packed(bool(packed))
src = api.Image.from_pixels(pix, maxval, ctx)
if c >= 3:
    assert src.do_transform(api.Transform(1))
assert src.do_transform(api.Transform(7, default_squeeze_parameters(w, h, c)))
inf = src.info()
planes = src.channels()
trs = src.transform
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
times, report = [], []
for it in range(iters + 3):
    g = api.Image.from_planes(inf.w, inf.h, inf.minval, inf.maxval, inf.nb_channels, inf.real_nb_channels, inf.nb_meta_channels, inf.colormodel, planes, trs, ctx)
    flush.zero_()
    torch.cuda.synchronize()
    timing = it == iters + 2
    if timing:
        ctx.enable_kernel_timing(True)
        ctx.timing_report()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches
    e0.record(stream); g.undo_transforms(0); e1.record(stream)
    torch.cuda.synchronize()
    if timing:
        report = ctx.timing_report()
    elif it >= 2:
        times.append(e0.elapsed_time(e1))
    nl = ctx.launches - l0
ok = bool(np.array_equal(g.pixels(), pix))
alg = 4.0 * w * h * c
ms = float(np.mean(times))
print(json.dumps({"shape": [w, h, c, maxval], "packed": packed, "chain_ms_mean": ms, "chain_ms_min": float(np.min(times)), "launches": nl,
                  "GBps": alg / ms / 1e6, "frac_of_6538": alg / ms / 1e6 / 6538.0, "exact": ok, "pk_repaired": ctx.pk_repaired, "pk_range_flagged": ctx.pk_range_flagged}))
for name, us, b in report:
    print(f"  {name:34s} {us:9.1f} us  {b / 1e6:9.2f} MB  {b / us / 1e3 if us > 0 else 0:8.1f} GB/s")
