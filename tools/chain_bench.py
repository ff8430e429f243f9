"""Times Image::undo_transforms alone (inverse transform chain) on a workload: decode once, keep the transformed planes
on the host, then per iteration upload them, flush L2 and time the chain with CUDA events."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fuif_b200 import api
import bench_workloads as wl
from fuif_b200.synth import synth_image
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
spec = wl.WORKLOADS[name]
im = wl.prepare_image(name)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = api.Context(0, stream.cuda_stream)
img = api.fuif_decode(im["fuif"], ctx=ctx, group_index=im["index"])
inf = img.info()
planes = img.channels()
trs = img.transform
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
times = []
for it in range(iters + 2):
    g = api.Image.from_planes(inf.w, inf.h, inf.minval, inf.maxval, inf.nb_channels, inf.real_nb_channels, inf.nb_meta_channels, inf.colormodel, planes, trs, ctx)
    flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches
    e0.record(stream); g.undo_transforms(0); e1.record(stream)
    torch.cuda.synchronize()
    if it >= 2: times.append(e0.elapsed_time(e1))
    nl = ctx.launches - l0
ok = bool(np.array_equal(g.pixels(), synth_image(spec[0], spec[1], spec[2], spec[3], spec[4]))) if name in ("cfg1", "cfg2", "mid") else None
alg = 4.0 * spec[0] * spec[1] * spec[2]
ms = float(np.mean(times))
print(json.dumps({"workload": name, "chain_ms_mean": ms, "chain_ms_min": float(np.min(times)), "launches": nl, "GBps": alg / ms / 1e6, "frac_of_6570": alg / ms / 1e6 / 6570.3, "exact": ok}))
