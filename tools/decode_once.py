"""Decodes one workload image once on cuda:0 (profiling helper: run under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fuif_b200 import api
import bench_workloads as wl
name = sys.argv[1] if len(sys.argv) > 1 else "mid"
undo = "--undo" in sys.argv
noindex = "--no-index" in sys.argv
reps = 1
for a in sys.argv:
    if a.startswith("--reps="):
        reps = int(a.split("=")[1])
im = wl.prepare_image(name, want_index=not noindex)
ctx = api.Context(0)
import time
for _ in range(reps):
    t0 = time.perf_counter()
    img = api.fuif_decode(im["fuif"], ctx=ctx, group_index=im["index"])
    print("fuif_decode wall %.3f s" % (time.perf_counter() - t0))
    if undo:
        img.undo_transforms(0)
    ctx.synchronize()
print("decoded", name, img.info().w, img.info().h, "launches", ctx.launches)
