#!/bin/bash
mkdir -p gpurun_out
FB_KERNEL_TIMING=1 timeout 300 python - > gpurun_out/timing.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, '.')
import numpy as np, torch
from fuif_b200 import api
import bench_workloads as wl
im = wl.prepare_image("cfg2")
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx = api.Context(0, st.cuda_stream)
img = api.fuif_decode(im["fuif"], ctx=ctx, group_index=im["index"])
inf = img.info(); planes = img.channels(); trs = img.transform
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(3):
    g = api.Image.from_planes(inf.w, inf.h, inf.minval, inf.maxval, inf.nb_channels, inf.real_nb_channels, inf.nb_meta_channels, inf.colormodel, planes, trs, ctx)
    flush.zero_(); torch.cuda.synchronize()
    ctx.synchronize()
    print("---- iteration", it, file=sys.stderr)
    g.undo_transforms(0)
    ctx.synchronize()
PY
grep -A16 "iteration 2" gpurun_out/timing.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/chain_bench.py cfg2 5 2>&1 | tail -1
