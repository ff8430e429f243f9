#!/bin/bash
mkdir -p gpurun_out
FB_MANIAC_DEBUG=1 timeout 120 python - > gpurun_out/debug.log 2>&1 <<'PY'
import sys
sys.path.insert(0,'.')
from fuif_b200 import api
from tests.util import load_golden, gpu_plane_image
from oracle import pyoracle as po
blob = load_golden("odd")
ctx = api.Context(0)
try:
    img = api.fuif_decode(blob["fuif"], ctx=ctx)
    ref = po.OracleImage.decode(blob["fuif"])
    po.compare_plane_images(gpu_plane_image(po, img), ref.to_plane_image(), "odd")
    print("OK")
except Exception as e:
    print("ERR", e)
_, offs = po.OracleImage.decode(blob["fuif"], want_offsets=True)
print(offs)
PY
tail -60 gpurun_out/debug.log
