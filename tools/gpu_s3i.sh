#!/bin/bash
# chain: hybrid (direct h + tiled v) vs direct v with deeper prefetch, a few shapes
mkdir -p gpurun_out
: > gpurun_out/sweep_c.jsonl
run() { echo "== $*" >> gpurun_out/sweep_c.jsonl; env "$@" timeout 120 python tools/chain_once.py ${SHAPE:-4096 4096 3} 5 >> gpurun_out/sweep_c.jsonl 2>> gpurun_out/sweep_c.err; }
run FB_X=hybrid
run FB_SQUEEZE_DIRECT_V=1
run FB_SQUEEZE_DIRECT_V=1 FB_DQ_VS=16
run FB_SQUEEZE_DIRECT_V=1 FB_DQ_VS=16 FB_DQ_VT=256
run FB_SQUEEZE_DIRECT_V=1 FB_DQ_VS=32 FB_DQ_VT=256
run FB_SQUEEZE_DIRECT_V=1 FB_DQ_VS=64 FB_DQ_VT=256
python - <<'PY'
import json
for ln in open('gpurun_out/sweep_c.jsonl'):
    if ln.startswith('=='): print(ln.strip()); continue
    try: d = json.loads(ln)
    except Exception: print(ln[:200]); continue
    print("  %s chain %.1f us (min %.1f) launches %d frac %.3f exact %s" % (d['shape'], d['chain_ms_mean']*1e3, d['chain_ms_min']*1e3, d['launches'], d['chain_frac'], d['exact']))
    print("  ", {k.split(':')[0]+k.split(':')[1][6:12]: v['us'] for k, v in d['kernels'].items()})
PY
tail -3 gpurun_out/sweep_c.err
