#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout -s KILL 400 python bench.py --workload cfg4 --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "bench cfg4 rc=$?"; tail -c 300 gpurun_out/bench_cfg4.err; python tools/show_bench.py gpurun_out/bench_cfg4.json | cut -c1-400
: > gpurun_out/sweep_d.jsonl
run() { echo "== $*" >> gpurun_out/sweep_d.jsonl; env "$@" timeout 120 python tools/chain_once.py ${SHAPE:-4096 4096 3} 5 >> gpurun_out/sweep_d.jsonl 2>> gpurun_out/sweep_d.err; }
run FB_DQ_HS=64
run FB_DQ_HS=128
python - <<'PY'
import json
for ln in open('gpurun_out/sweep_d.jsonl'):
    if ln.startswith('=='): print(ln.strip()); continue
    try: d = json.loads(ln)
    except Exception: print(ln[:200]); continue
    print("  %s chain %.1f us (min %.1f) launches %d frac %.3f exact %s" % (d['shape'], d['chain_ms_mean']*1e3, d['chain_ms_min']*1e3, d['launches'], d['chain_frac'], d['exact']))
    print("  ", {k.split(':')[0]+k.split(':')[1][6:12]: v['us'] for k, v in d['kernels'].items()})
PY
