#!/bin/bash
# Round-1 GPU job B: parity tests + chain timings of the three unsqueeze paths.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep_b.jsonl
run() { echo "== $*" >> gpurun_out/sweep_b.jsonl; env "$@" timeout 120 python tools/chain_once.py ${SHAPE:-4096 4096 3} 5 >> gpurun_out/sweep_b.jsonl 2>> gpurun_out/sweep_b.err; }
run FB_SQUEEZE_MODE=direct
run FB_SQUEEZE_MODE=perlevel
SHAPE="1920 1080 3" run FB_SQUEEZE_MODE=direct
SHAPE="8192 8192 4" run FB_SQUEEZE_MODE=direct
SHAPE="8192 8192 4" run FB_SQUEEZE_MODE=perlevel
python - <<'PY'
import json
for ln in open('gpurun_out/sweep_b.jsonl'):
    if ln.startswith('=='): print(ln.strip()); continue
    try: d = json.loads(ln)
    except Exception: print(ln[:200]); continue
    print("  %s chain %.1f us (min %.1f) launches %d frac %.3f exact %s" % (d['shape'], d['chain_ms_mean']*1e3, d['chain_ms_min']*1e3, d['launches'], d['chain_frac'], d['exact']))
    print("  ", {k: (v['us'], v['GBps']) for k, v in d['kernels'].items()})
PY
tail -5 gpurun_out/sweep_b.err
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/chain_once.py 1024 512 3 1 > gpurun_out/sanitizer_memcheck_b.log 2>&1; echo "memcheck rc=$?"; tail -1 gpurun_out/sanitizer_memcheck_b.log
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/chain_once.py 512 256 3 1 > gpurun_out/sanitizer_racecheck_b.log 2>&1; echo "racecheck rc=$?"; tail -1 gpurun_out/sanitizer_racecheck_b.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_inv_hsq_direct -s 8 -c 1 -f -o gpurun_out/hsq_direct_ycocg python tools/chain_once.py 4096 4096 3 1 > gpurun_out/ncu_hd.log 2>&1; echo "ncu h rc=$?"; tail -2 gpurun_out/ncu_hd.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_inv_vsq_direct -s 8 -c 1 -f -o gpurun_out/vsq_direct_luma python tools/chain_once.py 4096 4096 3 1 > gpurun_out/ncu_vd.log 2>&1; echo "ncu v rc=$?"; tail -2 gpurun_out/ncu_vd.log
