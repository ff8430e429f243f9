#!/bin/bash
# Round-1 GPU job A: parity tests, tile-shape sweep of the fused unsqueeze chain, bench cfg2, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep.jsonl
run() { echo "== $*" >> gpurun_out/sweep.jsonl; env "$@" timeout 120 python tools/chain_once.py 4096 4096 3 5 >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err; }
run FB_SQUEEZE_MODE=perlevel
run FB_FQ_TILE=64x64
run FB_FQ_TILE=64x64 FB_FQ_THREADS=64
run FB_FQ_TILE=32x32 FB_FQ_THREADS=64
run FB_FQ_TILE=128x64
run FB_FQ_TILE=64x128
run FB_FQ_TILE=128x128
run FB_FQ_TILE=64x64 FB_FQ_COARSE=256
run FB_FQ_TILE=64x64 FB_FQ_LEVELS=6
run FB_FQ_TILE=64x64 FB_FQ_FORCE_FALLBACK=2
run FB_FQ_TILE=64x64 FB_FQ_FORCE_FALLBACK=1
python - <<'PY'
import json
for ln in open('gpurun_out/sweep.jsonl'):
    if ln.startswith('=='): print(ln.strip()); continue
    try: d = json.loads(ln)
    except Exception: print(ln[:200]); continue
    print("  chain %.1f us (min %.1f) launches %d frac %.3f exact %s rep %s fb %s" % (d['chain_ms_mean']*1e3, d['chain_ms_min']*1e3, d['launches'], d['chain_frac'], d['exact'], d['repaired_tiles'], d['serial_fallbacks']))
    print("  ", {k: (v['us'], v['GBps']) for k, v in d['kernels'].items()})
PY
tail -5 gpurun_out/sweep.err
timeout 60 python tools/chain_once.py 1920 1080 3 5 | tee gpurun_out/chain_1080p.json | cut -c1-600
timeout 60 python tools/chain_once.py 8192 8192 4 3 | tee gpurun_out/chain_cfg5.json | cut -c1-600
timeout 60 python tools/chain_once.py 4096 4096 3 3 --dct | tee gpurun_out/chain_dct.json | cut -c1-600
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/chain_once.py 520 392 3 1 > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/sanitizer_memcheck.log; tail -2 gpurun_out/sanitizer_memcheck.log
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/chain_once.py 264 200 3 1 > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/sanitizer_racecheck.log
# ncu: full capture of the last fused launch (4th k_fq_tiles launch of the first undo)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fq_tiles -s 3 -c 1 -f -o gpurun_out/fq_last_cfg2 \
    python tools/chain_once.py 4096 4096 3 1 > gpurun_out/ncu_fq.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_fq.log
timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench cfg2 rc=$?"; tail -c 400 gpurun_out/bench_cfg2.err; python tools/show_bench.py gpurun_out/bench_cfg2.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_cfg2.csv \
    python bench.py --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches_bench_cfg2.csv
ls -la gpurun_out | tail -12
