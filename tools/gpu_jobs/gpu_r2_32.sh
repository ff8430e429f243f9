#!/bin/bash
# host entropy backend on the box: its GPU tests, the default bench line (now with e2e_host_entropy), host-only timings
mkdir -p gpurun_out
nproc; grep -m1 "model name" /proc/cpuinfo
timeout -s KILL 600 python -m pytest tests/test_gpu_host_backend.py -x -q 2>&1 | tail -4
timeout -s KILL 600 python bench.py > gpurun_out/bench_cfg2_final.json 2> gpurun_out/bench_cfg2_final.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_cfg2_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cfg2_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","value_no_index","e2e_no_index","gpu_launches")}); print("e2e", d["e2e"]); print("host", d["e2e_host_entropy"]); print("roofline", d["roofline"].get("kernel"), d["roofline"].get("frac"), d["roofline"].get("chain"))
print("cpu", d["cpu_baseline"])
PY
timeout -s KILL 300 python tools/host_entropy_perf.py .bench_cache/cfg2_s7.fuif 16 8 4 2 2>&1 | tee gpurun_out/host_entropy_perf_cfg2.txt
