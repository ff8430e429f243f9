#!/bin/bash
# final validation of the round: batch bench with the three entropy backends, the whole GPU suite, smoke()
mkdir -p gpurun_out
timeout -s KILL 300 python bench.py --workload cfg4 --steps 2 --warmup 1 --no-index-steps 0 > gpurun_out/bench_cfg4_final.json 2> gpurun_out/bench_cfg4_final.err; echo "cfg4 bench rc=$?"; tail -c 600 gpurun_out/bench_cfg4_final.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_cfg4_final.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}); print("e2e", d["e2e"]["value"]); print("host", d["e2e_host_entropy"]); print("hybrid", d["e2e_hybrid_entropy"])
except Exception as e: print("no line", e)
PY
timeout -s KILL 560 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_final.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
