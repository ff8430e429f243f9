#!/bin/bash
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout -s KILL 900 $R --master-port 29551 bench.py --gpus 8 --workload cfg4 --steps 3 --warmup 2 --no-index-steps 1 > gpurun_out/bench_cfg4_n8_weak.json 2> gpurun_out/bench_cfg4_n8_weak.err; echo "cfg4 weak n8 rc=$?"; tail -c 300 gpurun_out/bench_cfg4_n8_weak.err; tail -1 gpurun_out/bench_cfg4_n8_weak.json | cut -c1-400
timeout -s KILL 600 $R --master-port 29552 bench.py --gpus 8 --steps 1 --warmup 1 --no-index-steps 0 > gpurun_out/bench_cfg2_n8.json 2> gpurun_out/bench_cfg2_n8.err; echo "cfg2 n8 rc=$?"; tail -1 gpurun_out/bench_cfg2_n8.json | cut -c1-300; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cfg2_n8.json').read().strip().splitlines()[-1])
print([round(r['entropy_ms']) for r in d['per_rank']])
PY
