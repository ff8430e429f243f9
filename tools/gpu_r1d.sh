#!/bin/bash
# Round-1 final check: bench on cfg2 (both numbers the driver reproduces), smoke(), batch workload
mkdir -p gpurun_out
timeout -s KILL 400 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench cfg2 rc=$?"; tail -c 300 gpurun_out/bench_cfg2.err; python tools/show_bench.py gpurun_out/bench_cfg2.json
timeout -s KILL 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout -s KILL 200 python bench.py --workload cfg4 --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "bench cfg4 rc=$?"; python tools/show_bench.py gpurun_out/bench_cfg4.json | cut -c1-200
