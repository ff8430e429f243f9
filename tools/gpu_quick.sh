#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --workload mid --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_mid.json 2> gpurun_out/bench_mid.err; echo "mid rc=$?"; tail -c 300 gpurun_out/bench_mid.err; python tools/show_bench.py gpurun_out/bench_mid.json
timeout 600 python bench.py --workload cfg2 --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_cfg2q.json 2> gpurun_out/bench_cfg2q.err; echo "cfg2 rc=$?"; tail -c 300 gpurun_out/bench_cfg2q.err; python tools/show_bench.py gpurun_out/bench_cfg2q.json
