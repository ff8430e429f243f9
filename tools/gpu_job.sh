#!/bin/bash
# Runs on the GPU box via gpurun; everything of interest goes to gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload mid --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_mid.log 2>&1; echo "bench mid rc=$?"
tail -c 1500 gpurun_out/bench_mid.log
if [ "$1" == "full" ]; then
timeout 900 python bench.py --workload cfg2 --steps 2 --warmup 1 > gpurun_out/bench_cfg2.log 2>&1; echo "bench cfg2 rc=$?"
tail -c 2500 gpurun_out/bench_cfg2.log
fi
