#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_batch_many.py tests/test_gpu_parity.py -q -x -k "batch or many" > gpurun_out/pytest_batch.log 2>&1; echo "pytest batch rc=$?"; tail -3 gpurun_out/pytest_batch.log
timeout -s KILL 900 python bench.py --workload cfg4 --steps 3 --warmup 2 > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err; echo "cfg4 n1 rc=$?"; tail -c 300 gpurun_out/bench_cfg4_n1.err; python tools/show_bench.py gpurun_out/bench_cfg4_n1.json | cut -c1-700
