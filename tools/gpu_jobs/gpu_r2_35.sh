#!/bin/bash
# HEAD records: the whole GPU suite, the default bench line, the reference arm (short), the batch workload
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
timeout -s KILL 300 python bench.py > gpurun_out/bench_cfg2_final.json 2> gpurun_out/bench_cfg2_final.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_cfg2_final.err; cut -c1-200 gpurun_out/bench_cfg2_final.json
timeout -s KILL 120 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_cfg2_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/bench_cfg2_reference_arm.json
timeout -s KILL 200 python bench.py --workload cfg4 --steps 2 --warmup 1 --no-index-steps 0 > gpurun_out/bench_cfg4_final.json 2> gpurun_out/bench_cfg4_final.err; echo "cfg4 rc=$?"; cut -c1-200 gpurun_out/bench_cfg4_final.json
