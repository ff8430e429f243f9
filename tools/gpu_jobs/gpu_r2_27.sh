#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout -s KILL 120 python tools/chain_synth.py 4096 4096 3 5 1 > gpurun_out/chain_synth_4096_p1.txt 2>&1; cat gpurun_out/chain_synth_4096_p1.txt | cut -c1-330
timeout -s KILL 120 python tools/chain_synth.py 1920 1080 3 5 1 > gpurun_out/chain_synth_1080_p1.txt 2>&1; head -1 gpurun_out/chain_synth_1080_p1.txt | cut -c1-330
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --no-index-steps 0 --skip-cpu-baseline > gpurun_out/bench_cfg2_quick.json 2> gpurun_out/bench_cfg2_quick.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_cfg2_quick.err; python tools/show_bench.py gpurun_out/bench_cfg2_quick.json | cut -c1-1400
