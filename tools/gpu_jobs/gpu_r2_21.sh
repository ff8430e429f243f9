#!/bin/bash
# 1 GPU: the batch configuration (cfg4: 64 x 1080p) and the big 14-bit configuration (cfg5), ours and the reference arm
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --workload cfg4 --steps 3 --warmup 2 > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err; echo "cfg4 n1 rc=$?"; tail -c 300 gpurun_out/bench_cfg4_n1.err; python tools/show_bench.py gpurun_out/bench_cfg4_n1.json | cut -c1-1200
timeout -s KILL 600 python bench.py --impl reference --workload cfg4 --steps 2 --warmup 1 > gpurun_out/bench_cfg4_n1_ref.json 2> gpurun_out/bench_cfg4_n1_ref.err; echo "cfg4 ref rc=$?"; tail -1 gpurun_out/bench_cfg4_n1_ref.json | cut -c1-500
nproc
timeout -s KILL 1500 python bench.py --workload cfg5 --steps 1 --warmup 1 --no-index-steps 0 > gpurun_out/bench_cfg5_n1.json 2> gpurun_out/bench_cfg5_n1.err; echo "cfg5 n1 rc=$?"; tail -c 600 gpurun_out/bench_cfg5_n1.err; python tools/show_bench.py gpurun_out/bench_cfg5_n1.json | cut -c1-1500
