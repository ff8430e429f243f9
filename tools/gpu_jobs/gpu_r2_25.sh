#!/bin/bash
# 8 GPUs: BASELINE config 4 (batch of 64 1080p images): weak scaling (64 images on every GPU), strong scaling (the 64 split 8 ways), reference arm
mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout -s KILL 900 $R --master-port 29541 bench.py --gpus 8 --workload cfg4 --steps 2 --warmup 1 --no-index-steps 1 > gpurun_out/bench_cfg4_n8_weak.json 2> gpurun_out/bench_cfg4_n8_weak.err; echo "cfg4 weak n8 rc=$?"; tail -c 300 gpurun_out/bench_cfg4_n8_weak.err; tail -1 gpurun_out/bench_cfg4_n8_weak.json | cut -c1-700
timeout -s KILL 600 $R --master-port 29542 bench.py --gpus 8 --workload cfg4 --strong --steps 2 --warmup 1 --no-index-steps 0 > gpurun_out/bench_cfg4_n8_strong.json 2> gpurun_out/bench_cfg4_n8_strong.err; echo "cfg4 strong n8 rc=$?"; tail -1 gpurun_out/bench_cfg4_n8_strong.json | cut -c1-400
timeout -s KILL 600 $R --master-port 29543 bench.py --impl reference --gpus 8 --workload cfg4 --strong --steps 1 --warmup 0 > gpurun_out/bench_cfg4_n8_strong_ref.json 2> gpurun_out/bench_cfg4_n8_strong_ref.err; echo "cfg4 strong ref rc=$?"; tail -1 gpurun_out/bench_cfg4_n8_strong_ref.json | cut -c1-400
timeout -s KILL 900 $R --master-port 29544 bench.py --impl reference --gpus 8 --workload cfg4 --steps 1 --warmup 0 > gpurun_out/bench_cfg4_n8_weak_ref.json 2> gpurun_out/bench_cfg4_n8_weak_ref.err; echo "cfg4 weak ref rc=$?"; tail -1 gpurun_out/bench_cfg4_n8_weak_ref.json | cut -c1-400
