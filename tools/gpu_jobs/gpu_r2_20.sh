#!/bin/bash
# 2 GPUs: the N>1 path of bench.py (NCCL init, replicas, per-rank gather, un-indexed sample) on the small cfg1 workload, both arms
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 --workload cfg1 > gpurun_out/bench_cfg1_n2.json 2> gpurun_out/bench_cfg1_n2.err; echo "bench n2 rc=$?"; tail -c 400 gpurun_out/bench_cfg1_n2.err; tail -1 gpurun_out/bench_cfg1_n2.json | cut -c1-1200
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --workload cfg1 > gpurun_out/bench_cfg1_n2_ref.json 2> gpurun_out/bench_cfg1_n2_ref.err; echo "ref n2 rc=$?"; tail -1 gpurun_out/bench_cfg1_n2_ref.json | cut -c1-600
timeout -s KILL 300 python -m pytest tests/test_gpu_two_devices.py -q > gpurun_out/pytest_2dev.log 2>&1; echo "2dev rc=$?"; tail -3 gpurun_out/pytest_2dev.log
