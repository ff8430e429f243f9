#!/bin/bash
# round 2, call 1: instruction issue rates (tools/ubench_pipes) + the GPU suite of HEAD as the round's baseline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout -s KILL 120 tools/ubench_pipes > gpurun_out/r02_ubench_pipes.txt 2>&1; echo "ubench rc=$?"; cat gpurun_out/r02_ubench_pipes.txt
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
