#!/bin/bash
# round 2, call 2: first hardware run of the packed unsqueeze kernels (parity, then chain timings packed vs 32-bit)
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_pk_squeeze.py -x -q > gpurun_out/pytest_pk.log 2>&1; echo "pytest pk rc=$?"; tail -15 gpurun_out/pytest_pk.log
for p in 1 0; do
  timeout -s KILL 120 python tools/chain_synth.py 4096 4096 3 5 $p > gpurun_out/chain_synth_4096_p$p.txt 2>&1; echo "chain p=$p rc=$?"; cat gpurun_out/chain_synth_4096_p$p.txt | cut -c1-220
done
timeout -s KILL 120 python tools/chain_synth.py 1920 1080 3 5 1 > gpurun_out/chain_synth_1080_p1.txt 2>&1; cat gpurun_out/chain_synth_1080_p1.txt | cut -c1-200
