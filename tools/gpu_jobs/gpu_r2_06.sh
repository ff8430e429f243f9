#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_pk_squeeze.py -q > gpurun_out/pytest_pk.log 2>&1; echo "pytest pk rc=$?"; grep -E "^(FAILED|PASSED|ERROR)|passed|failed|AssertionError" gpurun_out/pytest_pk.log | cut -c1-250 | head -40
timeout -s KILL 120 python tools/chain_synth.py 4096 4096 3 5 1 > gpurun_out/chain_synth_4096_p1.txt 2>&1; echo "chain rc=$?"; cat gpurun_out/chain_synth_4096_p1.txt | cut -c1-330
timeout -s KILL 120 python tools/chain_synth.py 1920 1080 3 5 1 > gpurun_out/chain_synth_1080_p1.txt 2>&1; cat gpurun_out/chain_synth_1080_p1.txt | cut -c1-330
