#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pk_squeeze.py -q -x > gpurun_out/pytest_par.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|Error|differs" gpurun_out/pytest_par.log | cut -c1-300 | head -10
timeout -s KILL 120 python tools/chain_synth.py 4096 4096 3 5 1 > gpurun_out/chain_synth_4096_p1.txt 2>&1; echo "chain rc=$?"; cat gpurun_out/chain_synth_4096_p1.txt | cut -c1-330
timeout -s KILL 120 python tools/chain_synth.py 1920 1080 3 5 1 > gpurun_out/chain_synth_1080_p1.txt 2>&1; cat gpurun_out/chain_synth_1080_p1.txt | cut -c1-330
