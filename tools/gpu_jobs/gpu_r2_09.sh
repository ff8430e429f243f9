#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_pk_squeeze.py -q > gpurun_out/pytest_pk.log 2>&1; echo "pytest pk rc=$?"; grep -E "^(FAILED|PASSED|ERROR)|passed|failed|AssertionError" gpurun_out/pytest_pk.log | cut -c1-250 | head -40
timeout -s KILL 120 python tools/chain_synth.py 4096 4096 3 5 1 > gpurun_out/chain_synth_4096_p1.txt 2>&1; echo "chain rc=$?"; cat gpurun_out/chain_synth_4096_p1.txt | cut -c1-330
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_pk_hsq -s 33 -c 3 -f -o gpurun_out/r02_pk_hsq_v3 python tools/chain_synth.py 4096 4096 3 1 1 > gpurun_out/ncu_pk3.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_pk3.log
