#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_pk_squeeze.py -q > gpurun_out/pytest_pk.log 2>&1; echo "pytest pk rc=$?"; grep -E "^(FAILED|PASSED|ERROR)|passed|failed|AssertionError" gpurun_out/pytest_pk.log | cut -c1-250 | head -40
FB_PK_DEBUG=1 timeout -s KILL 120 python tools/chain_synth.py 4096 4096 3 1 1 > gpurun_out/chain_dbg.txt 2>&1; grep -n "pk\]" gpurun_out/chain_dbg.txt | head -60
