#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_gpu_at_size.py tests/test_gpu_corrupt.py tests/test_gpu_pk_squeeze.py -q -x > gpurun_out/pytest_new.log 2>&1; echo "pytest new rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|Error|differs" gpurun_out/pytest_new.log | cut -c1-300 | head -30
