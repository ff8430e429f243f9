#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
FB_MANIAC_DEBUG=1 timeout -s KILL 100 python tools/decode_once.py cfg2 > gpurun_out/dbg_cfg2.log 2>&1; echo "rc=$?"; grep -E "ch (48|54|59|60) .*cycles/symbol" gpurun_out/dbg_cfg2.log; grep wall gpurun_out/dbg_cfg2.log
