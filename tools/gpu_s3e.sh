#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 90 python tools/decode_once.py cfg1 --undo > gpurun_out/s3_cfg1.log 2>&1; echo "cfg1 rc=$?"; tail -2 gpurun_out/s3_cfg1.log
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
FB_MANIAC_DEBUG=1 timeout -s KILL 120 python tools/decode_once.py mid > gpurun_out/dbg_mid.log 2>&1; echo "rc=$?"; grep -E "ch (30|36|42|45|47|48|51|53|54) .*cycles/symbol" gpurun_out/dbg_mid.log; grep wall gpurun_out/dbg_mid.log
