#!/bin/bash
# Session-3 job B: run-ahead walkers: quick decode checks, parity tests, cycles/symbol, cfg2 wall time.
mkdir -p gpurun_out
timeout -s KILL 90 python tools/decode_once.py cfg1 --undo > gpurun_out/s3b_cfg1.log 2>&1; echo "cfg1 rc=$?"; tail -2 gpurun_out/s3b_cfg1.log
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
FB_MANIAC_DEBUG=1 timeout -s KILL 120 python tools/decode_once.py mid > gpurun_out/dbg_mid_ahead.log 2>&1; echo "dbg rc=$?"; grep -E "cycles/symbol|wall" gpurun_out/dbg_mid_ahead.log | tail -12
timeout -s KILL 200 python tools/decode_once.py cfg2 --reps=2 > gpurun_out/s3b_cfg2.log 2>&1; echo "cfg2 rc=$?"; tail -3 gpurun_out/s3b_cfg2.log
