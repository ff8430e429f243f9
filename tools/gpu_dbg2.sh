#!/bin/bash
mkdir -p gpurun_out
FB_MANIAC_DEBUG=1 timeout 200 python tools/decode_once.py cfg2 > gpurun_out/dbg_cfg2.log 2>&1; echo rc=$?
grep "cycles/symbol" gpurun_out/dbg_cfg2.log | sort -t: -k2 -n | tail -8
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
