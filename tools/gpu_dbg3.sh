#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clk.csv &
SMI=$!
FB_MANIAC_DEBUG=1 timeout 200 python tools/decode_once.py cfg2 --reps=2 > gpurun_out/dbg_cfg2.log 2>&1; echo rc=$?
kill $SMI
grep "wall" gpurun_out/dbg_cfg2.log
grep "cycles/symbol" gpurun_out/dbg_cfg2.log | sed 's/\[maniac\]   //' | sort -k2 -n | tail -4
sort gpurun_out/clk.csv | uniq -c | sort -rn | head -5
