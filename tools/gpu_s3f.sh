#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" FB_MANIAC_DEBUG=1 timeout -s KILL 120 python tools/decode_once.py mid > gpurun_out/dbg_mid_w.log 2>&1; grep -E "ch (45|48|51|53|54) .*cycles/symbol" gpurun_out/dbg_mid_w.log; grep wall gpurun_out/dbg_mid_w.log; }
run FB_MANIAC_WUSED=8
run FB_MANIAC_WUSED=4
timeout -s KILL 300 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats \
   --import-source on --clock-control none -k regex:k_maniac_decode -c 1 -f -o gpurun_out/maniac_ahead_mid python tools/decode_once.py mid > gpurun_out/ncu_maniac.log 2>&1; echo "ncu maniac rc=$?"; tail -2 gpurun_out/ncu_maniac.log
