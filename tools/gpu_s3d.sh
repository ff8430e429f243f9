#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats \
   --import-source on --clock-control none -k regex:k_maniac_decode -c 1 -f -o gpurun_out/maniac_ahead_cfg1 python tools/decode_once.py cfg1 > gpurun_out/ncu_maniac.log 2>&1; echo "ncu maniac rc=$?"; tail -2 gpurun_out/ncu_maniac.log
