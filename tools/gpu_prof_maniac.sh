#!/bin/bash
mkdir -p gpurun_out
WL=${1:-cfg1}
timeout 60 python tools/decode_once.py $WL --reps=2 > gpurun_out/decode_once_$WL.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/decode_once_$WL.log
timeout 400 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight \
   --import-source on --clock-control none -k regex:k_maniac_decode -c 1 -f -o gpurun_out/maniac_$WL python tools/decode_once.py $WL > gpurun_out/ncu_maniac_$WL.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_maniac_$WL.log; ls -la gpurun_out/ | tail -5
