#!/bin/bash
# MANIAC decode kernel: source-level profile of the shipped kernel on the 2048^2 workload (indexed: the three largest groups dominate)
mkdir -p gpurun_out
timeout -s KILL 300 python tools/decode_once.py mid > gpurun_out/decode_mid.txt 2>&1; tail -2 gpurun_out/decode_mid.txt
FB_MANIAC_DEBUG=1 timeout -s KILL 300 python tools/decode_once.py mid > gpurun_out/r02_maniac_cycles_mid.log 2>&1; grep -iE "cyc|sym" gpurun_out/r02_maniac_cycles_mid.log | tail -15
timeout -s KILL 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section InstructionStats \
   --clock-control none --import-source on -k regex:k_maniac_decode -c 1 -f -o gpurun_out/r02_maniac_mid python tools/decode_once.py mid > gpurun_out/ncu_maniac.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_maniac.log
cp .bench_cache/mid_s7.index.json gpurun_out/ 2>/dev/null
