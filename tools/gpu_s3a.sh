#!/bin/bash
# Session-3 job A: MANIAC cycles/symbol (walkers on/off), ncu of the MANIAC kernel, ncu --set full of the big unsqueeze launches.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
FB_MANIAC_DEBUG=1 timeout 120 python tools/decode_once.py mid > gpurun_out/dbg_mid_walk.log 2>&1; echo "dbg walk rc=$?"
FB_MANIAC_DEBUG=1 FB_MANIAC_NO_WALKERS=1 timeout 120 python tools/decode_once.py mid > gpurun_out/dbg_mid_nowalk.log 2>&1; echo "dbg nowalk rc=$?"
timeout 300 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section InstructionStats \
   --import-source on --clock-control none -k regex:k_maniac_decode -c 1 -f -o gpurun_out/maniac_walk_cfg1 python tools/decode_once.py cfg1 > gpurun_out/ncu_maniac.log 2>&1; echo "ncu maniac rc=$?"
FB_SQUEEZE_MODE=direct timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_inv_.sq_direct -s 10 -c 4 -f -o gpurun_out/direct_big4 \
   python tools/chain_once.py 4096 4096 3 1 > gpurun_out/ncu_direct_big4.log 2>&1; echo "ncu direct rc=$?"
FB_SQUEEZE_MODE=perlevel timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_inv_.squeeze_tiled -s 5 -c 5 -f -o gpurun_out/tiled_big \
   python tools/chain_once.py 4096 4096 3 1 > gpurun_out/ncu_tiled_big.log 2>&1; echo "ncu tiled rc=$?"
ls -la gpurun_out | tail -8
