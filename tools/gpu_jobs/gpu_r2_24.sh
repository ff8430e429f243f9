#!/bin/bash
mkdir -p gpurun_out
for cfg in "4 32 8" "6 32 8" "8 32 8" "4 64 16" "6 64 16" "8 64 16" "8 128 24"; do
  set -- $cfg
  FB_PK_VDEPTH=$1 FB_PK_VMAXSEG=$2 FB_PK_VWARPS=$3 timeout -s KILL 120 python tools/chain_synth.py 4096 4096 3 3 1 > gpurun_out/chain_v_$1_$2_$3.txt 2>&1; echo "VDEPTH=$1 VMAXSEG=$2 VWARPS=$3"; grep -E "k_pk_vsq|chain_ms" gpurun_out/chain_v_$1_$2_$3.txt | cut -c1-250 | tail -3
done
