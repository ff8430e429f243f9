#!/bin/bash
mkdir -p gpurun_out
for vw in 8 12 16 24; do
  FB_PK_VWARPS=$vw timeout -s KILL 120 python tools/chain_synth.py 4096 4096 3 3 1 > gpurun_out/chain_vw$vw.txt 2>&1; echo "VWARPS=$vw"; grep -E "k_pk_vsq|chain_ms" gpurun_out/chain_vw$vw.txt | cut -c1-200 | tail -4
done
bash tools/gpu_r2_22.sh
