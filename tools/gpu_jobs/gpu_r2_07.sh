#!/bin/bash
mkdir -p gpurun_out
FB_PK_DEBUG=1 timeout -s KILL 120 python tools/dbg_pk.py 4096 256 1 > gpurun_out/dbg_pk.txt 2>&1; grep -v "^\[pk\]" gpurun_out/dbg_pk.txt | cut -c1-400; grep "^\[pk\]" gpurun_out/dbg_pk.txt | head -30
sed -i 's/^iters = .*/iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5/' tools/chain_synth.py
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_pk_hsq -s 33 -c 3 -f -o gpurun_out/r02_pk_hsq_v2 python tools/chain_synth.py 4096 4096 3 1 1 > gpurun_out/ncu_pk2.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_pk2.log
