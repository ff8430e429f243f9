#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 tools/ubench_tma > gpurun_out/r02_ubench_tma.txt 2>&1; echo "ubench_tma rc=$?"; cat gpurun_out/r02_ubench_tma.txt
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_pk_hsq -s 12 -c 3 -f -o gpurun_out/r02_pk_hsq_v1 python tools/chain_synth.py 4096 4096 3 0 1 > gpurun_out/ncu_pk1.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_pk1.log
