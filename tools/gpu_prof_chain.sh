#!/bin/bash
# Launch list of the bench command + one full capture of the biggest unsqueeze launch.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_cfg2.csv \
    python bench.py --workload cfg2 --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_bench_cfg2.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_inv_hsqueeze_tiled -s 9 -c 1 -f -o gpurun_out/hsq_cfg2 \
    python tools/decode_once.py cfg2 --undo > gpurun_out/ncu_hsq.log 2>&1
echo "hsq rc=$?"; tail -2 gpurun_out/ncu_hsq.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_inv_vsqueeze_tiled -s 9 -c 1 -f -o gpurun_out/vsq_cfg2 \
    python tools/decode_once.py cfg2 --undo > gpurun_out/ncu_vsq.log 2>&1
echo "vsq rc=$?"; tail -2 gpurun_out/ncu_vsq.log
ls -la gpurun_out | tail -8
