#!/bin/bash
# Round-1 GPU job C: parity tests, bench (both arms) on cfg2, launch list, full ncu captures of the dominant chain kernels, cfg4 batch.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench cfg2 rc=$?"; tail -c 300 gpurun_out/bench_cfg2.err; python tools/show_bench.py gpurun_out/bench_cfg2.json
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_cfg2_ref.json 2> gpurun_out/bench_cfg2_ref.err; echo "ref arm rc=$?"; cut -c1-300 gpurun_out/bench_cfg2_ref.json
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_cfg2.csv \
    python bench.py --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches_bench_cfg2.csv
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:k_inv_hsq_direct -s 6 -c 1 -f -o gpurun_out/hsq_ycocg_cfg2 \
    python tools/chain_once.py 4096 4096 3 1 > gpurun_out/ncu_h.log 2>&1; echo "ncu h rc=$?"; tail -1 gpurun_out/ncu_h.log
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:k_inv_vsqueeze_tiled -c 7 -f -o gpurun_out/vsq_tiled_cfg2 \
    python tools/chain_once.py 4096 4096 3 1 > gpurun_out/ncu_v.log 2>&1; echo "ncu v rc=$?"; tail -1 gpurun_out/ncu_v.log
timeout -s KILL 400 python bench.py --workload cfg4 --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "bench cfg4 rc=$?"; tail -c 300 gpurun_out/bench_cfg4.err; python tools/show_bench.py gpurun_out/bench_cfg4.json
ls -la gpurun_out | tail -8
