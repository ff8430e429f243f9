#!/bin/bash
# final profile pass of round 2: full GPU suite, launch list of the bench command, ncu --set full of the dominant chain kernels, bench lines
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_cfg2.err; python tools/show_bench.py gpurun_out/bench_cfg2.json | cut -c1-1400
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_cfg2_ref.json 2> gpurun_out/bench_cfg2_ref.err; echo "ref rc=$?"; tail -1 gpurun_out/bench_cfg2_ref.json | cut -c1-300
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_cfg2.csv python bench.py --steps 1 --warmup 1 --no-index-steps 0 --skip-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r02_launches_bench_cfg2.csv
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k "regex:k_pk_hsq<.int.2" -s 2 -c 1 -f -o gpurun_out/r02_pk_hsq_ycocg_cfg2 python tools/chain_synth.py 4096 4096 3 1 1 > gpurun_out/ncu_pk_h.log 2>&1; echo "ncu h rc=$?"; grep -E "ycocg" gpurun_out/ncu_pk_h.log | tail -1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_pk_vsq -s 11 -c 2 -f -o gpurun_out/r02_pk_vsq_cfg2 python tools/chain_synth.py 4096 4096 3 1 1 > gpurun_out/ncu_pk_v.log 2>&1; echo "ncu v rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_idct_ycbcr -s 1 -c 1 -f -o gpurun_out/r02_idct_ycbcr_cfg3 python bench.py --workload cfg3 --steps 1 --warmup 2 --no-index-steps 0 --skip-cpu-baseline > gpurun_out/ncu_idct.log 2>&1; echo "ncu idct rc=$?"; tail -2 gpurun_out/ncu_idct.log | cut -c1-300
