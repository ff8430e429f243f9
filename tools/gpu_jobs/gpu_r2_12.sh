#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/pytest_gpu.log
# launches: per undo 7 H (the 7th = colour kernel) + 7 V; chain_synth with iters=1 runs 4 undos; profile the 3rd one
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_pk_hsq -s 20 -c 1 -f -o gpurun_out/r02_pk_hsq_ycocg_cfg2 python tools/chain_synth.py 4096 4096 3 1 1 > gpurun_out/ncu_pk_h.log 2>&1; echo "ncu h rc=$?"; grep -E "k_pk_hsq|Report" gpurun_out/ncu_pk_h.log | tail -3
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_pk_vsq -s 19 -c 2 -f -o gpurun_out/r02_pk_vsq_cfg2 python tools/chain_synth.py 4096 4096 3 1 1 > gpurun_out/ncu_pk_v.log 2>&1; echo "ncu v rc=$?"
timeout -s KILL 600 python tools/encode_bench.py 1024 2048 > gpurun_out/encode_bench.jsonl 2> gpurun_out/encode_bench.err; echo "encode bench rc=$?"; cat gpurun_out/encode_bench.jsonl; tail -c 300 gpurun_out/encode_bench.err
timeout -s KILL 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_cfg2.err; python tools/show_bench.py gpurun_out/bench_cfg2.json | head -60
cp .bench_cache/*.index.json gpurun_out/ 2>/dev/null
