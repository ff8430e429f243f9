#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_at_size.py tests/test_gpu_parity.py -q -x -k "dct or DCT or golden or responsive" > gpurun_out/pytest_dct.log 2>&1; echo "pytest dct rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|Error|differs" gpurun_out/pytest_dct.log | cut -c1-300 | head -20
timeout -s KILL 600 python bench.py --workload cfg3 --steps 3 --warmup 2 --no-index-steps 1 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "bench cfg3 rc=$?"; tail -c 500 gpurun_out/bench_cfg3.err; python tools/show_bench.py gpurun_out/bench_cfg3.json | cut -c1-1500
cp .bench_cache/*.index.json gpurun_out/ 2>/dev/null
