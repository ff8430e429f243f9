#!/bin/bash
# Full round-end style run: tests, smoke, both bench arms on cfg2, extra workloads.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench cfg2 rc=$?"; tail -c 600 gpurun_out/bench_cfg2.err; python tools/show_bench.py gpurun_out/bench_cfg2.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_cfg2_ref.json 2> gpurun_out/bench_cfg2_ref.err; echo "ref arm rc=$?"; python tools/show_bench.py gpurun_out/bench_cfg2_ref.json
if [ "$1" == "extra" ]; then
timeout 900 python bench.py --workload cfg4 --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "bench cfg4 rc=$?"; tail -c 400 gpurun_out/bench_cfg4.err; python tools/show_bench.py gpurun_out/bench_cfg4.json
timeout 300 python bench.py --workload cfg4 --impl reference --steps 1 --warmup 1 > gpurun_out/bench_cfg4_ref.json 2> gpurun_out/bench_cfg4_ref.err; echo "ref cfg4 rc=$?"; python tools/show_bench.py gpurun_out/bench_cfg4_ref.json
timeout 1200 python bench.py --workload cfg3 --steps 2 --warmup 1 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "bench cfg3 rc=$?"; tail -c 400 gpurun_out/bench_cfg3.err; python tools/show_bench.py gpurun_out/bench_cfg3.json
fi
