#!/bin/bash
# Round-2 GPU job A (first call of the round): everything written after the round-1 GPU budget ran out gets its first
# hardware run here -- fb_encode (kernel + host), the trimmed YCoCg epilogue of k_inv_hsq_direct -- then the bench and a fresh
# ncu capture of the dominant chain kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout -s KILL 900 python -m pytest tests -m gpu -q -rxX > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
# the kernels that have never run on hardware, once under memcheck (small cases; XPASS = clean and correct)
timeout -s KILL 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_subsample.py tests/test_zz_gpu_approximate.py \
    tests/test_zz_gpu_palette.py tests/test_zz_gpu_match.py tests/test_zz_gpu_permute.py -q -rxX > gpurun_out/memcheck_new_kernels.log 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/memcheck_new_kernels.log
timeout -s KILL 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/gpu_encode_child.py odd tiny unc > gpurun_out/memcheck_encode.log 2>&1; echo "memcheck encode rc=$?"; tail -8 gpurun_out/memcheck_encode.log
timeout -s KILL 300 python tests/gpu_encode_child.py odd dct noise synth512 > gpurun_out/encode_child.log 2>&1; echo "encode child rc=$?"; tail -6 gpurun_out/encode_child.log
timeout -s KILL 600 python tools/encode_bench.py 256 512 1024 2048 > gpurun_out/encode_bench.jsonl 2> gpurun_out/encode_bench.err; echo "encode bench rc=$?"; cat gpurun_out/encode_bench.jsonl; tail -c 400 gpurun_out/encode_bench.err
timeout -s KILL 600 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench cfg2 rc=$?"; tail -c 300 gpurun_out/bench_cfg2.err; python tools/show_bench.py gpurun_out/bench_cfg2.json
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:k_inv_hsq_direct -s 6 -c 1 -f -o gpurun_out/r02_hsq_ycocg_cfg2 \
    python tools/chain_once.py 4096 4096 3 1 > gpurun_out/ncu_h.log 2>&1; echo "ncu h rc=$?"; tail -1 gpurun_out/ncu_h.log
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:k_maniac_encode -c 1 -f -o gpurun_out/r02_maniac_encode_512 \
    python tools/encode_bench.py 512 > gpurun_out/ncu_enc.log 2>&1; echo "ncu enc rc=$?"; tail -1 gpurun_out/ncu_enc.log
ls -la gpurun_out | tail -8
