#!/bin/bash
mkdir -p gpurun_out
for spb in 0 4 8 16; do
  FB_MANIAC_SPB=$spb timeout -s KILL 600 python bench.py --workload cfg4 --steps 1 --warmup 1 --no-index-steps 0 --skip-cpu-baseline > gpurun_out/bench_cfg4_spb$spb.json 2> gpurun_out/bench_cfg4_spb$spb.err; echo "SPB=$spb rc=$?"; tail -c 200 gpurun_out/bench_cfg4_spb$spb.err; python tools/show_bench.py gpurun_out/bench_cfg4_spb$spb.json | cut -c1-260
done
timeout -s KILL 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_pk_hsq<.int.2" -s 2 -c 1 -f -o gpurun_out/r02_pk_hsq_ycocg_cfg2 python tools/chain_synth.py 4096 4096 3 1 1 > gpurun_out/ncu_pk_h.log 2>&1; echo "ncu h rc=$?"; ls -la gpurun_out/r02_pk_hsq_ycocg_cfg2.ncu-rep
