#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_chain.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_chain.log
timeout 300 python tools/chain_bench.py cfg2 5 2>&1 | tail -2
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_inv|k_ycocg|k_clamp" -c 30 --csv --log-file gpurun_out/launches_chain_mid.csv python tools/decode_once.py mid --undo > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_chain_mid.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index("Kernel Name"); iv=hdr.index("Metric Value"); ig=hdr.index("Grid Size")
tot=0
for r in rows[1:]:
    v=float(r[iv].replace(',','')); tot+=v
    print(r[ik].split('::')[-1][:28], round(v/1000,1), r[ig])
print("sum us", tot/1000)
PY
