"""Times the host-threads entropy backend (fb_host_decode, no GPU) against the unmodified reference decoder on the same file.
usage: python tools/host_entropy_perf.py file.fuif [threads ...]"""
import json
import subprocess
import sys
import time

sys.path.insert(0, ".")
from fuif_b200 import api  # noqa: E402

path = sys.argv[1]
threads = [int(a) for a in sys.argv[2:]] or [0]
data = open(path, "rb").read()
info = api.peek_header(data) if hasattr(api, "peek_header") else None
t0 = time.perf_counter()
seq = api.fuif_host_decode(data, threads=1)
t_seq = time.perf_counter() - t0
gi = seq.group_index()
px = seq.info().w * seq.info().h
print(json.dumps({"file": path, "groups": len(gi[0]), "sequential_s": round(t_seq, 3), "Mpx/s": round(px / t_seq / 1e6, 2)}))
for t in threads:
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        par = api.fuif_host_decode(data, group_index=gi, threads=t)
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({"threads": t, "indexed_s": round(best, 3), "Mpx/s": round(px / best / 1e6, 2)}))
try:
    r = subprocess.run(["oracle/_ref/ref_driver", "time", path, "1"], capture_output=True, text=True, check=True)
    print("reference:", r.stdout.strip().splitlines()[-1])
except Exception as e:  # noqa: BLE001
    print("reference not available:", e)
