import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("no json:", e); sys.exit(0)
keys = {k: d.get(k) for k in ("impl", "value", "ms_per_step", "gpu_launches")}
keys["e2e"] = d.get("e2e", {}).get("value")
keys["workload"] = d.get("config", {}).get("workload", "")[:40]
if d.get("stages"):
    keys["stages"] = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["stages"].items() if k != "kernels"}
    if d["stages"].get("kernels"):
        keys["kernels"] = {k: [round(v["launches_per_step"], 1), round(v["mean_us"], 1)] for k, v in d["stages"]["kernels"].items()}
if d.get("roofline"): keys["roofline"] = {k: d["roofline"].get(k) for k in ("achieved", "frac", "ms", "us", "share_of_chain", "chain")}
if d.get("cpu_baseline"): keys["cpu"] = {k: d["cpu_baseline"].get(k) for k in ("value", "kind", "cores", "entropy_s", "chain_s")}
keys["clocks"] = d.get("clocks")
print(json.dumps(keys))
