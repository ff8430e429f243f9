"""Compact text summary of an ncu report for profiles/: per captured launch the launch shape, duration, instruction and
memory counters, pipe utilisation, warp-stall breakdown, and the SASS instructions with the most stall samples.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_ncu.txt        (needs `ncu` on PATH; no GPU)"""
import csv
import collections
import subprocess
import sys

WANT = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__cycles_active.avg",
]


def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1]] + args, capture_output=True, text=True).stdout


def main():
    raw = list(csv.reader(run(["--page", "raw", "--csv"]).splitlines()))
    hdr, units = raw[0], raw[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {sys.argv[1]}: ncu capture (numbers under the profiler are never bench values)")
    for r in raw[2:]:
        print()
        for w in WANT:
            if w in idx and r[idx[w]] != "":
                print(f"{w:72s} {r[idx[w]]} {units[idx[w]]}")
        st = [(h, r[idx[h]]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
        st = sorted(st, key=lambda x: -float(x[1].replace(",", "") or 0))[:10]
        print("warp stall cycles per issued instruction (top 10):")
        for h, v in st:
            print("   ", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
    src = list(csv.reader(run(["--page", "source", "--csv", "--print-source", "sass"]).splitlines()))
    kern, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kern.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    seen = set()
    for k in kern:
        if k["name"] in seen or not k["hdr"] or "Warp Stall Sampling (All Samples)" not in k["hdr"]:
            continue
        seen.add(k["name"])
        h = k["hdr"]
        si, ei = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
        tot = sum(int(r[si] or 0) for r in k["rows"])
        print(f"\nsource page, {k['name']}: {len(k['rows'])} SASS instructions, {tot} stall samples; the 30 instructions with the most samples:")
        top = sorted(enumerate(k["rows"]), key=lambda x: -int(x[1][si] or 0))[:30]
        for i, r in sorted(top):
            print(f"  [{i:5d}] {r[1].strip()[:90]:90s} samples {r[si]:>6s}  executed {r[ei]}")
        ops = collections.Counter()
        for r in k["rows"]:
            ops[r[1].strip().split()[0 if not r[1].strip().startswith("@") else 1].split(".")[0]] += int(r[ei] or 0)
        print("  executed warp instructions by opcode (top 16):", ", ".join(f"{o} {n}" for o, n in ops.most_common(16)))


if __name__ == "__main__":
    main()
