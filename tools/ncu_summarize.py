#!/usr/bin/env python
"""Summarise an ncu capture (raw CSV page) and a launch list into profiles/.
usage: tools/ncu_summarize.py <tag> <raw.csv> <launches.csv> [kernel-substring]"""
import collections
import csv
import json
import os
import sys

tag, raw, launches = sys.argv[1:4]
want = sys.argv[4] if len(sys.argv) > 4 else "k_env_step32"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
caps = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if want not in d.get("Kernel Name", ""):
        continue
    m = {"kernel": d["Kernel Name"], "grid": d.get("Grid Size"), "block": d.get("Block Size")}
    for k in KEYS:
        if k in d:
            try:
                m[k] = float(d[k].replace(",", ""))
            except ValueError:
                m[k] = d[k]
            m[k + "#unit"] = units[hdr.index(k)]
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
            m[h] = float(d[h])
    caps.append(m)

agg = collections.defaultdict(list)
lr = list(csv.reader(open(launches)))
lh = None
for r in lr:
    if r and r[0] == "ID":
        lh = r
        continue
    if lh and len(r) == len(lh):
        d = dict(zip(lh, r))
        agg[d["Kernel Name"]].append(float(d["Metric Value"].replace(",", "")))
tot = sum(sum(v) for v in agg.values()) or 1.0
launch_tab = [{"kernel": k, "launches": len(v), "total_us": sum(v) / 1e3, "avg_us": sum(v) / len(v) / 1e3,
               "share_pct": 100 * sum(v) / tot} for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))]

c = caps[0]
unit = lambda k: c.get(k + "#unit", "")
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
dram = (c["dram__bytes_read.sum"] * scale.get(unit("dram__bytes_read.sum"), 1.0) +
        c["dram__bytes_write.sum"] * scale.get(unit("dram__bytes_write.sum"), 1.0))
summary = {
    "tag": tag, "kernel": c["kernel"], "grid": c["grid"], "block": c["block"],
    "duration_us": c["gpu__time_duration.sum"], "dram_bytes_per_launch": dram,
    "registers_per_thread": c["launch__registers_per_thread"],
    "warp_inst_per_launch": c["smsp__inst_executed.sum"],
    "pipes": {
        "issue_active_pct": c["smsp__issue_active.avg.pct_of_peak_sustained_active"],
        "alu_pct": c["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"],
        "fma_pct": c["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"],
        "fp64_pct": c["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"],
        "xu_pct": c["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"],
        "lsu_pct": c["sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"],
        "dram_pct": c["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"],
        "warps_active_pct": c["sm__warps_active.avg.pct_of_peak_sustained_active"],
    },
    "stalls_per_issue": {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): v
                         for k, v in c.items() if k.startswith("smsp__average_warps_issue_stalled")},
    "launch_list": launch_tab,
}
json.dump(summary, open(os.path.join(OUT, f"{tag}_ncu_summary.json"), "w"), indent=1)
json.dump(summary, open(os.path.join(OUT, "ncu_summary.json"), "w"), indent=1)  # the latest one; bench.py reads it
with open(os.path.join(OUT, f"{tag}_ncu_summary.md"), "w") as f:
    f.write(f"# ncu summary {tag}\n\nKernel `{c['kernel']}` grid {c['grid']} block {c['block']}\n\n")
    f.write("| metric | value |\n|---|---|\n")
    for k in KEYS:
        if k in c:
            f.write(f"| {k} | {c[k]} {unit(k)} |\n")
    f.write(f"| dram bytes per launch (read+write) | {dram:.4g} B |\n\n")
    f.write("Stall reasons (warps stalled per issue):\n\n| reason | ratio |\n|---|---|\n")
    for k, v in sorted(summary["stalls_per_issue"].items(), key=lambda kv: -kv[1]):
        f.write(f"| {k} | {v:.3f} |\n")
    f.write("\nLaunch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache serialised):\n\n")
    f.write("| kernel | launches | total us | avg us | share % |\n|---|---|---|---|---|\n")
    for r in launch_tab:
        f.write(f"| {r['kernel'][:80]} | {r['launches']} | {r['total_us']:.1f} | {r['avg_us']:.1f} | {r['share_pct']:.1f} |\n")
print(json.dumps(summary["pipes"]), summary["duration_us"], dram)
