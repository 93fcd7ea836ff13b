#!/usr/bin/env python
"""Summarise gpurun_out/launches.csv (ncu launch list) and a full .ncu-rep into profiles/<tag>_*.{csv,md}."""
import csv
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

tag = sys.argv[1]
rep = sys.argv[2] if len(sys.argv) > 2 else None
out = Path("profiles")
out.mkdir(exist_ok=True)

rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 10 and r[0].isdigit()]
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    unit, val = r[-2], float(r[-1])
    us = val / 1000.0 if unit in ("ns", "nsecond") else val * (1000.0 if unit in ("ms", "msecond") else 1.0)
    agg[name][0] += 1
    agg[name][1] += us
total = sum(v[1] for v in agg.values()) or 1.0
lines = ["kernel,launches,total_us,share"]
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{name},{n},{us:.1f},{us / total:.4f}")
(out / f"{tag}_launches.csv").write_text("\n".join(lines) + "\n")
print("\n".join(lines[:14]))

if rep:
    # (a .csv argument is the raw page already exported on the GPU box: `ncu -i x.ncu-rep --page raw --csv`)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    data = list(csv.reader(raw.splitlines()))
    hdr, units = data[0], data[1]
    keep = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
            "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum",
            "smsp__inst_executed_op_shared_atom.sum"]
    keep += [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled") or h.startswith("smsp__average_warps_issue_stalled")]
    md = [f"# ncu --set full: {Path(rep).name}", ""]
    seen = defaultdict(int)
    for r in data[2:]:
        seen[r[hdr.index('Kernel Name')]] += 1
        if seen[r[hdr.index('Kernel Name')]] > 3:                 # (repeated launches of one kernel: the first three are kept)
            continue
        md.append(f"## {r[hdr.index('Kernel Name')][:120]}")
        md.append("")
        md.append("| metric | value | unit |")
        md.append("|---|---|---|")
        for k in keep:
            if k in hdr:
                i = hdr.index(k)
                md.append(f"| {k} | {r[i][:100]} | {units[i]} |")
        md.append("")
    (out / f"{tag}_full.md").write_text("\n".join(md))
    print((out / f"{tag}_full.md"))
