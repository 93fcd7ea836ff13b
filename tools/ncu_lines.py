#!/usr/bin/env python
"""Top source lines by warp-stall samples: tools/ncu_lines.py <rep> <kernel regex> [n]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
hdr, agg, launches = None, {}, 0
for x in csv.reader(raw.splitlines()):
    if x and x[0] == "Line No":
        hdr = x
        launches += 1
        continue
    if hdr and len(x) > 10 and x[0].isdigit():
        key = (x[0], x[1][:130])
        smp = int(x[4]) if x[4].isdigit() else 0
        ins = int(x[7]) if x[7].isdigit() else 0
        a = agg.setdefault(key, [0, 0])
        a[0] += smp
        a[1] += ins
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print(f"{launches} launches, {tot} samples, {toti} warp instructions")
for (ln, src), (smp, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{100*smp/tot:5.1f}%  inst {100*ins/toti:5.1f}%  L{ln}: {src}")
