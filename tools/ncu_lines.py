#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from a full ncu capture (compiled with -lineinfo, captured with --import-source on):
python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N]"""
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}", "--launch-count", "1"]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
# (the Source column holds unescaped quotes, so the csv module cannot split these rows: CUDA-line rows look like
#  "<line>","<source text>","-","-","<all samples>","<not issued>","<# samples>","<inst executed>","<thread inst>",...)
fname, out = "", []
for line in raw.splitlines():
    if line.startswith('"File Path"'):
        fname = line.split('","')[1].rstrip('"').split("/")[-1]
        continue
    m = re.match(r'^"(\d+)","(.*?)","-","-","(.*)$', line)
    if not m:
        continue
    tail = m.group(3).split('","')
    if len(tail) < 8:
        continue
    out.append({"file": fname, "Line No": m.group(1), "Source": m.group(2), "# Samples": tail[2], "Instructions Executed": tail[3],
                "Thread Instructions Executed": tail[4], "Avg. Threads Executed": str(float(tail[4] or 0) / max(1.0, float(tail[3] or 0)))})
if not out:
    print("no rows", file=sys.stderr)
    sys.exit(1)


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


tot_i = sum(num(r["Instructions Executed"]) for r in out) or 1
tot_s = sum(num(r["# Samples"]) for r in out) or 1
print(f"total warp instructions {tot_i:.0f}, samples {tot_s:.0f}")
print("share_inst share_samp avg_thr  file:line  source")
for r in sorted(out, key=lambda r: -num(r["# Samples"]))[:top]:
    print(f"{num(r['Instructions Executed']) / tot_i:9.3f} {num(r['# Samples']) / tot_s:9.3f} {num(r['Avg. Threads Executed']):7.1f}  {r['file']}:{r['Line No']:>4}  {r['Source'].strip()[:130]}")
