#!/bin/bash
for r in 1 2 3; do MC2_DEBUG_HASH=1 python -m pytest tests/test_gpu_parity.py -q -s -k "test_synthetic_reads_vs_oracle" -p no:cacheprovider > gpurun_out/dbg_hash_$r.log 2>&1; echo "run $r: $(grep -c "\[hash\] surv" gpurun_out/dbg_hash_$r.log) checks, $(grep -c MISMATCH gpurun_out/dbg_hash_$r.log) mismatches; $(tail -1 gpurun_out/dbg_hash_$r.log)"; done
python - <<'PY'
import re,glob
bad=0; tot=0
for f in glob.glob('gpurun_out/dbg_hash_[0-9].log'):
    for line in open(f, errors='replace'):
        m = re.search(r'STRESS rep (\d+): pass2 hits (\d+), counts read back (\d+), claims (\d+) \(first (\d+)\), empty-key slots (\d+), survivors (\d+) \(first (\d+)\)', line)
        if m:
            rep,h,rb,cl,fcl,es,sv,fsv = map(int, m.groups()); tot+=1
            if h != rb or sv != fsv or es: bad+=1; print(line.strip()[:170]) if bad < 6 else None
print("stress lines", tot, "inconsistent", bad)
PY
