#!/bin/bash
# A/B of engine options on the quick bench: tools/gpu_ab.sh "opt=val,.." "opt=val,.." ...
mkdir -p gpurun_out
: > gpurun_out/ab.log
for o in "$@"; do
  echo "=== $o" >> gpurun_out/ab.log
  MERCAT2_B200_OPTIONS="$o" timeout 300 python bench.py --reads-per-gpu ${MC2_READS:-6000000} --steps 3 --warmup 2 --no-cpu --no-e2e ${MC2_BENCH_ARGS:-} 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('value %.2f G/s  ms/step %.2f  dev ms %.2f' % (d['value']/1e9, d['ms_per_step'], d['device_ms_per_step']))
    print('  '+'  '.join('%s %.2f' % (k.replace('_kernel',''),v['ms']/d['steps']) for k,v in d['kernels'].items()))
except Exception as ex: print('failed', ex)
" >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
