#!/bin/bash
# ncu launch list + one full capture of the named kernels (each only after the plain command exits 0)
# usage: tools/gpu_ncu.sh '<kernel regex>' <launches to skip> [bench args...]
mkdir -p gpurun_out
KERNEL=${1:-rc_count}
SKIP=${2:-8}
shift; shift
CMD="python bench.py --reads-per-gpu ${MC2_NCU_READS:-6000000} --genome-scale ${MC2_GENOME_SCALE:-0.3} --steps 1 --warmup 1 --no-cpu --no-e2e --no-check --no-secondary $*"
TAG=${MC2_NCU_TAG:-prof}
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$KERNEL" -s $SKIP -c ${MC2_NCU_COUNT:-6} -f -o gpurun_out/$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
tail -2 gpurun_out/plain_$TAG.log | cut -c1-1500
ls -la gpurun_out/ | tail -5
