#!/bin/bash
# ncu launch list + one full capture of the named kernel (run only after the plain command exits 0)
mkdir -p gpurun_out
KERNEL=${1:-rs_scatter}
SKIP=${2:-8}
CMD="python bench.py --reads-per-gpu 1400000 --genome-scale ${MC2_GENOME_SCALE:-1.0} --steps 1 --warmup 1 --no-cpu --no-e2e"
NAMES='regex:parse_|rs_|scan_|rle_|extract_|chunk_|dense_|gather_|iota_|wide_|seg_|mt_|symbol_|hc_|fn_|fill_'
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$NAMES" -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s $SKIP -c ${MC2_NCU_COUNT:-3} -f -o gpurun_out/prof_$KERNEL $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/plain.log
ls -la gpurun_out/
