export MC2_PROFILE_PASSES=1
python tools/profile_paths.py > gpurun_out/paths_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/paths_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source off -k regex:"fn_dense|dense_|mt_|tsv_|chunk_|rle_|wide_|mg_|fq_|rs_scatter|rs_hist|parse_emit|parse_classify" -c 70 -f -o gpurun_out/r02_other python tools/profile_paths.py > gpurun_out/ncu_other.log 2>&1
echo "capture exit $?"
ncu -i gpurun_out/r02_other.ncu-rep --page raw --csv > gpurun_out/r02_other_raw.csv 2>/dev/null
ls -la gpurun_out/r02_other*
if [ $(stat -c%s gpurun_out/r02_other.ncu-rep) -gt 30000000 ]; then rm gpurun_out/r02_other.ncu-rep; fi
du -sh gpurun_out
