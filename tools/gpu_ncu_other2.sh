# second pass over tools/profile_paths.py: the kernels the first 70-launch capture did not reach
export MC2_PROFILE_PASSES=1
timeout 600 ncu --set full --clock-control none --import-source off -k regex:"mt_|mg_|fq_|rle_|chunk_candidates|chunk_select|dense_smem|dense_global|seg_" -c 32 -f -o gpurun_out/r02_other2 python tools/profile_paths.py > gpurun_out/ncu_other2.log 2>&1
echo "capture exit $?"
ncu -i gpurun_out/r02_other2.ncu-rep --page raw --csv > gpurun_out/r02_other2_raw.csv 2>/dev/null
ls -la gpurun_out/r02_other2*
rm -f gpurun_out/r02_other2.ncu-rep
