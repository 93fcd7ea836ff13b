#!/bin/bash
# Round-end style run: GPU tests, the default bench (both arms), then ncu launch list + full capture.
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench full"; timeout 900 python bench.py > gpurun_out/bench_full.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_full.log | cut -c1-1500
echo "== bench reference"; timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-600
MC2_NCU_COUNT=${MC2_NCU_COUNT:-6} bash tools/gpu_ncu.sh "${1:-hc_count2|fn_scatter1|hc_scatter2}" ${2:-6} 2>&1 | tail -8
