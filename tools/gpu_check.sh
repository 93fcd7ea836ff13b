#!/bin/bash
# One GPU round trip: smoke, the GPU parity suite, a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/pytest_gpu.log
echo "== bench small"; timeout 600 python bench.py --reads-per-gpu ${MC2_READS:-4000000} --steps 2 --warmup 1 > gpurun_out/bench_small.log 2>&1; echo "bench exit $?"; tail -5 gpurun_out/bench_small.log
