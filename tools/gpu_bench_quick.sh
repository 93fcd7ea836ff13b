#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --reads-per-gpu ${MC2_READS:-6000000} --steps 3 --warmup 2 --no-cpu ${MC2_BENCH_ARGS:-} > gpurun_out/bench_quick.log 2>&1; echo "bench exit $?"; tail -3 gpurun_out/bench_quick.log
