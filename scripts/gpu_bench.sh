#!/bin/bash
# bench + ncu launch list (+ optional full capture of one kernel).  Usage: scripts/gpu_bench.sh [bench args]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== bench"; timeout 1200 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
