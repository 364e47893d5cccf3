#!/bin/bash
# one --set full capture of the ALS kernel (both half-steps of one iteration).  Usage: scripts/gpu_als_ncu.sh TAG WORKLOAD
cd "$(dirname "$0")/.."
TAG=$1; W=$2
mkdir -p gpurun_out
CMD="python bench.py --workload $W --steps 1 --warmup 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_als -c 2 -o gpurun_out/prof_als_$TAG -f $CMD > gpurun_out/ncu_full_als_$TAG.log 2>&1
echo "als full capture $TAG exit $?"; tail -1 gpurun_out/ncu_full_als_$TAG.log
