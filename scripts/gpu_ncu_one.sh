#!/bin/bash
# one --set full capture of a few launches of one kernel.  Usage: scripts/gpu_ncu_one.sh TAG KERNEL_REGEX SKIP COUNT [bench args]
cd "$(dirname "$0")/.."
TAG=$1; KRE=$2; SKIP=$3; CNT=$4; shift 4
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline $*"
$CMD > gpurun_out/prof_plain_$TAG.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"; tail -2 gpurun_out/ncu_full_$TAG.log
