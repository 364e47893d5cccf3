#!/bin/bash
# ncu evidence for the bench command: launch list (all kernels, device time) + one --set full capture of the
# CCD sweep kernels of one rank (fused CSC, fused CSR, solve CSC, solve CSR, ...).  Usage: scripts/gpu_profile.sh TAG [bench args]
cd "$(dirname "$0")/.."
TAG=$1; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline $*"
$CMD > gpurun_out/prof_plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/prof_plain2_$TAG.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_panel_sweep -s 246 -c 6 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"; tail -3 gpurun_out/ncu_full_$TAG.log
