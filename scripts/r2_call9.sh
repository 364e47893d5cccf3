#!/bin/bash
# round 2, call 9: ncu full profile of the STREAM kernel (one item per warp)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --legs none --no-cpu-baseline --no-e2e --pipeline stream"
timeout 300 $CMD > gpurun_out/r2k_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_panel_sweep_stream -s 62 -c 2 -o gpurun_out/r2k_prof $CMD > gpurun_out/r2k_ncu.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/r2k_ncu.log
