#!/bin/bash
# round 2, call 23: ncu evidence of the final configuration — launch list + full capture of one rank's six sweeps (Netflix),
# full capture of the two ALS k=100 half-steps, full capture of two Yahoo-shape sweeps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_profile.sh r2z --legs none
CMD="python bench.py --workload als_netflix_k100 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --legs none"
timeout 300 $CMD > gpurun_out/prof_plain_als_r2z.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_als -s 2 -c 2 -o gpurun_out/prof_als_r2z -f $CMD > gpurun_out/ncu_full_als_r2z.log 2>&1
echo "als capture exit $?"; tail -1 gpurun_out/ncu_full_als_r2z.log
CMD="python bench.py --workload yahoo_k100 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --legs none"
timeout 600 $CMD > gpurun_out/prof_plain_yahoo_r2z.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_panel_sweep -s 8 -c 2 -o gpurun_out/prof_yahoo_r2z -f $CMD > gpurun_out/ncu_full_yahoo_r2z.log 2>&1
echo "yahoo capture exit $?"; tail -1 gpurun_out/ncu_full_yahoo_r2z.log
ls -la gpurun_out/*r2z*
