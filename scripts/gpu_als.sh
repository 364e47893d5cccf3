#!/bin/bash
# ALS evidence: bench of the ALS workloads + one --set full capture of k_als_half (both half-steps) at k=100.
# Usage: scripts/gpu_als.sh TAG
cd "$(dirname "$0")/.."
TAG=$1; shift
mkdir -p gpurun_out
for w in als_ml20m_k10 als_netflix_k40 als_netflix_k100; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 1 2> gpurun_out/als_$w.err | tee gpurun_out/als_${w}_$TAG.json | python -c "
import sys,json
for line in sys.stdin:
    if line.startswith('{'):
        l=json.loads(line); print('$w', round(l['ms_per_step'],2),'ms', l['als'], l['rmse_after_run'])"
done
CMD="python bench.py --workload als_netflix_k100 --steps 1 --warmup 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_als -c 2 -o gpurun_out/prof_als_$TAG -f $CMD > gpurun_out/ncu_full_als_$TAG.log 2>&1
echo "als full capture exit $?"; tail -2 gpurun_out/ncu_full_als_$TAG.log
