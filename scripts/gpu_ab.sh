#!/bin/bash
# A/B of library variants (scripts/build_variant.sh) on a list of workloads.  Usage: scripts/gpu_ab.sh "v0 b c" "als_ml20m_k10 ..." [bench args]
cd "$(dirname "$0")/.."
VARS=$1; WLS=$2; shift 2
fmt='import sys,json
for line in sys.stdin:
    if line.startswith("{"):
        l=json.loads(line); r=l.get("roofline") or {}
        print(sys.argv[1], sys.argv[2], round(l["ms_per_step"],3),"ms", {k:round(v,2) for k,v in (r.get("families_ms_per_step") or {}).items()}, "frac", r.get("frac") and round(r.get("frac"),3), "rmse", l["rmse_after_run"])'
for w in $WLS; do
  for v in main $VARS; do
    LIB=cuda-recommender_b200/libmfb200.so; [ $v != main ] && LIB=cuda-recommender_b200/libmfb200_$v.so
    MF_LIB=$PWD/$LIB timeout 300 python bench.py --workload $w --steps 2 --warmup 1 --no-e2e --no-cpu-baseline "$@" 2>gpurun_out/ab.err | python -c "$fmt" $w $v || tail -3 gpurun_out/ab.err
  done
done
