#!/bin/bash
# Round-end evidence on one GPU: tests, smoke, the default bench (both arms), the other workloads, ncu launch list +
# full capture of the sweep kernels.  Usage: scripts/gpu_final.sh TAG
cd "$(dirname "$0")/.."
TAG=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | head -20 >> gpurun_out/host.txt; free -g >> gpurun_out/host.txt
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_$TAG.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke_$TAG.log
echo "== bench (default)"; timeout 1200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench_$TAG.json
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit $?"; cut -c1-200 gpurun_out/bench_ref_$TAG.json
fmt='import sys,json
for line in sys.stdin:
    if line.startswith("{"):
        l=json.loads(line); r=l.get("roofline") or {}
        print(sys.argv[1], round(l["ms_per_step"],3),"ms", {k:round(v,3) for k,v in (r.get("families_ms_per_step") or {}).items()}, "avg_launch_ms", r.get("avg_launch_ms"), "frac", r.get("frac"), l.get("als"), "rmse", l["rmse_after_run"])'
for w in ml100k_k10 ml20m_k10 als_ml20m_k10 als_netflix_k40 als_netflix_k100; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --timing-stride 1 2>>gpurun_out/bench_other.err | tee gpurun_out/bench_${w}_$TAG.json | python -c "$fmt" $w
done
MF_SEPARATE_FINALIZE=1 timeout 300 python bench.py --workload ml100k_k10 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --timing-stride 1 2>>gpurun_out/bench_other.err | python -c "$fmt" ml100k_k10_separate_finalize
scripts/gpu_profile.sh $TAG
