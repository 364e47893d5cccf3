#!/bin/bash
# One gpurun call: GPU tests, then short benches of the headline + ALS workloads.  Usage: scripts/gpu_round.sh TAG [pytest args]
cd "$(dirname "$0")/.."
TAG=$1; shift
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 "$@" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_$TAG.log
fmt='import sys,json
for line in sys.stdin:
    if line.startswith("{"):
        l=json.loads(line); r=l.get("roofline") or {}
        print(sys.argv[1], round(l["ms_per_step"],3),"ms", {k:round(v,2) for k,v in (r.get("families_ms_per_step") or {}).items()}, "frac", r.get("frac"), l.get("als"), "rmse", l["rmse_after_run"], "e2e", (l.get("e2e") or {}).get("value"))'
for w in als_ml20m_k10 als_netflix_k40 als_netflix_k100; do
  timeout 300 python bench.py --workload $w --steps 2 --warmup 1 2>gpurun_out/bench_$w.err | tee gpurun_out/bench_${w}_$TAG.json | python -c "$fmt" $w
done
timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>gpurun_out/bench_ccd.err | tee gpurun_out/bench_ccd_$TAG.json | python -c "$fmt" ccd_inkernel_finalize
MF_SEPARATE_FINALIZE=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>>gpurun_out/bench_ccd.err | python -c "$fmt" ccd_separate_finalize
MF_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/e2e_trace_$TAG.err | python -c "$fmt" ccd_e2e
grep "mf trace" gpurun_out/e2e_trace_$TAG.err | tail -24
