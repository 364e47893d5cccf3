#!/bin/bash
# Multi-GPU evidence on an 8-GPU box.  Usage: scripts/gpu_scale8.sh TAG
cd "$(dirname "$0")/.."
TAG=$1
mkdir -p gpurun_out
fmt='import sys,json
for line in sys.stdin:
    if line.startswith("{"):
        l=json.loads(line); r=l.get("roofline") or {}
        print(sys.argv[1], round(l["ms_per_step"],3),"ms", {k:round(v,2) for k,v in (r.get("families_ms_per_step") or {}).items()}, l.get("als"), "rmse", l["rmse_after_run"], "launches", l["gpu_launches"], l["clocks"])'
run() {  # N workload
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2952$1 bench.py --gpus $1 --workload $2 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/scale_$2_$1.err | tee gpurun_out/scale_$2_$1_$TAG.json | python -c "$fmt" "$2 x$1" || tail -5 gpurun_out/scale_$2_$1.err
}
run 4 netflix_k40
run 8 netflix_k40
run 8 als_netflix_k100
run 8 yahoo_k100
