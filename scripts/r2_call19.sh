#!/bin/bash
# round 2, call 19 (8 GPUs): scaling with the panel-entry charge, A/B against MF_PANEL_COST=0; Yahoo and ALS legs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
fmt='import sys,json
for line in sys.stdin:
    if line.startswith("{"):
        l=json.loads(line); r=l.get("roofline") or {}
        print(sys.argv[1], round(l["ms_per_step"],3),"ms", {k:round(v,2) for k,v in (r.get("families_ms_per_step") or {}).items()}, "bitwise", l.get("multi_gpu_bitwise"), "rmse", l["rmse_after_run"], "launches", l["gpu_launches"])'
run() {  # N workload tag extra-args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2952$1 bench.py --gpus $1 --workload $2 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --legs none $4 2>gpurun_out/r2u_$2_$1_$3.err | tee gpurun_out/r2u_$2_$1_$3.json | python -c "$fmt" "$2 x$1 $3" || tail -5 gpurun_out/r2u_$2_$1_$3.err
}
run 8 netflix_k40 pc8000
MF_PANEL_COST=0 run 8 netflix_k40 pc0 --no-bitwise-check
MF_PANEL_COST=16000 run 8 netflix_k40 pc16000 --no-bitwise-check
run 4 netflix_k40 pc8000 --no-bitwise-check
run 2 netflix_k40 pc8000 --no-bitwise-check
run 8 yahoo_k100 new --no-bitwise-check
run 8 als_netflix_k100 x --no-bitwise-check
