#!/bin/bash
# round 2, call 24 (8 GPUs): where an end-to-end multi-GPU call spends its time (MF_TRACE on every rank)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
MF_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline --legs none --no-bitwise-check > gpurun_out/r2y_8.json 2> gpurun_out/r2y_8.err; echo "exit $?"
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r2y_8.json') if x.startswith('{')][-1])
print(round(l['ms_per_step'],3), json.dumps(l['e2e'])[:900])
PY
grep -c . gpurun_out/r2y_8.err; grep -n "trace\|\[mf" gpurun_out/r2y_8.err | tail -150 | cut -c1-160
