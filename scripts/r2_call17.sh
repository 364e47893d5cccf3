#!/bin/bash
# round 2, call 17: panel-entry charge in the work partition — sweep of the charge on 1 GPU (same box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for pc in 0 4000 8000 12000 16000; do
  export MF_PANEL_COST=$pc
  timeout 300 python bench.py --steps 8 --warmup 3 --legs none --no-cpu-baseline --no-e2e > gpurun_out/r2s_pc$pc.json 2> gpurun_out/r2s_pc$pc.err || echo "FAILED $pc"
  python - <<PY
import json
try:
    l=json.loads([x for x in open('gpurun_out/r2s_pc$pc.json') if x.startswith('{')][-1])
    r=l['roofline']
    print('panel_cost $pc', round(l['ms_per_step'],3), 'ms', {k:round(x,2) for k,x in r['families_ms_per_step'].items()}, 'frac', round(r['frac'],3), 'rmse', l['rmse_after_run'])
except Exception as e:
    print('panel_cost $pc', 'ERR', e)
PY
done
unset MF_PANEL_COST
timeout 600 python -m pytest tests/test_gpu_ccd.py tests/test_gpu_integer_tier.py -m gpu -q -x 2>&1 | tail -3
