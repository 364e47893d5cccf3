#!/bin/bash
# round 2, call 12: full GPU suite (options f3/f4, PDL + metadata preload), headline bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/r2n_pytest.log
for wl in netflix_k40 ml100k_k10; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --legs none --no-cpu-baseline --no-e2e > gpurun_out/r2n_${wl}.json 2> gpurun_out/r2n_${wl}.err || echo "FAILED $wl"
  python - <<PY
import json
try:
    l=json.loads([x for x in open('gpurun_out/r2n_${wl}.json') if x.startswith('{')][-1])
    r=l['roofline']
    print('$wl', round(l['ms_per_step'],3), 'ms', {k:round(x,2) for k,x in r['families_ms_per_step'].items()}, 'frac', round(r['frac'],3), 'rmse', l['rmse_after_run'])
except Exception as e:
    print('$wl', 'ERR', e)
PY
done
