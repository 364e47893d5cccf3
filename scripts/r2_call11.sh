#!/bin/bash
# round 2, call 11: programmatic dependent launch of the sweeps — bitwise tests, then A/B (same box) on three shapes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ccd.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2m_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2m_pytest.log
for wl in netflix_k40 ml100k_k10; do
for v in nopdl pdl nopdl pdl; do
  if [ $v = nopdl ]; then export MF_NO_PDL=1; else unset MF_NO_PDL; fi
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --legs none --no-cpu-baseline --no-e2e > gpurun_out/r2m_${wl}_$v.json 2> gpurun_out/r2m_${wl}_$v.err || echo "FAILED $v"
  python - <<PY
import json
try:
    l=json.loads([x for x in open('gpurun_out/r2m_${wl}_$v.json') if x.startswith('{')][-1])
    r=l['roofline']
    print('$wl $v', round(l['ms_per_step'],3), 'ms', {k:round(x,2) for k,x in r['families_ms_per_step'].items()}, 'frac', round(r['frac'],3), 'rmse', l['rmse_after_run'])
except Exception as e:
    print('$wl $v', 'ERR', e)
PY
done
done
