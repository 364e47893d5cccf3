#!/bin/bash
# round 2, call 10: register-ring sweep with an L2 prefetch of the CTA's own stream (work-list-order storage), distance sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ccd.py -m gpu -q -x > gpurun_out/r2l_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2l_pytest.log
for pf in 0 4096 8192 16384 32768 65536; do
for pad in 32 8; do
  export MF_L2_PREFETCH=$pf
  timeout 300 python bench.py --steps 5 --warmup 3 --legs none --no-cpu-baseline --no-e2e --pipeline registers --pad $pad > gpurun_out/r2l_pf${pf}_pad$pad.json 2> gpurun_out/r2l_pf${pf}_pad$pad.err || echo "FAILED $pf"
  python - <<PY
import json
try:
    l=json.loads([x for x in open('gpurun_out/r2l_pf${pf}_pad$pad.json') if x.startswith('{')][-1])
    r=l['roofline']
    print('pf $pf pad $pad', round(l['ms_per_step'],2), 'ms', {k:round(x,2) for k,x in r['families_ms_per_step'].items()}, 'frac', round(r['frac'],3), 'avg_launch_ms', round(r['avg_launch_ms'],4), 'rmse', l['rmse_after_run'])
except Exception as e:
    print('pf $pf', 'ERR', e)
PY
done
done
