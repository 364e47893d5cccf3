#!/bin/bash
# round 2, call 18: lane-per-item path for short-piece copies — bitwise tests, Yahoo-shape A/B on 1 GPU, headline unchanged?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ccd.py tests/test_gpu_integer_tier.py "tests/test_gpu_config_scale.py::test_c5_yahoo_shape_layout_and_step_parity" -m gpu -q -x > gpurun_out/r2t_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2t_pytest.log
for v in old new new_pad16; do
  case $v in
    old) export MF_SHORT_ITEMS=0; unset MF_SHORT_PAD;;
    new) unset MF_SHORT_ITEMS; unset MF_SHORT_PAD;;
    new_pad16) unset MF_SHORT_ITEMS; export MF_SHORT_PAD=16;;
  esac
  timeout 600 python bench.py --workload yahoo_k100 --steps 2 --warmup 1 --legs none --no-cpu-baseline --no-e2e > gpurun_out/r2t_yahoo_$v.json 2> gpurun_out/r2t_yahoo_$v.err || echo "FAILED $v"
  python - <<PY
import json
try:
    l=json.loads([x for x in open('gpurun_out/r2t_yahoo_$v.json') if x.startswith('{')][-1])
    r=l['roofline']
    print('yahoo $v', round(l['ms_per_step'],1), 'ms', {k:round(x,1) for k,x in r['families_ms_per_step'].items()}, 'frac', round(r['frac'],3), 'avg_launch_ms', round(r['avg_launch_ms'],3), 'rmse', l['rmse_after_run'])
except Exception as e:
    print('yahoo $v', 'ERR', e)
PY
done
unset MF_SHORT_ITEMS MF_SHORT_PAD
timeout 300 python bench.py --steps 8 --warmup 3 --legs none --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
for x in sys.stdin:
    if x.startswith('{'):
        l=json.loads(x); print('netflix', round(l['ms_per_step'],3), l['roofline']['families_ms_per_step'], round(l['roofline']['frac'],3))"
