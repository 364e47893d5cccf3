#!/bin/bash
# round 2, call 8: STREAM pipeline v4 (one item per warp) — warp sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ccd.py tests/test_gpu_integer_tier.py -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2j_pytest.log
for lib in "" p4c16 p4c28 p2c30; do
  if [ -z "$lib" ]; then unset MF_LIB; name=p4c24; else export MF_LIB=$PWD/cuda-recommender_b200/libmfb200_$lib.so; name=$lib; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --legs none --no-cpu-baseline --no-e2e --pipeline stream > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.err || echo "FAILED $name"
  python - <<PY
import json
try:
    l=json.loads([x for x in open('gpurun_out/r2j_$name.json') if x.startswith('{')][-1])
    r=l['roofline']
    print('$name', round(l['ms_per_step'],2), 'ms', {k:round(x,2) for k,x in r['families_ms_per_step'].items()}, 'frac', round(r['frac'],3), 'avg_launch_ms', round(r['avg_launch_ms'],4), 'rmse', l['rmse_after_run'])
except Exception as e:
    print('$name', 'ERR', e)
PY
done
