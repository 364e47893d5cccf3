#!/bin/bash
# round 2, call 4: STREAM pipeline (work-list-order storage + TMA bulk tile ring) — bitwise tests, then A/B on the headline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ccd.py tests/test_gpu_integer_tier.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r2d_pytest.log
for v in registers stream stream_pad32 registers_pad8; do
  case $v in
    registers) a="--pipeline registers";;
    stream) a="--pipeline stream";;
    stream_pad32) a="--pipeline stream --pad 32";;
    registers_pad8) a="--pipeline registers --pad 8";;
  esac
  timeout 300 python bench.py --steps 5 --warmup 3 --legs none --no-cpu-baseline --no-e2e $a > gpurun_out/r2d_$v.json 2> gpurun_out/r2d_$v.err || echo "FAILED $v"
  python - <<PY
import json
try:
    l=json.loads([x for x in open('gpurun_out/r2d_$v.json') if x.startswith('{')][-1])
    r=l['roofline']
    print('$v', round(l['ms_per_step'],2), 'ms', {k:round(x,2) for k,x in r['families_ms_per_step'].items()}, 'frac', round(r['frac'],3), 'avg_launch_ms', round(r['avg_launch_ms'],4), 'rmse', l['rmse_after_run'])
except Exception as e:
    print('$v', 'ERR', e)
PY
done
