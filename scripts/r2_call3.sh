#!/bin/bash
# round 2, call 3: geometry A/B of the register-ring sweep (threads per CTA x ring depth), per-launch and persistent
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in "" t768r5 t768r6 t640r6 t512r8; do
  if [ -z "$lib" ]; then unset MF_LIB; name=t1024r4; else export MF_LIB=$PWD/cuda-recommender_b200/libmfb200_$lib.so; name=$lib; fi
  for v in nopersist persist; do
    if [ $v = nopersist ]; then export MF_NO_PERSISTENT=1; else unset MF_NO_PERSISTENT; fi
    timeout 300 python bench.py --steps 5 --warmup 2 --legs none --no-cpu-baseline --no-e2e > gpurun_out/r2c_${name}_$v.json 2> gpurun_out/r2c_${name}_$v.err || echo "FAILED $name $v"
    python - <<PY
import json
try:
    l=json.loads([x for x in open('gpurun_out/r2c_${name}_$v.json') if x.startswith('{')][-1])
    r=l['roofline']
    ph=r.get('phases') or {}
    print('$name $v', round(l['ms_per_step'],2), 'ms', {k:round(x,2) for k,x in r['families_ms_per_step'].items()}, {k:round(x['avg_ms']*1e3,1) for k,x in ph.items()}, 'rmse', l['rmse_after_run'])
except Exception as e:
    print('$name $v', 'ERR', e)
PY
  done
done
