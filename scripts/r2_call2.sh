#!/bin/bash
# round 2, call 2: persistent kernel — bitwise tests, then the headline with and without it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ccd.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r2b_pytest.log
for v in persist nopersist; do
  if [ $v = nopersist ]; then export MF_NO_PERSISTENT=1; else unset MF_NO_PERSISTENT; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --legs none --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_$v.json 2> gpurun_out/r2b_bench_$v.err; echo "bench $v exit $?"
  python - <<PY
import json
l=json.loads([x for x in open('gpurun_out/r2b_bench_$v.json') if x.startswith('{')][-1])
print('$v', l['ms_per_step'], l['gpu_launches'], json.dumps(l['roofline'])[:900], l['rmse_after_run'])
PY
done
unset MF_NO_PERSISTENT
timeout 300 python bench.py --workload ml100k_k10 --steps 10 --warmup 3 --legs none --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-400
