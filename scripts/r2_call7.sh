#!/bin/bash
# round 2, call 7: STREAM pipeline (4 producers, 24 consumers, no fence) — tests, bench, ncu full profile of three sweeps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ccd.py -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2h_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --legs none --no-cpu-baseline --no-e2e --pipeline stream > gpurun_out/r2h_stream.json 2> gpurun_out/r2h_stream.err || echo FAILED
python - <<PY
import json
l=json.loads([x for x in open('gpurun_out/r2h_stream.json') if x.startswith('{')][-1])
r=l['roofline']
print('stream', round(l['ms_per_step'],2), 'ms', {k:round(x,2) for k,x in r['families_ms_per_step'].items()}, 'frac', round(r['frac'],3), 'avg_launch_ms', round(r['avg_launch_ms'],4), 'rmse', l['rmse_after_run'])
PY
CMD="python bench.py --steps 1 --warmup 1 --legs none --no-cpu-baseline --no-e2e --pipeline stream"
timeout 300 $CMD > gpurun_out/r2h_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_panel_sweep_stream -s 60 -c 3 -o gpurun_out/r2h_prof $CMD > gpurun_out/r2h_ncu.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/r2h_ncu.log
