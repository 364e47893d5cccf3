#!/bin/bash
# round 2, final: the default bench line (as the driver runs it)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2fin_bench.json 2> gpurun_out/r2fin_bench.err; echo "bench exit $? in $(( $(date +%s) - t0 )) s"
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r2fin_bench.json') if x.startswith('{')][-1])
print('value', l['value'], 'launches', l['gpu_launches'], 'e2e', json.dumps(l['e2e'])[:500])
print('roofline', json.dumps(l['roofline'])[:800])
print('cpu', json.dumps(l['cpu_baseline'])[:300])
for k,v in (l.get('als') or {}).items():
    print(k, v['ms_per_step'], json.dumps(v.get('roofline'))[:400], 'e2e', (v.get('e2e') or {}).get('value'))
PY
