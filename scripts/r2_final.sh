#!/bin/bash
# round 2, final 1-GPU validation: the whole GPU suite, smoke(), the default bench line, the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/r2fin_pytest.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/r2fin_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2fin_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2fin_smoke.log
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/r2fin_bench.json 2> gpurun_out/r2fin_bench.err; echo "bench exit $?"; grep -E "Elapsed|Maximum resident" gpurun_out/r2fin_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2fin_ref.json 2> gpurun_out/r2fin_ref.err; echo "ref exit $?"
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r2fin_bench.json') if x.startswith('{')][-1])
print('value', l['value'], 'e2e', json.dumps(l['e2e'])[:400])
print('roofline', json.dumps(l['roofline'])[:700])
print('cpu', json.dumps(l['cpu_baseline'])[:300])
for k,v in (l.get('als') or {}).items():
    print(k, v['ms_per_step'], json.dumps(v.get('roofline'))[:300], 'e2e', (v.get('e2e') or {}).get('value'))
r=json.loads([x for x in open('gpurun_out/r2fin_ref.json') if x.startswith('{')][-1])
print('ref', r['value'], r.get('per_step_seconds'))
PY
