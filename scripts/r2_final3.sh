#!/bin/bash
# round 2, last call: the whole GPU suite and the default bench line on the final tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2fin3_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2fin3_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2fin3_bench.json 2> gpurun_out/r2fin3_bench.err; echo "bench exit $?"
python - <<'PY'
import json
l=json.loads([x for x in open('gpurun_out/r2fin3_bench.json') if x.startswith('{')][-1])
print('value', l['value'], 'frac', l['roofline']['frac'], 'e2e', l['e2e']['value'], 'cpu', l['cpu_baseline']['value'], {k:v['ms_per_step'] for k,v in (l.get('als') or {}).items()})
PY
