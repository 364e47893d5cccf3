#!/bin/bash
# parameter sweep of the bench (short runs, no e2e / CPU legs); one JSON line per variant in gpurun_out/sweep.jsonl
# stdin: one variant per line: bench args, optionally prefixed by VAR=value assignments (e.g. MF_LIB=... --chunk 512)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
while read -r LINE; do
  [ -z "$LINE" ] && continue
  ENVS=""; ARGS=""
  for tok in $LINE; do
    if [[ -z "$ARGS" && "$tok" == *=* && "$tok" != --* ]]; then ENVS="$ENVS $tok"; else ARGS="$ARGS $tok"; fi
  done
  echo "== $LINE"
  OUT=$(timeout 600 env $ENVS python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline $ARGS 2> gpurun_out/sweep.err | tail -1)
  echo "{\"args\": \"$LINE\", \"line\": $OUT}" >> gpurun_out/sweep.jsonl
  echo "$OUT" | python -c "import sys,json; l=json.loads(sys.stdin.read()); r=l['roofline']; print(round(l['ms_per_step'],2),'ms/step', {k:round(v,2) for k,v in r['families_ms_per_step'].items()})" 2>/dev/null || tail -3 gpurun_out/sweep.err
done
