#!/bin/bash
# round 2, call 16: per-CTA timeline of one rank's six sweeps on 1 GPU (who reaches the grid barrier last, and why)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2r_trace.*
MF_PANEL_COST=${PC:-16000} MF_SWEEP_TRACE=$PWD/gpurun_out/r2r_trace MF_SWEEP_TRACE_CTA=606 timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --legs none --no-launch-timing > gpurun_out/r2r_1.json 2> gpurun_out/r2r_1.err; echo "exit $?"
python - <<'PY'
import collections
rows=[list(map(int,l.split())) for l in open('gpurun_out/r2r_trace.cta.rank0') if not l.startswith('#')]
by=collections.defaultdict(list)
for l,c,a,b,ib,ie,p in rows: by[l].append((c,a,b,ib,ie,p))
for l,v in sorted(by.items()):
    dur=[(b-a)/1e3 for c,a,b,ib,ie,p in v]
    end=[b/1e3 for c,a,b,ib,ie,p in v]
    srt=sorted(v,key=lambda x:-(x[2]))
    print(f"launch {l}: ctas {len(v)} items-phase dur min {min(dur):.1f} med {sorted(dur)[len(dur)//2]:.1f} max {max(dur):.1f} us; end min {min(end):.1f} max {max(end):.1f}; start spread {max(a for c,a,b,ib,ie,p in v)/1e3:.1f}")
    print("   slowest:", [(c, round((b-a)/1e3,1), ie-ib, p) for c,a,b,ib,ie,p in srt[:6]], " fastest:", [(c, round((b-a)/1e3,1), ie-ib, p) for c,a,b,ib,ie,p in srt[-4:]])
PY
