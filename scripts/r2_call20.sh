#!/bin/bash
# round 2, call 20 (8 GPUs): per-CTA timeline of one rank's sweeps at 8 GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2v_trace*
MF_SWEEP_TRACE=$PWD/gpurun_out/r2v_trace MF_SWEEP_TRACE_CTA=606 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --legs none --no-bitwise-check --no-launch-timing > gpurun_out/r2v_8.json 2> gpurun_out/r2v_8.err; echo "exit $?"
python scripts/trace_summary.py gpurun_out/r2v_trace.rank0 480
python - <<'PY'
import collections
for rank in (0, 5):
    rows=[list(map(int,l.split())) for l in open(f'gpurun_out/r2v_trace.cta.rank{rank}') if not l.startswith('#')]
    by=collections.defaultdict(list)
    for l,c,a,b,ib,ie,p in rows: by[l].append((c,a,b,ib,ie,p))
    for l,v in sorted(by.items())[:6]:
        dur=[(b-a)/1e3 for c,a,b,ib,ie,p in v]
        st=[a/1e3 for c,a,b,ib,ie,p in v]
        srt=sorted(v,key=lambda x:-(x[2]))
        print(f"rank {rank} launch {l}: dur min {min(dur):.1f} med {sorted(dur)[len(dur)//2]:.1f} max {max(dur):.1f} us; start spread {max(st):.1f}; end min {min(b for c,a,b,ib,ie,p in v)/1e3:.1f} max {max(b for c,a,b,ib,ie,p in v)/1e3:.1f}")
        print("   slowest:", [(c, round(a/1e3,1), round((b-a)/1e3,1), ie-ib, p) for c,a,b,ib,ie,p in srt[:6]], " fastest:", [(c, round(a/1e3,1), round((b-a)/1e3,1), ie-ib, p) for c,a,b,ib,ie,p in srt[-3:]])
PY
rm -f gpurun_out/r2v_trace.rank[1-7] gpurun_out/r2v_trace.cta.rank[1-467]
