#!/usr/bin/env python
"""Summarises MF_SWEEP_TRACE dumps (one file per rank): per sweep mode, the median time CTA 0 spent in each phase and the gap
between consecutive launches.  Usage: scripts/trace_summary.py gpurun_out/trace.rank0 [skip_first_n]"""
import statistics
import sys

rows = [list(map(int, l.split())) for l in open(sys.argv[1]) if l.strip() and not l.startswith("#")]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 300
rows = rows[skip:]
names = ["wait(prologue->dependency)", "items(stage+stream)", "barrier", "finalize+push", "unpack(poll)"]
by = {}
for i, r in enumerate(rows):
    e, w, it, ba, fi, un, mode = r
    if ba < 0:
        continue
    seg = [w, it - w, ba - it, fi - ba, (un - fi) if un >= 0 else 0]
    total = (un if un >= 0 else fi)
    gap = rows[i + 1][0] - (e + total) if i + 1 < len(rows) else None
    by.setdefault(mode, []).append((seg, total, gap))
for mode, v in sorted(by.items()):
    med = [statistics.median(x[0][j] for x in v) / 1e3 for j in range(5)]
    tot = statistics.median(x[1] for x in v) / 1e3
    gaps = [x[2] for x in v if x[2] is not None]
    print(f"mode {mode:2d} n={len(v):4d} total {tot:6.1f} us | " + " | ".join(f"{n} {m:5.1f}" for n, m in zip(names, med)) +
          f" | next launch entry - this exit {statistics.median(gaps) / 1e3:5.1f} us")
