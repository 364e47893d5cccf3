#!/usr/bin/env python
"""Key metrics + stall breakdown + hottest source lines of every kernel in an .ncu-rep.  Usage: scripts/ncu_keys.py REPORT [nlines]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("====", d.get("Kernel Name", "")[:110])
    for k in keys:
        if k in d:
            print(f"  {k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
    stalls = [(float(d[h].replace(",", "") or 0), h) for h in hdr if "average_warp_latency_issue_stalled" in h or ("issue_stalled" in h and h.endswith("per_warp_active.pct"))]
    stalls = [(v, h) for v, h in stalls if v == v]
    for v, h in sorted(stalls, reverse=True)[:8]:
        print(f"  stall {h.split('issue_stalled_')[1][:40]:42s} {v:10.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True).stdout
# one table per kernel, separated by blank lines / repeated headers
blocks, cur = [], []
for line in src.splitlines():
    if line.startswith('"#"') or line.startswith('"Source"') or line.startswith('"Line'):
        if cur:
            blocks.append(cur)
        cur = [line]
    elif cur:
        cur.append(line)
if cur:
    blocks.append(cur)
for b in blocks:
    rr = list(csv.reader(io.StringIO("\n".join(b))))
    h = rr[0]
    try:
        si = h.index("Source")
        smp = [i for i, x in enumerate(h) if x.startswith("# Samples") or x == "Warp Stall Sampling (All Samples)"][0]
    except (ValueError, IndexError):
        print("source page header:", h[:12])
        continue
    tot = 0
    items = []
    for r in rr[1:]:
        try:
            v = float(r[smp].replace(",", ""))
        except (ValueError, IndexError):
            continue
        tot += v
        items.append((v, r[0] if h[0] in ("#", "Line") else "", r[si].strip()[:130]))
    print(f"---- source hot spots (total samples {tot:.0f})")
    for v, ln, text in sorted(items, reverse=True)[:nlines]:
        print(f"  {v / max(tot, 1) * 100:5.1f}%  L{ln:>4s}  {text}")
