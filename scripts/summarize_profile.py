#!/usr/bin/env python
"""Turns the raw gpurun_out/ evidence of one profiling call into the small text files kept under profiles/:
  profiles/<tag>_launches.txt  per-kernel totals of the ncu launch list (gpu__time_duration.sum)
  profiles/<tag>_ncu.txt       key metrics of every kernel in the --set full capture
Usage: scripts/summarize_profile.py TAG"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

lst = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lst):
    rows = [r for r in csv.reader(open(lst)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi])
        except ValueError:
            continue
        if r[ui] == "us":
            v *= 1e3
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_  (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# command: python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline ; {sum(a[0] for a in agg.values())} launches, {tot/1e6:.2f} ms total\n")
        f.write(f"{'share':>7} {'launches':>9} {'avg_us':>10}  kernel\n")
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{t/tot*100:6.2f}% {c:9d} {t/c/1e3:10.2f}  {n}\n")
    print("wrote", f"{tag}_launches.txt")

rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
            "lts__t_sector_hit_rate.pct"]
    with open(os.path.join(out_dir, f"{tag}_ncu.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:k_panel_sweep  (one rank of a steady-state outer iteration:\n")
        f.write("# fused CSC, fused CSR, solve CSC, solve CSR, solve CSC, solve CSR)\n")
        for r in rows[2:]:
            f.write("---\n")
            for w in want:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"{w} = {r[i]} {units[i]}\n")
            st = []
            for i, h in enumerate(hdr):
                if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                    try:
                        v = float(r[i])
                    except ValueError:
                        continue
                    if v > 0.3:
                        st.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            f.write("stall cycles per issued instruction: " + ", ".join(f"{h} {v:.2f}" for v, h in sorted(st, reverse=True)) + "\n")
    print("wrote", f"{tag}_ncu.txt")
