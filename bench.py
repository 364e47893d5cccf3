#!/usr/bin/env python
"""bench.py — seconds per CCD++ outer iteration on synthetic Netflix-shape ratings (BASELINE.json
configs[2]: 480 189 x 17 770, 100 M nnz, k=40, lambda=0.05, T=3 inner iterations).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] ...
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one steady-state outer iteration (oiter >= 2, so the add-back of src/CCD.cpp:100 is
included) over the whole rating matrix: k ranks x (fused first sweep of each copy + (T-1) x (v-solve,
u-solve)).  Prints ONE JSON line (see the contract in the task statement):
  value / ms_per_step  device time (CUDA events on the library's stream), ratings resident in HBM, max over ranks
  e2e                  the same metric through the drop-in C-ABI call with ordinary (pageable) HOST buffers — what the
                       reference's loader hands kernel_wrapper_ccdpp_NV: mf_ccdpp_train(..., maxiter=E) wall time / E,
                       uploads, layout build, per-iteration RMSE and the factor download included (median of
                       --e2e-calls calls); `pinned_call_seconds` is the same call from page-locked buffers
  roofline             dominant kernel family: compulsory HBM bytes of one launch / average launch time
  cpu_baseline         the reference's own OpenMP path (oracle/_ref, unmodified sources) on this box's host cores:
                       a REAL run of the benchmarked configuration (all k ranks), a few outer iterations, the
                       reference's own per-iteration rank_time + update_time (src/CCD.cpp:158) — nothing scaled
  als                  the two ALS configurations of BASELINE.json (configs[1] ML-20M k=10, configs[3] Netflix k=100):
                       seconds per iteration, FP32 roofline, CPU baseline, end to end through mf_als_train
  multi_gpu_bitwise    (--gpus N > 1) factors after the run are bit-identical to a single-GPU run of the same problem
`--impl reference` prints the reference-arm line: that CPU path alone, same config/metric/unit.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ccdpp_seconds_per_outer_iteration"
UNIT = "s"
FP32_SIMT_PEAK_TFLOPS = 74.0  # B200 FP32 FMA peak (148 SMs x 128 lanes x 2 flop x 1.965 GHz), SURVEY.md 8d

WORKLOADS = {
    # name: (shape, k, lambda, T)
    "netflix_k40": ("netflix", 40, 0.05, 3),
    "ml20m_k10": ("ml20m", 10, 0.05, 3),
    "ml100k_k10": ("ml100k", 10, 0.05, 3),
    "yahoo_k100": ("yahoo", 100, 0.05, 3),
    # ALS workloads (BASELINE.json configs[1] and configs[3]); T is unused.
    "als_ml100k_k10": ("ml100k", 10, 0.05, 0),
    "als_ml20m_k10": ("ml20m", 10, 0.05, 0),
    "als_netflix_k100": ("netflix", 100, 0.05, 0),
    "als_netflix_k40": ("netflix", 40, 0.05, 0),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="netflix_k40", choices=sorted(WORKLOADS))
    ap.add_argument("--schedule", default="fused", choices=["fused", "reference"])
    ap.add_argument("--layout", default="panel", choices=["panel", "direct"])
    ap.add_argument("--e2e-iters", type=int, default=3)
    ap.add_argument("--e2e-calls", type=int, default=3, help="end-to-end calls (the median wall time is reported)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-iters", type=int, default=3, help="steady-state outer iterations of the reference CPU path to time")
    ap.add_argument("--ref-budget-s", type=float, default=110.0, help="reference arm: CPU seconds to spend on timed iterations")
    ap.add_argument("--legs", default="auto", help="extra workloads measured into the same line: auto | none | comma list")
    ap.add_argument("--no-bitwise-check", action="store_true")
    ap.add_argument("--no-launch-timing", action="store_true")
    ap.add_argument("--panel-rows", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--pad", type=int, default=0)
    ap.add_argument("--pipeline", default="registers", choices=["registers", "async", "tma", "stream"])
    ap.add_argument("--timing-stride", type=int, default=8, help="per-launch CUDA events on every n-th rank only")
    return ap.parse_args()


def config_dict(workload, gpus):
    """The workload both arms run; identical keys and values in both JSON lines."""
    shape, k, lam, T = WORKLOADS[workload]
    from __graft_entry__ import load_package
    load_package()
    import cuda_recommender_b200.datagen as dg
    rows, cols, nnz, nt = dg.SHAPES[shape]
    als = workload.startswith("als_")
    return {"workload": (f"ALS k={k} lambda={lam}" if als else f"CCD++ k={k} lambda={lam} T={T}") +
                        f" on synthetic {shape}-shape ratings ({rows}x{cols}, {nnz} nnz)",
            "solver": "als" if als else "ccd++", "k": k, "lambda": lam, "inner_iters": T, "rows": rows, "cols": cols, "nnz": nnz,
            "nnz_test": nt, "step": "one ALS iteration (W then H half-step)" if als else "one steady-state outer iteration over all ratings",
            "l2_policy": "inputs larger than L2 (every sweep streams >= 0.6 GB; L2 is 126 MB)",
            "parallelism": f"rowblock-csr x colblock-csc over {gpus} gpu(s)"}


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling in the background (100 ms period).  It is started before the warm-up so that it is
    already running when the (short) timed region begins; only samples whose timestamp falls inside the timed
    region are used."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self):
        self.proc = None
        self.path = None

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                     stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def stop(self, gpu_indices, t_begin, t_end):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, sm_all, mx, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lo = datetime.datetime.fromtimestamp(t_begin - 0.11)
        hi = datetime.datetime.fromtimestamp(t_end + 0.11)
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                if int(f[1]) not in gpu_indices:
                    continue
                clk, cmax = float(f[2]), float(f[3])
            except ValueError:
                continue
            sm_all.append(clk)
            if not (lo <= ts <= hi):
                continue
            sm.append(clk); mx.append(cmax)
            for name, v in zip(names, f[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_whole_run": len(sm_all), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# the reference CPU path (oracle/_ref: the unmodified /root/reference/src sources)
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(dataset_dir, als, k, lam, T, steady_iters, threads):
    """ONE real run of the unmodified reference solver with all k ranks: ccdr1_OMP runs 1 + steady_iters outer
    iterations (the first has no add-back, src/CCD.cpp:100, and is not a steady-state step), ALS_OMP runs
    steady_iters.  Returns (seconds per step = mean of the reference's own rank_time + update_time of the
    steady-state iterations — its omp_get_wtime timers, src/CCD.cpp:158 / src/ALS.cpp:229 —, per-step list, run)."""
    from oracle import ref
    n = max(1, int(steady_iters))
    out = ref.train(dataset_dir, int(als), k, lam, n if als else 1 + n, max(T, 1), threads=threads)
    its = out["iters"] if als else out["iters"][1:]
    per = [it["rank_time"] + it["update_time"] for it in its]
    return sum(per) / len(per), per, out


def write_timing_dataset(data_np, tmpdir):
    import cuda_recommender_b200.datagen as dg
    # the reference's per-iteration RMSE is serial over the test set (src/tools.cpp:235-248) and not part
    # of the metric: keep 1000 test ratings so it costs nothing
    return dg.write_dataset(tmpdir, data_np, nnz_test_limit=1000)


def reference_iteration_estimate(workload):
    """Rough seconds per iteration of the CPU path (only to size the reference arm's iteration count)."""
    shape, k, lam, T = WORKLOADS[workload]
    import cuda_recommender_b200.datagen as dg
    rows, cols, nnz, nt = dg.SHAPES[shape]
    if workload.startswith("als_"):
        return (2.0 * nnz * k * (k + 1) + (rows + cols) * (k ** 3)) / 4e9 + 0.05
    return k * (48 + 16 * T) * nnz / 60e9 + 0.01


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from __graft_entry__ import load_package
    load_package()
    import cuda_recommender_b200.datagen as dg
    from oracle import ref
    shape, k, lam, T = WORKLOADS[args.workload]
    als = args.workload.startswith("als_")
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libmfref.so was not built"}))
        return 0
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    seed = 1 + list(dg.SHAPES).index(shape)
    data = dg.to_numpy(dg.synth_named(shape, seed=seed, device=dev))
    threads = os.cpu_count() or 1
    # as many of the requested steps as fit the CPU budget — every one of them a real, full iteration
    n = max(1, min(args.steps, int(args.ref_budget_s / reference_iteration_estimate(args.workload))))
    tmp = tempfile.mkdtemp(prefix="mfref_", dir=os.environ.get("TMPDIR", "/tmp"))
    try:
        write_timing_dataset(data, tmp)
        del data
        t0 = time.time()
        val, per, out = cpu_reference_run(tmp, als, k, lam, T, n, threads)
        wall = time.time() - t0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    warm = 0 if als else 1
    sample = (f"{'ALS_OMP' if als else 'ccdr1_OMP'} (unmodified reference sources), the benchmarked configuration itself (all k={k} ranks): "
              f"{warm + n} outer iterations in one call, mean of the reference's own rank_time+update_time over the {n} steady-state ones; "
              f"test set cut to 1000 ratings (its serial RMSE is not part of the metric)")
    line = {"impl": "reference", "metric": "als_seconds_per_iteration" if als else METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": n, "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": val * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args.workload, args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "per_step_seconds": [round(x, 4) for x in per], "solver_call_seconds": round(out["seconds"], 3),
            "wall_seconds_with_load": round(wall, 3), "extrapolated": False}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------
class Ctx:
    """process-wide plumbing of one bench run: torch.distributed, barrier, max over ranks"""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if args.gpus > 1 and self.world == 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist_mod
            self.dist = dist_mod
            self.dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier(device_ids=[self.local_rank])
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, flag):
        if self.dist is None:
            return bool(flag)
        t = self.torch.tensor([int(bool(flag))], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def shared_nccl_id(self, pkg):
        if self.world == 1:
            return None
        idt = self.torch.zeros(128, dtype=self.torch.uint8, device=self.dev)
        if self.rank == 0:
            idt.copy_(self.torch.tensor(list(pkg.nccl_unique_id()), dtype=self.torch.uint8))
        self.dist.broadcast(idt, 0)
        return bytes(idt.cpu().tolist())


def factor_digest(W, H):
    """order-sensitive integer digest of the factor bit patterns (exact: no floating-point reduction)"""
    import numpy as np
    out = []
    for a in (W, H):
        b = np.ascontiguousarray(a).view(np.uint32).astype(np.uint64).reshape(-1)
        idx = np.arange(1, b.size + 1, dtype=np.uint64)
        out += [int(b.sum() & np.uint64(0xFFFFFFFFFFFF)), int(((b * (idx | np.uint64(1))) & np.uint64(0xFFFFFFFF)).sum() & np.uint64(0xFFFFFFFFFFFF))]
    return out


def pin_dict(torch, host):
    import numpy as np
    pinned = {}
    for key, v in host.items():
        if isinstance(v, np.ndarray):
            tns = torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v).pin_memory()
            pinned[key] = tns.numpy().view(v.dtype)
            pinned["_keep_" + key] = tns
        else:
            pinned[key] = v
    return pinned


def e2e_calls(ctx, pkg, host, als, k, cols, rows, W0, H0, p2, nccl_id, ncalls):
    """wall seconds of `ncalls` one-shot training calls from the host arrays in `host` (median first)"""
    import numpy as np
    walls, rmse = [], None
    for _ in range(max(1, ncalls)):
        W = W0.copy()
        H = H0.copy() if als else np.zeros((k, cols), np.float32)
        ctx.barrier()
        t0 = time.perf_counter()
        if ctx.world == 1:
            st = (pkg.als_train if als else pkg.ccdpp_train)(host, W, H, p2)
        else:
            s2 = pkg.Session(host, p2, rank=ctx.rank, nranks=ctx.world, nccl_id=nccl_id)  # same id: the communicator is reused
            s2.set_factors(W, H if als else None)
            st = s2.iterate(p2.maxiter)
            s2.get_factors()
            s2.close()
        ctx.barrier()
        walls.append(ctx.max_over_ranks(time.perf_counter() - t0))
        rmse = st[-1]["rmse"]
    return sorted(walls)[len(walls) // 2], walls, rmse


def run_leg(ctx, pkg, dg, args, workload, headline):
    """One workload on the current process group: device-timed steps, kernel-family times, end to end, CPU baseline.
    Returns (fields for the JSON line / the leg's sub-object, clocks)."""
    import numpy as np
    torch = ctx.torch
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    shape, k, lam, T = WORKLOADS[workload]
    als = workload.startswith("als_")
    seed = 1 + list(dg.SHAPES).index(shape)
    t0 = time.time()
    data = dg.synth_named(shape, seed=seed, device=dev)  # same seed on every rank -> identical ratings
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    rows, cols, nnz = data["rows"], data["cols"], data["nnz"]
    for key in ("coo_row", "coo_col", "coo_val"):
        data.pop(key, None)

    params = pkg.make_params(pkg.SOLVER_ALS if als else pkg.SOLVER_CCD, k=k, lam=lam, maxiter=args.steps, maxinner=max(T, 1), device=ctx.local_rank,
                             schedule=pkg.SCHEDULE_REFERENCE if args.schedule == "reference" else pkg.SCHEDULE_FUSED,
                             layout=pkg.LAYOUT_DIRECT if args.layout == "direct" else pkg.LAYOUT_PANEL,
                             panel_rows=args.panel_rows, chunk=args.chunk, no_launch_timing=int(args.no_launch_timing),
                             pipeline={"registers": 0, "async": 1, "tma": 2, "stream": 3}[args.pipeline], timing_stride=args.timing_stride, pad_entries=args.pad)
    nccl_id = ctx.shared_nccl_id(pkg)

    # factors exactly as the reference seeds them (tools.cpp:165-173): libc srand(0)/rand()
    W0 = pkg.initial_col(rows, k) if als else pkg.initial_col(k, rows)
    H0 = pkg.initial_col(cols, k) if als else None

    sess = pkg.Session(data, params, rank=rank, nranks=world, nccl_id=nccl_id)
    sess.set_factors(W0, H0)
    want_e2e = (not args.no_e2e) and (headline or world == 1)
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    want_bitwise = world > 1 and not args.no_bitwise_check and shape != "yahoo"
    need_host = want_e2e or want_cpu
    host = dg.to_numpy(data) if need_host else None  # ordinary (pageable) numpy arrays
    keep_dev = data if (want_bitwise and rank == 0) else None
    del data
    torch.cuda.empty_cache()

    # ---- warm-up, then K timed steps bracketed by barrier + synchronize
    sampler = ClockSampler()
    if rank == 0 and headline:
        sampler.start()
    if args.warmup > 0:
        sess.iterate(args.warmup, want_stats=False)
    ctx.barrier()
    t_begin = time.time()
    t0 = time.perf_counter()
    sess.iterate(args.steps, want_stats=False)
    ctx.barrier()
    wall = time.perf_counter() - t0
    t_end = time.time()
    clocks = sampler.stop(set(range(world)), t_begin, t_end) if (rank == 0 and headline) else None
    dev_s = ctx.max_over_ranks(sess.last_seconds())
    wall = ctx.max_over_ranks(wall)
    kt = sess.kernel_times()
    rmse = sess.rmse()
    sec_per_iter = dev_s / args.steps
    launches = int(kt["total_launches"])  # every kernel the library launched in the timed region

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roofline = None
    if als and kt["als_launches"]:
        # Gram + RHS + factorisation flops of one iteration (symmetric count)
        flops = 2.0 * nnz * k * (k + 1) + 4.0 * nnz * k + (rows + cols) * (k ** 3 / 3.0 + 2.0 * k * k)
        avg = kt["als_s"] / kt["als_launches"]  # one half-step
        ach = flops / world / (2 * avg) / 1e12  # per GPU: every rank's launch works on 1/world of the segments
        roofline = {"bound": "fp32", "kernel": "k_als_tile (one half-step per launch)", "achieved": ach, "peak": FP32_SIMT_PEAK_TFLOPS,
                    "unit": "TFLOP/s", "frac": ach / FP32_SIMT_PEAK_TFLOPS,
                    "peak_source": "nominal FP32 FMA peak of the part (148 SMs x 128 lanes x 2 x 1.965 GHz; scripts/ubench/mma_tf32.cu measures 72 TFLOP/s of FFMA on it); MEASURED_PEAKS.json holds no FP32-SIMT figure",
                    "traffic": None, "flops_per_iteration": flops, "flops_per_launch": flops / world / 2, "avg_launch_ms": avg * 1e3, "launches": int(kt["als_launches"]),
                    "share_of_step": (2 * avg) / sec_per_iter if sec_per_iter > 0 else None,
                    "parity_note": "ALS factors are compared with an FP64 yardstick (the reference's explicit FP32 inverse is itself up to "
                                   "~2e-3 relative l2 away from it, SURVEY App. D); test RMSE within 1e-4 of the reference CPU path"}
    elif not als:
        fam = {}
        for name in ("solve", "fused", "update"):
            if kt[name + "_launches"]:
                fam[name] = (kt[name + "_s"], kt[name + "_launches"], kt[name + "_bytes"])
        # launches of each family in one outer iteration of the schedule that ran
        persistent = kt.get("persistent_launches", 0) > 0
        if args.schedule == "fused":
            per_step = {"solve": 2 * k * (T - 1), "fused": 2 * k, "update": 0, "finalize": 2 * k * T}
        else:
            per_step = {"solve": 2 * k * T, "fused": 0, "update": 4 * k, "finalize": 2 * k * T}
        if kt["finalize_launches"] == 0:
            # finalize (and, multi-GPU, the exchange) ran inside the sweep kernels; what is left is the once-per-iteration barrier
            per_step["finalize"] = 0
            per_step["collective"] = 1 if world > 1 else 0
        else:
            per_step["collective"] = 2 * k * T if world > 1 else 0
        peak = float(peaks.get("hbm_gbs", 6650.0))
        fam_ms = {f: fam[f][0] / fam[f][1] * per_step[f] * 1e3 for f in fam} | \
                 {"finalize": kt["finalize_s"] / max(kt["finalize_launches"], 1) * per_step["finalize"] * 1e3,
                  "collective": kt["collective_s"] / max(kt["collective_launches"], 1) * per_step["collective"] * 1e3}
        traffic_tab = {}
        try:
            traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(workload, {})
        except Exception:
            pass
        if persistent:
            # the dominant kernel IS the step: one cooperative launch per outer iteration (k ranks x 2T sweep phases)
            n = int(kt["persistent_launches"])
            avg = kt["persistent_s"] / n
            nbytes = int(kt["persistent_bytes"])
            achieved = nbytes / avg / 1e9
            phases = {f: {"avg_ms": fam[f][0] / fam[f][1] * 1e3, "bytes": int(fam[f][2]), "gbs": fam[f][2] / (fam[f][0] / fam[f][1]) / 1e9,
                          "frac": fam[f][2] / (fam[f][0] / fam[f][1]) / 1e9 / peak, "per_step": per_step[f]} for f in fam}
            roofline = {"bound": "hbm", "kernel": f"k_ccd_persistent (one launch = one outer iteration: {2 * k * T} sweep phases, {args.layout} layout)",
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": "measured" if peaks else "fallback",
                        "traffic": traffic_tab.get("persistent"), "bytes_per_launch": nbytes, "avg_launch_ms": avg * 1e3, "launches": n,
                        "share_of_step": avg / sec_per_iter if sec_per_iter > 0 else None,
                        "timing": "CUDA events around every launch in the timed region; `phases` = in-kernel %globaltimer stamps at the grid barriers",
                        "achieved_at_survey_bytes": k * (20 + 16 * (T - 1) + 4) * (nnz / world) / avg / 1e9,
                        "phases": phases, "families_ms_per_step": fam_ms}
        elif fam:
            top = max(fam, key=lambda n: fam[n][0] / fam[n][1] * per_step[n])
            secs, n, nbytes = fam[top]
            avg = secs / n
            achieved = nbytes / avg / 1e9
            survey_bytes = {"solve": 8, "fused": 12, "update": 12}[top] * (nnz / world)
            roofline = {"bound": "hbm", "kernel": f"ccd {top} sweep ({args.layout} layout)",
                        "achieved": achieved, "peak": peak,
                        "unit": "GB/s", "frac": achieved / peak, "peak_source": "measured" if peaks else "fallback",
                        # (the ncu DRAM bytes were captured on one GPU: per launch they only compare with a one-GPU launch)
                        "traffic": traffic_tab.get(top) if world == 1 else None, "bytes_per_launch": nbytes, "avg_launch_ms": avg * 1e3, "launches": n,
                        "share_of_step": (avg * per_step[top]) / sec_per_iter if sec_per_iter > 0 else None,
                        "timing": f"CUDA events around the launches of every {max(args.timing_stride, 1)}-th rank in the timed region",
                        "achieved_at_survey_bytes": survey_bytes / avg / 1e9,
                        "families_ms_per_step": fam_ms}

    # ---- multi-GPU == single GPU, bit for bit: rank 0 repeats the same iterations on a single-GPU session
    bitwise = None
    if want_bitwise:
        W, H = sess.get_factors()
        dig = factor_digest(W, H)
        t = torch.tensor(dig, dtype=torch.int64, device=dev)
        t0_ = t.clone()
        ctx.dist.broadcast(t0_, 0)
        ranks_agree = ctx.all_true(bool(torch.equal(t, t0_)))
        same = True
        if rank == 0:
            p1 = pkg.make_params(pkg.SOLVER_ALS if als else pkg.SOLVER_CCD, k=k, lam=lam, maxiter=1, maxinner=max(T, 1), device=ctx.local_rank,
                                 schedule=params.schedule, layout=params.layout, panel_rows=args.panel_rows, chunk=args.chunk,
                                 no_launch_timing=1, pipeline=params.pipeline, pad_entries=args.pad)
            with pkg.Session(keep_dev, p1) as s1:
                s1.set_factors(W0, H0)
                s1.iterate(args.warmup + args.steps, want_stats=False)
                W1, H1 = s1.get_factors()
            same = bool(np.array_equal(W.view(np.uint32), W1.view(np.uint32)) and np.array_equal(H.view(np.uint32), H1.view(np.uint32)))
            del W1, H1
        keep_dev = None
        torch.cuda.empty_cache()
        bitwise = ctx.all_true(same) and ranks_agree
        del W, H

    # ---- end to end through the public drop-in call, ordinary host buffers, every copy inside the timed region
    e2e = None
    if want_e2e:
        E = max(1, args.e2e_iters)
        sess.close()
        p2 = pkg.make_params(pkg.SOLVER_ALS if als else pkg.SOLVER_CCD, k=k, lam=lam, maxiter=E, maxinner=max(T, 1), device=ctx.local_rank,
                             schedule=params.schedule, layout=params.layout, panel_rows=args.panel_rows, chunk=args.chunk,
                             no_launch_timing=1, pipeline=params.pipeline, pad_entries=args.pad)
        nfac = W0.size + (H0.size if als else 0)
        h2d = sum(v.nbytes for v in host.values() if isinstance(v, np.ndarray)) + nfac * 4
        d2h = (rows + cols) * k * 4 + 8 * E
        e2e_wall, walls, e2e_rmse = e2e_calls(ctx, pkg, host, als, k, cols, rows, W0, H0, p2, nccl_id, args.e2e_calls)
        pinned = pin_dict(torch, host)
        pin_wall, pin_walls, _ = e2e_calls(ctx, pkg, pinned, als, k, cols, rows, W0, H0, p2, nccl_id, args.e2e_calls)
        del pinned
        e2e = {"value": e2e_wall / E, "unit": UNIT, "h2d_bytes_per_step": int(h2d / E), "d2h_bytes_per_step": int(d2h / E),
               "call": ("mf_als_train" if als else "mf_ccdpp_train") if world == 1 else "mf_session_create_dist+iterate+get_factors",
               "host_buffers": "pageable (ordinary numpy arrays, as the reference's loader allocates them); the library stages them itself",
               "outer_iters_per_call": E, "call_seconds": e2e_wall, "calls": len(walls), "call_seconds_all": [round(w, 4) for w in walls],
               "pinned_call_seconds": pin_wall, "pinned_call_seconds_all": [round(w, 4) for w in pin_walls],
               "statistic": "median over calls (sessions of one process reuse the pooled arena block)", "rmse": e2e_rmse}
    else:
        sess.close()

    # ---- the reference's CPU path on this box's host cores (rank 0, single-GPU runs only)
    cpu = None
    if want_cpu:
        from oracle import ref, port
        threads = os.cpu_count() or 1
        if workload == "als_netflix_k100" or (als and reference_iteration_estimate(workload) > 20):
            # a full iteration of ALS_OMP takes ~15 minutes here: time the oracle's restatement of the two half-steps
            # (bit-equal to the reference, tests/test_oracle_vs_ref.py) on a seeded sample of rows and of columns and
            # scale by the share of ratings covered — per-segment work is |O| k^2 + k^3, so a random sample scales
            rng = np.random.default_rng(5)
            est, parts = 0.0, []
            for name, ptr, idx, val, Y, nseg in (("rows", host["csr_ptr"], host["csr_idx"], host["csr_val"], H0, rows),
                                                 ("cols", host["csc_ptr"], host["csc_idx"], host["csc_val"], W0, cols)):
                p64 = ptr.astype(np.int64)
                pick = np.sort(rng.choice(nseg, size=max(1, nseg // 256), replace=False))
                lens = p64[pick + 1] - p64[pick]
                take = np.concatenate([np.arange(p64[s], p64[s + 1]) for s in pick]) if lens.sum() else np.zeros(0, np.int64)
                sp = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
                t0 = time.perf_counter()
                port.als_half_step(sp, idx[take], val[take], Y, k, lam)
                dt = time.perf_counter() - t0
                cover = (lens.sum() * k * k + len(pick) * k ** 3) / (float(nnz) * k * k + nseg * k ** 3)
                est += dt / cover
                parts.append(f"{name}: {len(pick)} of {nseg} segments, {int(lens.sum())} ratings, {dt:.2f} s")
            cpu = {"value": est, "unit": UNIT, "cores": threads, "kind": "port", "extrapolated": True,
                   "sample": "oracle restatement of ALS_OMP's half-steps on 1/256 of the rows and of the columns (" + "; ".join(parts) +
                             "), each scaled by its share of the |O|k^2 + k^3 work"}
        elif ref.available():
            tmp = tempfile.mkdtemp(prefix="mfref_", dir=os.environ.get("TMPDIR", "/tmp"))
            try:
                write_timing_dataset(host, tmp)
                val, per, _ = cpu_reference_run(tmp, als, k, lam, T, args.cpu_iters, threads)
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
            cpu = {"value": val, "unit": UNIT, "cores": threads, "kind": "reference", "extrapolated": False,
                   "sample": f"unmodified reference {'ALS_OMP' if als else 'ccdr1_OMP'}, all k={k} ranks, {len(per)} steady-state iterations "
                             f"(its own rank_time+update_time: {[round(x, 3) for x in per]})"}
    del host

    out = {"metric": "als_seconds_per_iteration" if als else METRIC, "value": sec_per_iter, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": sec_per_iter * 1e3, "config": config_dict(workload, world),
           "impl_config": {"schedule": args.schedule, "layout": args.layout, "datagen_s": round(gen_s, 2)},
           "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
           "wall_ms_per_step": wall / args.steps * 1e3, "rmse_after_run": rmse, "outer_iterations_done": args.warmup + args.steps}
    if bitwise is not None:
        out["multi_gpu_bitwise"] = bitwise
    return out, clocks


def run_b200_arm(args):
    from __graft_entry__ import load_package
    pkg = load_package()
    import cuda_recommender_b200.datagen as dg
    ctx = Ctx(args)
    head, clocks = run_leg(ctx, pkg, dg, args, args.workload, headline=True)
    legs = {}
    if args.legs == "auto":
        names = []
        if args.workload == "netflix_k40":
            names = ["als_netflix_k100"] + (["als_ml20m_k10"] if ctx.world == 1 else []) + (["yahoo_k100"] if ctx.world == 8 else [])
    elif args.legs in ("none", ""):
        names = []
    else:
        names = [n for n in args.legs.split(",") if n in WORKLOADS]
    for name in names:
        ctx.torch.cuda.empty_cache()
        leg, _ = run_leg(ctx, pkg, dg, args, name, headline=False)
        legs[name] = leg
    if ctx.rank == 0:
        line = {"metric": head["metric"], "value": head["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": False, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": head["config"], "impl_config": head["impl_config"],
                "clocks": clocks, "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
                "cpu_baseline": head["cpu_baseline"], "wall_ms_per_step": head["wall_ms_per_step"], "rmse_after_run": head["rmse_after_run"],
                "outer_iterations_done": head["outer_iterations_done"]}
        if "multi_gpu_bitwise" in head:
            line["multi_gpu_bitwise"] = head["multi_gpu_bitwise"]
        als_legs = {n: v for n, v in legs.items() if n.startswith("als_")}
        other = {n: v for n, v in legs.items() if not n.startswith("als_")}
        line["als"] = als_legs or None
        if other:
            line["other_workloads"] = other
        print(json.dumps(line))
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
