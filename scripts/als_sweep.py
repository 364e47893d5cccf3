"""ALS tuning sweep on one GPU: per workload the ratings are generated once, then every (MF_ALS_KS, MF_ALS_REGS)
combination runs 1 warm-up + 2 timed iterations.  Usage: python scripts/als_sweep.py [workload ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
import cuda_recommender_b200.datagen as dg  # noqa: E402

WL = {"ml20m_k10": ("ml20m", 10, [(0, 0), (4, 0), (5, 0)]),
      "netflix_k40": ("netflix", 40, [(0, 0), (2, 0), (8, 0)]),
      "netflix_k100": ("netflix", 100, [(0, 0), (2, 0)]),
      "netflix_k64": ("netflix", 64, [(0, 0), (1, 0), (4, 0)])}
for name in sys.argv[1:] or ["ml20m_k10", "netflix_k40", "netflix_k100"]:
    shape, k, configs = WL[name]
    data = dg.synth_named(shape, seed=1 + list(dg.SHAPES).index(shape), device="cuda")
    W0, H0 = pkg.initial_col(data["rows"], k), pkg.initial_col(data["cols"], k)
    for ks, regs in configs:
        for key, v in (("MF_ALS_KS", ks), ("MF_ALS_REGS", regs)):
            if v:
                os.environ[key] = str(v)
            else:
                os.environ.pop(key, None)
        with pkg.Session(data, pkg.make_params(pkg.SOLVER_ALS, k=k, lam=0.05)) as s:
            s.set_factors(W0, H0)
            s.iterate(1, want_stats=False)
            s.iterate(2, want_stats=False)
            ms = s.last_seconds() / 2 * 1e3
            print(f"{name} ks={ks or 'auto'} regs={regs or 'auto'}: {ms:.2f} ms/iter rmse {s.rmse():.6f}", flush=True)
    del data
    torch.cuda.empty_cache()
