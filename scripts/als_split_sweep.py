"""ALS split-threshold sweep (MF_ALS_SPLIT) on one GPU.  Usage: python scripts/als_split_sweep.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
import cuda_recommender_b200.datagen as dg  # noqa: E402

for shape, k in (("netflix", 100), ("netflix", 40), ("ml20m", 10)):
    data = dg.synth_named(shape, seed=1 + list(dg.SHAPES).index(shape), device="cuda")
    W0, H0 = pkg.initial_col(data["rows"], k), pkg.initial_col(data["cols"], k)
    for split in (4096, 8192, 16384, 65536, 1 << 30):
        os.environ["MF_ALS_SPLIT"] = str(split)
        with pkg.Session(data, pkg.make_params(pkg.SOLVER_ALS, k=k, lam=0.05)) as s:
            s.set_factors(W0, H0)
            s.iterate(1, want_stats=False)
            s.iterate(2, want_stats=False)
            print(f"{shape} k={k} split={split}: {s.last_seconds() / 2 * 1e3:.2f} ms/iter", flush=True)
    del data
    torch.cuda.empty_cache()
