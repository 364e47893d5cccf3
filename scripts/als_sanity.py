"""Small ALS run for compute-sanitizer (memcheck / racecheck): python scripts/als_sanity.py [k ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
import cuda_recommender_b200.datagen as dg  # noqa: E402

d = dg.to_numpy(dg.synth_named("tiny"))
for k in [int(a) for a in sys.argv[1:]] or [10, 40]:
    W, H = pkg.initial_col(d["rows"], k), pkg.initial_col(d["cols"], k)
    st = pkg.als_train(d, W, H, pkg.make_params(pkg.SOLVER_ALS, k=k, lam=0.05, maxiter=2, quiet=1))
    print("k", k, "rmse", [round(x["rmse"], 6) for x in st], "finite", bool(np.isfinite(W).all() and np.isfinite(H).all()))
