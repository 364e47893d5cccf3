import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
import cuda_recommender_b200.datagen as dg
d = dg.to_numpy(dg.synth_named("ml20m"))
for k, iters, kw in ((2, 1, dict()), (3, 2, dict()), (3, 2, dict(chunk=64))):
    W0 = pkg.initial_col(k, d["rows"])
    try:
        with pkg.Session(d, pkg.make_params(k=k, lam=0.05, maxinner=1, pipeline=0, **kw)) as s:
            s.set_factors(W0)
            st = s.iterate(iters)
        print("k", k, kw, "rmse", st[-1]["rmse"], flush=True)
    except Exception as e:
        print("k", k, kw, "ERROR", str(e)[-160:], flush=True)
        sys.exit(1)
