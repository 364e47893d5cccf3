import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
pkg = load_package()
import cuda_recommender_b200.datagen as dg
shape = sys.argv[1] if len(sys.argv) > 1 else "ml100k"
d = dg.to_numpy(dg.synth_named(shape))
k = 3
W0 = pkg.initial_col(k, d["rows"])
for kw in (dict(chunk=8), dict(chunk=8, panel_rows=64), dict(chunk=16, panel_rows=256), dict()):
    ref = None
    for pipeline in (1, 2, 0):
        try:
            with pkg.Session(d, pkg.make_params(k=k, lam=0.05, maxinner=2, pipeline=pipeline, **kw)) as s:
                s.set_factors(W0)
                st = s.iterate(2)
                W, H = s.get_factors()
            if ref is None:
                ref = (W, H)
            print(shape, kw, "pipeline", pipeline, "rmse", st[-1]["rmse"], "equal_to_registers", bool(np.array_equal(W, ref[0]) and np.array_equal(H, ref[1])), flush=True)
        except Exception as e:
            print(shape, kw, "pipeline", pipeline, "ERROR", str(e)[-200:], flush=True)
            sys.exit(1)
