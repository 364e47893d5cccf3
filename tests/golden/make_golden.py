"""Generates tests/golden/*.npz from the UNMODIFIED reference CPU path (oracle/_ref/libmfref.so,
compiled from /root/reference/src by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

Each fixture stores the input ratings (COO, CSR/CSC are re-derived deterministically), the
hyper-parameters, and the reference's outputs: final factors, the per-iteration RMSE lines it
printed (src/CCD.cpp:158, src/ALS.cpp:229), its final calrmse and, for CCD++, the residual it
leaves in R.  The oracle restatement and the CUDA path are both checked against these.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
import cuda_recommender_b200.datagen as dg  # noqa: E402
from oracle import ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, shape name, seed, als, k, lambda, maxiter, maxinner
    ("ccd_c1_ml100k", "ml100k", 1, 0, 10, 0.05, 3, 3),   # BASELINE.json configs[0]
    ("ccd_small_T1", "small", 11, 0, 6, 0.1, 2, 1),
    ("ccd_tiny_k1", "tiny", 12, 0, 1, 0.05, 3, 2),
    ("als_ml100k", "ml100k", 1, 1, 10, 0.05, 3, 1),
    ("als_small_k24", "small", 13, 1, 24, 0.05, 2, 1),
]


def main():
    for name, shape, seed, als, k, lam, maxiter, maxinner in CASES:
        d = dg.to_numpy(dg.synth_named(shape, seed=seed))
        with tempfile.TemporaryDirectory() as tmp:
            dg.write_dataset(tmp, d)
            per_iter = []
            out = ref.train(tmp, als, k, lam, maxiter, maxinner, threads=4, want_residual=not als)
            per_iter = [it["rmse"] for it in out["iters"]]
            first = ref.train(tmp, als, k, lam, 1, maxinner, threads=4)  # state after ONE outer iteration
        rows = np.repeat(np.arange(d["rows"], dtype=np.uint32), np.diff(d["csr_ptr"].astype(np.int64)))
        fix = dict(rows=d["rows"], cols=d["cols"], coo_row=rows.astype(np.uint16 if d["rows"] < 65536 else np.uint32),
                   coo_col=d["csr_idx"].astype(np.uint16 if d["cols"] < 65536 else np.uint32),
                   coo_val=d["csr_val"].astype(np.uint8),
                   test_row=d["test_row"].astype(np.uint16), test_col=d["test_col"].astype(np.uint16),
                   test_val=d["test_val"].astype(np.uint8),
                   als=als, k=k, lam=np.float32(lam), maxiter=maxiter, maxinner=maxinner,
                   W=out["W"], H=out["H"], W_iter1=first["W"], H_iter1=first["H"], rmse_iter1=first["rmse"],
                   rmse_printed=np.array(per_iter), rmse_final=out["rmse"])
        if not als:
            fix["csr_resid"] = out["csr_val"]
            fix["csc_resid"] = out["csc_val"]
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **fix)
        print(name, "rmse", per_iter, "->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
