#!/usr/bin/env python
"""Run under torch.distributed.run with N >= 2 ranks (one per GPU): every rank builds the same synthetic
ratings, trains CCD++ (and ALS) through a multi-GPU session (row-block CSR x column-block CSC, NCCL
all-gather of the fresh factor blocks), and rank 0 also trains the same problem on a single-GPU session.
Checks: factors identical on every rank, and identical BIT FOR BIT to the single-GPU run (CCD++: the
reduction tree of a segment does not depend on the shard it sits in; ALS: one CTA per segment either way).
Prints one JSON line on rank 0; exit code 1 on mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "ml100k"
    pkg = load_package()
    import cuda_recommender_b200.datagen as dg
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    def shared_id():
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.tensor(list(pkg.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())

    d = dg.to_numpy(dg.synth_named(shape, seed=5))
    ok = True
    report = {"world": world, "shape": shape}
    for solver, k, inner in ((pkg.SOLVER_CCD, 6, 2), (pkg.SOLVER_ALS, 8, 1)):
        p = pkg.make_params(solver, k=k, lam=0.05, maxinner=inner, device=local)
        if solver == pkg.SOLVER_CCD:
            W0, H0 = pkg.initial_col(k, d["rows"]), None
        else:
            W0, H0 = pkg.initial_col(d["rows"], k), pkg.initial_col(d["cols"], k)
        with pkg.Session(d, p, rank=rank, nranks=world, nccl_id=shared_id()) as s:
            s.set_factors(W0, H0)
            st = s.iterate(3)
            W, H = s.get_factors()
        # every rank must hold the same full factors
        wt, ht = torch.from_numpy(W).to(dev), torch.from_numpy(H).to(dev)
        w0, h0 = wt.clone(), ht.clone()
        dist.broadcast(w0, 0)
        dist.broadcast(h0, 0)
        same = bool(torch.equal(w0.view(torch.int32), wt.view(torch.int32)) and torch.equal(h0.view(torch.int32), ht.view(torch.int32)))
        flag = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        name = "ccd" if solver == pkg.SOLVER_CCD else "als"
        report[name + "_ranks_agree"] = bool(flag.item())
        ok &= bool(flag.item())
        if rank == 0:
            with pkg.Session(d, p) as s1:
                s1.set_factors(W0, H0)
                st1 = s1.iterate(3)
                W1, H1 = s1.get_factors()
            bit = bool(np.array_equal(W.view(np.uint32), W1.view(np.uint32)) and np.array_equal(H.view(np.uint32), H1.view(np.uint32)))
            report[name + "_bitwise_equal_to_1gpu"] = bit
            report[name + "_rmse"] = [st[-1]["rmse"], st1[-1]["rmse"]]
            report[name + "_max_abs_diff"] = float(max(np.abs(W - W1).max(), np.abs(H - H1).max()))
            ok &= bit
    okt = torch.tensor([int(ok)], device=dev)
    dist.broadcast(okt, 0)
    if rank == 0:
        report["ok"] = bool(okt.item())
        print(json.dumps(report))
    dist.destroy_process_group()
    return 0 if okt.item() else 1


if __name__ == "__main__":
    sys.exit(main())
