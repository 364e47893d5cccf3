#!/usr/bin/env python
"""Run under torch.distributed.run with N >= 2 ranks: several CCD++ sessions IN A ROW on ONE NCCL unique id — the second
adopts the peer-to-peer state (exported receive buffers, CUDA IPC mappings, exchange epoch) cached with the communicator,
the third has another shape (every rank drops the stale state behind a barrier and sets up a new one), the fourth goes
back to the first shape — each checked bit for bit against a single-GPU session on rank 0.  One JSON line on rank 0."""
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
    pkg = load_package()
    import cuda_recommender_b200.datagen as dg
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t.copy_(torch.tensor(list(pkg.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    nccl_id = bytes(t.cpu().tolist())
    data = {name: dg.to_numpy(dg.synth_named(name, seed=5)) for name in ("ml100k", "small")}
    k = 5
    ok, seq = True, []
    for step, name in enumerate(("ml100k", "ml100k", "small", "ml100k")):
        d = data[name]
        p = pkg.make_params(pkg.SOLVER_CCD, k=k, lam=0.05, maxinner=2, device=local)
        W0 = pkg.initial_col(k, d["rows"])
        with pkg.Session(d, p, rank=rank, nranks=world, nccl_id=nccl_id) as s:
            s.set_factors(W0)
            st = s.iterate(2 + step % 2)
            W, H = s.get_factors()
        same = True
        if rank == 0:
            with pkg.Session(d, p) as s1:
                s1.set_factors(W0)
                s1.iterate(2 + step % 2)
                W1, H1 = s1.get_factors()
            same = bool(np.array_equal(W, W1) and np.array_equal(H, H1))
        flag = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        seq.append(bool(flag.item()))
        ok &= bool(flag.item())
    pkg.release_cached_memory(local)
    dist.barrier()
    if rank == 0:
        print(json.dumps({"ok": ok, "sessions_bitwise_equal_to_1gpu": seq, "world": world}))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
