"""CPU, world_size 2 over gloo: the multi-GPU sharding scheme of the CUDA path (session.cu / dist.cu),
replayed with the oracle's step functions, one process per rank:
  rank r keeps CSR row block r and CSC column block r (nnz-balanced bounds, same rule as mf_partition),
  solves v on its columns -> all-gather v -> solves u on its rows -> all-gather u, and applies residual
  updates to its own blocks only.
Claim checked: no summation order changes, so the sharded run equals the single-process oracle BIT FOR
BIT (factors on every rank, and each rank's residual blocks)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _allgather_blocks(vec, bound, rank):
    """in-place all-gather of unequal blocks: grouped broadcasts, one per owner (dist.cu does the same with NCCL)"""
    for r in range(len(bound) - 1):
        lo, hi = int(bound[r]), int(bound[r + 1])
        if hi > lo:
            t = torch.from_numpy(vec[lo:hi].copy()) if r == rank else torch.empty(hi - lo, dtype=torch.float32)
            dist.broadcast(t, r)
            vec[lo:hi] = t.numpy()


def _block(ptr, idx, val, lo, hi):
    e0, e1 = int(ptr[lo]), int(ptr[hi])
    return (ptr[lo:hi + 1] - ptr[lo]).astype(np.uint32), idx[e0:e1].copy(), val[e0:e1].copy(), e0, e1


def _worker(rank, world, port, k, lam, iters, inner, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from __graft_entry__ import load_package
    load_package()
    import cuda_recommender_b200.datagen as dg
    from oracle import port as orc
    d = dg.to_numpy(dg.synth_named("small", seed=77))
    rows, cols = d["rows"], d["cols"]
    rb, cb = orc.partition(d["csr_ptr"], world), orc.partition(d["csc_ptr"], world)
    rptr, ridx, rval, r0, r1 = _block(d["csr_ptr"], d["csr_idx"], d["csr_val"], rb[rank], rb[rank + 1])
    cptr, cidx, cval, c0, c1 = _block(d["csc_ptr"], d["csc_idx"], d["csc_val"], cb[rank], cb[rank + 1])
    W = orc.initial_col(k, rows)
    H = np.zeros((k, cols), np.float32)
    for oiter in range(iters):
        for t in range(k):
            u, v = W[t], H[t]
            if oiter > 0:
                cval = orc.ccd_update_sweep(cptr, cidx, cval, u, v[cb[rank]:cb[rank + 1]], add=True)
                rval = orc.ccd_update_sweep(rptr, ridx, rval, v, u[rb[rank]:rb[rank + 1]], add=True)
            for _ in range(inner):
                v[cb[rank]:cb[rank + 1]] = orc.ccd_solve_sweep(cptr, cidx, cval, u, lam)
                _allgather_blocks(v, cb, rank)
                u[rb[rank]:rb[rank + 1]] = orc.ccd_solve_sweep(rptr, ridx, rval, v, lam)
                _allgather_blocks(u, rb, rank)
            cval = orc.ccd_update_sweep(cptr, cidx, cval, u, v[cb[rank]:cb[rank + 1]], add=False)
            rval = orc.ccd_update_sweep(rptr, ridx, rval, v, u[rb[rank]:rb[rank + 1]], add=False)
    # timing plumbing of bench.py: max over ranks
    tmax = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), W=W, H=H, rval=rval, cval=cval, r=np.array([r0, r1]), c=np.array([c0, c1]),
             tmax=tmax.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_ccdpp_equals_single_process_bitwise(tmp_path, port, datagen):
    world, k, lam, iters, inner = 2, 4, 0.05, 2, 2
    mp.spawn(_worker, args=(world, _free_port(), k, lam, iters, inner, str(tmp_path)), nprocs=world, join=True)
    d = datagen.to_numpy(datagen.synth_named("small", seed=77))
    full = port.ccdpp(d["rows"], d["cols"], (d["csr_ptr"], d["csr_idx"], d["csr_val"]), (d["csc_ptr"], d["csc_idx"], d["csc_val"]),
                      port.initial_col(k, d["rows"]), k, lam, iters, inner)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["W"], full["W"]) and np.array_equal(z["H"], full["H"])
        assert np.array_equal(z["rval"], full["csr_val"][z["r"][0]:z["r"][1]])
        assert np.array_equal(z["cval"], full["csc_val"][z["c"][0]:z["c"][1]])
        assert z["tmax"][0] == world
