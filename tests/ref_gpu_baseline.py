#!/usr/bin/env python
"""Secondary baselines on the GPU box, both through the reference's UNMODIFIED main.cpp on a dataset directory in the
reference's on-disk format (so the rating arrays are the loader's ordinary pageable `new[]` memory):
  * oracle/_ref/cuda_andre_refgpu  — the reference's own CUDA path (cuda_src/*.cu) rebuilt for sm_100a: GPU vs GPU
  * oracle/_ref/cuda_andre_dropin  — the same main.cpp linked against this repo's shim + libmfb200.so
Prints one JSON line with the per-iteration times each binary reports and its "CUDA Training time".
Test infrastructure (it executes binaries under oracle/_ref): kept under tests/.
Usage: tests/ref_gpu_baseline.py [shape=netflix] [k=40] [outer=3] [inner=3]"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def run(binary, args, timeout):
    t0 = time.time()
    out = subprocess.run([binary] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout).stdout
    it = re.findall(r"\[-INFO-\] iteration num (\d+) \trank_time ([\d.]+)\|[\d.]+ s \tupdate_time ([\d.]+)\|[\d.]+s \tRMSE=([\d.]+)", out)
    train = re.search(r"CUDA Training time: ([\d.]+) s", out)
    return {"iterations": [{"rank_time": float(a), "update_time": float(b), "rmse": float(c)} for _, a, b, c in it],
            "cuda_training_time_s": float(train.group(1)) if train else None, "wall_s": round(time.time() - t0, 2),
            "tail": out[-400:] if not it else None,
            "trace": [ln for ln in out.splitlines() if ln.startswith("[mf trace]")] if os.environ.get("MF_TRACE") else None}


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "netflix"
    k, outer, inner = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((2, 40), (3, 3), (4, 3)))
    load_package()
    import cuda_recommender_b200.datagen as dg
    import torch
    seed = 1 + list(dg.SHAPES).index(shape)
    data = dg.synth_named(shape, seed=seed, device=torch.device("cuda", 0))
    tmp = tempfile.mkdtemp(prefix="mf_refgpu_", dir=os.environ.get("TMPDIR", "/tmp"))
    dg.write_dataset(tmp, data, nnz_test_limit=100000)
    del data
    torch.cuda.empty_cache()
    args = ["-CUDA", "-k", str(k), "-l", "0.05", "-t", str(outer), "-T", str(inner), "-n", "16", tmp]
    rep = {"shape": shape, "k": k, "outer": outer, "inner": inner}
    names = ("cuda_andre_dropin", "cuda_andre_refgpu") if not os.environ.get("MF_ONLY_DROPIN") else ("cuda_andre_dropin", "cuda_andre_dropin")
    for name in names:
        b = os.path.join(ROOT, "oracle", "_ref", name)
        rep[name] = run(b, args, 900) if os.path.exists(b) else None
    print(json.dumps(rep))
    subprocess.run(["rm", "-rf", tmp])


if __name__ == "__main__":
    main()
