"""Synthetic ratings of MovieLens / Netflix / Yahoo-Music shape and the
reference's on-disk dataset format (reader: src/tools.cpp:3-85,
src/pmf_util.h:108-136,171-196).

The recipe: unique (row, col) pairs drawn with log-normal row and column popularity,
values from a planted rank-8 model with geometrically decaying component scales plus
N(0, 0.5^2) noise, rounded and clipped to 1..5, held-out test pairs from the same
distribution.  One torch implementation runs on CPU (tests, small shapes) and on the
GPU (bench shapes: 100 M+ nnz in a few seconds); the value of a pair is a pure function
of (row, col, seed), so the CSR and CSC copies are valued independently and agree.

torch is plumbing here (device memory, sort/unique for the synthetic input); none of it
is on the measured path.
"""
import math
import os

import numpy as np
import torch

# name -> (rows, cols, nnz, nnz_test)   BASELINE.json configs / SURVEY.md §8d
SHAPES = {
    "ml100k": (943, 1682, 100_000, 10_000),
    "ml20m": (138_493, 26_744, 20_000_000, 200_000),
    "netflix": (480_189, 17_770, 100_000_000, 1_400_000),
    "yahoo": (1_000_990, 624_961, 252_800_275, 4_003_960),
    # small shapes for CPU-side tests
    "tiny": (60, 90, 1_500, 200),
    "small": (300, 500, 12_000, 1_000),
}

_M64 = (1 << 64) - 1


def _wrap64(x):
    """python int -> the int64 with the same low 64 bits (torch has no uint64 math)."""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _mix64(x):
    """splitmix64 finaliser on int64 tensors (wrap-around arithmetic)."""
    x = (x ^ ((x >> 30) & 0x3FFFFFFFF)) * _wrap64(0xBF58476D1CE4E5B9)
    x = (x ^ ((x >> 27) & 0x1FFFFFFFFF)) * _wrap64(0x94D049BB133111EB)
    return x ^ ((x >> 31) & 0x1FFFFFFFF)


def _uniform01(h):
    """int64 hash -> float64 uniform in (0, 1)."""
    return (((h >> 11) & ((1 << 53) - 1)).to(torch.float64) + 0.5) * (1.0 / (1 << 53))


class _Planted:
    def __init__(self, rows, cols, seed, device, rank=8, decay=0.7, noise=0.5):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed * 7919 + 17)
        scale = torch.tensor([decay ** d for d in range(rank)], dtype=torch.float32)
        scale = scale * math.sqrt(1.0 / float((scale ** 2).sum()))
        self.U = (torch.randn(rows, rank, generator=g) * scale).to(device)
        self.V = torch.randn(cols, rank, generator=g).to(device)
        self.seed = seed
        self.noise = noise
        self.cols = cols

    def values(self, r, c, chunk=1 << 24):
        out = torch.empty(r.numel(), dtype=torch.float32, device=r.device)
        for lo in range(0, r.numel(), chunk):
            rr = r[lo:lo + chunk].long()
            cc = c[lo:lo + chunk].long()
            dot = (self.U[rr] * self.V[cc]).sum(1)
            key = rr * self.cols + cc
            h1 = _mix64(key + _wrap64(0x9E3779B97F4A7C15 * (self.seed + 1)))
            h2 = _mix64(h1 + _wrap64(0xD1B54A32D192ED03))
            z = torch.sqrt(-2.0 * torch.log(_uniform01(h1))) * torch.cos(2.0 * math.pi * _uniform01(h2))
            val = 3.5 + dot + self.noise * z.to(torch.float32)
            out[lo:lo + chunk] = torch.clamp(torch.round(val), 1.0, 5.0)
        return out


def _draw_keys(rows, cols, need, seed, device, sigma_row, sigma_col):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    wr = torch.exp(sigma_row * torch.randn(rows, generator=g, device=device, dtype=torch.float64))
    wc = torch.exp(sigma_col * torch.randn(cols, generator=g, device=device, dtype=torch.float64))
    cdf_r = torch.cumsum(wr / wr.sum(), 0)
    cdf_c = torch.cumsum(wc / wc.sum(), 0)
    if need > rows * cols:
        raise ValueError("more ratings requested than cells")
    keys = torch.empty(0, dtype=torch.int64, device=device)
    while keys.numel() < need:
        draw = int((need - keys.numel()) * 1.25) + 4096
        r = torch.searchsorted(cdf_r, torch.rand(draw, generator=g, device=device, dtype=torch.float64)).clamp_(max=rows - 1)
        c = torch.searchsorted(cdf_c, torch.rand(draw, generator=g, device=device, dtype=torch.float64)).clamp_(max=cols - 1)
        keys = torch.unique(torch.cat([keys, r * cols + c]))
        del r, c
    perm = torch.randperm(keys.numel(), generator=g, device=device)[:need]
    return keys[perm]


def synth_ratings(rows, cols, nnz, nnz_test, seed=1, device="cpu", sigma_row=1.05, sigma_col=1.5):
    """Returns a dict of torch tensors on `device`:
         csr_ptr int32[rows+1], csr_idx int32[nnz], csr_val f32[nnz]   sorted by (row, col)
         csc_ptr int32[cols+1], csc_idx int32[nnz], csc_val f32[nnz]   sorted by (col, row)
         coo_row/coo_col/coo_val (train, in (row, col) order)
         test_row int32[nt], test_col int32[nt], test_val f32[nt]
       plus rows, cols, nnz, nnz_test.  int32 bit patterns are the uint32 the reference uses."""
    device = torch.device(device)
    keys = _draw_keys(rows, cols, nnz + nnz_test, seed, device, sigma_row, sigma_col)
    train = torch.sort(keys[:nnz]).values
    test = keys[nnz:nnz + nnz_test]
    del keys
    model = _Planted(rows, cols, seed, device)

    r = torch.div(train, cols, rounding_mode="floor")
    c = train - r * cols
    del train
    out = dict(rows=rows, cols=cols, nnz=nnz, nnz_test=nnz_test)
    ptr = torch.zeros(rows + 1, dtype=torch.int64, device=device)
    ptr[1:] = torch.cumsum(torch.bincount(r, minlength=rows), 0)
    out["csr_ptr"] = ptr.to(torch.int32)
    out["csr_idx"] = c.to(torch.int32)
    out["csr_val"] = model.values(r, c)
    out["coo_row"] = r.to(torch.int32)
    out["coo_col"] = out["csr_idx"]
    out["coo_val"] = out["csr_val"]

    k2 = torch.sort(c * rows + r).values
    del r, c
    c2 = torch.div(k2, rows, rounding_mode="floor")
    r2 = k2 - c2 * rows
    del k2
    ptr = torch.zeros(cols + 1, dtype=torch.int64, device=device)
    ptr[1:] = torch.cumsum(torch.bincount(c2, minlength=cols), 0)
    out["csc_ptr"] = ptr.to(torch.int32)
    out["csc_idx"] = r2.to(torch.int32)
    out["csc_val"] = model.values(r2, c2)
    del r2, c2

    tr = torch.div(test, cols, rounding_mode="floor")
    tc = test - tr * cols
    out["test_row"] = tr.to(torch.int32)
    out["test_col"] = tc.to(torch.int32)
    out["test_val"] = model.values(tr, tc)
    return out


def synth_named(name, seed=None, device="cpu"):
    rows, cols, nnz, nt = SHAPES[name]
    if seed is None:
        seed = 1 + list(SHAPES).index(name)
    return synth_ratings(rows, cols, nnz, nt, seed=seed, device=device)


def to_numpy(data):
    """torch dict -> numpy dict; index arrays become uint32 views."""
    out = {}
    for key, v in data.items():
        if torch.is_tensor(v):
            a = v.detach().cpu().numpy()
            out[key] = a.view(np.uint32) if a.dtype == np.int32 else a
        else:
            out[key] = v
    return out


def from_coo(rows, cols, r, c, v, test=None):
    """numpy COO (unique pairs, any order) -> numpy dataset dict with CSR sorted by
    (row, col) and CSC by (col, row).  Host-side convenience for tests (numpy lexsort)."""
    r = np.asarray(r, np.int64); c = np.asarray(c, np.int64); v = np.asarray(v, np.float32)
    o = np.lexsort((c, r))
    csr_ptr = np.zeros(rows + 1, np.int64); np.add.at(csr_ptr, r + 1, 1); csr_ptr = np.cumsum(csr_ptr)
    o2 = np.lexsort((r, c))
    csc_ptr = np.zeros(cols + 1, np.int64); np.add.at(csc_ptr, c + 1, 1); csc_ptr = np.cumsum(csc_ptr)
    d = dict(rows=rows, cols=cols, nnz=len(v),
             csr_ptr=csr_ptr.astype(np.uint32), csr_idx=c[o].astype(np.uint32), csr_val=v[o].copy(),
             csc_ptr=csc_ptr.astype(np.uint32), csc_idx=r[o2].astype(np.uint32), csc_val=v[o2].copy(),
             coo_row=r[o].astype(np.uint32), coo_col=c[o].astype(np.uint32), coo_val=v[o].copy())
    if test is None:
        test = (np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.float32))
    d["test_row"], d["test_col"], d["test_val"] = (np.asarray(test[0], np.uint32), np.asarray(test[1], np.uint32),
                                                   np.asarray(test[2], np.float32))
    d["nnz_test"] = len(d["test_val"])
    return d


# ---------------------------------------------------------------------------------------
# on-disk format
# ---------------------------------------------------------------------------------------
_FILES = dict(csr_ptr="R_train_csr.indptr.bin", csr_idx="R_train_csr.indices.bin", csr_val="R_train_csr.data.bin",
              csc_ptr="R_train_csc.indptr.bin", csc_idx="R_train_csc.indices.bin", csc_val="R_train_csc.data.bin",
              test_val="R_test_coo.data.bin", test_row="R_test_coo.row.bin", test_col="R_test_coo.col.bin")


def write_dataset(dirname, data, nnz_test_limit=None):
    """Writes `data` (numpy or torch dict) as a directory the reference binary, the
    reference harness (oracle/_ref) and this repo's CLI all read: `meta`,
    `meta_modified_all`, little-endian headerless binaries (ptr int32, idx uint32,
    val float32).  nnz_test_limit truncates the test set (used when timing the
    reference: its per-iteration RMSE is serial, src/tools.cpp:235-248)."""
    d = to_numpy(data) if any(torch.is_tensor(v) for v in data.values()) else data
    os.makedirs(dirname, exist_ok=True)
    nt = int(d["nnz_test"]) if nnz_test_limit is None else min(int(d["nnz_test"]), int(nnz_test_limit))
    for key, fname in _FILES.items():
        a = np.ascontiguousarray(d[key])
        if key.startswith("test_"):
            a = a[:nt]
        if key.endswith("_val"):
            a = a.astype(np.float32, copy=False)
        elif key.endswith("_ptr"):
            a = a.astype(np.int32, copy=False) if a.dtype != np.uint32 else a.view(np.int32)
        else:
            a = a.astype(np.uint32, copy=False)
        a.tofile(os.path.join(dirname, fname))
    with open(os.path.join(dirname, "meta_modified_all"), "w") as f:
        f.write(f"{d['rows']} {d['cols']} {d['nnz']}\n")
        f.write("R_train_coo.data.bin R_train_coo.row.bin R_train_coo.col.bin\n")  # parsed, never opened (tools.cpp:30-35)
        f.write(f"{_FILES['csr_ptr']} {_FILES['csr_idx']} {_FILES['csr_val']}\n")
        f.write(f"{_FILES['csc_ptr']} {_FILES['csc_idx']} {_FILES['csc_val']}\n")
        f.write(f"{nt}\n")
        f.write(f"{_FILES['test_val']} {_FILES['test_row']} {_FILES['test_col']}\n")
    with open(os.path.join(dirname, "meta"), "w") as f:  # extras.cpp:24-44
        f.write(f"{d['rows']} {d['cols']}\n{d['nnz']} train.ratings\n{nt} test.ratings\n")
    open(os.path.join(dirname, "test.ratings"), "w").close()  # opened "r", never parsed (extras.cpp:5-9)
    return dirname


def read_dataset(dirname):
    """Reads a directory written by write_dataset (or for the reference) -> numpy dict."""
    tok = open(os.path.join(dirname, "meta_modified_all")).read().split()
    rows, cols, nnz = int(tok[0]), int(tok[1]), int(tok[2])
    names = tok[3:12]
    nt = int(tok[12])
    tnames = tok[13:16]
    p = lambda n: os.path.join(dirname, n)
    d = dict(rows=rows, cols=cols, nnz=nnz, nnz_test=nt)
    d["csr_ptr"] = np.fromfile(p(names[3]), np.int32, rows + 1).view(np.uint32)
    d["csr_idx"] = np.fromfile(p(names[4]), np.uint32, nnz)
    d["csr_val"] = np.fromfile(p(names[5]), np.float32, nnz)
    d["csc_ptr"] = np.fromfile(p(names[6]), np.int32, cols + 1).view(np.uint32)
    d["csc_idx"] = np.fromfile(p(names[7]), np.uint32, nnz)
    d["csc_val"] = np.fromfile(p(names[8]), np.float32, nnz)
    d["test_val"] = np.fromfile(p(tnames[0]), np.float32, nt)
    d["test_row"] = np.fromfile(p(tnames[1]), np.uint32, nt)
    d["test_col"] = np.fromfile(p(tnames[2]), np.uint32, nt)
    return d
