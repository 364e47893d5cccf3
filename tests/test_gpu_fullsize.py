"""GPU (-m gpu): BASELINE.json's full sizes, checked through size-independent properties (the CPU oracle would take
minutes per iteration here).  Ratings are generated on the GPU (datagen.py) and handed over as device pointers.

Netflix shape (480 189 x 17 770, 100 M nnz), CCD++ k = 6, T = 2, two outer iterations:
  * fused schedule == reference launch order, bit for bit (factors and both residual copies);
  * the CSR and CSC residual copies hold the same multiset of bit patterns (checksum of checksums: wrap-around
    sums of the float bit patterns and of their squares agree) — u*v == v*u, so they never drift apart;
  * the residual IS R - sum_t u_t v_t^T: recomputed with torch in fp64 on a 2 M-entry sample of the CSR copy;
  * test RMSE decreases monotonically and is finite; empty rows / columns solve to exactly 0.
ML-20M shape, ALS k = 10: RMSE decreases, factors finite, empty segments exactly 0.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device visible")
    return torch


def _bit_checksums(torch, val_np):
    """Order-independent integer checksums of the float bit patterns (exact: no floating-point reduction involved)."""
    v = torch.from_numpy(val_np).cuda().view(torch.int32).to(torch.int64)
    return int(v.sum().item()), int((v * v % 1000003).sum().item()), int(((v >> 9) * 2654435761 % 998244353).sum().item())


def test_netflix_shape_ccdpp_properties(gpu, datagen, torch_cuda):
    torch = torch_cuda
    d = datagen.synth_named("netflix", device="cuda")
    k, lam = 6, 0.05
    W0 = gpu.initial_col(k, d["rows"])
    outs = []
    for schedule in (0, 1):
        with gpu.Session(d, gpu.make_params(k=k, lam=lam, maxinner=2, schedule=schedule)) as s:
            s.set_factors(W0)
            st = s.iterate(2)
            W, H = s.get_factors()
            rv, cv = s.get_values()
            outs.append((W, H, rv, cv, [x["rmse"] for x in st]))
    (W, H, rv, cv, rmse), (W2, H2, rv2, cv2, rmse2) = outs
    # fused == reference launch order, bit for bit
    assert np.array_equal(W, W2) and np.array_equal(H, H2) and np.array_equal(rv, rv2) and np.array_equal(cv, cv2)
    assert rmse == rmse2
    # CSR copy and CSC copy: same multiset of bit patterns
    assert _bit_checksums(torch, rv) == _bit_checksums(torch, cv)
    # the residual is R - W^T H on a sample of the CSR copy (fp64 recomputation; fp32 rounding accumulates per rank)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    e = torch.randint(0, d["nnz"], (2_000_000,), device="cuda", generator=g)
    rows = torch.searchsorted(d["csr_ptr"].long(), e, right=True) - 1
    cols = d["csr_idx"].long()[e]
    Wt, Ht = torch.from_numpy(W).cuda().double(), torch.from_numpy(H).cuda().double()
    want = d["csr_val"][e].double() - (Wt[:, rows] * Ht[:, cols]).sum(0)
    got = torch.from_numpy(rv).cuda()[e].double()
    assert float((got - want).abs().max().item()) < 2e-5
    # RMSE finite and decreasing; empty segments exactly zero
    assert np.all(np.isfinite(rmse)) and rmse[1] < rmse[0] < 1.5
    deg_r = np.diff(d["csr_ptr"].cpu().numpy().astype(np.int64))
    deg_c = np.diff(d["csc_ptr"].cpu().numpy().astype(np.int64))
    assert np.all(W[:, deg_r == 0] == 0.0) and np.all(H[:, deg_c == 0] == 0.0)
    assert np.all(np.isfinite(W)) and np.all(np.isfinite(H))


def test_ml20m_shape_als_properties(gpu, datagen, torch_cuda):
    d = datagen.synth_named("ml20m", device="cuda")
    k = 10
    W0, H0 = gpu.initial_col(d["rows"], k), gpu.initial_col(d["cols"], k)
    with gpu.Session(d, gpu.make_params(gpu.SOLVER_ALS, k=k, lam=0.05)) as s:
        s.set_factors(W0, H0)
        st = s.iterate(3)
        W, H = s.get_factors()
    rmse = [x["rmse"] for x in st]
    assert np.all(np.isfinite(rmse)) and rmse[2] < rmse[1] < rmse[0]
    assert np.all(np.isfinite(W)) and np.all(np.isfinite(H))
    deg_r = np.diff(d["csr_ptr"].cpu().numpy().astype(np.int64))
    deg_c = np.diff(d["csc_ptr"].cpu().numpy().astype(np.int64))
    assert np.all(W[deg_r == 0] == 0.0) and np.all(H[deg_c == 0] == 0.0)
