"""CPU: the oracle restatement (oracle/mf_oracle.c) reproduces, bit for bit, what the unmodified
reference CPU path produced for the committed fixtures (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import GOLDEN_ALS, GOLDEN_CCD, sides


@pytest.mark.parametrize("name", GOLDEN_CCD)
def test_ccdpp_matches_reference_fixture(name, golden, port):
    d, z = golden(name)
    csr, csc, test = sides(d)
    k, lam, it, inner = int(z["k"]), float(z["lam"]), int(z["maxiter"]), int(z["maxinner"])
    W0 = port.initial_col(k, d["rows"])
    out = port.ccdpp(d["rows"], d["cols"], csr, csc, W0, k, lam, it, inner, test=test)
    assert np.array_equal(out["W"], z["W"]) and np.array_equal(out["H"], z["H"])
    assert np.array_equal(out["csr_val"], z["csr_resid"]) and np.array_equal(out["csc_val"], z["csc_resid"])
    assert out["rmse"][-1] == pytest.approx(float(z["rmse_final"]), abs=1e-12)
    # the reference prints RMSE with 6 decimals (src/CCD.cpp:158)
    assert np.allclose(out["rmse"], z["rmse_printed"], atol=5.1e-7, rtol=0)
    one = port.ccdpp(d["rows"], d["cols"], csr, csc, W0, k, lam, 1, inner, test=test)
    assert np.array_equal(one["W"], z["W_iter1"]) and np.array_equal(one["H"], z["H_iter1"])


@pytest.mark.parametrize("name", GOLDEN_ALS)
def test_als_matches_reference_fixture(name, golden, port):
    d, z = golden(name)
    csr, csc, test = sides(d)
    k, lam, it = int(z["k"]), float(z["lam"]), int(z["maxiter"])
    W0, H0 = port.initial_col(d["rows"], k), port.initial_col(d["cols"], k)  # main.cpp:86-87 argument order
    out = port.als(d["rows"], d["cols"], csr, csc, W0, H0, k, lam, it, test=test)
    assert np.array_equal(out["W"], z["W"]) and np.array_equal(out["H"], z["H"])
    assert np.allclose(out["rmse"], z["rmse_printed"], atol=5.1e-7, rtol=0)
    assert out["bad_pivots"] == 0


def test_f64_yardstick_is_close_to_reference(golden, port):
    """Appendix D of SURVEY.md: reference-FP32 vs FP64 accumulation after one outer iteration on the
    C1 shape stays below 1e-4 in Frobenius norm — the size of the tolerance the GPU is held to."""
    d, z = golden("ccd_c1_ml100k")
    csr, csc, test = sides(d)
    W0 = port.initial_col(10, d["rows"])
    hi = port.ccdpp(d["rows"], d["cols"], csr, csc, W0, 10, 0.05, 1, 3, test=test, f64acc=True)
    relW = np.linalg.norm(hi["W"] - z["W_iter1"]) / np.linalg.norm(z["W_iter1"])
    relH = np.linalg.norm(hi["H"] - z["H_iter1"]) / np.linalg.norm(z["H_iter1"])
    assert relW < 1e-4 and relH < 1e-4


def test_ccdpp_ex_with_options_off_is_ccdpp(port, data_factory):
    """orc_ccdpp_ex (the checker of the §8 f4 options) with every option off is the pinned restatement, bit for bit; its
    per-rank incremental RMSE ends every outer iteration on the iteration's RMSE (calrmse_r1 vs calrmse: FP32 drift only)."""
    from conftest import sides
    d = data_factory("ml100k")
    csr, csc, test = sides(d)
    k = 4
    W0 = port.initial_col(k, d["rows"])
    a = port.ccdpp(d["rows"], d["cols"], csr, csc, W0, k, 0.05, 3, 2, test=test)
    b = port.ccdpp_ex(d["rows"], d["cols"], csr, csc, W0, k, 0.05, 3, 2, test=test)
    for key in ("W", "H", "rmse", "csr_val", "csc_val"):
        assert np.array_equal(a[key], b[key])
    assert np.abs(b["rank_rmse"][:, -1] - b["rmse"]).max() < 1e-5
    assert (b["inner_done"] == 2).all()
    c = port.ccdpp_ex(d["rows"], d["cols"], csr, csc, W0, k, 0.05, 3, 6, test=test, early_stop=True, eps=1e-2)
    assert c["inner_done"].sum() < 3 * k * 6 and c["inner_done"].min() >= 1
    n = port.ccdpp_ex(d["rows"], d["cols"], csr, csc, W0 - 0.05, k, 0.05, 2, 2, test=test, nmf=True)
    assert (n["W"] >= 0).all() and (n["H"] >= 0).all()
