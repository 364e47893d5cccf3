"""GPU (-m gpu): mf_session_predict — predictions of the current factors for arbitrary pairs, BIT-EXACT against the
oracle's restatement of src/extras.cpp:165-168 (FP32 products, FP64 sum in rank order) on the factors the session
holds, for both factor layouts (CCD++ k x rows, ALS rows x k)."""
import numpy as np
import pytest

from conftest import sides

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("solver,k", [("ccd", 10), ("ccd", 1), ("als", 10), ("als", 24)])
def test_predict_bitwise(gpu, port, data_factory, solver, k):
    d = data_factory("small")
    _, _, (trow, tcol, tval) = sides(d)
    rng = np.random.default_rng(7)
    row = np.concatenate([trow, rng.integers(0, d["rows"], 5000).astype(np.uint32)])
    col = np.concatenate([tcol, rng.integers(0, d["cols"], 5000).astype(np.uint32)])
    als = solver == "als"
    prm = gpu.make_params(gpu.SOLVER_ALS if als else gpu.SOLVER_CCD, k=k, lam=0.05, maxinner=2)
    with gpu.Session(d, prm) as s:
        if als:
            s.set_factors(port.initial_col(d["rows"], k), port.initial_col(d["cols"], k))
        else:
            s.set_factors(port.initial_col(k, d["rows"]))
        st = s.iterate(2)
        W, H = s.get_factors()
        got = s.predict(row, col)
        assert got.dtype == np.float64 and got.shape == row.shape
        want = port.predict(row, col, W, H, d["rows"], d["cols"], k, als)
        assert np.array_equal(got, want)
        # the session's own RMSE is the RMSE of these predictions on the test pairs
        err = got[: len(tval)] - tval.astype(np.float64)
        assert np.sqrt(np.mean(err * err)) == pytest.approx(st[-1]["rmse"], abs=1e-9)
        assert s.predict(np.zeros(0, np.uint32), np.zeros(0, np.uint32)).shape == (0,)
