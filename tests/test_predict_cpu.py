"""CPU: the oracle's prediction op (oracle/mf_oracle.c orc_predict, restating src/extras.cpp:165-168 / the dot() of
src/tools.cpp:184-198) against plain numpy of the same arithmetic, and against the live reference: the test RMSE the
reference prints is the RMSE of these predictions."""
import math

import numpy as np
import pytest

from conftest import sides


def _numpy_predict(row, col, W, H, als_layout):
    """FP32 products, summed in rank order in FP64 (cumsum is a sequential sum)."""
    if als_layout:
        prod = W[row.astype(np.int64), :] * H[col.astype(np.int64), :]
    else:
        prod = (W[:, row.astype(np.int64)] * H[:, col.astype(np.int64)]).T
    assert prod.dtype == np.float32
    return np.cumsum(prod.astype(np.float64), axis=1)[:, -1]


@pytest.mark.parametrize("als_layout", [0, 1])
@pytest.mark.parametrize("k", [1, 7, 40])
def test_predict_is_fp32_products_summed_in_fp64(port, als_layout, k):
    rng = np.random.default_rng(3 + k)
    rows, cols, n = 57, 91, 400
    W = rng.standard_normal((rows, k) if als_layout else (k, rows)).astype(np.float32)
    H = rng.standard_normal((cols, k) if als_layout else (k, cols)).astype(np.float32)
    row = rng.integers(0, rows, n).astype(np.uint32)
    col = rng.integers(0, cols, n).astype(np.uint32)
    got = port.predict(row, col, W, H, rows, cols, k, als_layout)
    assert got.dtype == np.float64 and np.array_equal(got, _numpy_predict(row, col, W, H, als_layout))


def test_predict_empty(port):
    z = np.zeros(0, np.uint32)
    assert port.predict(z, z, np.ones((2, 3), np.float32), np.ones((2, 4), np.float32), 3, 4, 2, 0).shape == (0,)


@pytest.mark.parametrize("als", [0, 1])
def test_rmse_of_predictions_is_the_reference_rmse(als, port, ref, datagen, data_factory, tmp_path):
    d = data_factory("small", seed=31)
    datagen.write_dataset(str(tmp_path), d)
    k = 6
    r = ref.train(str(tmp_path), als, k, 0.05, 2, 2, threads=2)
    _, _, (trow, tcol, tval) = sides(d)
    pred = port.predict(trow, tcol, r["W"], r["H"], d["rows"], d["cols"], k, als)
    acc = 0.0
    for p, v in zip(pred.tolist(), tval.tolist()):  # calrmse's own order: err = -v; err += pred; acc += err*err
        err = -v
        err += p
        acc += err * err
    assert math.sqrt(acc / len(tval)) == pytest.approx(r["rmse"], abs=1e-12)
