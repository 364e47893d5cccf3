"""CPU: the oracle restatement against the live reference library (oracle/_ref/libmfref.so, the
unmodified /root/reference/src sources) on freshly generated inputs, including the reference's own
loader reading the on-disk format this repo writes."""
import numpy as np
import pytest

from conftest import sides


def test_initial_col_same_libc_sequence(port, ref):
    for k, n in [(10, 943), (943, 10), (1, 7), (5, 1)]:
        assert np.array_equal(port.initial_col(k, n), ref.initial_col(k, n))


def test_reference_loader_reads_our_dataset(ref, datagen, data_factory, tmp_path):
    d = data_factory("small")
    datagen.write_dataset(str(tmp_path), d)
    info = ref.probe(str(tmp_path))
    assert (info["rows"], info["cols"], info["nnz"], info["nnz_test"]) == (d["rows"], d["cols"], d["nnz"], d["nnz_test"])
    assert info["max_row_nnz"] == int(np.diff(d["csr_ptr"].astype(np.int64)).max())
    assert info["max_col_nnz"] == int(np.diff(d["csc_ptr"].astype(np.int64)).max())
    back = datagen.read_dataset(str(tmp_path))
    for key in ("csr_ptr", "csr_idx", "csr_val", "csc_ptr", "csc_idx", "csc_val", "test_row", "test_col", "test_val"):
        assert np.array_equal(back[key], d[key]), key


@pytest.mark.parametrize("shape,k,lam,iters,inner,threads", [("small", 7, 0.05, 2, 3, 1), ("small", 3, 0.2, 3, 1, 4), ("tiny", 12, 0.01, 2, 2, 2)])
def test_ccdpp_bitwise_vs_reference(shape, k, lam, iters, inner, threads, port, ref, datagen, data_factory, tmp_path):
    d = data_factory(shape, seed=100 + k)
    datagen.write_dataset(str(tmp_path), d)
    r = ref.train(str(tmp_path), 0, k, lam, iters, inner, threads=threads, want_residual=True)
    csr, csc, test = sides(d)
    o = port.ccdpp(d["rows"], d["cols"], csr, csc, port.initial_col(k, d["rows"]), k, lam, iters, inner, test=test, threads=threads)
    assert np.array_equal(o["W"], r["W"]) and np.array_equal(o["H"], r["H"])
    assert np.array_equal(o["csr_val"], r["csr_val"]) and np.array_equal(o["csc_val"], r["csc_val"])
    assert o["rmse"][-1] == pytest.approx(r["rmse"], abs=1e-12)
    assert len(r["iters"]) == iters  # one "[-INFO-] iteration num" line per outer iteration


@pytest.mark.parametrize("shape,k,lam,iters", [("small", 5, 0.05, 2), ("tiny", 16, 0.1, 2)])
def test_als_bitwise_vs_reference(shape, k, lam, iters, port, ref, datagen, data_factory, tmp_path):
    d = data_factory(shape, seed=200 + k)
    datagen.write_dataset(str(tmp_path), d)
    r = ref.train(str(tmp_path), 1, k, lam, iters, threads=3)
    csr, csc, test = sides(d)
    o = port.als(d["rows"], d["cols"], csr, csc, port.initial_col(d["rows"], k), port.initial_col(d["cols"], k), k, lam, iters, test=test)
    assert np.array_equal(o["W"], r["W"]) and np.array_equal(o["H"], r["H"])
    assert o["rmse"][-1] == pytest.approx(r["rmse"], abs=1e-12)


def test_step_functions_compose_to_the_driver(port, data_factory):
    """The step-level oracle entry points (what the GPU step-parity tests compare against) are the
    same arithmetic as the full driver: replaying one rank by hand reproduces orc_ccdpp bit for bit."""
    d = data_factory("tiny", seed=5)
    csr, csc, _ = sides(d)
    k, lam = 3, 0.05
    W0 = port.initial_col(k, d["rows"])
    full = port.ccdpp(d["rows"], d["cols"], csr, csc, W0, k, lam, 1, 2)
    W, H = W0.copy(), np.zeros((k, d["cols"]), np.float32)
    cv, rv = csc[2].copy(), csr[2].copy()
    for t in range(k):
        for _ in range(2):
            H[t] = port.ccd_solve_sweep(csc[0], csc[1], cv, W[t], lam)
            W[t] = port.ccd_solve_sweep(csr[0], csr[1], rv, H[t], lam)
        cv = port.ccd_update_sweep(csc[0], csc[1], cv, W[t], H[t], add=False)
        rv = port.ccd_update_sweep(csr[0], csr[1], rv, H[t], W[t], add=False)
    assert np.array_equal(W, full["W"]) and np.array_equal(H, full["H"])
    assert np.array_equal(cv, full["csc_val"]) and np.array_equal(rv, full["csr_val"])
