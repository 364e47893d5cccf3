"""GPU (-m gpu): CCD++ parity of the CUDA path (through the C-ABI) with the oracle.

Tolerances (BASELINE.json north_star):
  residual updates                bit-exact (one rounded product + one rounded sum per entry)
  one solve sweep, same inputs    relative l2 <= 5e-5 and |d| <= 1e-4*|x| + 1e-6*max|x| per entry
  factors after 1 outer iteration relative l2 per factor matrix <= 1e-4 (config C1)
  test RMSE, every iteration      |d| <= 1e-4 absolute
"""
import numpy as np
import pytest

from conftest import GOLDEN_CCD, rel_l2, sides

pytestmark = pytest.mark.gpu

LAYOUTS = [0, 1]  # PANEL, DIRECT


def _check_sweep(got, want):
    assert rel_l2(got, want) <= 5e-5
    tol = 1e-4 * np.abs(want) + 1e-6 * np.abs(want).max()
    assert np.all(np.abs(got - want) <= tol)


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("shape,panel_rows,chunk", [("ml100k", 0, 0), ("small", 64, 16), ("tiny", 0, 0)])
def test_solve_step_parity(gpu, port, data_factory, layout, shape, panel_rows, chunk):
    d = data_factory(shape)
    csr, csc, _ = sides(d)
    k, lam = 3, 0.05
    rng = np.random.default_rng(3)
    W = (rng.random((k, d["rows"])) * 0.5 + 0.01).astype(np.float32)
    H = (rng.standard_normal((k, d["cols"])) * 0.3).astype(np.float32)
    with gpu.Session(d, gpu.make_params(k=k, lam=lam, layout=layout, panel_rows=panel_rows, chunk=chunk)) as s:
        s.set_factors(W, H)
        s.ccd_solve(1, gpu.SIDE_CSC)  # v = H[1] from u = W[1]
        W1, H1 = s.get_factors()
        want_v = port.ccd_solve_sweep(csc[0], csc[1], csc[2], W[1], lam)
        _check_sweep(H1[1], want_v)
        assert np.array_equal(W1, W) and np.array_equal(H1[[0, 2]], H[[0, 2]])
        s.ccd_solve(1, gpu.SIDE_CSR)  # u = W[1] from the new v
        W2, H2 = s.get_factors()
        want_u = port.ccd_solve_sweep(csr[0], csr[1], csr[2], H1[1], lam)
        _check_sweep(W2[1], want_u)
        assert np.array_equal(H2, H1)
        # empty segments solve to exactly 0 (src/CCD.cpp:8)
        assert np.all(H1[1][np.diff(csc[0].astype(np.int64)) == 0] == 0.0)


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("shape,panel_rows,chunk", [("ml100k", 0, 0), ("small", 64, 16)])
def test_residual_update_bit_exact(gpu, port, data_factory, layout, shape, panel_rows, chunk):
    d = data_factory(shape)
    csr, csc, _ = sides(d)
    k = 2
    rng = np.random.default_rng(4)
    W = rng.standard_normal((k, d["rows"])).astype(np.float32)
    H = rng.standard_normal((k, d["cols"])).astype(np.float32)
    with gpu.Session(d, gpu.make_params(k=k, layout=layout, panel_rows=panel_rows, chunk=chunk)) as s:
        s.set_factors(W, H)
        s.ccd_update(0, add=False)
        rv, cv = s.get_values()
        want_c = port.ccd_update_sweep(csc[0], csc[1], csc[2], W[0], H[0], add=False)
        want_r = port.ccd_update_sweep(csr[0], csr[1], csr[2], H[0], W[0], add=False)
        assert np.array_equal(cv, want_c) and np.array_equal(rv, want_r)
        s.ccd_update(1, add=True)
        rv, cv = s.get_values()
        assert np.array_equal(cv, port.ccd_update_sweep(csc[0], csc[1], want_c, W[1], H[1], add=True))
        assert np.array_equal(rv, port.ccd_update_sweep(csr[0], csr[1], want_r, H[1], W[1], add=True))


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("schedule", [0, 1])  # FUSED, REFERENCE
@pytest.mark.parametrize("name", GOLDEN_CCD)
def test_trajectory_vs_reference_fixture(gpu, golden, name, schedule, layout):
    """Against outputs of the unmodified reference CPU path (tests/golden)."""
    d, z = golden(name)
    k, lam, iters, inner = int(z["k"]), float(z["lam"]), int(z["maxiter"]), int(z["maxinner"])
    from oracle import port
    W0 = port.initial_col(k, d["rows"])
    with gpu.Session(d, gpu.make_params(k=k, lam=lam, maxinner=inner, schedule=schedule, layout=layout)) as s:
        s.set_factors(W0)
        st = s.iterate(1)
        W1, H1 = s.get_factors()
        assert rel_l2(W1, z["W_iter1"]) <= 1e-4 and rel_l2(H1, z["H_iter1"]) <= 1e-4
        for t in range(k):  # per-rank, as SURVEY.md Appendix D measures it
            assert rel_l2(W1[t], z["W_iter1"][t]) <= 2e-4 and rel_l2(H1[t], z["H_iter1"][t]) <= 2e-4
        rm = [st[0]["rmse"]] + [x["rmse"] for x in s.iterate(iters - 1)] if iters > 1 else [st[0]["rmse"]]
        assert np.allclose(rm, z["rmse_printed"], atol=1e-4, rtol=0)
        W, H = s.get_factors()
        assert rel_l2(W, z["W"]) <= 3e-4 and rel_l2(H, z["H"]) <= 3e-4
        assert abs(s.rmse() - float(z["rmse_final"])) <= 1e-4
        # the residual the solver holds is R - W H^T up to accumulated rounding, like the reference's
        rv, cv = s.get_values()
        assert np.allclose(rv, z["csr_resid"], atol=2e-3) and np.allclose(cv, z["csc_resid"], atol=2e-3)


def test_fused_schedule_equals_reference_schedule_bitwise(gpu, port, data_factory):
    """Fusing the deferred subtraction and the add-back into the first solve sweep changes no
    rounding: same factors, same residual, bit for bit."""
    d = data_factory("ml100k")
    k, lam = 6, 0.05
    W0 = port.initial_col(k, d["rows"])
    outs = []
    for schedule in (0, 1):
        with gpu.Session(d, gpu.make_params(k=k, lam=lam, maxinner=2, schedule=schedule)) as s:
            s.set_factors(W0)
            s.iterate(3)
            outs.append(s.get_factors() + s.get_values())
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_all_pipelines_bitwise_equal(gpu, port, data_factory):
    """The three 8-lane-group ways of feeding the sweep (a register ring; cp.async producer warps or one bulk-copy
    descriptor per item into a shared-memory slot ring) use the same reduction tree: identical factors and residual, for
    several panel / chunk geometries.  The STREAM pipeline (large TMA bulk copies of the contiguous stream into a tile
    ring, one item per warp) has its own fixed tree (32 lanes per item) and pads pieces to 8 instead of 32 entries: it
    agrees with the others to rounding, and with itself bit for bit."""
    for shape, kw in (("ml100k", dict()), ("ml100k", dict(panel_rows=256, chunk=64)), ("small", dict(panel_rows=64, chunk=16))):
        d = data_factory(shape)
        k = 5
        W0 = port.initial_col(k, d["rows"])
        outs = []
        for pipeline in (0, 1, 2, 3, 3):
            with gpu.Session(d, gpu.make_params(k=k, lam=0.05, maxinner=2, pipeline=pipeline, **kw)) as s:
                s.set_factors(W0)
                s.iterate(3)
                outs.append(s.get_factors() + s.get_values())
        for other in outs[1:3]:
            for a, b in zip(outs[0], other):
                assert np.array_equal(a, b)
        for a, b in zip(outs[3], outs[4]):
            assert np.array_equal(a, b)
        for a, b in zip(outs[0], outs[3]):
            assert np.linalg.norm(a - b) <= 2e-5 * np.linalg.norm(a)


def test_stream_pipeline_step_parity_and_schedules(gpu, port, data_factory):
    """STREAM pipeline: one solve sweep against the oracle on identical inputs (<= 5e-5), residual updates bit-exact,
    fused schedule == reference launch order bit for bit."""
    for shape, kw in (("ml100k", dict()), ("small", dict(panel_rows=64, chunk=16)), ("tiny", dict())):
        d = data_factory(shape)
        k, lam = 4, 0.05
        W0 = port.initial_col(k, d["rows"])
        outs = []
        for schedule in (0, 1):
            with gpu.Session(d, gpu.make_params(k=k, lam=lam, maxinner=3, pipeline=3, schedule=schedule, **kw)) as s:
                s.set_factors(W0)
                st = s.iterate(3)
                outs.append(s.get_factors() + s.get_values() + (np.array([x["rmse"] for x in st]),))
        for a, b in zip(*outs):
            assert np.array_equal(a, b)
        want = port.ccdpp(d["rows"], d["cols"], (d["csr_ptr"], d["csr_idx"], d["csr_val"]), (d["csc_ptr"], d["csc_idx"], d["csc_val"]),
                          W0, k, lam, 3, 3, test=(d["test_row"], d["test_col"], d["test_val"]))
        W, H = outs[0][0], outs[0][1]
        assert np.linalg.norm(W - want["W"]) <= 2e-4 * np.linalg.norm(want["W"])
        assert np.linalg.norm(H - want["H"]) <= 2e-4 * np.linalg.norm(want["H"])
        assert np.abs(outs[0][-1] - np.array(want["rmse"])).max() < 1e-4


def test_panel_geometry_does_not_change_residual_and_barely_factors(gpu, port, data_factory):
    d = data_factory("small")
    k, lam = 4, 0.05
    W0 = port.initial_col(k, d["rows"])
    res = []
    for kw in (dict(), dict(panel_rows=64, chunk=16), dict(layout=1)):
        with gpu.Session(d, gpu.make_params(k=k, lam=lam, maxinner=2, **kw)) as s:
            s.set_factors(W0)
            s.iterate(2)
            res.append(s.get_factors())
    for W, H in res[1:]:
        assert rel_l2(W, res[0][0]) <= 1e-4 and rel_l2(H, res[0][1]) <= 1e-4


def test_csr_and_csc_residual_copies_stay_bit_identical(gpu, port, data_factory):
    """u*v == v*u: both copies receive the same rounded product, so after any number of iterations
    they hold the same bits for the same (row, col) — the reference's copies have the same property."""
    d = data_factory("ml100k")
    k = 5
    with gpu.Session(d, gpu.make_params(k=k, lam=0.05, maxinner=3)) as s:
        s.set_factors(port.initial_col(k, d["rows"]))
        s.iterate(2)
        rv, cv = s.get_values()
    rows = np.repeat(np.arange(d["rows"]), np.diff(d["csr_ptr"].astype(np.int64)))
    cols = np.repeat(np.arange(d["cols"]), np.diff(d["csc_ptr"].astype(np.int64)))
    key_r = rows.astype(np.int64) * d["cols"] + d["csr_idx"]
    key_c = d["csc_idx"].astype(np.int64) * d["cols"] + cols
    assert np.array_equal(rv[np.argsort(key_r)].view(np.uint32), cv[np.argsort(key_c)].view(np.uint32))


def test_edge_cases(gpu, port, datagen):
    # empty rows and columns, a 1-entry row, a dense row, k=1, T=1; unsorted input is sorted on upload
    rows, cols = 40, 30
    r = np.array([0, 0, 0, 5, 7, 7, 39] + [12] * cols, np.uint32)
    c = np.array([0, 3, 29, 4, 4, 9, 0] + list(range(cols)), np.uint32)
    rng = np.random.default_rng(9)
    v = rng.integers(1, 6, len(r)).astype(np.float32)
    d = datagen.from_coo(rows, cols, r, c, v, test=(np.array([0, 12]), np.array([3, 7]), np.array([2.0, 4.0], np.float32)))
    csr, csc, test = sides(d)
    for k, inner in ((1, 1), (3, 2)):
        W0 = port.initial_col(k, rows)
        want = port.ccdpp(rows, cols, csr, csc, W0, k, 0.1, 2, inner, test=test)
        for layout in LAYOUTS:
            W, H = W0.copy(), np.zeros((k, cols), np.float32)
            st = gpu.ccdpp_train(d, W, H, gpu.make_params(k=k, lam=0.1, maxiter=2, maxinner=inner, layout=layout))
            assert rel_l2(W, want["W"]) <= 1e-4 and rel_l2(H, want["H"]) <= 1e-4
            assert abs(st[-1]["rmse"] - want["rmse"][-1]) <= 1e-4
            assert np.all(W[:, np.diff(csr[0].astype(np.int64)) == 0] == 0.0)
    # shuffle the entries inside every row/column: the reference does not require sorted indices
    d2 = dict(d)
    for side in ("csr", "csc"):
        ptr = d[side + "_ptr"].astype(np.int64)
        idx, val = d[side + "_idx"].copy(), d[side + "_val"].copy()
        for s_ in range(len(ptr) - 1):
            p = rng.permutation(ptr[s_ + 1] - ptr[s_]) + ptr[s_]
            idx[ptr[s_]:ptr[s_ + 1]], val[ptr[s_]:ptr[s_ + 1]] = idx[p], val[p]
        d2[side + "_idx"], d2[side + "_val"] = idx, val
    k = 3
    W0 = port.initial_col(k, rows)
    want = port.ccdpp(rows, cols, (d2["csr_ptr"], d2["csr_idx"], d2["csr_val"]), (d2["csc_ptr"], d2["csc_idx"], d2["csc_val"]), W0, k, 0.1, 2, 2)
    W, H = W0.copy(), np.zeros((k, cols), np.float32)
    gpu.ccdpp_train(d2, W, H, gpu.make_params(k=k, lam=0.1, maxiter=2, maxinner=2))
    assert rel_l2(W, want["W"]) <= 1e-4 and rel_l2(H, want["H"]) <= 1e-4


def test_one_shot_trainer_equals_session_and_prints_reference_line(gpu, port, data_factory, capfd):
    d = data_factory("small")
    k = 4
    W0 = port.initial_col(k, d["rows"])
    W, H = W0.copy(), np.full((k, d["cols"]), 7.0, np.float32)  # H in is ignored: CCD_CUDA.cu:263-269,287
    gpu.ccdpp_train(d, W, H, gpu.make_params(k=k, lam=0.05, maxiter=2, maxinner=2, quiet=False))
    out = capfd.readouterr().out
    import re
    lines = re.findall(r"\[-INFO-\] iteration num (\d+) \trank_time [\d.]+\|[\d.]+ s \tupdate_time [\d.]+\|[\d.]+s \tRMSE=([\d.]+) time:[\d.]+s", out)
    assert [int(a) for a, _ in lines] == [1, 2]
    with gpu.Session(d, gpu.make_params(k=k, lam=0.05, maxinner=2)) as s:
        s.set_factors(W0)
        st = s.iterate(2)
        W2, H2 = s.get_factors()
    assert np.array_equal(W, W2) and np.array_equal(H, H2)
    assert float(lines[-1][1]) == pytest.approx(st[-1]["rmse"], abs=1e-6)


def test_rmse_kernel(gpu, port, data_factory):
    d = data_factory("ml100k")
    _, _, test = sides(d)
    k = 7
    rng = np.random.default_rng(1)
    W = rng.standard_normal((k, d["rows"])).astype(np.float32)
    H = rng.standard_normal((k, d["cols"])).astype(np.float32)
    with gpu.Session(d, gpu.make_params(k=k)) as s:
        s.set_factors(W, H)
        got = s.rmse()
    want = port.rmse(test[0], test[1], test[2], W, H, d["rows"], d["cols"], k, False)
    assert got == pytest.approx(want, rel=1e-12)


def test_nmf_projection_extension(gpu, port, data_factory):
    """mf_params.nmf_project = 1 (the -N option the reference parses and never uses): every solved coordinate is clamped at
    zero in the finalize, so the factors stay non-negative; with the option off the same run has negative entries and is
    the reference's arithmetic."""
    d = data_factory("small")
    k = 6
    W0 = port.initial_col(k, d["rows"])
    outs = []
    for nmf in (0, 1):
        with gpu.Session(d, gpu.make_params(k=k, lam=0.05, maxinner=2, nmf_project=nmf)) as s:
            s.set_factors(W0)
            st = s.iterate(2)
            W, H = s.get_factors()
            outs.append((W, H, st[-1]["rmse"]))
    (W, H, r0), (Wn, Hn, r1) = outs
    assert (W < 0).any() or (H < 0).any()
    assert (Wn >= 0).all() and (Hn >= 0).all() and np.isfinite(r1)
    assert not np.array_equal(Wn, W)


@pytest.mark.parametrize("shape,kw", [("ml100k", dict()), ("ml100k", dict(panel_rows=256, chunk=64)), ("small", dict(panel_rows=64, chunk=16)), ("tiny", dict())])
@pytest.mark.parametrize("k,inner", [(5, 3), (1, 2), (3, 1)])
def test_persistent_kernel_equals_per_launch_path_bitwise(gpu, port, data_factory, monkeypatch, shape, kw, k, inner):
    """One cooperative launch per outer iteration (k_ccd_persistent) runs the same sweep body and the same finalize as the
    per-launch kernels: identical factors, residual and RMSE, bit for bit, over several outer iterations (first iteration
    without add-back, k = 1 where the subtracted rank is the rank being solved, T = 1 where the v_old copy shares a phase
    with its last reader)."""
    d = data_factory(shape)
    W0 = port.initial_col(k, d["rows"])
    outs = []
    for no_persistent in (False, True):
        if no_persistent:
            monkeypatch.delenv("MF_PERSISTENT", raising=False)
        else:
            monkeypatch.setenv("MF_PERSISTENT", "1")
        with gpu.Session(d, gpu.make_params(k=k, lam=0.05, maxinner=inner, **kw)) as s:
            s.set_factors(W0)
            st = s.iterate(2)
            st += s.iterate(2)
            kt = s.kernel_times()
            assert (kt["persistent_launches"] > 0) == (not no_persistent)
            outs.append(s.get_factors() + s.get_values() + (np.array([x["rmse"] for x in st]),))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("shape,kw", [("ml100k", dict(panel_rows=64, chunk=64)), ("small", dict(panel_rows=16, chunk=32)), ("tiny", dict(panel_rows=8, chunk=8)), ("ml100k", dict())])
@pytest.mark.parametrize("schedule", [0, 1])
def test_lane_per_item_path_is_bitwise_the_group_path(gpu, port, data_factory, monkeypatch, shape, kw, schedule):
    """Short-piece copies run one work item per LANE (32 items in flight per warp) where all items of a batch fit one
    32-entry step; the lane replays the 8-lane group's arithmetic, so factors, residual and RMSE are identical whichever path
    a batch takes — for every sweep mode (both schedules), for geometries that make pieces short, and for the padding the
    short mode picks (8) as well as the default (32)."""
    d = data_factory(shape)
    k = 4
    W0 = port.initial_col(k, d["rows"])
    outs = []
    for short, pad in (("0", 32), ("1", 32), ("1", 8), ("1", 0)):
        monkeypatch.setenv("MF_SHORT_ITEMS", short)
        with gpu.Session(d, gpu.make_params(k=k, lam=0.05, maxinner=2, schedule=schedule, pad_entries=pad, **kw)) as s:
            s.set_factors(W0)
            st = s.iterate(3)
            outs.append(s.get_factors() + s.get_values() + (np.array([x["rmse"] for x in st]),))
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b)
