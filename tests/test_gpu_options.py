"""GPU (-m gpu): the solver options the reference parses and leaves inert (-N, -e, -q/-p) and the
predict-from-saved-model path (§8 f3), each against the oracle's restatement (oracle/mf_oracle.c orc_ccdpp_ex,
orc_predict) through the C-ABI."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import sides

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cuda-recommender_b200", "host", "b200_recommender")


def _run(cmd):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600).stdout


@pytest.mark.parametrize("shape,schedule", [("ml100k", 0), ("ml100k", 1), ("small", 0)])
def test_per_rank_report_matches_calrmse_r1(gpu, port, data_factory, shape, schedule):
    """verbose + do_predict: time and incremental test RMSE after every rank — the reference's commented-out verbose block
    (src/CCD.cpp:141-148, calrmse_r1 src/tools.cpp:260-270); the factors are the ones a silent run gives, bit for bit."""
    d = data_factory(shape)
    csr, csc, test = sides(d)
    k, lam, outer, inner = 6, 0.05, 3, 2
    W0 = port.initial_col(k, d["rows"])
    want = port.ccdpp_ex(d["rows"], d["cols"], csr, csc, W0, k, lam, outer, inner, test=test)
    with gpu.Session(d, gpu.make_params(k=k, lam=lam, maxinner=inner, schedule=schedule, verbose=1, do_predict=1)) as s:
        s.set_factors(W0)
        for it in range(outer):
            st = s.iterate(1)
            rs = s.rank_stats()
            assert np.abs(rs["rmse"] - want["rank_rmse"][it]).max() < 1e-4
            assert abs(rs["rmse"][-1] - st[0]["rmse"]) < 1e-4  # after the last rank: the iteration's RMSE (FP32 drift only)
            assert (rs["seconds"] > 0).all() and (rs["inner_iters"] == inner).all()
        loud = s.get_factors()
    with gpu.Session(d, gpu.make_params(k=k, lam=lam, maxinner=inner, schedule=schedule)) as s:
        s.set_factors(W0)
        s.iterate(outer)
        for a, b in zip(loud, s.get_factors()):
            assert np.array_equal(a, b)


def test_do_nmf_clamps_like_the_oracle(gpu, port, data_factory):
    d = data_factory("ml100k")
    csr, csc, test = sides(d)
    k, lam = 5, 0.05
    W0 = port.initial_col(k, d["rows"]) - 0.05  # some negative coordinates to start from
    want = port.ccdpp_ex(d["rows"], d["cols"], csr, csc, W0, k, lam, 2, 3, test=test, nmf=True)
    outs = []
    for kw in (dict(do_nmf=1), dict(nmf_project=1)):
        with gpu.Session(d, gpu.make_params(k=k, lam=lam, maxinner=3, **kw)) as s:
            s.set_factors(W0)
            st = s.iterate(2)
            outs.append(s.get_factors())
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    W, H = outs[0]
    assert (W >= 0).all() and (H >= 0).all() and (want["W"] >= 0).all()
    assert np.linalg.norm(W - want["W"]) <= 2e-4 * np.linalg.norm(want["W"])
    assert np.linalg.norm(H - want["H"]) <= 2e-4 * np.linalg.norm(want["H"])
    assert abs(st[-1]["rmse"] - want["rmse"][-1]) < 1e-4


@pytest.mark.parametrize("shape,eps", [("ml100k", 1e-3), ("ml100k", 5e-2), ("small", 1e-2)])
def test_early_stop_rule(gpu, port, data_factory, shape, eps):
    """early_stop = 1: the -e rule (function decrease below eps x the largest seen ends a rank's inner iterations).  Same
    decisions as the oracle wherever the margin is not within rounding; same RMSE trajectory."""
    d = data_factory(shape)
    csr, csc, test = sides(d)
    k, lam, outer, inner = 5, 0.05, 4, 5
    W0 = port.initial_col(k, d["rows"])
    want = port.ccdpp_ex(d["rows"], d["cols"], csr, csc, W0, k, lam, outer, inner, test=test, early_stop=True, eps=eps)
    done = []
    with gpu.Session(d, gpu.make_params(k=k, lam=lam, maxinner=inner, early_stop=1, eps=eps)) as s:
        s.set_factors(W0)
        for it in range(outer):
            st = s.iterate(1)
            done.append(s.rank_stats()["inner_iters"].copy())
            assert abs(st[0]["rmse"] - want["rmse"][it]) < 2e-4
    done = np.array(done)
    assert done.sum() < outer * k * inner, "the rule never fired"
    assert (done != want["inner_done"]).sum() <= 1, (done, want["inner_done"])
    # inert by default, exactly like the reference: eps alone changes nothing
    outs = []
    for e in (1e-3, 0.5):
        with gpu.Session(d, gpu.make_params(k=k, lam=lam, maxinner=inner, eps=e)) as s:
            s.set_factors(W0)
            s.iterate(2)
            outs.append(s.get_factors())
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


def test_predict_pairs_bit_exact(gpu, port):
    rng = np.random.default_rng(5)
    rows, cols, k, n = 700, 300, 13, 5000
    W = rng.standard_normal((rows, k)).astype(np.float32)
    H = rng.standard_normal((cols, k)).astype(np.float32)
    r = rng.integers(0, rows, n).astype(np.uint32)
    c = rng.integers(0, cols, n).astype(np.uint32)
    got = gpu.predict_pairs(W, H, r, c)
    assert np.array_equal(got, port.predict(r, c, W, H, rows, cols, k, True))
    with pytest.raises(gpu.MFError):
        gpu.predict_pairs(W, H, np.array([rows], np.uint32), np.array([0], np.uint32))


@pytest.mark.parametrize("als", [False, True])
def test_cli_save_then_load_round_trip(gpu, datagen, data_factory, tmp_path, als):
    """-save then -load: the predict-only run (calculate_rmse_from_file, src/extras.cpp:143-180) reproduces the training
    run's final RMSE and its -p 1 prediction file from the saved model alone."""
    d = data_factory("small")
    datagen.write_dataset(str(tmp_path), d)
    out = _run([CLI, "-CUDA", "-k", "5", "-l", "0.05", "-t", "2", "-T", "2", "-p", "1", "-save"] + (["-ALS"] if als else []) + [str(tmp_path)])
    assert "FAILED" not in out, out
    trained = float(re.search(r"Test RMSE = (\d+\.\d+)", out).group(1))
    first = open(os.path.join(str(tmp_path), "output")).read()
    os.remove(os.path.join(str(tmp_path), "output"))
    out2 = _run([CLI, "-load", str(tmp_path)])
    m = re.search(r"\[FINAL INFO\] Test RMSE = (\d+\.\d+)", out2)
    assert m, out2
    assert abs(float(m.group(1)) - trained) < 2e-6
    assert open(os.path.join(str(tmp_path), "output")).read() == first
    if not als:  # the per-rank lines of the verbose block
        ranks = re.findall(r"iter (\d+) rank (\d+) time [\d.]+ rmse ([\d.]+)", out)
        assert len(ranks) == 2 * 5, out
        last = re.findall(r"RMSE=([\d.]+)", out)
        assert abs(float(ranks[4][2]) - float(last[0])) < 1e-4


def test_cli_nmf_and_eps_flags(gpu, datagen, data_factory, tmp_path):
    d = data_factory("small")
    datagen.write_dataset(str(tmp_path), d)
    base = [CLI, "-CUDA", "-k", "4", "-l", "0.05", "-t", "3", "-T", "4", "-save"]
    plain = _run(base + [str(tmp_path)])
    nmf = _run(base + ["-N", "1", str(tmp_path)])
    raw = open(os.path.join(str(tmp_path), "model"), "rb").read()
    W = np.frombuffer(raw, np.float32, 300 * 4, offset=16)
    assert (W >= 0).all()
    stop = _run(base + ["-e", "0.05", str(tmp_path)])
    for o in (plain, nmf, stop):
        assert "FAILED" not in o and re.search(r"Test RMSE = [\d.]+", o), o


def _shuffled(d, seed=3):
    """The same matrix with the entries of every row / column in random order (the reference never asks for sorted ones)."""
    rng = np.random.default_rng(seed)
    d2 = dict(d)
    for side in ("csr", "csc"):
        ptr = d[side + "_ptr"].astype(np.int64)
        idx, val = d[side + "_idx"].copy(), d[side + "_val"].copy()
        for s in range(len(ptr) - 1):
            p = rng.permutation(ptr[s + 1] - ptr[s]) + ptr[s]
            idx[ptr[s]:ptr[s + 1]], val[ptr[s]:ptr[s + 1]] = idx[p], val[p]
        d2[side + "_idx"], d2[side + "_val"] = idx, val
    return d2


def test_unsorted_segments_are_sorted_on_upload_not_degraded(gpu, port, data_factory):
    """Segments that are not index-sorted are sorted once at session creation and still get the panel layout: the factors
    are the sorted input's, bit for bit, and the residual comes back in the CALLER's order."""
    d = data_factory("small")
    d2 = _shuffled(d)
    k = 4
    W0 = port.initial_col(k, d["rows"])
    outs = []
    for data in (d, d2):
        with gpu.Session(data, gpu.make_params(k=k, lam=0.05, maxinner=2)) as s:
            s.set_factors(W0)
            s.iterate(2)
            assert s.panel_layout(gpu.SIDE_CSC)["n_items"] > 0  # the panel layout, not the caller-order fallback
            outs.append(s.get_factors() + s.get_values())
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    for side, got_sorted, got_shuffled in (("csr", outs[0][2], outs[1][2]), ("csc", outs[0][3], outs[1][3])):
        ptr = d[side + "_ptr"].astype(np.int64)
        for s_ in range(len(ptr) - 1):
            lo, hi = ptr[s_], ptr[s_ + 1]
            order = np.argsort(d2[side + "_idx"][lo:hi], kind="stable")
            assert np.array_equal(got_shuffled[lo:hi][order], got_sorted[lo:hi])


def test_bad_input_is_refused(gpu, port, data_factory):
    d = data_factory("small")
    p = gpu.make_params(k=3, lam=0.05, maxinner=1)
    for solver in (gpu.SOLVER_CCD, gpu.SOLVER_ALS):
        for layout in ((0, 1) if solver == gpu.SOLVER_CCD else (0,)):
            q = gpu.make_params(solver=solver, k=3, lam=0.05, maxinner=1, layout=layout)
            bad = dict(d)
            bad["csc_idx"] = d["csc_idx"].copy()
            bad["csc_idx"][7] = d["rows"] + 5           # not a row
            with pytest.raises(gpu.MFError, match="index"):
                gpu.Session(bad, q)
            bad = dict(d)
            bad["csr_idx"] = d["csr_idx"].copy()
            bad["csr_idx"][-1] = 0xfffffff0             # not a column
            with pytest.raises(gpu.MFError, match="index"):
                gpu.Session(bad, q)
    bad = dict(d)
    bad["csr_ptr"] = d["csr_ptr"].copy()
    bad["csr_ptr"][5], bad["csr_ptr"][6] = d["csr_ptr"][6], d["csr_ptr"][5] - 1 if d["csr_ptr"][5] else 0
    if not np.all(np.diff(bad["csr_ptr"].astype(np.int64)) >= 0):
        with pytest.raises(gpu.MFError, match="non-decreasing"):
            gpu.Session(bad, p)
    bad = dict(d)
    bad["test_col"] = d["test_col"].copy()
    bad["test_col"][0] = d["cols"]
    with pytest.raises(gpu.MFError, match="test set"):
        gpu.Session(bad, p)
    with gpu.Session(d, p) as s:
        s.set_factors(port.initial_col(3, d["rows"]))
        with pytest.raises(gpu.MFError, match="outside"):
            s.predict(np.array([0, d["rows"]], np.uint32), np.array([0, 0], np.uint32))
    with pytest.raises(gpu.MFError, match="outside"):
        gpu.build_csr_csc(4, 4, np.array([0, 4], np.uint32), np.array([1, 1], np.uint32), np.array([1, 2], np.float32))
