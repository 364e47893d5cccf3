"""CPU: the C-ABI shared library loads and exports every function include/mf_abi.h declares; the
Python binding lists exactly those; without a GPU the compute entry points fail loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest


def _declared_functions(header_path):
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mf_[a-z0-9_]+)\s*\(", text)))


def test_header_library_and_binding_agree(pkg):
    declared = _declared_functions(pkg.HEADER_PATH)
    assert len(declared) >= 20
    lib = pkg.lib()
    for name in declared:
        assert hasattr(lib, name), f"libmfb200.so does not export {name}"
    assert sorted(pkg.ABI_SYMBOLS) == declared
    assert lib.mf_abi_version() == 3


def test_params_default_matches_reference_defaults(pkg):
    p = pkg.mf_params()
    pkg.lib().mf_params_default(C.byref(p))
    # src/pmf.h:26-42
    assert (p.solver_type, p.k, p.threads, p.maxiter, p.maxinneriter) == (0, 10, 4, 5, 1)
    assert p.lambda_ == pytest.approx(0.1) and p.eps == pytest.approx(1e-3)
    assert (p.do_predict, p.verbose, p.do_nmf, p.nBlocks, p.nThreadsPerBlock) == (0, 0, 0, 32, 256)


def test_struct_sizes_match_header_layout(pkg):
    assert C.sizeof(pkg.mf_ratings) == 3 * 8 + 6 * 8
    assert C.sizeof(pkg.mf_testset) == 8 + 3 * 8
    assert C.sizeof(pkg.mf_params) == 4 * 28
    assert C.sizeof(pkg.mf_iter_stats) == 32
    assert C.sizeof(pkg.mf_kernel_times) == 14 * 8 + 4 * 8 + 3 * 8


def test_no_cpu_fallback_without_gpu(pkg, data_factory):
    """Where no CUDA device exists the trainers return an error; they never compute on the host."""
    try:
        n = pkg.device_count()
    except pkg.MFError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present")
    d = data_factory("tiny")
    W = np.full((2, d["rows"]), 0.5, np.float32)
    H = np.zeros((2, d["cols"]), np.float32)
    with pytest.raises(pkg.MFError):
        pkg.ccdpp_train(d, W, H, pkg.make_params(k=2, maxiter=1))
    assert np.all(W == 0.5) and np.all(H == 0.0)
    with pytest.raises(pkg.MFError):
        pkg.Session(d, pkg.make_params(k=2))


def test_product_never_imports_oracle():
    """The shipped path must not reach into oracle/ (tests, smoke() and bench.py's CPU legs only)."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cuda-recommender_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "libmforacle" not in text and "libmfref" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert not re.search(r'#include\s+"[^"]*oracle/', text), f


def test_extension_fields_sit_behind_the_reference_fields(pkg):
    """mf_params keeps the reference's `parameter` fields first (src/pmf.h:10-24) and the extensions behind them; early_stop
    took one of the reserved words (the struct size did not change between ABI versions 2 and 3)."""
    names = [n for n, _ in pkg.mf_params._fields_]
    assert names[:12] == ["solver_type", "k", "threads", "maxiter", "maxinneriter", "lambda_", "eps", "do_predict", "verbose",
                          "do_nmf", "nBlocks", "nThreadsPerBlock"]
    assert names.index("early_stop") == names.index("pad_entries") + 1 and names[-1] == "reserved"
    p = pkg.make_params(k=3, eps=0.25, early_stop=1, do_nmf=1, verbose=1, do_predict=1)
    assert (p.early_stop, p.do_nmf, p.verbose, p.do_predict) == (1, 1, 1, 1) and p.eps == pytest.approx(0.25)
    q = pkg.mf_params()
    pkg.lib().mf_params_default(C.byref(q))
    assert q.early_stop == 0 and q.nmf_project == 0  # inert by default, exactly like the reference
