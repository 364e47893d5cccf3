import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from __graft_entry__ import load_package  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def datagen(pkg):
    import importlib
    return importlib.import_module("cuda_recommender_b200.datagen")


@pytest.fixture(scope="session")
def port():
    from oracle import port as p
    p.lib()
    return p


@pytest.fixture(scope="session")
def ref():
    from oracle import ref as r
    if not r.available():
        pytest.skip("oracle/_ref/libmfref.so not built (needs /root/reference at build time)")
    return r


@pytest.fixture(scope="session")
def gpu(pkg):
    """The CUDA library on a real device.  Fails (does not skip) when the extension is missing."""
    pkg.lib()
    if pkg.device_count() < 1:
        pytest.fail("no CUDA device visible")
    return pkg


_cache = {}


@pytest.fixture(scope="session")
def data_factory(datagen):
    def make(name, seed=None):
        key = (name, seed)
        if key not in _cache:
            _cache[key] = datagen.to_numpy(datagen.synth_named(name, seed=seed))
        return _cache[key]
    return make


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def sides(d):
    return ((d["csr_ptr"], d["csr_idx"], d["csr_val"]), (d["csc_ptr"], d["csc_idx"], d["csc_val"]),
            (d["test_row"], d["test_col"], d["test_val"]))


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden(datagen):
    """name -> (dataset dict rebuilt from the stored COO triples, fixture npz)."""
    def load(name):
        key = ("golden", name)
        if key not in _cache:
            z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
            d = datagen.from_coo(int(z["rows"]), int(z["cols"]), z["coo_row"], z["coo_col"], z["coo_val"].astype(np.float32),
                                 test=(z["test_row"], z["test_col"], z["test_val"].astype(np.float32)))
            _cache[key] = (d, z)
        return _cache[key]
    return load


GOLDEN_CCD = ["ccd_c1_ml100k", "ccd_small_T1", "ccd_tiny_k1"]
GOLDEN_ALS = ["als_ml100k", "als_small_k24"]
