"""GPU (-m gpu): integer / index work must be bit-exact — CSR/CSC build, degree bins, partition and
the panel layout the session builds, against the oracle and the numpy restatement."""
import numpy as np
import pytest

import panel_ref
from conftest import sides

pytestmark = pytest.mark.gpu


def _random_coo(rng, rows, cols, nnz):
    keys = rng.choice(rows * cols, size=nnz, replace=False)
    rng.shuffle(keys)
    return (keys // cols).astype(np.uint32), (keys % cols).astype(np.uint32), rng.random(nnz).astype(np.float32)


@pytest.mark.parametrize("rows,cols,nnz", [(1, 1, 1), (5, 7, 0), (50, 40, 600), (300, 1000, 20000), (1000, 3, 1500), (4000, 70000, 300000)])
def test_build_csr_csc_bit_exact(gpu, port, rows, cols, nnz):
    rng = np.random.default_rng(rows + 7 * cols)
    r, c, v = _random_coo(rng, rows, cols, nnz)
    csr, csc = gpu.build_csr_csc(rows, cols, r, c, v)
    wcsr, wcsc = port.coo_to_csr_csc(rows, cols, r, c, v)
    for got, want in zip(csr + csc, wcsr + wcsc):
        assert np.array_equal(got, want)


def test_build_wide_minor_dimension_and_duplicates(gpu, port):
    """A minor dimension wider than one shared-memory bitmap (> ~928 K keys: the Yahoo-Music shape's 1 000 990 rows on the
    CSC build) is ranked in several windows of the key range; a duplicate (row, col) pair is refused."""
    rng = np.random.default_rng(11)
    rows, cols, nnz = 2_100_000, 37, 60_000
    r, c, v = _random_coo(rng, rows, cols, nnz)
    csr, csc = gpu.build_csr_csc(rows, cols, r, c, v)
    wcsr, wcsc = port.coo_to_csr_csc(rows, cols, r, c, v)
    for got, want in zip(csr + csc, wcsr + wcsc):
        assert np.array_equal(got, want)
    r2, c2, v2 = np.concatenate([r, r[:1]]), np.concatenate([c, c[:1]]), np.concatenate([v, v[:1]])
    with pytest.raises(gpu.MFError, match="more than once"):
        gpu.build_csr_csc(rows, cols, r2, c2, v2)


def test_degree_bins_and_partition_bit_exact(gpu, port, data_factory):
    d = data_factory("ml100k")
    for ptr in (d["csr_ptr"], d["csc_ptr"]):
        a, b = gpu.degree_bins(ptr)
        wa, wb = port.degree_bins(ptr)
        assert np.array_equal(a, wa) and np.array_equal(b, wb)
        for P in (1, 2, 4, 8):
            assert np.array_equal(gpu.partition(ptr, P), port.partition(ptr, P))


@pytest.mark.parametrize("shape,panel_rows,chunk", [("small", 0, 0), ("small", 64, 16), ("ml100k", 0, 0), ("ml100k", 256, 64), ("tiny", 8, 8)])
def test_panel_layout_bit_exact(gpu, data_factory, shape, panel_rows, chunk):
    d = data_factory(shape)
    csr, csc, _ = sides(d)
    with gpu.Session(d, gpu.make_params(k=2, panel_rows=panel_rows, chunk=chunk)) as s:
        for side, (ptr, idx, val), gdim in ((gpu.SIDE_CSC, csc, d["rows"]), (gpu.SIDE_CSR, csr, d["cols"])):
            got = s.panel_layout(side)
            pr = panel_ref.session_panel_rows(gdim, side == gpu.SIDE_CSR, panel_rows)
            pad = panel_ref.session_pad(len(ptr) - 1, len(idx), gdim, pr)
            want = panel_ref.panel_layout(ptr, idx, val, gdim, pr, chunk if chunk else 512, pad=pad)
            assert got["n_panels"] == want["n_panels"] and got["n_padded"] == want["n_padded"] and got["n_items"] == want["n_items"]
            panel_ref.check_layout(got, want)
        # and the values come back in the caller's order, untouched
        rv, cv = s.get_values()
        assert np.array_equal(rv, csr[2]) and np.array_equal(cv, csc[2])
