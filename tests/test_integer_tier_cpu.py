"""CPU: the integer-tier oracle (COO -> CSR/CSC stable counting sort, degree bins, nnz-balanced
partition) against independent numpy statements, including empty rows/columns and shuffled input."""
import numpy as np
import pytest


def _random_coo(rng, rows, cols, nnz):
    keys = rng.choice(rows * cols, size=nnz, replace=False)
    rng.shuffle(keys)
    return (keys // cols).astype(np.uint32), (keys % cols).astype(np.uint32), rng.random(nnz).astype(np.float32)


@pytest.mark.parametrize("rows,cols,nnz", [(1, 1, 1), (5, 7, 0), (50, 40, 600), (300, 1000, 20000), (1000, 3, 1500)])
def test_coo_to_csr_csc(port, datagen, rows, cols, nnz):
    rng = np.random.default_rng(rows * 31 + cols)
    r, c, v = _random_coo(rng, rows, cols, nnz)
    csr, csc = port.coo_to_csr_csc(rows, cols, r, c, v)
    want = datagen.from_coo(rows, cols, r, c, v)
    assert np.array_equal(csr[0], want["csr_ptr"]) and np.array_equal(csr[1], want["csr_idx"]) and np.array_equal(csr[2], want["csr_val"])
    assert np.array_equal(csc[0], want["csc_ptr"]) and np.array_equal(csc[1], want["csc_idx"]) and np.array_equal(csc[2], want["csc_val"])
    assert csr[0][0] == 0 and csr[0][-1] == nnz and csc[0][0] == 0 and csc[0][-1] == nnz  # pmf_util.h:119-129 invariants


def test_degree_bins(port):
    deg = np.array([0, 0, 1, 2, 3, 4, 7, 8, 255, 256, 70000], np.int64)
    ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint32)
    seg, nnz = port.degree_bins(ptr)
    want_seg = np.zeros(33, np.uint64)
    want_nnz = np.zeros(33, np.uint64)
    for d in deg:
        b = int(d).bit_length()
        want_seg[b] += 1
        want_nnz[b] += d
    assert np.array_equal(seg, want_seg) and np.array_equal(nnz, want_nnz)


@pytest.mark.parametrize("P", [1, 2, 3, 8])
def test_partition_is_nnz_balanced(port, data_factory, P):
    d = data_factory("small")
    for ptr in (d["csr_ptr"], d["csc_ptr"]):
        b = port.partition(ptr, P)
        assert b[0] == 0 and b[-1] == len(ptr) - 1 and np.all(np.diff(b) >= 0)
        nnz = int(ptr[-1])
        for p in range(1, P):
            target = -(-nnz * p // P)
            assert int(ptr[b[p]]) >= target and (b[p] == 0 or int(ptr[b[p] - 1]) < target)
        per = np.diff(ptr[b].astype(np.int64))
        maxdeg = int(np.diff(ptr.astype(np.int64)).max())
        assert per.max() - per.min() <= 2 * maxdeg + 1
