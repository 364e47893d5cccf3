"""ctypes binding of oracle/libmforacle.so — the C restatement of the reference's CPU
hot path (oracle/mf_oracle.c; every function there cites the reference file:line).

TEST INFRASTRUCTURE ONLY: the checker, never the thing measured or shipped.
All arrays are numpy; sparse copies are (ptr uint32[nseg+1], idx uint32[nnz], val f32[nnz]).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmforacle.so")
_lib = None

u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "mf_oracle.c")):
            build()
        L = C.CDLL(_SO)
        L.orc_initial_col.argtypes = [f32p, C.c_long, C.c_long]
        L.orc_ccd_solve_sweep.argtypes = [C.c_long, u32p, u32p, f32p, f32p, C.c_float, f32p]
        L.orc_ccd_solve_sweep_f64.argtypes = L.orc_ccd_solve_sweep.argtypes
        L.orc_ccd_update_sweep.argtypes = [C.c_long, u32p, u32p, f32p, f32p, f32p, C.c_int]
        L.orc_rmse.argtypes = [C.c_long, u32p, u32p, f32p, f32p, f32p, C.c_long, C.c_long, C.c_long, C.c_int]
        L.orc_predict.argtypes = [C.c_long, u32p, u32p, f32p, f32p, C.c_long, C.c_long, C.c_long, C.c_int, f64p]
        L.orc_predict.restype = None
        L.orc_rmse.restype = C.c_double
        L.orc_ccdpp.argtypes = [C.c_long, C.c_long, u32p, u32p, f32p, u32p, u32p, f32p, f32p, f32p,
                                C.c_long, C.c_float, C.c_int, C.c_int,
                                C.c_long, u32p, u32p, f32p, f64p, C.c_int, C.c_int]
        L.orc_ccdpp_ex.argtypes = [C.c_long, C.c_long, u32p, u32p, f32p, u32p, u32p, f32p, f32p, f32p,
                                   C.c_long, C.c_float, C.c_int, C.c_int,
                                   C.c_long, u32p, u32p, f32p, f64p, f64p, C.c_int, C.c_int, C.c_float,
                                   np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")]
        L.orc_ccdpp_ex.restype = None
        L.orc_als_half_step.argtypes = [C.c_long, u32p, u32p, f32p, f32p, f32p, C.c_long, C.c_float]
        L.orc_als_half_step.restype = C.c_long
        L.orc_als_half_step_f64.argtypes = L.orc_als_half_step.argtypes
        L.orc_als.argtypes = [C.c_long, C.c_long, u32p, u32p, f32p, u32p, u32p, f32p, f32p, f32p,
                              C.c_long, C.c_float, C.c_int,
                              C.c_long, u32p, u32p, f32p, f64p, C.c_int, C.c_int]
        L.orc_als.restype = C.c_long
        L.orc_coo_to_csr_csc.argtypes = [C.c_long, C.c_long, C.c_long, u32p, u32p, f32p,
                                         u32p, u32p, f32p, u32p, u32p, f32p]
        L.orc_degree_bins.argtypes = [C.c_long, u32p, u64p, u64p]
        L.orc_partition.argtypes = [C.c_long, u32p, C.c_int, i64p]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _u32(a):
    return np.ascontiguousarray(a).view(np.uint32) if np.asarray(a).dtype == np.int32 else np.ascontiguousarray(a, dtype=np.uint32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def initial_col(k, n):
    """Factors as the reference seeds them (tools.cpp:165): returns X[k, n]."""
    X = np.empty((k, n), np.float32)
    lib().orc_initial_col(X, k, n)
    return X


def ccd_solve_sweep(ptr, idx, val, gather, lam, f64=False):
    ptr, idx, val, gather = _u32(ptr), _u32(idx), _f32(val), _f32(gather)
    out = np.empty(len(ptr) - 1, np.float32)
    fn = lib().orc_ccd_solve_sweep_f64 if f64 else lib().orc_ccd_solve_sweep
    fn(len(ptr) - 1, ptr, idx, val, gather, lam, out)
    return out


def ccd_update_sweep(ptr, idx, val, gather, seg_factor, add):
    """In place on a copy of val; returns the new values."""
    ptr, idx = _u32(ptr), _u32(idx)
    val = _f32(val).copy()
    lib().orc_ccd_update_sweep(len(ptr) - 1, ptr, idx, val, _f32(gather), _f32(seg_factor), int(bool(add)))
    return val


def rmse(trow, tcol, tval, W, H, rows, cols, k, als_layout):
    trow, tcol, tval = _u32(trow), _u32(tcol), _f32(tval)
    return float(lib().orc_rmse(len(tval), trow, tcol, tval, _f32(W).reshape(-1), _f32(H).reshape(-1),
                                rows, cols, k, int(bool(als_layout))))


def predict(row, col, W, H, rows, cols, k, als_layout):
    """Predictions w_i . h_j for arbitrary pairs (src/extras.cpp:165-168): float64 array."""
    row, col = _u32(row), _u32(col)
    out = np.zeros(len(row), np.float64)
    lib().orc_predict(len(row), row, col, _f32(W).reshape(-1), _f32(H).reshape(-1), rows, cols, k, int(bool(als_layout)), out)
    return out


def _test_arrays(test):
    if test is None:
        z = np.zeros(1, np.uint32)
        return 0, z, z, np.zeros(1, np.float32)
    trow, tcol, tval = test
    return len(tval), _u32(trow), _u32(tcol), _f32(tval)


def ccdpp(rows, cols, csr, csc, W, k, lam, maxiter, maxinner, test=None, f64acc=False, threads=0):
    """Full CCD++ run (CCD.cpp:45).  W: [k, rows] initial factors.  Returns dict with final
    W[k,rows], H[k,cols], per-iteration rmse and the residual copies."""
    csr_ptr, csr_idx, csr_val = _u32(csr[0]), _u32(csr[1]), _f32(csr[2]).copy()
    csc_ptr, csc_idx, csc_val = _u32(csc[0]), _u32(csc[1]), _f32(csc[2]).copy()
    Wf = _f32(W).reshape(k, rows).copy()
    Hf = np.zeros((k, cols), np.float32)
    nt, trow, tcol, tval = _test_arrays(test)
    r = np.zeros(max(maxiter, 1), np.float64)
    lib().orc_ccdpp(rows, cols, csr_ptr, csr_idx, csr_val, csc_ptr, csc_idx, csc_val,
                    Wf.reshape(-1), Hf.reshape(-1), k, lam, maxiter, maxinner,
                    nt, trow, tcol, tval, r, int(f64acc), threads)
    return dict(W=Wf, H=Hf, rmse=r[:maxiter], csr_val=csr_val, csc_val=csc_val)


def ccdpp_ex(rows, cols, csr, csc, W, k, lam, maxiter, maxinner, test=None, nmf=False, early_stop=False, eps=1e-3):
    """CCD++ with the options the reference parses but never acts on switched on (orc_ccdpp_ex): per-rank incremental
    test RMSE (calrmse_r1), the non-negativity clamp and the -e stop rule.  Returns the dict of ccdpp plus
    rank_rmse [maxiter, k] and inner_done [maxiter, k]."""
    csr_ptr, csr_idx, csr_val = _u32(csr[0]), _u32(csr[1]), _f32(csr[2]).copy()
    csc_ptr, csc_idx, csc_val = _u32(csc[0]), _u32(csc[1]), _f32(csc[2]).copy()
    Wf = _f32(W).reshape(k, rows).copy()
    Hf = np.zeros((k, cols), np.float32)
    nt, trow, tcol, tval = _test_arrays(test)
    r = np.zeros(max(maxiter, 1), np.float64)
    rr = np.zeros(max(maxiter, 1) * k, np.float64)
    done = np.zeros(max(maxiter, 1) * k, np.int32)
    lib().orc_ccdpp_ex(rows, cols, csr_ptr, csr_idx, csr_val, csc_ptr, csc_idx, csc_val,
                       Wf.reshape(-1), Hf.reshape(-1), k, lam, maxiter, maxinner,
                       nt, trow, tcol, tval, r, rr, int(nmf), int(early_stop), eps, done)
    return dict(W=Wf, H=Hf, rmse=r[:maxiter], csr_val=csr_val, csc_val=csc_val,
                rank_rmse=rr.reshape(-1, k)[:maxiter], inner_done=done.reshape(-1, k)[:maxiter])


def als_half_step(ptr, idx, val, Y, k, lam, f64=False):
    ptr, idx, val = _u32(ptr), _u32(idx), _f32(val)
    X = np.empty((len(ptr) - 1, k), np.float32)
    fn = lib().orc_als_half_step_f64 if f64 else lib().orc_als_half_step
    fn(len(ptr) - 1, ptr, idx, val, _f32(Y).reshape(-1), X.reshape(-1), k, lam)
    return X


def als(rows, cols, csr, csc, W, H, k, lam, maxiter, test=None, f64=False, threads=0):
    """Full ALS run (ALS.cpp:81).  W: [rows,k], H: [cols,k] initial factors."""
    csr_ptr, csr_idx, csr_val = _u32(csr[0]), _u32(csr[1]), _f32(csr[2])
    csc_ptr, csc_idx, csc_val = _u32(csc[0]), _u32(csc[1]), _f32(csc[2])
    Wf = _f32(W).reshape(rows, k).copy()
    Hf = _f32(H).reshape(cols, k).copy()
    nt, trow, tcol, tval = _test_arrays(test)
    r = np.zeros(max(maxiter, 1), np.float64)
    bad = lib().orc_als(rows, cols, csr_ptr, csr_idx, csr_val, csc_ptr, csc_idx, csc_val,
                        Wf.reshape(-1), Hf.reshape(-1), k, lam, maxiter, nt, trow, tcol, tval, r,
                        int(f64), threads)
    return dict(W=Wf, H=Hf, rmse=r[:maxiter], bad_pivots=int(bad))


def coo_to_csr_csc(rows, cols, r, c, v):
    r, c, v = _u32(r), _u32(c), _f32(v)
    nnz = len(v)
    csr = (np.empty(rows + 1, np.uint32), np.empty(max(nnz, 1), np.uint32)[:nnz], np.empty(max(nnz, 1), np.float32)[:nnz])
    csc = (np.empty(cols + 1, np.uint32), np.empty(max(nnz, 1), np.uint32)[:nnz], np.empty(max(nnz, 1), np.float32)[:nnz])
    csr = tuple(np.ascontiguousarray(a) for a in csr)
    csc = tuple(np.ascontiguousarray(a) for a in csc)
    lib().orc_coo_to_csr_csc(rows, cols, nnz, r, c, v, csr[0], csr[1], csr[2], csc[0], csc[1], csc[2])
    return csr, csc


def degree_bins(ptr):
    ptr = _u32(ptr)
    a = np.zeros(33, np.uint64)
    b = np.zeros(33, np.uint64)
    lib().orc_degree_bins(len(ptr) - 1, ptr, a, b)
    return a, b


def partition(ptr, P):
    ptr = _u32(ptr)
    out = np.zeros(P + 1, np.int64)
    lib().orc_partition(len(ptr) - 1, ptr, P, out)
    return out


def max_threads():
    return int(lib().orc_max_threads())
