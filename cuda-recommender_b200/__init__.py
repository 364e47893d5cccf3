"""cuda-recommender_b200 — B200-native CCD++ / ALS matrix-factorization training.

Python is only the thinnest host layer here: a ctypes binding of the C-ABI shared library
(`libmfb200.so`, declared in include/mf_abi.h) that replaces the reference's GPU entry
points `kernel_wrapper_ccdpp_NV` / `kernel_wrapper_als_NV` (cuda_src/CCD_CUDA.h:49,
cuda_src/ALS_CUDA.h:40).  The product is the CUDA library and the C++ CLI in host/.

There is no CPU fallback: importing works without a GPU (so the build and the symbol
table can be checked), but every compute call goes to the CUDA library and raises
`MFError` if it, or a CUDA device, is missing.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MF_LIB") or os.path.join(_HERE, "libmfb200.so")  # MF_LIB: an experimental build of the same library
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mf_abi.h")

MF_OK = 0
SOLVER_CCD, SOLVER_ALS = 0, 1
SCHEDULE_FUSED, SCHEDULE_REFERENCE = 0, 1
LAYOUT_PANEL, LAYOUT_DIRECT = 0, 1
PIPELINE_REGISTERS, PIPELINE_ASYNC, PIPELINE_TMA_BULK = 0, 1, 2
SIDE_CSC, SIDE_CSR = 0, 1


class MFError(RuntimeError):
    pass


class mf_ratings(C.Structure):
    _fields_ = [("rows", C.c_int64), ("cols", C.c_int64), ("nnz", C.c_int64),
                ("csr_row_ptr", C.c_void_p), ("csr_col_idx", C.c_void_p), ("csr_val", C.c_void_p),
                ("csc_col_ptr", C.c_void_p), ("csc_row_idx", C.c_void_p), ("csc_val", C.c_void_p)]


class mf_testset(C.Structure):
    _fields_ = [("nnz", C.c_int64), ("row", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p)]


class mf_params(C.Structure):
    _fields_ = [("solver_type", C.c_int32), ("k", C.c_uint32), ("threads", C.c_int32), ("maxiter", C.c_int32),
                ("maxinneriter", C.c_int32), ("lambda_", C.c_float), ("eps", C.c_float), ("do_predict", C.c_int32),
                ("verbose", C.c_int32), ("do_nmf", C.c_int32), ("nBlocks", C.c_uint32), ("nThreadsPerBlock", C.c_uint32),
                ("device", C.c_int32), ("schedule", C.c_int32), ("layout", C.c_int32), ("quiet", C.c_int32),
                ("panel_rows", C.c_int32), ("chunk", C.c_int32), ("nmf_project", C.c_int32),
                ("no_launch_timing", C.c_int32), ("pipeline", C.c_int32), ("timing_stride", C.c_int32), ("pad_entries", C.c_int32), ("early_stop", C.c_int32), ("reserved", C.c_int32 * 4)]


class mf_iter_stats(C.Structure):
    _fields_ = [("rank_time", C.c_double), ("update_time", C.c_double), ("rmse", C.c_double), ("rmse_time", C.c_double)]


class mf_kernel_times(C.Structure):
    _fields_ = [("solve_s", C.c_double), ("solve_launches", C.c_int64), ("fused_s", C.c_double), ("fused_launches", C.c_int64),
                ("update_s", C.c_double), ("update_launches", C.c_int64), ("finalize_s", C.c_double), ("finalize_launches", C.c_int64),
                ("als_s", C.c_double), ("als_launches", C.c_int64), ("rmse_s", C.c_double), ("rmse_launches", C.c_int64),
                ("collective_s", C.c_double), ("collective_launches", C.c_int64),
                ("solve_bytes", C.c_int64), ("fused_bytes", C.c_int64), ("update_bytes", C.c_int64),
                ("total_launches", C.c_int64),
                ("persistent_s", C.c_double), ("persistent_launches", C.c_int64), ("persistent_bytes", C.c_int64)]


# every symbol include/mf_abi.h declares (tests/test_abi_symbols.py checks header <-> library <-> this list)
ABI_SYMBOLS = [
    "mf_abi_version", "mf_last_error", "mf_device_count", "mf_params_default", "mf_host_initial_col",
    "mf_ccdpp_train", "mf_als_train", "mf_release_cached_memory",
    "mf_session_create", "mf_session_destroy", "mf_dist_unique_id", "mf_session_create_dist",
    "mf_session_set_factors", "mf_session_get_factors", "mf_session_get_values",
    "mf_session_ccdpp_iterate", "mf_session_als_iterate", "mf_session_rmse", "mf_session_predict", "mf_session_kernel_times",
    "mf_session_rank_stats", "mf_predict_pairs",
    "mf_session_last_seconds", "mf_session_ccd_solve", "mf_session_ccd_update", "mf_session_als_half",
    "mf_build_csr_csc", "mf_degree_bins", "mf_partition", "mf_session_panel_layout", "mf_als_plan",
]

_lib = None


def build(verbose=False):
    """Compile libmfb200.so (and the CLI) for sm_100a in-tree."""
    cmd = ["make", "-C", _HERE, "-j8", "all"] + ([] if verbose else ["-s"])
    subprocess.check_call(cmd)


def lib():
    """The loaded C-ABI library.  Raises MFError when it has not been built: there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MFError(f"{LIB_PATH} is missing: build it with `make -C {_HERE}` "
                          "(__graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.mf_last_error.restype = C.c_char_p
        vp = C.c_void_p
        L.mf_params_default.argtypes = [C.POINTER(mf_params)]
        L.mf_params_default.restype = None
        L.mf_device_count.argtypes = [C.POINTER(C.c_int)]
        L.mf_host_initial_col.argtypes = [vp, C.c_int64, C.c_int64]
        L.mf_host_initial_col.restype = None
        L.mf_ccdpp_train.argtypes = [C.POINTER(mf_ratings), C.POINTER(mf_testset), vp, vp, C.POINTER(mf_params), vp]
        L.mf_als_train.argtypes = L.mf_ccdpp_train.argtypes
        L.mf_session_create.argtypes = [C.POINTER(mf_ratings), C.POINTER(mf_testset), C.POINTER(mf_params), C.POINTER(vp)]
        L.mf_session_create_dist.argtypes = [C.POINTER(mf_ratings), C.POINTER(mf_testset), C.POINTER(mf_params),
                                             C.c_int, C.c_int, vp, C.POINTER(vp)]
        L.mf_dist_unique_id.argtypes = [vp]
        L.mf_session_destroy.argtypes = [vp]
        L.mf_session_set_factors.argtypes = [vp, vp, vp]
        L.mf_session_get_factors.argtypes = [vp, vp, vp]
        L.mf_session_get_values.argtypes = [vp, vp, vp]
        L.mf_session_ccdpp_iterate.argtypes = [vp, C.c_int, vp]
        L.mf_session_als_iterate.argtypes = [vp, C.c_int, vp]
        L.mf_session_rmse.argtypes = [vp, C.POINTER(C.c_double)]
        L.mf_release_cached_memory.argtypes = [C.c_int]
        L.mf_als_plan.argtypes = [C.c_int64, vp, C.c_uint32, vp, C.POINTER(C.c_int64), C.POINTER(C.c_uint32)]
        L.mf_session_predict.argtypes = [vp, C.c_int64, vp, vp, vp]
        L.mf_session_kernel_times.argtypes = [vp, C.POINTER(mf_kernel_times)]
        L.mf_session_last_seconds.argtypes = [vp, C.POINTER(C.c_double)]
        L.mf_session_ccd_solve.argtypes = [vp, C.c_int, C.c_int]
        L.mf_session_ccd_update.argtypes = [vp, C.c_int, C.c_int]
        L.mf_session_als_half.argtypes = [vp, C.c_int]
        L.mf_build_csr_csc.argtypes = [C.c_int64, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]
        L.mf_degree_bins.argtypes = [C.c_int64, vp, vp, vp, C.c_int]
        L.mf_partition.argtypes = [C.c_int64, vp, C.c_int, vp, C.c_int]
        L.mf_session_panel_layout.argtypes = [vp, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), vp, vp, vp]
        L.mf_session_rank_stats.argtypes = [vp, vp, vp, vp]
        L.mf_predict_pairs.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, vp, vp, vp, C.c_int]
        _lib = L
    return _lib


def _check(rc):
    if rc != MF_OK:
        raise MFError(f"mf error {rc}: {lib().mf_last_error().decode(errors='replace')}")


def initial_col(k, n):
    """Host-side factor seeding as the reference's initial_col(X, k, n) (tools.cpp:165): X[k, n]."""
    X = np.empty((k, n), np.float32)
    lib().mf_host_initial_col(X.ctypes.data, k, n)
    return X


def device_count():
    n = C.c_int(0)
    _check(lib().mf_device_count(C.byref(n)))
    return n.value


def als_plan(ptr, split=8192):
    """The ALS work list for a host pointer array: (items[n, 4] = {segment, part, nparts, slot}, n_slots).  Host-only."""
    ptr = np.ascontiguousarray(ptr, np.uint32)
    n, slots = C.c_int64(0), C.c_uint32(0)
    _check(lib().mf_als_plan(len(ptr) - 1, ptr.ctypes.data, int(split), None, C.byref(n), C.byref(slots)))
    items = np.zeros((n.value, 4), np.uint32)
    _check(lib().mf_als_plan(len(ptr) - 1, ptr.ctypes.data, int(split), items.ctypes.data, C.byref(n), C.byref(slots)))
    return items, int(slots.value)


def release_cached_memory(device=0):
    """Hand the device memory cached between sessions (rating arena, scratch pool) back to the driver."""
    _check(lib().mf_release_cached_memory(int(device)))


def _ptr(a):
    """Address of a numpy array or a torch tensor (host or device)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor


def _keep(a, dtype):
    """numpy arrays are made contiguous with the expected dtype (uint32 accepts int32 bit patterns);
    torch tensors must already be contiguous with a 4-byte dtype."""
    if isinstance(a, np.ndarray):
        if dtype == np.uint32 and a.dtype == np.int32:
            a = a.view(np.uint32)
        return np.ascontiguousarray(a, dtype=dtype)
    if not a.is_contiguous():
        a = a.contiguous()
    return a


class Ratings:
    """Holds references to the six arrays of a paired CSR+CSC rating matrix (numpy or torch, host
    or device) and the optional test triples, and exposes them as mf_ratings / mf_testset."""

    def __init__(self, data):
        g = lambda key, dt: _keep(data[key], dt)
        self.rows, self.cols, self.nnz = int(data["rows"]), int(data["cols"]), int(data["nnz"])
        self.arrays = dict(csr_ptr=g("csr_ptr", np.uint32), csr_idx=g("csr_idx", np.uint32), csr_val=g("csr_val", np.float32),
                           csc_ptr=g("csc_ptr", np.uint32), csc_idx=g("csc_idx", np.uint32), csc_val=g("csc_val", np.float32))
        self.nt = int(data.get("nnz_test", 0) or 0)
        if self.nt:
            self.arrays.update(test_row=g("test_row", np.uint32), test_col=g("test_col", np.uint32), test_val=g("test_val", np.float32))
        a = self.arrays
        self.c_ratings = mf_ratings(self.rows, self.cols, self.nnz, _ptr(a["csr_ptr"]), _ptr(a["csr_idx"]), _ptr(a["csr_val"]),
                                    _ptr(a["csc_ptr"]), _ptr(a["csc_idx"]), _ptr(a["csc_val"]))
        self.c_test = mf_testset(self.nt, _ptr(a.get("test_row")), _ptr(a.get("test_col")), _ptr(a.get("test_val")))


def make_params(solver=SOLVER_CCD, k=10, lam=0.1, maxiter=5, maxinner=1, device=0, schedule=SCHEDULE_FUSED,
                layout=LAYOUT_PANEL, quiet=True, panel_rows=0, chunk=0, nmf_project=0, no_launch_timing=0, pipeline=0, timing_stride=0, pad_entries=0,
                eps=None, do_predict=0, verbose=0, do_nmf=0, early_stop=0):
    p = mf_params()
    lib().mf_params_default(C.byref(p))
    p.solver_type, p.k, p.lambda_, p.maxiter, p.maxinneriter = solver, k, lam, maxiter, maxinner
    p.device, p.schedule, p.layout, p.quiet = device, schedule, layout, int(quiet)
    p.panel_rows, p.chunk, p.nmf_project, p.no_launch_timing = panel_rows, chunk, nmf_project, no_launch_timing
    p.pipeline, p.timing_stride, p.pad_entries = pipeline, timing_stride, pad_entries
    p.do_predict, p.verbose, p.do_nmf, p.early_stop = do_predict, verbose, do_nmf, early_stop
    if eps is not None:
        p.eps = eps
    return p


def _stats_list(arr, n):
    return [dict(rank_time=arr[i].rank_time, update_time=arr[i].update_time, rmse=arr[i].rmse, rmse_time=arr[i].rmse_time)
            for i in range(n)]


def ccdpp_train(data, W, H, params):
    """Drop-in for kernel_wrapper_ccdpp_NV(R, T, W, H, parameters) (CCD_CUDA.cu:164): host buffers in,
    W [k, rows] / H [k, cols] overwritten in place with the final factors.  Returns per-iteration stats."""
    R = data if isinstance(data, Ratings) else Ratings(data)
    assert W.dtype == np.float32 and H.dtype == np.float32 and W.flags.c_contiguous and H.flags.c_contiguous
    st = (mf_iter_stats * max(params.maxiter, 1))()
    _check(lib().mf_ccdpp_train(C.byref(R.c_ratings), C.byref(R.c_test), W.ctypes.data, H.ctypes.data, C.byref(params), C.addressof(st)))
    return _stats_list(st, params.maxiter)


def als_train(data, W, H, params):
    """Drop-in for kernel_wrapper_als_NV (ALS_CUDA.cu:183): W [rows, k], H [cols, k] in place."""
    R = data if isinstance(data, Ratings) else Ratings(data)
    assert W.dtype == np.float32 and H.dtype == np.float32 and W.flags.c_contiguous and H.flags.c_contiguous
    st = (mf_iter_stats * max(params.maxiter, 1))()
    _check(lib().mf_als_train(C.byref(R.c_ratings), C.byref(R.c_test), W.ctypes.data, H.ctypes.data, C.byref(params), C.addressof(st)))
    return _stats_list(st, params.maxiter)


class Session:
    """Device-resident training state (mf_session): ratings/residual, factors and test set stay in HBM."""

    def __init__(self, data, params, rank=0, nranks=1, nccl_id=None):
        self.R = data if isinstance(data, Ratings) else Ratings(data)
        self.params = params
        self.k = int(params.k)
        self.als = params.solver_type == SOLVER_ALS
        self.rows, self.cols, self.nnz = self.R.rows, self.R.cols, self.R.nnz
        h = C.c_void_p()
        if nranks > 1:
            idbuf = C.create_string_buffer(bytes(nccl_id), 128)
            _check(lib().mf_session_create_dist(C.byref(self.R.c_ratings), C.byref(self.R.c_test), C.byref(params),
                                                rank, nranks, C.addressof(idbuf), C.byref(h)))
        else:
            _check(lib().mf_session_create(C.byref(self.R.c_ratings), C.byref(self.R.c_test), C.byref(params), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            lib().mf_session_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _shapes(self):
        return ((self.rows, self.k), (self.cols, self.k)) if self.als else ((self.k, self.rows), (self.k, self.cols))

    def set_factors(self, W, H=None):
        sw, sh = self._shapes()
        W = np.ascontiguousarray(W, np.float32).reshape(sw)
        Hp = None
        if H is not None:
            H = np.ascontiguousarray(H, np.float32).reshape(sh)
            Hp = H.ctypes.data
        _check(lib().mf_session_set_factors(self.h, W.ctypes.data, Hp))

    def get_factors(self):
        sw, sh = self._shapes()
        W = np.empty(sw, np.float32)
        H = np.empty(sh, np.float32)
        _check(lib().mf_session_get_factors(self.h, W.ctypes.data, H.ctypes.data))
        return W, H

    def get_values(self, n_csr=None, n_csc=None):
        """(csr_val, csc_val) as currently held, in the caller's order (CCD++: the residual)."""
        a = np.empty(self.nnz if n_csr is None else n_csr, np.float32)
        b = np.empty(self.nnz if n_csc is None else n_csc, np.float32)
        _check(lib().mf_session_get_values(self.h, a.ctypes.data, b.ctypes.data))
        return a, b

    def iterate(self, n=1, want_stats=True):
        st = (mf_iter_stats * max(n, 1))()
        fn = lib().mf_session_als_iterate if self.als else lib().mf_session_ccdpp_iterate
        _check(fn(self.h, n, C.addressof(st) if want_stats else None))
        return _stats_list(st, n) if want_stats else None

    def rmse(self):
        r = C.c_double()
        _check(lib().mf_session_rmse(self.h, C.byref(r)))
        return r.value

    def predict(self, row, col):
        """w_i . h_j of the current factors for arbitrary pairs: float64 array (mf_session_predict)."""
        row = np.ascontiguousarray(row, np.uint32)
        col = np.ascontiguousarray(col, np.uint32)
        out = np.zeros(len(row), np.float64)
        _check(lib().mf_session_predict(self.h, len(row), row.ctypes.data, col.ctypes.data, out.ctypes.data))
        return out

    def rank_stats(self):
        """Per-rank report of the last CCD++ outer iteration (mf_session_rank_stats): dict of seconds / rmse (None unless the
        session was created with verbose and do_predict) and inner_iters."""
        k = self.k
        inner = np.zeros(k, np.int32)
        sec = np.zeros(k, np.float64)
        rm = np.zeros(k, np.float64)
        if lib().mf_session_rank_stats(self.h, sec.ctypes.data, rm.ctypes.data, inner.ctypes.data) != 0:
            _check(lib().mf_session_rank_stats(self.h, None, None, inner.ctypes.data))
            return dict(seconds=None, rmse=None, inner_iters=inner)
        return dict(seconds=sec, rmse=rm, inner_iters=inner)

    def last_seconds(self):
        r = C.c_double()
        _check(lib().mf_session_last_seconds(self.h, C.byref(r)))
        return r.value

    def kernel_times(self):
        kt = mf_kernel_times()
        _check(lib().mf_session_kernel_times(self.h, C.byref(kt)))
        return {name: getattr(kt, name) for name, _ in mf_kernel_times._fields_}

    def ccd_solve(self, t, side):
        _check(lib().mf_session_ccd_solve(self.h, t, side))

    def ccd_update(self, t, add):
        _check(lib().mf_session_ccd_update(self.h, t, int(bool(add))))

    def als_half(self, side):
        _check(lib().mf_session_als_half(self.h, side))

    def panel_layout(self, side):
        npad, nit, npan = C.c_int64(), C.c_int64(), C.c_int64()
        _check(lib().mf_session_panel_layout(self.h, side, C.byref(npad), C.byref(nit), C.byref(npan), None, None, None))
        idx16 = np.empty(npad.value, np.uint16)
        val = np.empty(npad.value, np.float32)
        items = np.empty((nit.value, 4), np.uint32)
        _check(lib().mf_session_panel_layout(self.h, side, C.byref(npad), C.byref(nit), C.byref(npan),
                                             idx16.ctypes.data, val.ctypes.data, items.ctypes.data))
        return dict(n_padded=npad.value, n_items=nit.value, n_panels=npan.value, idx16=idx16, val=val, items=items)


def predict_pairs(W, H, row, col, device=0):
    """Predict-only path for a saved model (mf_predict_pairs): W [rows, k], H [cols, k] row-major -> float64 predictions."""
    W = np.ascontiguousarray(W, np.float32)
    H = np.ascontiguousarray(H, np.float32)
    row = np.ascontiguousarray(row, np.uint32)
    col = np.ascontiguousarray(col, np.uint32)
    out = np.zeros(len(row), np.float64)
    _check(lib().mf_predict_pairs(W.ctypes.data, H.ctypes.data, W.shape[0], H.shape[0], W.shape[1], len(row), row.ctypes.data,
                              col.ctypes.data, out.ctypes.data, device))
    return out


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().mf_dist_unique_id(C.addressof(buf)))
    return bytes(buf.raw)


def degree_bins(ptr, device=0):
    ptr = _keep(ptr, np.uint32)
    a = np.zeros(33, np.uint64)
    b = np.zeros(33, np.uint64)
    _check(lib().mf_degree_bins(len(ptr) - 1, _ptr(ptr), a.ctypes.data, b.ctypes.data, device))
    return a, b


def partition(ptr, P, device=0):
    ptr = _keep(ptr, np.uint32)
    out = np.zeros(P + 1, np.int64)
    _check(lib().mf_partition(len(ptr) - 1, _ptr(ptr), P, out.ctypes.data, device))
    return out


def build_csr_csc(rows, cols, coo_row, coo_col, coo_val, device=0):
    r, c, v = _keep(coo_row, np.uint32), _keep(coo_col, np.uint32), _keep(coo_val, np.float32)
    nnz = len(v)
    csr = (np.empty(rows + 1, np.uint32), np.empty(nnz, np.uint32), np.empty(nnz, np.float32))
    csc = (np.empty(cols + 1, np.uint32), np.empty(nnz, np.uint32), np.empty(nnz, np.float32))
    _check(lib().mf_build_csr_csc(rows, cols, nnz, _ptr(r), _ptr(c), _ptr(v), csr[0].ctypes.data, csr[1].ctypes.data,
                                  csr[2].ctypes.data, csc[0].ctypes.data, csc[1].ctypes.data, csc[2].ctypes.data, device))
    return csr, csc
