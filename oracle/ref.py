"""ctypes binding of oracle/_ref/libmfref.so — the UNMODIFIED reference CPU path
(/root/reference/src/{CCD,ALS,tools,extras}.cpp) behind oracle/ref_harness.cpp.

TEST INFRASTRUCTURE ONLY.  The library is built in the build container by
oracle/Makefile (`make ref`) and travels to the GPU box as a prebuilt file; nothing
here reads /root/reference at run time.  The reference only reads datasets from a
directory in its own on-disk format, so every call takes a dataset directory.
"""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libmfref.so")
_lib = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libmfref.so missing (build it with `make -C oracle ref`)")
        L = C.CDLL(_SO)
        L.ref_initial_col.argtypes = [f32p, C.c_long, C.c_long]
        L.ref_probe.argtypes = [C.c_char_p] + [C.POINTER(C.c_long)] * 6
        L.ref_train.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_double), C.c_char_p, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def initial_col(k, n):
    X = np.empty((k, n), np.float32)
    lib().ref_initial_col(X, k, n)
    return X


def probe(dirname):
    v = [C.c_long() for _ in range(6)]
    lib().ref_probe(dirname.encode(), *[C.byref(x) for x in v])
    return dict(zip(["rows", "cols", "nnz", "nnz_test", "max_row_nnz", "max_col_nnz"], [x.value for x in v]))


_LINE = re.compile(r"iteration num (\d+)\s+(?:rank_time ([\d.]+)\|[\d.]+ s\s+)?update_time ([\d.]+)\|[\d.]+s\s+RMSE=([\d.naif+-]+) time:([\d.]+)s")


def parse_log(text):
    """The reference's per-iteration stdout lines (CCD.cpp:158, ALS.cpp:229) ->
    list of dict(iter, rank_time, update_time, rmse, rmse_time)."""
    out = []
    for m in _LINE.finditer(text):
        out.append(dict(iter=int(m.group(1)), rank_time=float(m.group(2) or 0.0), update_time=float(m.group(3)),
                        rmse=float(m.group(4)), rmse_time=float(m.group(5))))
    return out


def train(dirname, als, k, lam, maxiter, maxinner=1, threads=1, W=None, H=None, want_residual=False):
    """Run ccdr1_OMP / ALS_OMP on the dataset directory exactly as main.cpp does.
    Returns dict(W, H, rmse, seconds, iters=[per-iteration lines], csr_val, csc_val)."""
    info = probe(dirname)
    m, n, nnz = info["rows"], info["cols"], info["nnz"]
    shape_w = (m, k) if als else (k, m)
    shape_h = (n, k) if als else (k, n)
    Wout = np.empty(shape_w, np.float32)
    Hout = np.empty(shape_h, np.float32)
    Win = None if W is None else np.ascontiguousarray(W, np.float32).reshape(shape_w)
    Hin = None if H is None else np.ascontiguousarray(H, np.float32).reshape(shape_h)
    csr_val = np.empty(nnz, np.float32) if want_residual else None
    csc_val = np.empty(nnz, np.float32) if want_residual else None
    rmse = C.c_double()
    secs = C.c_double()
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    with tempfile.NamedTemporaryFile(suffix=".log") as tf:
        lib().ref_train(dirname.encode(), int(als), k, lam, maxiter, maxinner, threads,
                        p(Win), p(Hin), p(Wout), p(Hout), p(csr_val), p(csc_val),
                        C.byref(rmse), tf.name.encode(), C.byref(secs))
        text = open(tf.name).read()
    return dict(W=Wout, H=Hout, rmse=rmse.value, seconds=secs.value, iters=parse_log(text), log=text,
                csr_val=csr_val, csc_val=csc_val)
