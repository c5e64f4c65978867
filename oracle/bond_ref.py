"""ctypes loader of oracle/_build/libbond_ref.so (C restatement of Loss_Grad_KLD + update_caches!).
Bench / test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbond_ref.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _lib = C.CDLL(_SO)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        _lib.loss_grad_kld_ref.restype = C.c_double
        _lib.loss_grad_kld_ref.argtypes = [dp, dp, dp, dp, dp, C.c_int64, C.c_int, C.c_int, C.c_int, ip, C.c_int, C.c_int,
                                           C.c_int, dp]
        _lib.env_update_ref.restype = None
        _lib.env_update_ref.argtypes = [dp, dp, dp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, dp]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def loss_grad_kld(B, L, R, xl, xr, counts, train_sep=False, nthreads=1):
    lib = load()
    B = np.asfortranarray(B, dtype=np.float64)
    L, R, xl, xr = (np.ascontiguousarray(a, dtype=np.float64) for a in (L, R, xl, xr))
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    G = np.empty(B.shape, dtype=np.float64, order="F")
    lo = lib.loss_grad_kld_ref(_p(B), _p(L), _p(R), _p(xl), _p(xr), xl.shape[0], xl.shape[1], L.shape[1], R.shape[1],
                               counts.ctypes.data_as(C.POINTER(C.c_int64)), len(counts), int(train_sep), int(nthreads), _p(G))
    return lo, np.ascontiguousarray(G)


def env_update(x, env, W_flat, chi_new, nthreads=1):
    """W_flat: [s + d*(a + chi*k)]"""
    lib = load()
    x, env = np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(env, dtype=np.float64)
    W_flat = np.ascontiguousarray(W_flat, dtype=np.float64)
    out = np.empty((x.shape[0], chi_new))
    lib.env_update_ref(_p(x), _p(env), _p(W_flat), x.shape[0], x.shape[1], env.shape[1], chi_new, int(nthreads), _p(out))
    return out
