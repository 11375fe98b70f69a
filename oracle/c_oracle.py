"""TEST INFRASTRUCTURE — ctypes front-end of oracle/liboracle.so (plain-C, OpenMP, chunked CPU oracle).

See oracle/oracle.c for what each entry point restates and which reference lines it follows.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        c_f, c_d, c_i32 = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_num_threads.argtypes = [ctypes.c_int]
        L.oracle_supcon_fwd_bwd.argtypes = [c_f, c_i32, ctypes.c_int64, ctypes.c_int64, ctypes.c_double, ctypes.c_int,
                                            c_d, c_d, c_d, c_f]
        L.oracle_iic_raw_joint.argtypes = [c_f, c_f] + [ctypes.c_int] * 6 + [c_d]
        L.oracle_iic_epilogue.argtypes = [c_d, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                          ctypes.c_double, c_d, c_d, c_d]
        L.oracle_iic_input_grads.argtypes = [c_f, c_f, c_d] + [ctypes.c_int] * 5 + [c_f, c_f]
        L.oracle_iic_fwd_bwd.argtypes = [c_f, c_f] + [ctypes.c_int] * 6 + [ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                                                         c_d, c_d, c_f, c_f]
        _LIB = L
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n: int):
    """explicit OpenMP team size (torchrun forces OMP_NUM_THREADS=1 on its ranks)"""
    lib().oracle_set_num_threads(int(n))


def supcon_fwd_bwd(z, labels, t=0.07, prec=1, want_grad=True):
    """z [N,d] float32 (both views stacked), labels [N] int32 (tiled).  prec 1 = float64 accumulation."""
    z = np.ascontiguousarray(z, dtype=np.float32)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    N, d = z.shape
    loss = np.zeros(1); lse = np.zeros(N); cnt = np.zeros(N)
    dz = np.zeros((N, d), dtype=np.float32) if want_grad else None
    rc = lib().oracle_supcon_fwd_bwd(_p(z, ctypes.c_float), _p(labels, ctypes.c_int32), N, d, float(t), int(prec),
                                     _p(loss, ctypes.c_double), _p(lse, ctypes.c_double), _p(cnt, ctypes.c_double),
                                     _p(dz, ctypes.c_float))
    assert rc == 0, rc
    return dict(loss=float(loss[0]), grad=dz, row_lse=lse, row_count=cnt)


def iic_fwd_bwd(x, y, padding, symmetric=False, lamda=1.0, eps=1e-5, prec=1, want_grad=True):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    B, K, H, W = x.shape
    loss = np.zeros(1); P00 = np.zeros((K, K))
    dx = np.zeros_like(x) if want_grad else None
    dy = np.zeros_like(y) if want_grad else None
    rc = lib().oracle_iic_fwd_bwd(_p(x, ctypes.c_float), _p(y, ctypes.c_float), B, K, H, W, int(padding), int(symmetric),
                                  float(lamda), float(eps), int(prec), _p(loss, ctypes.c_double),
                                  _p(P00, ctypes.c_double), _p(dx, ctypes.c_float), _p(dy, ctypes.c_float))
    assert rc == 0, rc
    return dict(loss=float(loss[0]), joint=P00, grad_x=dx, grad_y=dy)


def iic_raw_joint(x, y, padding, prec=1):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    B, K, H, W = x.shape
    T = 2 * padding + 1
    J = np.zeros((K, K, T, T))
    lib().oracle_iic_raw_joint(_p(x, ctypes.c_float), _p(y, ctypes.c_float), B, K, H, W, int(padding), int(prec),
                               _p(J, ctypes.c_double))
    return J
