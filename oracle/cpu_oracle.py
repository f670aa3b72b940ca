"""ctypes front end of oracle/sd_oracle.c (built by oracle/Makefile into oracle/_build/).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each wrapper names the reference routine the
C function restates; the C source carries the file:line citations.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsdoracle.so")
_lib = None

_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (no GPU needed)."""
    src = os.path.join(_HERE, "sd_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.sdo_num_threads.restype = C.c_int
    return _lib


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _q(idx, n):
    q = np.arange(n, dtype=np.int64) if idx is None else np.ascontiguousarray(idx, dtype=np.int64)
    return q


def num_threads() -> int:
    return int(lib().sdo_num_threads())


def set_num_threads(k: int) -> None:
    lib().sdo_set_num_threads(C.c_int(int(k)))


def band_counts_enum(X, queries=None, j=2, relax=False):
    """Reference algorithm (_functional.py:238-253 + _containment.py:68-80) in C.  X is [T, n]."""
    X = _c64(X)
    T, n = X.shape
    q = _q(queries, n)
    out = np.zeros(q.size, dtype=np.int64)
    rc = lib().sdo_band_counts_enum(X.ctypes.data_as(_f64p), C.c_int64(T), C.c_int64(n), C.c_int64(n),
                                    q.ctypes.data_as(_i64p), C.c_int64(q.size), C.c_int(j),
                                    C.c_int(int(bool(relax))), out.ctypes.data_as(_i64p))
    if rc:
        raise RuntimeError("sdo_band_counts_enum rc=%d" % rc)
    return out


def mbd_counts_all(X, j=2, want_ranks=False):
    """Closed-form relaxed numerator sum_t [C(n-1,j) - C(b,j) - C(a,j)] for every curve."""
    X = _c64(X)
    T, n = X.shape
    out = np.zeros(n, dtype=np.int64)
    rb = ra = None
    pb = pa = None
    if want_ranks:
        rb = np.zeros((T, n), dtype=np.int32)
        ra = np.zeros((T, n), dtype=np.int32)
        pb, pa = rb.ctypes.data_as(_i32p), ra.ctypes.data_as(_i32p)
    rc = lib().sdo_mbd_counts_all(X.ctypes.data_as(_f64p), C.c_int64(T), C.c_int64(n), C.c_int64(n),
                                  C.c_int(j), out.ctypes.data_as(_i64p), pb, pa)
    if rc:
        raise RuntimeError("sdo_mbd_counts_all rc=%d" % rc)
    return (out, rb, ra) if want_ranks else out


def bd_counts(X, queries=None, j=2):
    """Closed-form strict numerator (#J-subsets that never jointly violate)."""
    X = _c64(X)
    T, n = X.shape
    q = _q(queries, n)
    out = np.zeros(q.size, dtype=np.int64)
    rc = lib().sdo_bd_counts(X.ctypes.data_as(_f64p), C.c_int64(T), C.c_int64(n), C.c_int64(n),
                             q.ctypes.data_as(_i64p), C.c_int64(q.size), C.c_int(j),
                             out.ctypes.data_as(_i64p))
    if rc:
        raise RuntimeError("sdo_bd_counts rc=%d" % rc)
    return out


def l1_depth(P, queries=None):
    """_L1_depth (_pointcloud.py:125-150), sequential float64 accumulation."""
    P = _c64(P)
    n, d = P.shape
    q = _q(queries, n)
    out = np.zeros(q.size, dtype=np.float64)
    rc = lib().sdo_l1_depth(P.ctypes.data_as(_f64p), C.c_int64(n), C.c_int64(d), q.ctypes.data_as(_i64p),
                            C.c_int64(q.size), out.ctypes.data_as(_f64p))
    if rc:
        raise RuntimeError("sdo_l1_depth rc=%d" % rc)
    return out


def in_simplex(V, p, tol=1e-7) -> bool:
    """_is_in_simplex (_containment.py:138-176) restated with an explicit absolute tolerance."""
    V = _c64(V)
    p = _c64(p)
    d = p.size
    assert V.shape == (d + 1, d)
    return bool(lib().sdo_in_simplex(V.ctypes.data_as(_f64p), C.c_int(d), p.ctypes.data_as(_f64p),
                                     C.c_double(tol)))


def simplicial_counts(P, queries=None, tol=1e-7):
    """_pointwisedepth 'simplex' numerator (_pointcloud.py:44-56)."""
    P = _c64(P)
    n, d = P.shape
    q = _q(queries, n)
    out = np.zeros(q.size, dtype=np.int64)
    rc = lib().sdo_simplicial_counts(P.ctypes.data_as(_f64p), C.c_int64(n), C.c_int(d),
                                     q.ctypes.data_as(_i64p), C.c_int64(q.size), C.c_double(tol),
                                     out.ctypes.data_as(_i64p))
    if rc:
        raise RuntimeError("sdo_simplicial_counts rc=%d" % rc)
    return out


def simplex_depth_counts(F, queries=None, relax=False, tol=1e-7):
    """_simplex_depth numerator (_functional.py:257-286).  F is [N, T, d]."""
    F = _c64(F)
    N, T, d = F.shape
    q = _q(queries, N)
    out = np.zeros(q.size, dtype=np.int64)
    rc = lib().sdo_simplex_depth_counts(F.ctypes.data_as(_f64p), C.c_int64(N), C.c_int64(T), C.c_int(d),
                                        q.ctypes.data_as(_i64p), C.c_int64(q.size),
                                        C.c_int(int(bool(relax))), C.c_double(tol),
                                        out.ctypes.data_as(_i64p))
    if rc:
        raise RuntimeError("sdo_simplex_depth_counts rc=%d" % rc)
    return out


def oja(P, hull_volume, queries=None, pool=None):
    """_oja_depth (_pointcloud.py:176-204); hull_volume is supplied by the caller."""
    P = _c64(P)
    n, d = P.shape
    q = _q(queries, n)
    pl = _q(pool, n)
    out = np.zeros(q.size, dtype=np.float64)
    rc = lib().sdo_oja(P.ctypes.data_as(_f64p), C.c_int64(n), C.c_int(d), q.ctypes.data_as(_i64p),
                       C.c_int64(q.size), pl.ctypes.data_as(_i64p), C.c_int64(pl.size),
                       C.c_double(hull_volume), out.ctypes.data_as(_f64p))
    if rc:
        raise RuntimeError("sdo_oja rc=%d" % rc)
    return out


def triangle_counts_arcs(P, queries=None, tol=1e-7):
    """2-D simplicial numerator by O(n^2) arc counting (independent restatement of dist(p, triangle) <= tol)."""
    P = _c64(P)
    n, d = P.shape
    assert d == 2
    q = _q(queries, n)
    out = np.zeros(q.size, dtype=np.int64)
    rc = lib().sdo_triangle_counts_arcs(P.ctypes.data_as(_f64p), C.c_int64(n), C.c_int64(2), C.c_int64(1),
                                        q.ctypes.data_as(_i64p), C.c_int64(q.size), C.c_double(tol),
                                        out.ctypes.data_as(_i64p))
    if rc:
        raise RuntimeError("sdo_triangle_counts_arcs rc=%d" % rc)
    return out


def simplex2_relaxed_counts_arcs(F, queries=None, tol=1e-7):
    """Relaxed multivariate simplex numerator (d = 2) by O(n^2) arc counting per (query, time point)."""
    F = _c64(F)
    N, T, d = F.shape
    assert d == 2
    q = _q(queries, N)
    out = np.zeros(q.size, dtype=np.int64)
    rc = lib().sdo_triangle_counts_arcs(F.ctypes.data_as(_f64p), C.c_int64(N), C.c_int64(2 * T), C.c_int64(T),
                                        q.ctypes.data_as(_i64p), C.c_int64(q.size), C.c_double(tol),
                                        out.ctypes.data_as(_i64p))
    if rc:
        raise RuntimeError("sdo_triangle_counts_arcs rc=%d" % rc)
    return out


def simplex2_strict_fast(F, queries=None, tol=1e-7):
    """Strict multivariate simplex numerator (d = 2) with an exact pre-test per (triple, row): for config-4 sizes."""
    F = _c64(F)
    N, T, d = F.shape
    assert d == 2
    q = _q(queries, N)
    out = np.zeros(q.size, dtype=np.int64)
    rc = lib().sdo_simplex2_strict_fast(F.ctypes.data_as(_f64p), C.c_int64(N), C.c_int64(T), q.ctypes.data_as(_i64p),
                                        C.c_int64(q.size), C.c_double(tol), out.ctypes.data_as(_i64p))
    if rc:
        raise RuntimeError("sdo_simplex2_strict_fast rc=%d" % rc)
    return out
