"""ctypes front-end of oracle/mg_oracle.c (the plain-C, OpenMP restatement of the reference hot
path)  --  TEST INFRASTRUCTURE ONLY, same rules as np_oracle.py.

Exposes the same function names and signatures as ``np_oracle`` so that
``OracleMultigrid(..., ops=c_oracle)`` runs the reference recursion on the C kernels; that is the
large-grid CPU oracle and the timed CPU baseline ("port", all host threads)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmgoracle.so")
_lib = None
_RESTRICT = {"full_weighting": 0, "injection": 1, "half_weighting": 2}
_PROLONG = {"bilinear": 0, "injection": 1}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_sumsq_f64.restype = C.c_double
        _lib.orc_sumsq_f32.restype = C.c_double
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def _sfx(a: np.ndarray) -> str:
    if a.dtype == np.float64:
        return "f64"
    if a.dtype == np.float32:
        return "f32"
    raise TypeError(a.dtype)


def _p(a: np.ndarray):
    assert a.flags.c_contiguous
    return a.ctypes.data_as(C.c_void_p)


def _d(x):
    return C.c_double(float(x))


def apply_laplacian(u, hx, hy, coefficient=1.0):
    u = np.ascontiguousarray(u)
    out = np.empty_like(u)
    getattr(lib(), "orc_apply_" + _sfx(u))(_p(u), _p(out), u.shape[0], u.shape[1], _d(hx), _d(hy), _d(coefficient))
    return out


def residual(u, f, hx, hy, coefficient=1.0):
    u = np.ascontiguousarray(u)
    f = np.ascontiguousarray(f)
    if u.dtype != f.dtype:  # NumPy promotion of f - A u (only met on mixed-dtype levels)
        dt = np.result_type(u.dtype, f.dtype)
        return f.astype(dt) - apply_laplacian(u, hx, hy, coefficient).astype(dt)
    r = np.empty_like(u)
    getattr(lib(), "orc_residual_" + _sfx(u))(_p(u), _p(f), _p(r), u.shape[0], u.shape[1], _d(hx), _d(hy),
                                              _d(coefficient))
    return r


def _smooth(name, u, rhs, hx, hy, omega, sweeps):
    out = np.array(u, copy=True, order="C")
    rhs = np.ascontiguousarray(rhs, dtype=out.dtype)
    getattr(lib(), f"orc_{name}_" + _sfx(out))(_p(out), _p(rhs), out.shape[0], out.shape[1], _d(hx), _d(hy),
                                               _d(omega), int(sweeps))
    return out


def rbgs_smooth(u, rhs, hx, hy, omega=1.0, sweeps=1):
    return _smooth("rbgs", u, rhs, hx, hy, omega, sweeps)


def lexgs_smooth(u, rhs, hx, hy, omega=1.0, sweeps=1):
    return _smooth("lexgs", u, rhs, hx, hy, omega, sweeps)


def jacobi_smooth(u, rhs, hx, hy, omega=2.0 / 3.0, sweeps=1):
    out = np.array(u, copy=True, order="C")
    rhs = np.ascontiguousarray(rhs, dtype=out.dtype)
    tmp = np.empty_like(out)
    getattr(lib(), "orc_jacobi_" + _sfx(out))(_p(out), _p(tmp), _p(rhs), out.shape[0], out.shape[1], _d(hx), _d(hy),
                                              _d(omega), int(sweeps))
    return out


def restrict(field, method="full_weighting", out_dtype=None):
    field = np.ascontiguousarray(field)
    nf, mf = field.shape
    c = np.empty(((nf - 1) // 2 + 1, (mf - 1) // 2 + 1), dtype=field.dtype)
    getattr(lib(), "orc_restrict_" + _sfx(field))(_p(field), _p(c), nf, mf, _RESTRICT[method])
    return c if out_dtype is None or np.dtype(out_dtype) == c.dtype else c.astype(out_dtype)


def prolong(field, method="bilinear", out_dtype=None):
    src = np.ascontiguousarray(field if out_dtype is None else field.astype(out_dtype, copy=False))
    nc, mc = src.shape
    f = np.empty((2 * (nc - 1) + 1, 2 * (mc - 1) + 1), dtype=src.dtype)
    getattr(lib(), "orc_prolong_" + _sfx(src))(_p(src), _p(f), nc, mc, _PROLONG[method])
    return f


def sumsq(x) -> float:
    x = np.ascontiguousarray(x)
    return float(getattr(lib(), "orc_sumsq_" + _sfx(x))(_p(x), x.shape[0], x.shape[1]))


def l2_norm(field, hx, hy) -> float:
    return float(np.sqrt(hx * hy * sumsq(field)))
