"""ctypes binding of libmgb200.so (include/mgb200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, an
exception is raised.  The product path never routes through ``oracle/`` or any CPU code.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmgb200.so")

F32, F64 = 0, 1
RESTRICT = {"full_weighting": 0, "injection": 1, "half_weighting": 2}
PROLONG = {"bilinear": 0, "injection": 1}

_i, _l, _d, _p = C.c_int, C.c_int64, C.c_double, C.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPE); mirrors include/mgb200.h
SIGNATURES = {
    "mg_abi_version": [],
    "mg_status_string": [_i],
    "mg_device_sm_count": [],
    "mg_launch_count": [],
    "mg_read_doubles": [_p, _p, _i, _p],
    "mg_apply_laplacian": [_p, _p, _i, _i, _l, _l, _d, _d, _d, _i, _p],
    "mg_residual": [_p, _p, _p, _i, _i, _l, _l, _l, _d, _d, _d, _i, _i, _p],
    "mg_smooth_rbgs": [_p, _p, _i, _i, _l, _l, _d, _d, _d, _i, _i, _p],
    "mg_smooth_jacobi": [_p, _p, _p, _i, _i, _l, _l, _d, _d, _d, _i, _i, _p],
    "mg_smooth_lexgs": [_p, _p, _i, _i, _l, _l, _d, _d, _d, _i, _i, _i, _p],
    "mg_coarse_solve_lexgs": [_p, _p, _i, _i, _l, _l, _d, _d, _d, _d, _d, _i, _p, _i, _p],
    "mg_restrict": [_p, _p, _i, _i, _l, _l, _i, _i, _i, _p],
    "mg_prolong": [_p, _p, _i, _i, _l, _l, _i, _i, _i, _i, _p],
    "mg_sumsq": [_p, _i, _i, _l, _i, _p, _p, _p],
    "mg_sumsq_workspace_doubles": [],
    "mg_cast": [_p, _p, _i, _i, _l, _l, _i, _i, _p],
    "mg_axpy": [_d, _p, _p, _i, _i, _l, _l, _i, _i, _p],
    "mg_zero": [_p, _i, _l, _i, _p],
    "mg_zero_ring": [_p, _i, _i, _l, _i, _i, _i, _p],
    "mg_cm_workspace_doubles": [],
    "mg_cm_gs": [_p, _p, _i, _i, _l, _l, _d, _i, _p],
    "mg_cm_residual": [_p, _p, _p, _p, _p, _i, _i, _l, _l, _l, _d, _p],
    "mg_cm_diff_sumsq": [_p, _p, _p, _p, _i, _i, _l, _l, _p],
    "mg_cm_restrict": [_p, _p, _i, _i, _i, _i, _l, _l, _p],
    "mg_cm_prolong_add": [_p, _p, _i, _i, _i, _i, _l, _l, _p],
    "mg_heat_rhs": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _l, _l, _l, _l, _l, _d, _d, _d, _d, _d, _d, _i, _i, _i, _i, _p],
    "mg_fill_sinsin": [_p, _i, _i, _l, _d, _d, _d, _d, _d, _d, _d, _i, _p],
    "mg_maxerr_sinsin": [_p, _i, _i, _l, _d, _d, _d, _d, _d, _d, _d, _i, _p, _p, _p],
    "mg_vc_workspace_doubles": [_i, _i],
    "mg_vc_pass": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _l, _l, _l, _l, _l, _d, _d, _d, _d, _i, _i, _i, _p],
    "mg_vc_defect_pass": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _l, _l, _l, _l, _l, _d, _d, _d, _i, _p],
    "mg_vc_pass_slab": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _l, _l, _l, _l, _l, _d, _d, _d, _d, _i, _i, _i, _i, _i, _d, _p],
    "mg_vc_defect_pass_slab": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _l, _l, _l, _l, _l, _d, _d, _d, _i, _i, _i, _d, _p],
    "mg_varcoef_residual": [_p, _p, _p, _p, _i, _i, _l, _l, _l, _l, _d, _d, _d, _i, _i, _p],
    "mg_varcoef_smooth_rbgs": [_p, _p, _p, _i, _i, _l, _l, _l, _d, _d, _d, _d, _i, _i, _p],
    "mg_varcoef_coarse_solve": [_p, _p, _p, _i, _i, _l, _l, _l, _d, _d, _d, _d, _d, _i, _p, _i, _p],
    "mg_vcv_pass_slab": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _l, _l, _l, _l, _l, _l, _d, _d, _d, _i, _i, _i, _i, _i, _d, _p],
    "mg_vcv_defect_pass_slab": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _l, _l, _l, _l, _l, _l, _d, _d, _i, _i, _i, _d, _p],
    "mg_vc_defect_down_pass_slab": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _l, _l, _l, _l, _l, _l, _l, _d, _d, _d, _d, _i,
                                    _i, _i, _d, _p],
    "mg_residual_h": [_p, _p, _p, _i, _i, _l, _l, _l, _d, _d, _d, _d, _i, _i, _p],
    "mg_smooth_rbgs_h": [_p, _p, _i, _i, _l, _l, _d, _d, _d, _d, _i, _i, _p],
    "mg_coarse_solve_lexgs_h": [_p, _p, _i, _i, _l, _l, _d, _d, _d, _d, _d, _d, _i, _p, _i, _p],
    "mg_small_cycle_smem_bytes": [_i, _i, _i, _i, _i],
    "mg_small_cycle": [_p, _p, _i, _i, _l, _l, _d, _d, _i, _i, _i, _i, _d, _d, _d, _d, _i, _i, _p, _i, _i, _p],
    "mg_vc_smooth": [_p, _p, _p, _i, _i, _l, _l, _l, _d, _d, _d, _i, _i, _i, _p],
    "mg_vc_residual_restrict": [_p, _p, _p, _i, _i, _l, _l, _l, _d, _d, _d, _i, _i, _p],
    "mg_vc_prolong_correct_smooth": [_p, _p, _p, _p, _i, _i, _l, _l, _l, _l, _d, _d, _d, _i, _i, _i, _p],
}
VC_PROLONG, VC_RESTRICT, VC_NORM, VC_LOADER_CPASYNC, VC_NO_STORE, VC_U_ZERO, VC_JACOBI = 1, 2, 4, 16, 32, 64, 128
_RESTYPE = {"mg_status_string": C.c_char_p, "mg_launch_count": C.c_longlong}
_NO_STATUS = {"mg_abi_version", "mg_launch_count", "mg_status_string", "mg_device_sm_count", "mg_sumsq_workspace_doubles",
              "mg_vc_workspace_doubles", "mg_small_cycle_smem_bytes", "mg_cm_workspace_doubles"}

_lib: Optional[C.CDLL] = None


class MGLibraryError(RuntimeError):
    """libmgb200.so is missing or a kernel call returned an error status."""


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MGLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C mixed_precision_multigrid_solvers_for_pdes_b200/csrc`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, C.c_int)
        _lib = lib
    return _lib


def call(name: str, *args) -> int:
    """Invoke an entry point; raise MGLibraryError on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if name not in _NO_STATUS and rc != 0:
        msg = lib.mg_status_string(rc)
        raise MGLibraryError(f"{name} failed with status {rc}: {msg.decode() if msg else '?'}")
    return rc
