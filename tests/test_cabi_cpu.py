"""CPU-side checks of the boundary: the shared library loads and exports every symbol the public
header declares (no kernels are launched here)."""
import ctypes
import os
import re

import pytest

from mixed_precision_multigrid_solvers_for_pdes_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "mgb200.h")).read()
    return sorted(set(re.findall(r"MG_API\s+[\w\s\*]+?\b(mg_\w+)\s*\(", txt)))


def test_library_exists_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "build with __graft_entry__.build()"
    lib = _lib.load()
    assert lib.mg_abi_version() >= 1


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mgb200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound in _lib.py but not declared in the header"


def test_status_strings():
    lib = _lib.load()
    assert lib.mg_status_string(0) == b"ok"
    assert b"align" in lib.mg_status_string(-3)


def test_bad_arguments_rejected_without_gpu():
    # argument validation happens before any CUDA call
    with pytest.raises(_lib.MGLibraryError):
        _lib.call("mg_smooth_rbgs", None, None, 9, 9, 9, 9, 0.1, 0.1, 1.0, 1, _lib.F64, None)
    with pytest.raises(_lib.MGLibraryError):
        _lib.call("mg_restrict", 8, 8, 10, 9, 10, 5, 0, _lib.F64, _lib.F64, None)  # even nxf cannot be coarsened


def test_product_path_fails_loudly_without_gpu():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mixed_precision_multigrid_solvers_for_pdes_b200 import Grid, LaplacianOperator
    with pytest.raises(_lib.MGLibraryError):
        LaplacianOperator(-1.0).apply(Grid(9, 9), np.zeros((9, 9)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mixed_precision_multigrid_solvers_for_pdes_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
