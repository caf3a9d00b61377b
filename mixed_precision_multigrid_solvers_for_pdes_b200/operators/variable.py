"""Variable-coefficient operator  A u = -div(a grad u) + shift*u  and its red-black smoother.

The reference lists the problem class (README.md:175) but has no operator for it, so this follows the textbook
conservative 5-point discretisation with arithmetic-mean face coefficients of a nodal field ``a`` (see
csrc/mg_varcoef.cu).  Both classes plug into the unchanged ``MultigridSolver`` through its operator / smoother
protocol: the coefficient field of a coarse level is the injection of the next finer one and is looked up by grid
shape, because the reference hands the SAME operator object to every level (multigrid.py:163)."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from .. import _lib
from ..device import code, empty_field, ld, like_input, stream_ptr, to_device
from ..solvers.base import IterativeSolver
from .base import BaseOperator


class VariableCoefficientOperator(BaseOperator):
    def __init__(self, coefficient_field, shift: float = 0.0):
        super().__init__(f"VariableCoefficient(shift={shift})")
        a, _ = to_device(coefficient_field)
        if float(a.min().item()) <= 0.0:
            raise ValueError("the diffusion coefficient must be positive")
        self.shift = float(shift)
        self._a: Dict[Tuple[int, int, torch.dtype], torch.Tensor] = {(a.shape[0], a.shape[1], a.dtype): a}
        self._fine = a

    def can_apply(self, grid) -> bool:
        return grid.nx >= 3 and grid.ny >= 3

    def coefficients(self, nx: int, ny: int, dtype) -> torch.Tensor:
        """Nodal coefficient field for an (nx, ny) level: injection from the fine field, cached per shape/dtype."""
        key = (nx, ny, dtype)
        t = self._a.get(key)
        if t is None:
            fx, fy = self._fine.shape
            sx, sy = (fx - 1) // (nx - 1), (fy - 1) // (ny - 1)
            if (nx - 1) * sx != fx - 1 or (ny - 1) * sy != fy - 1 or sx != sy or sx & (sx - 1):
                raise ValueError(f"grid {nx}x{ny} is not a coarsening of the coefficient field {fx}x{fy}")
            t = empty_field(nx, ny, dtype, self._fine.device, zero=False)
            t.copy_(self._fine[::sx, ::sy])
            self._a[key] = t
        return t

    def _run(self, grid, u, f, apply_only: bool):
        if tuple(u.shape) != tuple(grid.shape):
            raise ValueError(f"Field shape {tuple(u.shape)} doesn't match grid shape {grid.shape}")
        du, was_np = to_device(u)
        df = None if apply_only else to_device(f, dtype=du.dtype)[0]
        a = self.coefficients(grid.nx, grid.ny, du.dtype)
        out = empty_field(grid.nx, grid.ny, du.dtype, du.device, zero=False)
        _lib.call("mg_varcoef_residual", du.data_ptr(), df.data_ptr() if df is not None else None, a.data_ptr(),
                  out.data_ptr(), grid.nx, grid.ny, ld(du), ld(df) if df is not None else 0, ld(a), ld(out), grid.hx,
                  grid.hy, self.shift, 1 if apply_only else 0, code(du.dtype), stream_ptr())
        return like_input(out, was_np)

    def apply(self, grid, field=None):
        return self._run(grid, grid.values if field is None else field, None, True)

    def residual(self, grid, u, f):
        r = self._run(grid, u, f, False)
        grid.residual = r.copy() if isinstance(r, np.ndarray) else r
        return r


class VariableCoefficientSmoother(IterativeSolver):
    """Red-black Gauss-Seidel for ``VariableCoefficientOperator`` (pass it as the `smoother` AND, for the coarsest
    level, as the `coarse_solver` of MultigridSolver.setup)."""
    kind = "rbgs_var"  # the cycle engine fuses it (mg_vcv_* passes) together with VariableCoefficientOperator

    def __init__(self, operator: VariableCoefficientOperator, max_iterations: int = 1000, tolerance: float = 1e-8,
                 relaxation_parameter: float = 1.0, verbose: bool = False):
        super().__init__(max_iterations, tolerance, relaxation_parameter, verbose, "VariableCoefficientRBGS")
        self.op = operator

    def _smooth_device_(self, grid, u, rhs, num_iterations: int):
        a = self.op.coefficients(grid.nx, grid.ny, u.dtype)
        _lib.call("mg_varcoef_smooth_rbgs", u.data_ptr(), rhs.data_ptr(), a.data_ptr(), grid.nx, grid.ny, ld(u), ld(rhs),
                  ld(a), grid.hx, grid.hy, self.op.shift, self.omega, num_iterations, code(u.dtype), stream_ptr())
        return u

    def smooth(self, grid, operator, u, rhs, num_iterations: int = 1):
        du, was_np = to_device(u)
        out = torch.empty_strided(du.shape, du.stride(), dtype=du.dtype, device=du.device)
        out.copy_(du)
        df, _ = to_device(rhs, dtype=out.dtype)
        self._smooth_device_(grid, out, df, num_iterations)
        return like_input(out, was_np)
