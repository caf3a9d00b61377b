"""Drop-in ``LaplacianOperator`` (reference operators/laplacian.py:15-158) running on the GPU.

``apply``/``residual`` accept NumPy arrays (copied to the device and back, returning NumPy like
the reference) or CUDA tensors (returned as CUDA tensors, no host traffic).  The arithmetic is
``mg_apply_laplacian`` / ``mg_residual``: coefficient*((u_e+u_w)/hx^2 + (u_n+u_s)/hy^2 - u*(2/hx^2+2/hy^2))
on the interior, 0 on the boundary, so ``residual`` equals f on the boundary exactly as in the
reference."""
from __future__ import annotations

import numpy as np

from .. import ops
from ..device import like_input, to_device
from .base import BaseOperator


class LaplacianOperator(BaseOperator):
    def __init__(self, coefficient: float = 1.0):
        super().__init__(f"Laplacian(coeff={coefficient})")
        self.coefficient = coefficient

    def can_apply(self, grid) -> bool:
        return grid.nx >= 3 and grid.ny >= 3

    def _check(self, grid, field):
        if not self.can_apply(grid):
            raise ValueError(f"Cannot apply Laplacian to grid {grid.shape}")
        if tuple(field.shape) != tuple(grid.shape):
            raise ValueError(f"Field shape {tuple(field.shape)} doesn't match grid shape {grid.shape}")

    def apply(self, grid, field=None):
        if field is None:
            field = grid.values
        self._check(grid, field)
        d, was_np = to_device(field)
        return like_input(ops.apply_laplacian(d, grid.hx, grid.hy, self.coefficient), was_np)

    def apply_stencil(self, grid, field, i: int, j: int) -> float:
        if i < 1 or i >= grid.nx - 1 or j < 1 or j >= grid.ny - 1:
            raise ValueError(f"Point ({i}, {j}) is not an interior point")
        f = lambda a, b: float(field[a, b])  # noqa: E731  (single-point host evaluation, not a hot path)
        return self.coefficient * ((f(i + 1, j) + f(i - 1, j)) / grid.hx ** 2 + (f(i, j + 1) + f(i, j - 1)) / grid.hy ** 2
                                   - f(i, j) * (2.0 / grid.hx ** 2 + 2.0 / grid.hy ** 2))

    def residual(self, grid, u, f):
        """r = f - A u; also stored on ``grid.residual`` like the reference (laplacian.py:121)."""
        self._check(grid, u)
        if tuple(f.shape) != tuple(grid.shape):
            raise ValueError(f"Field shape {tuple(f.shape)} doesn't match grid shape {grid.shape}")
        du, was_np = to_device(u)
        df, _ = to_device(f, dtype=du.dtype)
        r = like_input(ops.residual(du, df, grid.hx, grid.hy, self.coefficient), was_np)
        grid.residual = r.copy() if was_np else r
        return r

    def eigenvalues_1d(self, n: int, h: float) -> np.ndarray:
        k = np.arange(1, n + 1)
        return self.coefficient * (-4.0 / h ** 2) * np.sin(k * np.pi / (2 * (n + 1))) ** 2

    def condition_number(self, grid) -> float:
        ev = self.eigenvalues_1d(min(grid.nx - 2, grid.ny - 2), max(grid.hx, grid.hy))
        return float(np.abs(ev[-1] / ev[0]))


class HelmholtzOperator(LaplacianOperator):
    """coefficient * lap_h(u) + shift * u  (shift >= 0).  With coefficient = -1 this is -lap + lambda, the operator of
    an implicit heat step: (I - alpha*dt*lap) u = rhs  <=>  (-lap + 1/(alpha*dt)) u = rhs/(alpha*dt)
    (docs/methodology.md:710; the reference has the shifted stencil only inside applications/heat_equation.py:459-497,
    solved there with plain Gauss-Seidel).  The smoothers relax (-lap_h + shift) u = rhs when the cycle engine sees
    this operator."""

    def __init__(self, coefficient: float = -1.0, shift: float = 0.0):
        super().__init__(coefficient)
        if shift < 0:
            raise ValueError("shift must be >= 0")
        self.shift = float(shift)
        self.name = f"Helmholtz(coeff={coefficient}, shift={shift})"

    def apply(self, grid, field=None):
        if field is None:
            field = grid.values
        self._check(grid, field)
        d, was_np = to_device(field)
        out = ops.apply_laplacian(d, grid.hx, grid.hy, self.coefficient)
        if self.shift:
            out[1:-1, 1:-1] += self.shift * d[1:-1, 1:-1]
        return like_input(out, was_np)

    def residual(self, grid, u, f):
        self._check(grid, u)
        du, was_np = to_device(u)
        df, _ = to_device(f, dtype=du.dtype)
        r = like_input(ops.residual(du, df, grid.hx, grid.hy, self.coefficient, shift=self.shift), was_np)
        grid.residual = r.copy() if was_np else r
        return r
