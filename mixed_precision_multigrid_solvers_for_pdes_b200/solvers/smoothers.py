"""Drop-in smoothers (reference solvers/smoothers.py) backed by libmgb200 kernels.

``smooth(grid, operator, u, rhs, num_iterations)`` returns a NEW array and leaves its input
untouched, like the reference (smoothers.py:138); the ``operator`` argument is accepted and
ignored, as in the reference (every smoother relaxes -lap_h(u) = rhs).  ``kind`` is what the
cycle engine dispatches on."""
from __future__ import annotations

import torch

from .. import ops
from ..device import like_input, to_device
from .base import IterativeSolver


class _DeviceSmoother(IterativeSolver):
    kind = "custom"

    def smooth(self, grid, operator, u, rhs, num_iterations: int = 1):
        if tuple(u.shape) != tuple(grid.shape) or tuple(rhs.shape) != tuple(grid.shape):
            raise ValueError(f"Field shape {tuple(u.shape)} doesn't match grid shape {grid.shape}")
        du, was_np = to_device(u)
        out = torch.empty_strided(du.shape, du.stride(), dtype=du.dtype, device=du.device) if du.stride(1) == 1 \
            else torch.empty_like(du)
        out.copy_(du)
        df, _ = to_device(rhs, dtype=out.dtype)
        self._smooth_device_(grid, out, df, num_iterations)
        return like_input(out, was_np)


class JacobiSmoother(_DeviceSmoother):
    kind = "jacobi"

    def __init__(self, max_iterations: int = 1000, tolerance: float = 1e-8, relaxation_parameter: float = 2.0 / 3.0,
                 verbose: bool = False):
        super().__init__(max_iterations, tolerance, relaxation_parameter, verbose, "Jacobi")

    def _smooth_device_(self, grid, u, rhs, num_iterations: int):
        return ops.smooth_jacobi_(u, rhs, grid.hx, grid.hy, self.omega, num_iterations)


class WeightedJacobiSmoother(JacobiSmoother):
    def __init__(self, max_iterations: int = 1000, tolerance: float = 1e-8, verbose: bool = False):
        super().__init__(max_iterations, tolerance, 4.0 / 5.0, verbose)
        self.name = "WeightedJacobi"


class GaussSeidelSmoother(_DeviceSmoother):
    """solvers/smoothers.py:89-207.  ``red_black=True`` is the smoother this build is about (fused, temporally blocked,
    HBM-bound).  ``red_black=False`` -- the reference's default, and what `MultigridSolver.setup(smoother=None)` picks
    (multigrid.py:112-117) -- is the lexicographic sweep, run as a skewed wavefront pipelined over all warps of the grid
    (`mg_smooth_lexgs` -> lexgs_pipe_kernel: one warp per 32 rows, each row one step behind the row above): bit-exact
    with the sequential loop at any size.  A sweep is still a dependency chain of about ny + 2 * nx steps (0.6 ms at
    1025^2, 12 ms at 16385^2; a red-black sweep of 16385^2 takes 0.25 ms), so pass a red-black smoother whenever the
    ordering is free."""

    def __init__(self, max_iterations: int = 1000, tolerance: float = 1e-8, relaxation_parameter: float = 1.0,
                 verbose: bool = False, red_black: bool = False):
        super().__init__(max_iterations, tolerance, relaxation_parameter, verbose, "Gauss-Seidel")
        self.red_black = red_black

    @property
    def kind(self) -> str:
        return "rbgs" if self.red_black else "lexgs"

    def _smooth_device_(self, grid, u, rhs, num_iterations: int):
        if self.red_black:
            return ops.smooth_rbgs_(u, rhs, grid.hx, grid.hy, self.omega, num_iterations)
        return ops.smooth_lexgs_(u, rhs, grid.hx, grid.hy, self.omega, num_iterations, "forward")


class SymmetricGaussSeidelSmoother(GaussSeidelSmoother):
    def __init__(self, max_iterations: int = 1000, tolerance: float = 1e-8, relaxation_parameter: float = 1.0,
                 verbose: bool = False):
        super().__init__(max_iterations, tolerance, relaxation_parameter, verbose)
        self.name = "SymmetricGauss-Seidel"

    @property
    def kind(self) -> str:
        return "sgs"

    def _smooth_device_(self, grid, u, rhs, num_iterations: int):
        return ops.smooth_lexgs_(u, rhs, grid.hx, grid.hy, self.omega, num_iterations, "symmetric")
