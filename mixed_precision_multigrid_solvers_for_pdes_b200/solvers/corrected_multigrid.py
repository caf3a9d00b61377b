"""Device mirror of the reference's SECONDARY solver, `CorrectedMultigridSolver`
(src/multigrid/solvers/corrected_multigrid.py:24-418): same constructor, `setup_grid_hierarchy`, `solve(initial_guess,
rhs, grid, precision_manager=None)` returning the same dict, `create_test_problem`.  Every level lives in device memory;
the cycle is five kernels of libmgb200 (`mg_cm_*`, csrc/mg_corrected.cu) that keep the reference's operand order, so
the solution is the reference's bit for bit and the residual history agrees to the last digits of the summation tree
(tests/test_gpu_corrected.py against runs of the reference's class, tests/golden/corrected_golden.npz).

The coarsest-level iteration (:366-390) tests two norms after EVERY sweep; they are read back together, one host
synchronisation per sweep (at most `coarse_max_iterations`, 100 by default) -- this solver is the reference's
validation vehicle, not the HBM-bound path (`MixedPrecisionMultigrid`)."""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from ..core.grid import Grid
from ..core.precision import PrecisionLevel
from ..device import require_cuda, stream_ptr
from .base import BaseSolver


class _Level:
    def __init__(self, nx: int, ny: int, h: float, dev):
        self.nx, self.ny, self.h = nx, ny, h
        mk = lambda: torch.zeros((nx, ny), dtype=torch.float64, device=dev)  # noqa: E731
        self.u, self.f, self.r, self.old = mk(), mk(), mk(), mk()


class CorrectedMultigridSolver(BaseSolver):
    def __init__(self, max_levels: int = 4, max_iterations: int = 50, tolerance: float = 1e-8, cycle_type: str = "V",
                 pre_smooth_iterations: int = 2, post_smooth_iterations: int = 2, coarse_tolerance: float = 1e-12,
                 coarse_max_iterations: int = 100, verbose: bool = False, device=None):
        super().__init__(max_iterations, tolerance, verbose, "CorrectedMultigrid")
        self.max_levels, self.cycle_type = max_levels, cycle_type
        self.pre_smooth_iterations, self.post_smooth_iterations = pre_smooth_iterations, post_smooth_iterations
        self.coarse_tolerance, self.coarse_max_iterations = coarse_tolerance, coarse_max_iterations
        self.grids: List[Grid] = []
        self._device_arg = device
        self.device: Optional[torch.device] = None  # resolved at setup: constructing a solver needs no GPU
        self._levels: List[_Level] = []
        self._ss: Optional[torch.Tensor] = None
        self._ws: Optional[torch.Tensor] = None
        self.coarse_sweeps = 0  # GS sweeps spent on the coarsest level in the last solve

    # -- hierarchy (:70-107): n -> max(5, (n-1)//2 + 1), at most max_levels levels, stop at <= 5 -----------------
    def setup_grid_hierarchy(self, fine_grid: Grid) -> None:
        self.grids = [fine_grid]
        g = fine_grid
        for _ in range(1, self.max_levels):
            cx, cy = max(5, (g.nx - 1) // 2 + 1), max(5, (g.ny - 1) // 2 + 1)
            g = Grid(cx, cy, domain=g.domain, dtype=g.dtype)
            self.grids.append(g)
            if cx <= 5 or cy <= 5:
                break
        self.device = require_cuda(self._device_arg)
        dom = fine_grid.domain
        # the reference's kernels use ONE spacing per level, taken from the x extent (:255, :299)
        self._levels = [_Level(g.nx, g.ny, (dom[1] - dom[0]) / (g.nx - 1), self.device) for g in self.grids]
        self._ss = torch.zeros(2, dtype=torch.float64, device=self.device)
        self._ws = torch.empty(2 * _lib.call("mg_cm_workspace_doubles"), dtype=torch.float64, device=self.device)

    # -- kernels ----------------------------------------------------------------------------------------------------
    @staticmethod
    def _ld(t: torch.Tensor) -> int:
        return t.stride(0)

    def _gs(self, L: _Level, sweeps: int) -> None:
        if sweeps > 0:
            _lib.call("mg_cm_gs", L.u.data_ptr(), L.f.data_ptr(), L.nx, L.ny, self._ld(L.u), self._ld(L.f), L.h, sweeps,
                      stream_ptr())

    def _residual(self, L: _Level, store: bool, norm_slot: Optional[int]) -> None:
        half = self._ws.numel() // 2
        _lib.call("mg_cm_residual", L.u.data_ptr(), L.f.data_ptr(), L.r.data_ptr() if store else None,
                  self._ss[norm_slot:norm_slot + 1].data_ptr() if norm_slot is not None else None,
                  self._ws[norm_slot * half:].data_ptr() if norm_slot is not None else None,
                  L.nx, L.ny, self._ld(L.u), self._ld(L.f), self._ld(L.r), L.h, stream_ptr())

    def _residual_norm(self, L: _Level) -> float:
        self._residual(L, False, 0)
        return float(np.sqrt(self._ss[0].item()))

    def _solve_coarsest(self, L: _Level) -> None:
        half = self._ws.numel() // 2
        for _ in range(self.coarse_max_iterations):
            L.old.copy_(L.u)
            self._gs(L, 1)
            self.coarse_sweeps += 1
            self._residual(L, False, 0)
            _lib.call("mg_cm_diff_sumsq", L.u.data_ptr(), L.old.data_ptr(), self._ss[1:2].data_ptr(),
                      self._ws[half:].data_ptr(), L.nx, L.ny, self._ld(L.u), self._ld(L.old), stream_ptr())
            rr, dd = np.sqrt(self._ss.cpu().numpy())
            if rr < self.coarse_tolerance or dd < self.coarse_tolerance:
                break

    def _v_cycle(self, level: int) -> None:
        L = self._levels[level]
        if level == len(self._levels) - 1:
            self._solve_coarsest(L)
            return
        C = self._levels[level + 1]
        self._gs(L, self.pre_smooth_iterations)
        self._residual(L, True, None)
        _lib.call("mg_cm_restrict", L.r.data_ptr(), C.f.data_ptr(), L.nx, L.ny, C.nx, C.ny, self._ld(L.r), self._ld(C.f),
                  stream_ptr())
        C.u.zero_()
        self._v_cycle(level + 1)
        _lib.call("mg_cm_prolong_add", C.u.data_ptr(), L.u.data_ptr(), C.nx, C.ny, L.nx, L.ny, self._ld(C.u), self._ld(L.u),
                  stream_ptr())
        self._gs(L, self.post_smooth_iterations)

    # -- driver (:109-186) ------------------------------------------------------------------------------------------
    def solve(self, initial_guess, rhs, grid: Grid, precision_manager=None) -> Dict[str, Any]:
        if not self.grids or tuple(self.grids[0].shape) != tuple(grid.shape):
            self.setup_grid_hierarchy(grid)
        self.reset()
        self.coarse_sweeps = 0
        L0 = self._levels[0]
        was_tensor = isinstance(rhs, torch.Tensor)
        L0.f.copy_(torch.as_tensor(rhs, dtype=torch.float64))
        L0.u.copy_(torch.as_tensor(initial_guess, dtype=torch.float64))
        for t in (L0.u[0], L0.u[-1], L0.u[:, 0], L0.u[:, -1]):  # _apply_boundary_conditions (:392-397)
            t.zero_()
        residual_norm = self._residual_norm(L0)
        residual_history = [residual_norm]
        iteration = 0
        for iteration in range(1, self.max_iterations + 1):
            t0 = time.time()
            if precision_manager is not None:  # the reference only flips the manager's state (:146-151)
                if precision_manager.should_promote_precision(residual_history, precision_manager.current_precision):
                    precision_manager.current_precision = PrecisionLevel.DOUBLE
            self._v_cycle(0)
            residual_norm = self._residual_norm(L0)
            residual_history.append(residual_norm)
            level = precision_manager.current_precision.value if precision_manager is not None else "double"
            self.history.record_iteration(residual_norm, time.time() - t0, level, 0)
            self.log_iteration(iteration, residual_norm)
            if self.check_convergence(residual_norm, iteration):
                self.converged = True
                break
        self.iterations_performed, self.final_residual = iteration, residual_norm
        u = L0.u.clone() if was_tensor else L0.u.cpu().numpy()
        return {"solution": u, "converged": self.converged, "iterations": iteration, "final_residual": residual_norm,
                "residual_history": residual_history, "convergence_info": self.get_convergence_info()}

    # -- :399-430 ---------------------------------------------------------------------------------------------------
    def create_test_problem(self, grid: Grid, problem_type: str = "manufactured") -> Tuple[np.ndarray, np.ndarray]:
        x = np.linspace(grid.domain[0], grid.domain[1], grid.nx)
        y = np.linspace(grid.domain[2], grid.domain[3], grid.ny)
        X, Y = np.meshgrid(x, y, indexing="ij")
        if problem_type == "manufactured":
            u_exact = np.sin(np.pi * X) * np.sin(np.pi * Y)
            rhs = 2 * np.pi ** 2 * u_exact
        elif problem_type == "polynomial":
            u_exact = X * (1 - X) * Y * (1 - Y)
            rhs = 2 * X * (1 - X) + 2 * Y * (1 - Y)
        else:
            raise ValueError(f"Unknown problem type: {problem_type}")
        return rhs, u_exact
