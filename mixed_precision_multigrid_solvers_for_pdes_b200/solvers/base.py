"""Solver base classes, API-compatible with the reference's solvers/base.py.

``IterativeSolver.solve`` keeps the reference loop (base.py:258-285): one smoothing sweep, residual,
h-scaled L2 norm, stop below ``tolerance``.  Everything runs on the device; NumPy inputs are copied
in once and the result copied out once."""
from __future__ import annotations

import time
from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional, Tuple

import numpy as np


class ConvergenceHistory:
    def __init__(self):
        self.residual_norms: List[float] = []
        self.iteration_times: List[float] = []
        self.precision_levels: List[str] = []
        self.grid_levels: List[Optional[int]] = []

    def record_iteration(self, residual_norm: float, iteration_time: float, precision_level: str,
                         grid_level: Optional[int] = None) -> None:
        self.residual_norms.append(residual_norm)
        self.iteration_times.append(iteration_time)
        self.precision_levels.append(precision_level)
        self.grid_levels.append(grid_level)

    def get_convergence_rate(self) -> float:
        """Mean of the contracting ratios among the last five residuals (base.py:39-57)."""
        if len(self.residual_norms) < 3:
            return 0.0
        r = self.residual_norms[-5:]
        ratios = [r[k] / r[k - 1] for k in range(1, len(r)) if r[k - 1] > 0 and 0 < r[k] / r[k - 1] < 1]
        return float(np.mean(ratios)) if ratios else 0.0

    def clear(self) -> None:
        self.residual_norms.clear()
        self.iteration_times.clear()
        self.precision_levels.clear()
        self.grid_levels.clear()


class BaseSolver(ABC):
    def __init__(self, max_iterations: int = 1000, tolerance: float = 1e-8, verbose: bool = False,
                 name: str = "BaseSolver"):
        self.max_iterations, self.tolerance, self.verbose, self.name = max_iterations, tolerance, verbose, name
        self.history = ConvergenceHistory()
        self.converged = False
        self.final_residual = float("inf")
        self.iterations_performed = 0

    @abstractmethod
    def solve(self, grid, operator, rhs, initial_guess=None, precision_manager=None) -> Tuple[Any, Dict[str, Any]]:
        ...

    def check_convergence(self, residual_norm: float, iteration: int) -> bool:
        return residual_norm < self.tolerance

    def log_iteration(self, iteration: int, residual_norm: float, grid_level: Optional[int] = None) -> None:
        if self.verbose and iteration % max(1, self.max_iterations // 10) == 0:
            lvl = f" (level {grid_level})" if grid_level is not None else ""
            print(f"{self.name} iteration {iteration}{lvl}: residual = {residual_norm:.2e}")

    def get_convergence_info(self) -> Dict[str, Any]:
        t = self.history.iteration_times
        return {
            "converged": self.converged,
            "iterations": self.iterations_performed,
            "final_residual": self.final_residual,
            "convergence_rate": self.history.get_convergence_rate(),
            "residual_history": self.history.residual_norms.copy(),
            "total_time": sum(t),
            "average_time_per_iteration": (float(np.mean(t)) if t else 0.0),
            "precision_levels_used": list(set(self.history.precision_levels)),
        }

    def reset(self) -> None:
        self.history.clear()
        self.converged = False
        self.final_residual = float("inf")
        self.iterations_performed = 0


class IterativeSolver(BaseSolver):
    def __init__(self, max_iterations: int = 1000, tolerance: float = 1e-8, relaxation_parameter: float = 1.0,
                 verbose: bool = False, name: str = "IterativeSolver"):
        super().__init__(max_iterations, tolerance, verbose, name)
        self.omega = relaxation_parameter

    @abstractmethod
    def smooth(self, grid, operator, u, rhs, num_iterations: int = 1):
        ...

    def _smooth_device_(self, grid, u, rhs, num_iterations: int):
        """In-place device smoothing on pitched CUDA tensors (implemented by subclasses)."""
        raise NotImplementedError

    def solve(self, grid, operator, rhs, initial_guess=None, precision_manager=None):
        import torch

        from .. import ops
        from ..device import like_input, to_device
        self.reset()
        f, was_np = to_device(rhs)
        if initial_guess is None:
            u = torch.zeros_like(f)
        else:
            u0, _ = to_device(initial_guess, dtype=f.dtype)
            u = torch.empty_like(f)
            u.copy_(u0)
        coeff = getattr(operator, "coefficient", None)
        norm = float("inf")
        iteration = 0
        for iteration in range(1, self.max_iterations + 1):
            t0 = time.time()
            self._smooth_device_(grid, u, f, 1)
            if coeff is not None:
                r = ops.residual(u, f, grid.hx, grid.hy, coeff)
            else:
                r, _ = to_device(operator.residual(grid, u, f))
            norm = float(np.sqrt(grid.hx * grid.hy * ops.sumsq(r)))
            prec = precision_manager.current_precision.value if precision_manager else "unknown"
            self.history.record_iteration(norm, time.time() - t0, prec)
            self.log_iteration(iteration, norm)
            if self.check_convergence(norm, iteration):
                self.converged = True
                break
        self.iterations_performed = iteration
        self.final_residual = norm
        return like_input(u, was_np), self.get_convergence_info()
