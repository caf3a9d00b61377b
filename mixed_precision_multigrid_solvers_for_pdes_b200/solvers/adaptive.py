"""``AdaptivePrecisionSolver``: the reference's wrapper that starts in single precision and promotes to double when
convergence slows (solvers/iterative.py:379-552; the second statement of the precision-switching rule next to
PrecisionManager).  Mirrored rule for rule through the public protocol only -- ``base_solver.smooth``,
``operator.residual``, ``grid.l2_norm`` -- so it drives the device smoothers of this package (and any foreign object with
that protocol) unchanged:

  * per iteration: one sweep of the base solver, residual, h-scaled L2 norm, ratio to the previous norm;
  * from iteration ``min_iterations_before_switch`` on, once: promote to double if the mean of the last
    ``convergence_window`` ratios is >= ``precision_switch_threshold`` (iterative.py:533-552);
  * like the reference, "starting in single" only sets ``precision_manager.current_precision`` -- the iterate keeps
    the dtype of ``rhs`` -- and the promotion converts the iterate with ``convert_array``; ``info['switch_iteration']``
    reproduces the reference's formula (iterative.py:518-523), including its quirk of not being the actual iteration."""
from __future__ import annotations

import time
from typing import Any, Dict, List, Tuple

import numpy as np

from ..core.precision import PrecisionLevel
from .base import IterativeSolver


class AdaptivePrecisionSolver(IterativeSolver):
    def __init__(self, base_solver: IterativeSolver, precision_switch_threshold: float = 0.95, convergence_window: int = 5,
                 min_iterations_before_switch: int = 10):
        super().__init__(base_solver.max_iterations, base_solver.tolerance, base_solver.omega, base_solver.verbose,
                         f"Adaptive{base_solver.name}")
        self.base_solver = base_solver
        self.precision_switch_threshold = precision_switch_threshold
        self.convergence_window = convergence_window
        self.min_iterations_before_switch = min_iterations_before_switch
        self.convergence_rates: List[float] = []
        self.precision_switched = False

    def smooth(self, grid, operator, u, rhs, num_iterations: int = 1):
        return self.base_solver.smooth(grid, operator, u, rhs, num_iterations)

    def solve(self, grid, operator, rhs, initial_guess=None, precision_manager=None) -> Tuple[Any, Dict[str, Any]]:
        self.reset()
        # the reference never clears these between solves (iterative.py:419-421); a second solve() of the same object
        # therefore starts "already switched" -- kept, it is observable behaviour
        if precision_manager is not None and precision_manager.adaptive:
            precision_manager.current_precision = PrecisionLevel.SINGLE
        if initial_guess is None:
            u = rhs * 0 if not isinstance(rhs, np.ndarray) else np.zeros_like(rhs)
        else:
            u = initial_guess.copy() if isinstance(initial_guess, np.ndarray) else initial_guess.clone()
        previous = float("inf")
        norm, iteration = float("inf"), 0
        for iteration in range(1, self.max_iterations + 1):
            t0 = time.time()
            u = self.smooth(grid, operator, u, rhs, 1)
            norm = float(grid.l2_norm(operator.residual(grid, u, rhs)))
            if previous != float("inf"):
                self.convergence_rates.append(norm / previous)
                if (precision_manager is not None and not self.precision_switched
                        and iteration >= self.min_iterations_before_switch and self._should_switch_precision()):
                    precision_manager.current_precision = PrecisionLevel.DOUBLE
                    u = precision_manager.convert_array(u, PrecisionLevel.DOUBLE)
                    self.precision_switched = True
            level = precision_manager.current_precision.value if precision_manager is not None else "unknown"
            self.history.record_iteration(norm, time.time() - t0, level)
            self.log_iteration(iteration, norm)
            if self.check_convergence(norm, iteration):
                self.converged = True
                break
            previous = norm
        self.iterations_performed = iteration
        self.final_residual = norm
        info = self.get_convergence_info()
        info["precision_switched"] = self.precision_switched
        info["switch_iteration"] = (
            self.min_iterations_before_switch
            + len([r for r in self.convergence_rates[:self.min_iterations_before_switch]
                   if r >= self.precision_switch_threshold])
            if self.precision_switched else None)
        return u, info

    def _should_switch_precision(self) -> bool:
        if len(self.convergence_rates) < self.convergence_window:
            return False
        return bool(np.mean(self.convergence_rates[-self.convergence_window:]) >= self.precision_switch_threshold)
