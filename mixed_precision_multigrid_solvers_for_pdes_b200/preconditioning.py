"""``MultigridPreconditioner``: a fixed number of cycles from a zero initial guess, for an outer Krylov method
(reference preconditioning/multigrid_preconditioner.py:20-175, same constructor/setup/apply).
The reference drives ``MultigridSolver.solve`` with ``tolerance=1e-16`` and pays a residual norm per cycle
(:119-163); here ``MultigridSolver.apply_cycles`` runs the cycles without norms or host synchronisation, and
vectors that are CUDA tensors never leave the device."""
from __future__ import annotations

from typing import Optional

from .operators.transfer import ProlongationOperator, RestrictionOperator
from .solvers.multigrid import MultigridSolver
from .solvers.smoothers import GaussSeidelSmoother


class MultigridPreconditioner:
    def __init__(self, max_levels: int = 3, cycle_type: str = "V", pre_smooth_iterations: int = 1,
                 post_smooth_iterations: int = 1, num_cycles: int = 1, coarse_tolerance: float = 1e-6,
                 coarse_max_iterations: int = 100):
        self.name = "MultigridPreconditioner"
        self.setup_completed = False
        self.max_levels, self.cycle_type = max_levels, cycle_type
        self.pre_smooth_iterations, self.post_smooth_iterations = pre_smooth_iterations, post_smooth_iterations
        self.num_cycles, self.coarse_tolerance, self.coarse_max_iterations = num_cycles, coarse_tolerance, coarse_max_iterations
        self.mg_solver: Optional[MultigridSolver] = None
        self.grid = self.operator = self.precision_manager = None

    def setup(self, grid, operator, restriction_op=None, prolongation_op=None, smoother=None, coarse_solver=None,
              precision_manager=None) -> None:
        self.grid, self.operator, self.precision_manager = grid, operator, precision_manager
        restriction_op = restriction_op or RestrictionOperator("full_weighting")
        prolongation_op = prolongation_op or ProlongationOperator("bilinear")
        if smoother is None:  # reference default is lexicographic GS; red-black is the parallel kernel
            smoother = GaussSeidelSmoother(red_black=True)
        if coarse_solver is None:
            coarse_solver = GaussSeidelSmoother(max_iterations=self.coarse_max_iterations, tolerance=self.coarse_tolerance)
        self.mg_solver = MultigridSolver(max_levels=self.max_levels, max_iterations=self.num_cycles, tolerance=1e-16,
                                         cycle_type=self.cycle_type, pre_smooth_iterations=self.pre_smooth_iterations,
                                         post_smooth_iterations=self.post_smooth_iterations,
                                         coarse_tolerance=self.coarse_tolerance,
                                         coarse_max_iterations=self.coarse_max_iterations)
        self.mg_solver.setup(grid, operator, restriction_op, prolongation_op, smoother, coarse_solver)
        self.setup_completed = True

    def apply(self, x):
        """z ~= A^-1 x: `num_cycles` cycles from z = 0."""
        if not self.setup_completed:
            raise RuntimeError("Multigrid preconditioner not setup")
        return self.mg_solver.apply_cycles(x, self.num_cycles)

    def apply_transpose(self, x):
        return self.apply(x)  # the cycle of a symmetric operator is approximately symmetric (reference :165-175)
