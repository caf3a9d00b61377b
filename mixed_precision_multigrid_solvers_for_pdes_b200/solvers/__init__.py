from .adaptive import AdaptivePrecisionSolver
from .base import BaseSolver, ConvergenceHistory, IterativeSolver
from .corrected_multigrid import CorrectedMultigridSolver
from .mixed_precision import MixedPrecisionMultigrid, MixedPrecisionMultigridSolver
from .multigrid import MultigridCycle, MultigridSolver
from .smoothers import (GaussSeidelSmoother, JacobiSmoother, SymmetricGaussSeidelSmoother,
                        WeightedJacobiSmoother)

__all__ = ["AdaptivePrecisionSolver", "CorrectedMultigridSolver", "BaseSolver", "IterativeSolver", "ConvergenceHistory", "MultigridSolver", "MultigridCycle", "MixedPrecisionMultigrid", "MixedPrecisionMultigridSolver",
           "JacobiSmoother", "GaussSeidelSmoother", "WeightedJacobiSmoother", "SymmetricGaussSeidelSmoother"]
