from .base import BaseSolver, ConvergenceHistory, IterativeSolver
from .multigrid import MultigridCycle, MultigridSolver
from .smoothers import (GaussSeidelSmoother, JacobiSmoother, SymmetricGaussSeidelSmoother,
                        WeightedJacobiSmoother)

__all__ = ["BaseSolver", "IterativeSolver", "ConvergenceHistory", "MultigridSolver", "MultigridCycle",
           "JacobiSmoother", "GaussSeidelSmoother", "WeightedJacobiSmoother", "SymmetricGaussSeidelSmoother"]
