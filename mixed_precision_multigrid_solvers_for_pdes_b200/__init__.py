"""mgb200: B200-native (sm_100a) mixed-precision geometric multigrid V/W-cycle hot path.

Drop-in for the reference's ``multigrid.core`` / ``.operators`` / ``.solvers`` surface on this path
(see SURVEY.md section 8).  All arithmetic runs in hand-written CUDA kernels behind the C ABI of
``include/mgb200.h`` (libmgb200.so); there is no CPU fallback."""
from . import ops
from ._lib import LIB_PATH, MGLibraryError
from .core import Grid, PrecisionLevel, PrecisionManager
from .applications.heat_solver import HeatSolver2D
from .applications.poisson_solver import PoissonSolver2D
from .preconditioning import MultigridPreconditioner
from .operators import (BaseOperator, HelmholtzOperator, LaplacianOperator, ProlongationOperator,
                        RestrictionOperator, VariableCoefficientOperator, VariableCoefficientSmoother)
from .problems import (HeatProblem, HeatTestProblems, PoissonProblem, PoissonTestProblems, TimeSteppingConfig,
                       TimeSteppingMethod)
from .solvers import (AdaptivePrecisionSolver, CorrectedMultigridSolver, MixedPrecisionMultigrid, MixedPrecisionMultigridSolver, BaseSolver, ConvergenceHistory, GaussSeidelSmoother, IterativeSolver, JacobiSmoother,
                      MultigridCycle, MultigridSolver, SymmetricGaussSeidelSmoother, WeightedJacobiSmoother)

__version__ = "0.1.0"
GPU_AVAILABLE = True  # the only path there is

__all__ = ["Grid", "PrecisionManager", "PrecisionLevel", "BaseOperator", "LaplacianOperator", "HelmholtzOperator", "VariableCoefficientOperator", "VariableCoefficientSmoother", "HeatSolver2D", "PoissonSolver2D", "MultigridPreconditioner", "RestrictionOperator",
           "ProlongationOperator", "BaseSolver", "IterativeSolver", "ConvergenceHistory", "MultigridSolver",
           "MultigridCycle", "JacobiSmoother", "GaussSeidelSmoother", "WeightedJacobiSmoother",
           "SymmetricGaussSeidelSmoother", "AdaptivePrecisionSolver", "CorrectedMultigridSolver", "MixedPrecisionMultigrid", "MixedPrecisionMultigridSolver",
           "PoissonProblem", "HeatProblem", "TimeSteppingConfig", "TimeSteppingMethod", "PoissonTestProblems",
           "HeatTestProblems", "MGLibraryError", "ops", "LIB_PATH"]
