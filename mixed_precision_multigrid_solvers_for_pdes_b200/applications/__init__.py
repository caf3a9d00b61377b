from ..problems import (HeatProblem, HeatTestProblems, PoissonProblem, PoissonTestProblems, TimeSteppingConfig,
                        TimeSteppingMethod)
from .heat_solver import HeatSolver2D
from .poisson_solver import PoissonSolver2D

__all__ = ["HeatSolver2D", "PoissonSolver2D", "HeatProblem", "TimeSteppingConfig", "TimeSteppingMethod", "PoissonProblem",
           "PoissonTestProblems", "HeatTestProblems"]
