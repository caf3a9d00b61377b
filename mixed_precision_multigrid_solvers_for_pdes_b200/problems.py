"""Problem definitions: the API surface the north star keeps ("PoissonProblem and other problem
classes").

``PoissonProblem`` accepts BOTH spellings found in the reference:
  * the dataclass actually shipped  (applications/poisson_solver.py:24-32):
        PoissonProblem(name, source_function, analytical_solution=None, boundary_conditions=None,
                       domain=(0,1,0,1), description="")
  * the documented facade call       (README.md:81):
        PoissonProblem(source_term, nx=129, ny=129)
``HeatProblem`` / ``TimeSteppingConfig`` / ``TimeSteppingMethod`` mirror applications/heat_solver.py:26-55.
The catalogues hold manufactured problems with homogeneous Dirichlet data (the only boundary
condition the hot path implements); sources are derived here from the exact solutions."""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

Domain = Tuple[float, float, float, float]


class PoissonProblem:
    """-lap(u) = f on a rectangle with homogeneous Dirichlet data."""

    def __init__(self, *args, **kw):
        self.nx: Optional[int] = kw.pop("nx", None)
        self.ny: Optional[int] = kw.pop("ny", None)
        self.rhs_array = kw.pop("rhs", None)          # optional precomputed right-hand side (NumPy or CUDA tensor)
        self.device_mms = kw.pop("device_mms", None)  # optional (amplitude, kx, ky): generate f in HBM
        if args and callable(args[0]):                # README style: PoissonProblem(source_term, nx=..., ny=...)
            self.name = kw.pop("name", "poisson")
            self.source_function = args[0]
            rest = list(args[1:])
        else:                                         # dataclass style: PoissonProblem(name, source_function, ...)
            rest = list(args)
            self.name = rest.pop(0) if rest else kw.pop("name", "poisson")
            self.source_function = rest.pop(0) if rest else kw.pop("source_function", None)
        self.analytical_solution = rest.pop(0) if rest else kw.pop("analytical_solution", None)
        self.boundary_conditions = rest.pop(0) if rest else kw.pop("boundary_conditions", None)
        self.domain: Domain = tuple(rest.pop(0)) if rest else tuple(kw.pop("domain", (0, 1, 0, 1)))
        self.description: str = rest.pop(0) if rest else kw.pop("description", "")
        if rest or kw:
            raise TypeError(f"unexpected arguments: {rest} {sorted(kw)}")
        if self.nx is not None and self.ny is None:
            self.ny = self.nx
        if self.source_function is None and self.rhs_array is None and self.device_mms is None:
            raise TypeError("PoissonProblem needs a source_function, rhs= or device_mms=")

    @classmethod
    def manufactured(cls, nx: int, ny: Optional[int] = None, on_device: bool = True) -> "PoissonProblem":
        """u = sin(pi x) sin(pi y), f = 2 pi^2 u on the unit square (README.md:77-78); with
        ``on_device`` the right-hand side is generated in HBM instead of on the host."""
        p = cls("trigonometric", lambda x, y: 2 * np.pi ** 2 * np.sin(np.pi * x) * np.sin(np.pi * y),
                lambda x, y: np.sin(np.pi * x) * np.sin(np.pi * y), {"type": "dirichlet", "value": 0.0}, (0, 1, 0, 1),
                "u = sin(pi x) sin(pi y), homogeneous Dirichlet BC", nx=nx, ny=ny)
        if on_device:
            p.device_mms = (2 * np.pi ** 2, 1.0, 1.0)
        return p

    def __repr__(self) -> str:
        return f"PoissonProblem(name={self.name!r}, nx={self.nx}, ny={self.ny}, domain={self.domain})"


class TimeSteppingMethod(Enum):
    BACKWARD_EULER = "backward_euler"
    CRANK_NICOLSON = "crank_nicolson"
    THETA_METHOD = "theta_method"


@dataclass
class HeatProblem:
    name: str
    initial_condition: Callable[[np.ndarray, np.ndarray], np.ndarray]
    source_function: Optional[Callable[[np.ndarray, np.ndarray, float], np.ndarray]] = None
    analytical_solution: Optional[Callable[[np.ndarray, np.ndarray, float], np.ndarray]] = None
    boundary_conditions: Optional[Dict[str, Any]] = None
    thermal_diffusivity: float = 1.0
    domain: Domain = (0, 1, 0, 1)
    description: str = ""


@dataclass
class TimeSteppingConfig:
    method: TimeSteppingMethod
    dt: float
    t_final: float
    theta: float = 0.5
    adaptive_dt: bool = False
    cfl_max: float = 0.5
    save_frequency: int = 1


_DIRICHLET0 = {"type": "dirichlet", "value": 0.0}


class PoissonTestProblems:
    """Manufactured Poisson problems with zero boundary data (names follow applications/test_problems.py:27-57)."""

    def __init__(self):
        pi = np.pi
        self.problems: Dict[str, PoissonProblem] = {}

        def add(name, exact, source, desc):
            self.problems[name] = PoissonProblem(name, source, exact, dict(_DIRICHLET0), (0, 1, 0, 1), desc)

        add("trigonometric", lambda x, y: np.sin(pi * x) * np.sin(pi * y),
            lambda x, y: 2 * pi ** 2 * np.sin(pi * x) * np.sin(pi * y), "u = sin(pi x) sin(pi y)")

        def poly(x, y):
            return x ** 2 * (1 - x) ** 2 * y ** 2 * (1 - y) ** 2

        def poly_src(x, y):
            p = lambda t: t ** 2 * (1 - t) ** 2                 # noqa: E731
            d2 = lambda t: 2 - 12 * t + 12 * t ** 2             # noqa: E731  (p'')
            return -(d2(x) * p(y) + p(x) * d2(y))

        add("polynomial", poly, poly_src, "u = x^2 (1-x)^2 y^2 (1-y)^2")
        add("high_frequency", lambda x, y: np.sin(8 * pi * x) * np.sin(8 * pi * y),
            lambda x, y: 128 * pi ** 2 * np.sin(8 * pi * x) * np.sin(8 * pi * y), "u = sin(8 pi x) sin(8 pi y)")
        add("anisotropic", lambda x, y: np.sin(pi * x) * np.sin(4 * pi * y),
            lambda x, y: 17 * pi ** 2 * np.sin(pi * x) * np.sin(4 * pi * y), "u = sin(pi x) sin(4 pi y)")
        add("mixed", lambda x, y: x * (1 - x) * np.sin(pi * y),
            lambda x, y: (2 + pi ** 2 * x * (1 - x)) * np.sin(pi * y), "u = x (1-x) sin(pi y)")

    def get_problem(self, name: str) -> PoissonProblem:
        if name not in self.problems:
            raise ValueError(f"Unknown problem: {name}")
        return self.problems[name]

    def list_problems(self) -> List[str]:
        return list(self.problems)

    def get_problem_info(self) -> Dict[str, str]:
        return {k: v.description for k, v in self.problems.items()}


class HeatTestProblems:
    """Manufactured heat problems u_t = alpha lap(u) + f with zero boundary data
    (names follow applications/test_problems.py:325-349)."""

    def __init__(self):
        pi = np.pi
        self.problems: Dict[str, HeatProblem] = {}
        a = 1.0
        self.problems["pure_diffusion"] = HeatProblem(
            "pure_diffusion", lambda x, y: np.sin(pi * x) * np.sin(pi * y), lambda x, y, t: np.zeros_like(x),
            lambda x, y, t: np.sin(pi * x) * np.sin(pi * y) * np.exp(-2 * pi ** 2 * a * t), dict(_DIRICHLET0), a,
            (0, 1, 0, 1), "u = sin(pi x) sin(pi y) exp(-2 pi^2 alpha t)")
        # u = sin(pi x) sin(pi y) (1 + t): u_t - alpha lap u = sin sin (1 + 2 pi^2 alpha (1 + t))
        self.problems["heat_source"] = HeatProblem(
            "heat_source", lambda x, y: np.sin(pi * x) * np.sin(pi * y),
            lambda x, y, t: np.sin(pi * x) * np.sin(pi * y) * (1 + 2 * pi ** 2 * a * (1 + t)),
            lambda x, y, t: np.sin(pi * x) * np.sin(pi * y) * (1 + t), dict(_DIRICHLET0), a, (0, 1, 0, 1),
            "u = sin(pi x) sin(pi y) (1 + t)")
        self.problems["multiple_frequencies"] = HeatProblem(
            "multiple_frequencies",
            lambda x, y: np.sin(pi * x) * np.sin(pi * y) + 0.5 * np.sin(3 * pi * x) * np.sin(2 * pi * y),
            lambda x, y, t: np.zeros_like(x),
            lambda x, y, t: (np.sin(pi * x) * np.sin(pi * y) * np.exp(-2 * pi ** 2 * a * t)
                             + 0.5 * np.sin(3 * pi * x) * np.sin(2 * pi * y) * np.exp(-13 * pi ** 2 * a * t)),
            dict(_DIRICHLET0), a, (0, 1, 0, 1), "two decaying modes")

    def get_problem(self, name: str) -> HeatProblem:
        if name not in self.problems:
            raise ValueError(f"Unknown problem: {name}")
        return self.problems[name]

    def list_problems(self) -> List[str]:
        return list(self.problems)
