"""AdaptivePrecisionSolver (reference solvers/iterative.py:379-552) against golden runs of the reference's own class
(tests/golden/make_golden_adaptive.py).  The wrapper is pure protocol logic, so on the CPU it drives oracle-backed
stand-ins for the device smoother / operator; residual histories must match the reference to rounding, switch decisions,
recorded precision levels and the (quirky) switch_iteration formula exactly."""
import json
import os

import numpy as np
import pytest

from oracle import np_oracle as O

from mixed_precision_multigrid_solvers_for_pdes_b200 import AdaptivePrecisionSolver, Grid, PrecisionManager
from mixed_precision_multigrid_solvers_for_pdes_b200.solvers.base import IterativeSolver

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "adaptive_golden.json")))


class _OracleSmoother(IterativeSolver):
    def __init__(self, kind, max_iterations, tolerance):
        super().__init__(max_iterations, tolerance, 2.0 / 3.0 if kind == "jacobi" else 1.0, False,
                         "Jacobi" if kind == "jacobi" else "Gauss-Seidel")
        self.kind = kind

    def smooth(self, grid, operator, u, rhs, num_iterations=1):
        fn = O.jacobi_smooth if self.kind == "jacobi" else O.rbgs_smooth
        return fn(u, rhs, grid.hx, grid.hy, self.omega, num_iterations)


class _OracleOperator:
    coefficient = -1.0

    def residual(self, grid, u, f):
        return O.residual(u, f, grid.hx, grid.hy, -1.0)


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_matches_reference_runs(case):
    g = Grid(case["n"], case["n"])
    f = 2 * np.pi ** 2 * np.sin(np.pi * g.X) * np.sin(np.pi * g.Y)
    s = AdaptivePrecisionSolver(_OracleSmoother(case["base"], case["max_iterations"], case["tolerance"]),
                                precision_switch_threshold=case["threshold"], convergence_window=case["window"],
                                min_iterations_before_switch=case["min_it"])
    pm = PrecisionManager(default_precision="double", adaptive=True) if case["pm"] else None
    u, info = s.solve(g, _OracleOperator(), f, precision_manager=pm)
    assert s.name == case["solver_name"]
    assert info["iterations"] == case["iterations"] and info["converged"] == case["converged"]
    np.testing.assert_allclose(info["residual_history"], case["residual_history"], rtol=1e-13)
    assert info["precision_switched"] == case["precision_switched"]
    assert info["switch_iteration"] == case["switch_iteration"]
    assert s.history.precision_levels == case["precision_levels"]
    assert (pm.current_precision.value if pm else None) == case["final_precision"]
    assert str(u.dtype) == case["u_dtype"]
    assert abs(float(np.sum(u)) - case["u_sum"]) <= 1e-12 * abs(case["u_sum"])
    assert sorted(info.keys()) == case["keys"]


def test_second_solve_keeps_the_switch_state_like_the_reference():
    g = Grid(17, 17)
    f = np.ones((17, 17))
    s = AdaptivePrecisionSolver(_OracleSmoother("rbgs", 12, 1e-12), precision_switch_threshold=0.5, convergence_window=2,
                                min_iterations_before_switch=3)
    pm = PrecisionManager(default_precision="double", adaptive=True)
    _, a = s.solve(g, _OracleOperator(), f, precision_manager=pm)
    assert a["precision_switched"]
    _, b = s.solve(g, _OracleOperator(), f, precision_manager=pm)
    # convergence_rates / precision_switched survive reset() (iterative.py:419-421): every iteration of the second
    # solve is recorded at the level the manager was re-set to at its start
    assert b["precision_switched"] and set(s.history.precision_levels) == {"float32"}
