"""PoissonSolver2D wrapper and MultigridPreconditioner (callers of the hot path)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

from mixed_precision_multigrid_solvers_for_pdes_b200 import (Grid, LaplacianOperator, MultigridPreconditioner,  # noqa: E402
                                                             PoissonSolver2D, PoissonTestProblems)
from mixed_precision_multigrid_solvers_for_pdes_b200.device import to_device  # noqa: E402


@pytest.mark.parametrize("solver_type", ["multigrid", "gpu_multigrid"])
def test_poisson_solver_2d_result_dict_and_errors(solver_type):
    prob = PoissonTestProblems().get_problem("trigonometric")
    res = PoissonSolver2D(solver_type=solver_type, max_levels=6).solve_poisson_problem(prob, 129, 129)
    for key in ("problem_name", "grid_size", "domain", "solution", "solve_time", "solver_info", "errors", "solver_type",
                "use_gpu", "mixed_precision", "analytical_solution"):
        assert key in res
    assert res["solver_info"]["converged"] and res["solver_info"]["iterations"] == 8
    assert abs(res["errors"]["max_error"] - 5.0201e-5) < 1e-8        # 5.020084e-5 at 129^2
    for key in ("l2_error", "relative_l2_error", "max_error", "relative_max_error", "h1_semi_error", "grid_spacing"):
        assert key in res["errors"]


def test_convergence_study_is_second_order():
    st = PoissonSolver2D(max_levels=4).run_convergence_study(PoissonTestProblems().get_problem("trigonometric"),
                                                             grid_sizes=(33, 65, 129))
    assert abs(st["max_error_rate"] - 2.0) < 0.05 and abs(st["l2_error_rate"] - 2.0) < 0.05


def test_multigrid_preconditioner_is_fixed_cycles_from_zero():
    n = 65
    g = Grid(n, n)
    op = LaplacianOperator(-1.0)
    x = np.random.default_rng(4).standard_normal((n, n))
    x[0, :] = x[-1, :] = x[:, 0] = x[:, -1] = 0.0
    pc = MultigridPreconditioner(max_levels=5, num_cycles=2, pre_smooth_iterations=1, post_smooth_iterations=1,
                                 coarse_tolerance=1e-12, coarse_max_iterations=1000)
    with pytest.raises(RuntimeError):
        pc.apply(x)
    pc.setup(g, op)
    z = pc.apply(x)
    s = O.OracleMultigrid(n, max_levels=5, pre=1, post=1, max_iterations=2, tolerance=0.0)
    zo, _ = s.solve(x)
    assert np.max(np.abs(z - zo)) <= 1e-12 * np.max(np.abs(zo))
    zt = pc.apply(to_device(x)[0])
    assert isinstance(zt, torch.Tensor) and np.array_equal(zt.cpu().numpy(), z)
    # preconditioned Richardson on A u = x converges fast
    u = np.zeros_like(x)
    for _ in range(6):
        u = u + pc.apply(O.residual(u, x, g.hx, g.hy, -1.0))
    assert np.linalg.norm(O.residual(u, x, g.hx, g.hy, -1.0)[1:-1, 1:-1]) < 1e-6 * np.linalg.norm(x)


def test_poisson_solver_2d_benchmark_and_statistics_keys():
    """poisson_solver.py:398-480: same result keys."""
    prob = PoissonTestProblems().get_problem("trigonometric")
    ps = PoissonSolver2D(solver_type="gpu_multigrid")
    b = ps.benchmark_solver_performance(prob, [(65, 65), (129, 129)], num_runs=2)
    assert b["problem_name"] == prob.name and set(b["solver_configuration"]) == {"solver_type", "use_gpu", "mixed_precision",
                                                                                 "max_levels", "cycle_type"}
    assert [r["grid_size"] for r in b["benchmark_results"]] == [(65, 65), (129, 129)]
    for r in b["benchmark_results"]:
        assert set(r) == {"grid_size", "total_unknowns", "num_runs", "avg_time", "std_time", "min_time", "max_time",
                          "avg_iterations", "throughput", "solver_type"}
        assert r["avg_iterations"] == 8 and r["throughput"] > 0
    st = ps.get_solver_statistics()
    assert st["solve_history_count"] == 4 and st["configuration"]["tolerance"] == 1e-8
