"""SURVEY 8f-1 / BASELINE config 5 (small sizes): Helmholtz-shifted cycle and implicit heat time stepping.
The reference has no operator for these (parity UNPINNED): checked against the repo's own NumPy oracle (shifted
restatement of the same formulas) and against analytical solutions with the expected orders."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

from mixed_precision_multigrid_solvers_for_pdes_b200 import (Grid, HeatSolver2D, HeatTestProblems,  # noqa: E402
                                                             HelmholtzOperator, MixedPrecisionMultigrid, PoissonProblem,
                                                             TimeSteppingConfig, TimeSteppingMethod, ops)
from mixed_precision_multigrid_solvers_for_pdes_b200.device import empty_field, to_device, to_host  # noqa: E402


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_shifted_operator_kernels_match_oracle(dt):
    rng = np.random.default_rng(21)
    n, lam = 129, 4096.0
    g = Grid(n, n, dtype=dt)
    u, f = rng.uniform(-1, 1, (n, n)).astype(dt), rng.uniform(-1, 1, (n, n)).astype(dt)
    du, df = to_device(u)[0], to_device(f)[0]
    exp_r = O.residual(u, f, g.hx, g.hy, -1.0, lam)
    np.testing.assert_array_equal(to_host(ops.residual(du, df, g.hx, g.hy, -1.0, shift=lam)), exp_r)
    np.testing.assert_array_equal(HelmholtzOperator(-1.0, lam).residual(g, u, f), exp_r)
    exp_s = O.rbgs_smooth(u, f, g.hx, g.hy, 1.0, 2, lam)
    s = du.clone()
    ops.smooth_rbgs_(s, df, g.hx, g.hy, 1.0, 2, shift=lam)
    np.testing.assert_array_equal(to_host(s), exp_s)
    out, rc = empty_field(n, n, dt), empty_field(65, 65, dt)
    # strict kernels: bit-exact.  Fused kernels multiply by 1/(4/h^2 + lambda), not a power of two: few-ulp agreement
    ops.vc_pass(du, out, df, g.hx, g.hy, sweeps=2, coarse_out=rc, shift=lam)
    tol = 1e-13 if dt is np.float64 else 2e-6
    assert np.max(np.abs(to_host(out).astype(np.float64) - exp_s)) <= tol * np.max(np.abs(exp_s))
    exp_rc = O.restrict(O.residual(exp_s, f, g.hx, g.hy, -1.0, lam))
    assert np.max(np.abs(to_host(rc).astype(np.float64) - exp_rc)) <= (1e-12 if dt is np.float64 else 1e-4) * np.max(np.abs(exp_rc))


def test_shifted_solve_matches_oracle_cycle_for_cycle():
    n, lam = 129, 1000.0
    f = O.mms_rhs(n)
    ou, oinfo = O.OracleMultigrid(n, max_levels=6, shift=lam).solve(f)
    u, info = MixedPrecisionMultigrid("double", shift=lam, strict_reference_norm=True).solve(PoissonProblem(rhs=f, nx=n, ny=n))
    assert info["iterations"] == oinfo["iterations"]
    # 1/(4/h^2 + lambda) is not a power of two: operators agree to ~1 ulp, which f - A u amplifies to ~1e-5 relative in
    # the tail of the history (residual 1e-9 from terms of size 1e+1); counts and solution are pinned tightly
    np.testing.assert_allclose(info["residual_history"], oinfo["residual_history"], rtol=1e-3)
    assert np.max(np.abs(u - ou)) <= 1e-12 * np.max(np.abs(ou))
    um, im = MixedPrecisionMultigrid("adaptive", shift=lam).solve(PoissonProblem(rhs=f, nx=n, ny=n))
    assert im["converged"] and np.max(np.abs(um - ou)) < 1e-9


@pytest.mark.parametrize("method,order", [(TimeSteppingMethod.BACKWARD_EULER, 1), (TimeSteppingMethod.CRANK_NICOLSON, 2)])
def test_heat_pure_diffusion_temporal_order(method, order):
    prob = HeatTestProblems().get_problem("pure_diffusion")
    n, T = 129, 0.02
    errs = []
    for steps in (4, 8, 16):
        res = HeatSolver2D(tolerance=1e-10).solve_heat_problem(prob, n, n, TimeSteppingConfig(method, T / steps, T))
        assert res["total_steps"] == steps and abs(res["final_time"] - T) < 1e-12
        # the discrete spatial operator has eigenvalue mu_h for sin*sin: compare with the exact solution of the
        # SEMI-discrete problem so that only the time discretisation error remains
        h = 1.0 / (n - 1)
        mu = 2 * (4 / h ** 2) * np.sin(np.pi * h / 2) ** 2
        exact_semi = O.mms_exact(n) * np.exp(-mu * T)
        errs.append(np.max(np.abs(res["final_solution"] - exact_semi)))
        for key in ("problem_name", "grid_size", "time_config", "final_solution", "final_time", "total_steps",
                    "total_time", "total_solver_time", "avg_mg_iterations", "total_mg_iterations", "errors"):
            assert key in res
        assert res["errors"]["max_error"] < 2e-2  # 4 backward-Euler steps: 0.5 mu^2 dt T e^{-mu T} = 0.013
    rates = [np.log2(errs[k] / errs[k + 1]) for k in range(2)]
    assert all(abs(r - order) < 0.25 for r in rates), (errs, rates)


def test_heat_with_source_matches_analytical_solution():
    prob = HeatTestProblems().get_problem("heat_source")   # u = sin sin (1 + t): backward Euler is exact in time for
    res = HeatSolver2D().solve_heat_problem(prob, 65, 65, TimeSteppingConfig(TimeSteppingMethod.BACKWARD_EULER, 0.05, 0.2),
                                            save_solution_history=True)
    # a solution linear in t: only the O(h^2) spatial error remains
    assert res["errors"]["relative_max_error"] < 3e-4
    assert len(res["solution_history"]) == 5 and res["time_steps"][-1] == pytest.approx(0.2)
    assert 1 <= res["avg_mg_iterations"] <= 12
