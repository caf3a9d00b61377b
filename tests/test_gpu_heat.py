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


def test_heat_rhs_kernel_equals_the_eager_formula():
    """mg_heat_rhs: lam*(u + c_lap*L_h u + c_f1*f1 + c_f0*f0), ring zeroed, sum of squares -- one pass -- against NumPy,
    for lap_h and for div(a grad .)."""
    rng = np.random.default_rng(61)
    n, m = 129, 65
    g = Grid(n, m, (0.0, 2.0, 0.0, 1.0))
    u, f1, f0 = (rng.uniform(-1, 1, (n, m)) for _ in range(3))
    a = 1.0 + 0.5 * rng.uniform(0, 1, (n, m))
    du, d1, d0, da = (to_device(t)[0] for t in (u, f1, f0, a))
    for coef in (None, a):
        lap = np.zeros_like(u)
        if coef is None:
            lap[1:-1, 1:-1] = ((u[2:, 1:-1] + u[:-2, 1:-1]) / g.hx ** 2 + (u[1:-1, 2:] + u[1:-1, :-2]) / g.hy ** 2
                               - u[1:-1, 1:-1] * (2 / g.hx ** 2 + 2 / g.hy ** 2))
        else:
            lap = -O.varcoef_apply(u, a, g.hx, g.hy, 0.0)
        want = 7.5 * (u + 0.3 * lap + 0.2 * f1 + 0.1 * f0)
        want[0, :] = want[-1, :] = want[:, 0] = want[:, -1] = 0.0
        out = empty_field(n, m, np.float64)
        ss = ops.heat_rhs_(du, out, g.hx, g.hy, lam=7.5, c_lap=0.3, f1=d1, c_f1=0.2, f0=d0, c_f0=0.1,
                           a=da if coef is not None else None, norm_rows=(3, n - 5))
        got = to_host(out)
        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
        assert abs(ss.item() - float(np.sum(got[3:n - 5] ** 2))) <= 1e-12 * float(np.sum(got ** 2))
    # backward Euler without sources: a pure scaling, slab flags leave the first / last row alone
    out = empty_field(n, m, np.float64)
    ops.heat_rhs_(du, out, g.hx, g.hy, lam=2.0, zero_first_row=False, zero_last_row=False)
    want = 2.0 * u
    want[:, 0] = want[:, -1] = 0.0
    assert np.array_equal(to_host(out), want)


def test_heat_with_variable_diffusivity_converges_in_time_and_space():
    """u_t = div(a grad u) + f with a manufactured solution u = exp(-t) sin(pi x) sin(pi y) (BASELINE configs[4], the
    variable-coefficient half; no reference operator): backward Euler is O(dt) + O(h^2)."""
    from mixed_precision_multigrid_solvers_for_pdes_b200 import HeatProblem

    def a(X, Y):
        return 1.0 + 0.5 * np.sin(2 * np.pi * X) * np.cos(np.pi * Y) + X * Y

    def exact(X, Y, t):
        return np.exp(-t) * np.sin(np.pi * X) * np.sin(np.pi * Y)

    def source(X, Y, t):
        sx, cx, sy, cy = np.sin(np.pi * X), np.cos(np.pi * X), np.sin(np.pi * Y), np.cos(np.pi * Y)
        ax = np.pi * np.cos(2 * np.pi * X) * np.cos(np.pi * Y) + Y
        ay = -0.5 * np.pi * np.sin(2 * np.pi * X) * np.sin(np.pi * Y) + X
        div = (ax * np.pi * cx * sy + ay * np.pi * sx * cy) - a(X, Y) * 2 * np.pi ** 2 * sx * sy
        return np.exp(-t) * (-sx * sy - div)

    prob = HeatProblem("varcoef_mms", lambda X, Y: exact(X, Y, 0.0), source, exact, thermal_diffusivity=a)

    def run(method, n, steps, T):
        res = HeatSolver2D(tolerance=1e-10).solve_heat_problem(prob, n, n, TimeSteppingConfig(method, T / steps, T))
        assert res["total_steps"] == steps
        return res["errors"]["max_error"]

    # time: backward Euler with large steps (the O(dt) error dominates the O(h^2) one): halving dt halves the error
    e4, e8 = run(TimeSteppingMethod.BACKWARD_EULER, 129, 4, 0.4), run(TimeSteppingMethod.BACKWARD_EULER, 129, 8, 0.4)
    assert 1.7 < e4 / e8 < 2.3, (e4, e8)
    # space: Crank-Nicolson with small steps (O(dt^2) negligible): halving h quarters the error.  The explicit half of
    # the step goes through the div(a grad .) branch of mg_heat_rhs.
    c65, c129 = run(TimeSteppingMethod.CRANK_NICOLSON, 65, 16, 0.05), run(TimeSteppingMethod.CRANK_NICOLSON, 129, 16, 0.05)
    assert 3.2 < c65 / c129 < 4.8, (c65, c129)
