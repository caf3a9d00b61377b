"""GPU parity of full solves through the reference-shaped API (MultigridSolver.setup/solve) against
the golden runs of the reference: identical cycle counts, residual histories and solutions.
Basic kernel set => bit-exact operators; the only non-bitwise piece is the fp64 reduction order
of the residual norm (tolerance 1e-12 relative on the history)."""
import numpy as np
import pytest

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

from mixed_precision_multigrid_solvers_for_pdes_b200 import (GaussSeidelSmoother, Grid, JacobiSmoother,  # noqa: E402
                                                             LaplacianOperator, MultigridSolver, PrecisionManager,
                                                             ProlongationOperator, RestrictionOperator,
                                                             WeightedJacobiSmoother)


def _smoother(kind):
    return {"rbgs": lambda: GaussSeidelSmoother(red_black=True), "lexgs": lambda: None,
            "jacobi": JacobiSmoother, "wjacobi": WeightedJacobiSmoother}[kind]()


def _run(m, kernels):
    dt = np.dtype(m["dtype"]).type
    g = Grid(m["nx"], m["ny"], dtype=dt)
    op = LaplacianOperator(-1.0)
    s = MultigridSolver(max_levels=m["max_levels"], max_iterations=30 if m["dtype"] == "float64" else 12,
                        tolerance=1e-8, cycle_type=m["cycle"], kernels=kernels)
    s.setup(g, op, RestrictionOperator("full_weighting"), ProlongationOperator("bilinear"),
            smoother=_smoother(m["smoother"]))
    f = O.mms_rhs(m["nx"], m["ny"], dtype=dt)
    pm = PrecisionManager(default_precision=m["precision"], adaptive=True) if m["precision"] else None
    return s.solve(g, op, f, precision_manager=pm)


@pytest.mark.parametrize("kernels", ["basic"])
def test_solves_match_reference_runs(solve_golden, golden_meta, kernels):
    for m in golden_meta["solves"]:
        u, info = _run(m, kernels)
        hist = solve_golden[f"{m['name']}_hist"]
        assert info["iterations"] == m["iterations"], m["name"]
        assert info["converged"] == m["converged"], m["name"]
        assert info["num_levels"] == m["num_levels"], m["name"]
        strict = m["dtype"] == "float64" and m["precision"] is None
        np.testing.assert_allclose(info["residual_history"], hist, rtol=1e-12 if strict else 2e-2, err_msg=m["name"])
        if f"{m['name']}_u" in solve_golden:
            ref = solve_golden[f"{m['name']}_u"]
            assert u.dtype == ref.dtype
            if strict:
                # only the coarse-solve stopping test sees a differently ordered sum: allow 1e-12 relative
                assert np.max(np.abs(u - ref)) <= 1e-12 * np.max(np.abs(ref)), m["name"]
            else:
                assert np.max(np.abs(u.astype(np.float64) - ref)) < 2e-6, m["name"]
        err = np.max(np.abs(u.astype(np.float64) - O.mms_exact(m["nx"], m["ny"])))
        assert abs(err - m["max_error"]) <= (1e-12 if strict else 1e-6), m["name"]
        for key in ("converged", "iterations", "final_residual", "convergence_rate", "residual_history", "total_time",
                    "average_time_per_iteration", "precision_levels_used", "cycle_type", "num_levels",
                    "grid_hierarchy", "level_timings", "pre_smooth_iterations", "post_smooth_iterations"):
            assert key in info, key


def test_solve_errors_like_reference():
    s = MultigridSolver()
    g = Grid(17, 17)
    with pytest.raises(ValueError, match="not properly setup"):
        s.solve(g, LaplacianOperator(-1.0), np.zeros((17, 17)))
    s.setup(g, LaplacianOperator(-1.0), RestrictionOperator(), ProlongationOperator())
    with pytest.raises(ValueError, match="grid mismatch"):
        s.solve(Grid(33, 33), LaplacianOperator(-1.0), np.zeros((33, 33)))


def test_default_operator_sign_diverges_like_reference():
    # with LaplacianOperator() (+1) the reference diverges by ~2.3x per cycle
    g = Grid(33, 33)
    op = LaplacianOperator()
    s = MultigridSolver(max_levels=4, max_iterations=6)
    s.setup(g, op, RestrictionOperator(), ProlongationOperator())
    _, info = s.solve(g, op, O.mms_rhs(33))
    h = info["residual_history"]
    assert not info["converged"] and abs(h[0] - 23.8) < 0.2 and h[-1] > 10 * h[0] and h[5] > h[4] > h[3]
