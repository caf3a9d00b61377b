"""Device mirror of the reference's secondary solver (`CorrectedMultigridSolver`, corrected_multigrid.py) against
  (1) runs of the reference's own class (tests/golden/corrected_golden.npz, made by make_golden_corrected.py), and
  (2) the NumPy restatement oracle/corrected_oracle.py on larger grids.
Bar: the SOLUTION is bit-exact (every kernel keeps the reference's operand order in strict IEEE arithmetic); residual
norms are sums over a different tree than NumPy's pairwise summation, so histories are compared to 1e-12 relative and
cycle counts / convergence flags exactly."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from mixed_precision_multigrid_solvers_for_pdes_b200 import CorrectedMultigridSolver, Grid  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "corrected_golden.npz")


def test_reference_runs_are_reproduced_bit_for_bit():
    g = np.load(GOLD)
    meta = json.loads(str(g["meta"]))
    assert meta["cases"]
    for c in meta["cases"]:
        rhs = g[c["name"] + "_rhs"]
        s = CorrectedMultigridSolver(max_levels=c["levels"], max_iterations=c["max_iterations"], tolerance=c["tolerance"])
        r = s.solve(np.zeros_like(rhs), rhs, Grid(*rhs.shape))
        assert [list(x.shape) for x in s.grids] == c["hierarchy"], c["name"]
        assert r["iterations"] == c["iterations"] and r["converged"] == c["converged"], c["name"]
        np.testing.assert_allclose(r["residual_history"], g[c["name"] + "_hist"], rtol=1e-12, err_msg=c["name"])
        assert np.array_equal(r["solution"], g[c["name"] + "_u"]), c["name"]
        for k in ("solution", "converged", "iterations", "final_residual", "residual_history", "convergence_info"):
            assert k in r


@pytest.mark.parametrize("n,levels", [(129, 4), (257, 6), (65, 2), (100, 5)])
def test_larger_grids_against_the_oracle(n, levels):
    from oracle import corrected_oracle as CM
    grid = Grid(n, n)
    s = CorrectedMultigridSolver(max_levels=levels, max_iterations=12, tolerance=1e-9)
    rhs, u_exact = s.create_test_problem(grid, "manufactured")
    rng = np.random.default_rng(n)
    u0 = rng.uniform(-1, 1, (n, n))
    got = s.solve(u0, rhs, grid)
    o = CM.OracleCorrectedMultigrid(max_levels=levels, max_iterations=12, tolerance=1e-9)
    exp = o.solve(u0, rhs)
    assert [tuple(g.shape) for g in s.grids] == [tuple(x) for x in o.shapes]
    assert got["iterations"] == exp["iterations"] and got["converged"] == exp["converged"]
    np.testing.assert_allclose(got["residual_history"], exp["residual_history"], rtol=1e-12)
    assert np.array_equal(got["solution"], exp["solution"])
    if got["converged"]:
        assert np.max(np.abs(got["solution"] - u_exact)) < 5.0 / n ** 2 * np.pi ** 2


def test_kernels_one_by_one_against_the_oracle():
    import torch

    from mixed_precision_multigrid_solvers_for_pdes_b200 import _lib
    from mixed_precision_multigrid_solvers_for_pdes_b200.device import stream_ptr
    from oracle import corrected_oracle as CM
    rng = np.random.default_rng(3)
    for n, m in [(33, 33), (70, 41), (129, 65)]:
        h = 1.0 / (n - 1)
        u, f = rng.uniform(-1, 1, (n, m)), rng.uniform(-1, 1, (n, m))
        CM._bc(u)
        du, df = torch.tensor(u, device="cuda"), torch.tensor(f, device="cuda")
        _lib.call("mg_cm_gs", du.data_ptr(), df.data_ptr(), n, m, m, m, h, 3, stream_ptr())
        exp = u
        for _ in range(3):
            exp = CM.gs_sweep(exp, f, h)
        assert np.array_equal(du.cpu().numpy(), exp), "gs"
        dr = torch.empty_like(du)
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        ws = torch.empty(_lib.call("mg_cm_workspace_doubles"), dtype=torch.float64, device="cuda")
        _lib.call("mg_cm_residual", du.data_ptr(), df.data_ptr(), dr.data_ptr(), ss.data_ptr(), ws.data_ptr(), n, m, m, m, m,
                  h, stream_ptr())
        r = CM.residual(exp, f, h)
        assert np.array_equal(dr.cpu().numpy(), r), "residual"
        assert abs(np.sqrt(ss.item()) - CM.residual_norm(exp, f, h)) <= 1e-13 * CM.residual_norm(exp, f, h)
        nc, mc = max(5, (n - 1) // 2 + 1), max(5, (m - 1) // 2 + 1)
        dc = torch.full((nc, mc), float("nan"), dtype=torch.float64, device="cuda")
        _lib.call("mg_cm_restrict", dr.data_ptr(), dc.data_ptr(), n, m, nc, mc, m, mc, stream_ptr())
        rc = CM.restrict(r, nc, mc)
        assert np.array_equal(dc.cpu().numpy(), rc), "restrict"
        e = rng.uniform(-1, 1, (nc, mc))
        de = torch.tensor(e, device="cuda")
        _lib.call("mg_cm_prolong_add", de.data_ptr(), du.data_ptr(), nc, mc, n, m, mc, m, stream_ptr())
        exp2 = exp + CM.prolongate(e, n, m)
        CM._bc(exp2)
        assert np.array_equal(du.cpu().numpy(), exp2), "prolong + add"
