"""Pin the NumPy oracle against outputs of the reference itself (tests/golden/*.npz).

fp64 must agree BIT FOR BIT: the oracle evaluates the same IEEE operations per point."""
import numpy as np
import pytest

from oracle import np_oracle as O


def _cases(meta):
    return [(m["key"], m["nx"], m["ny"], tuple(m["domain"]), m["dtype"]) for m in meta["ops"]]


def _eq(a, b, what):
    assert a.dtype == b.dtype, (what, a.dtype, b.dtype)
    assert a.shape == b.shape, what
    assert np.array_equal(a, b), f"{what}: max abs diff {np.max(np.abs(a.astype(np.float64) - b))}"


def test_ops_bitwise(ops_golden, golden_meta):
    G = ops_golden
    for key, nx, ny, dom, dt in _cases(golden_meta):
        g = O.OGrid(nx, ny, dom, np.dtype(dt).type)
        u, f, uc = G[f"{key}_u"], G[f"{key}_f"], G[f"{key}_uc"]
        for coeff in (1.0, -1.0, 2.5):
            _eq(O.apply_laplacian(u, g.hx, g.hy, coeff), G[f"{key}_apply_{coeff}"], f"{key} apply {coeff}")
            _eq(O.residual(u, f, g.hx, g.hy, coeff), G[f"{key}_residual_{coeff}"], f"{key} residual {coeff}")
        assert O.l2_norm(f, g.hx, g.hy) == float(G[f"{key}_l2"])
        for omega in (1.0, 1.3):
            for sweeps in (1, 2, 3):
                _eq(O.rbgs_smooth(u, f, g.hx, g.hy, omega, sweeps), G[f"{key}_rbgs_{omega}_{sweeps}"],
                    f"{key} rbgs {omega} {sweeps}")
            _eq(O.lexgs_smooth(u, f, g.hx, g.hy, omega, 2), G[f"{key}_lexgs_{omega}_2"], f"{key} lexgs {omega}")
        for sweeps in (1, 3):
            _eq(O.jacobi_smooth(u, f, g.hx, g.hy, 2.0 / 3.0, sweeps), G[f"{key}_jacobi_{sweeps}"], f"{key} jacobi")
            _eq(O.jacobi_smooth(u, f, g.hx, g.hy, 4.0 / 5.0, sweeps), G[f"{key}_wjacobi_{sweeps}"], f"{key} wjacobi")
        for m in ("full_weighting", "injection", "half_weighting"):
            _eq(O.restrict(u, m), G[f"{key}_restrict_{m}"], f"{key} restrict {m}")
        for m in ("bilinear", "injection"):
            _eq(O.prolong(uc, m), G[f"{key}_prolong_{m}"], f"{key} prolong {m}")


def test_restrict_output_dtype_follows_coarse_grid(ops_golden):
    u32 = ops_golden["probe_restrict_f32_in_f64_grid_u"]
    out = O.restrict(u32, "full_weighting", out_dtype=np.float64)
    _eq(out, ops_golden["probe_restrict_f32_in_f64_grid"], "f32 field -> f64 coarse grid")


def test_prolong_last_row_col_quirk():
    # SURVEY 8a probe: ones(5,5) -> 9x9 gives 1 everywhere except odd points of last row/column
    out = O.prolong(np.ones((5, 5)))
    exp = np.ones((9, 9))
    exp[1::2, 8] = 0
    exp[8, 1::2] = 0
    assert np.array_equal(out, exp)


def test_rbgs_probe_values():
    v = O.rbgs_smooth(np.zeros((9, 9)), np.ones((9, 9)), 0.125, 0.125, 1.0, 1)
    assert v[1, 1] == 0.00390625 and v[1, 2] == 0.0068359375


def _solver_for(m):
    sm = {"rbgs": ("rbgs", None), "lexgs": ("lexgs", None), "jacobi": ("jacobi", 2.0 / 3.0),
          "wjacobi": ("jacobi", 4.0 / 5.0)}[m["smoother"]]
    dt = np.dtype(m["dtype"]).type
    level_dtypes = None
    if m["precision"] == "mixed":
        L = m["num_levels"]
        level_dtypes = [np.float32 if l >= L // 2 else np.float64 for l in range(L)]  # precision.py:351-357
    return O.OracleMultigrid(m["nx"], m["ny"], max_levels=m["max_levels"], cycle_type=m["cycle"],
                             max_iterations=30 if m["dtype"] == "float64" else 12, smoother=sm[0], omega=sm[1],
                             dtype=dt, level_dtypes=level_dtypes)


def test_solves_match_reference(solve_golden, golden_meta):
    for m in golden_meta["solves"]:
        if m["nx"] > 129:
            continue
        s = _solver_for(m)
        f = O.mms_rhs(m["nx"], m["ny"], dtype=np.dtype(m["dtype"]).type)
        u, info = s.solve(f)
        hist = solve_golden[f"{m['name']}_hist"]
        assert info["iterations"] == m["iterations"], m["name"]
        assert info["converged"] == m["converged"], m["name"]
        assert info["num_levels"] == m["num_levels"]
        if m["dtype"] == "float64" and m["precision"] is None:
            # identical arithmetic; only np.sum's pairwise order inside the coarse-solve stopping
            # test is shared too, so the histories agree to the last bit
            assert np.array_equal(np.array(info["residual_history"]), hist), m["name"]
            assert np.array_equal(u, solve_golden[f"{m['name']}_u"]), m["name"]
        else:
            np.testing.assert_allclose(info["residual_history"], hist, rtol=1e-5, err_msg=m["name"])
            np.testing.assert_allclose(u, solve_golden[f"{m['name']}_u"], rtol=0, atol=1e-6)
        assert abs(np.max(np.abs(u - O.mms_exact(m["nx"], m["ny"]))) - m["max_error"]) < 1e-7


@pytest.mark.slow
def test_solve_257(solve_golden, golden_meta):
    m = [x for x in golden_meta["solves"] if x["name"] == "v257"][0]
    u, info = _solver_for(m).solve(O.mms_rhs(257))
    assert info["iterations"] == 8
    assert np.array_equal(np.array(info["residual_history"]), solve_golden["v257_hist"])
    assert abs(np.max(np.abs(u - O.mms_exact(257))) - m["max_error"]) < 1e-15


def test_closed_form_error():
    assert abs(O.mms_discretisation_error(129) - 5.020091592e-5) < 1e-13
    assert abs(O.mms_discretisation_error(16385) - 3.063928466e-9) < 1e-15


# ----------------------------------------------------------------------------------------------
# The plain-C oracle (oracle/mg_oracle.c, OpenMP) against the same golden vectors
# ----------------------------------------------------------------------------------------------
from oracle import c_oracle as CO  # noqa: E402


def test_c_oracle_ops_bitwise(ops_golden, golden_meta):
    G = ops_golden
    for key, nx, ny, dom, dt in _cases(golden_meta):
        g = O.OGrid(nx, ny, dom, np.dtype(dt).type)
        u, f, uc = G[f"{key}_u"], G[f"{key}_f"], G[f"{key}_uc"]
        for coeff in (1.0, -1.0, 2.5):
            _eq(CO.apply_laplacian(u, g.hx, g.hy, coeff), G[f"{key}_apply_{coeff}"], f"{key} apply {coeff}")
            _eq(CO.residual(u, f, g.hx, g.hy, coeff), G[f"{key}_residual_{coeff}"], f"{key} residual {coeff}")
        for omega in (1.0, 1.3):
            for sweeps in (1, 2, 3):
                _eq(CO.rbgs_smooth(u, f, g.hx, g.hy, omega, sweeps), G[f"{key}_rbgs_{omega}_{sweeps}"], f"{key} rbgs")
            _eq(CO.lexgs_smooth(u, f, g.hx, g.hy, omega, 2), G[f"{key}_lexgs_{omega}_2"], f"{key} lexgs")
        for sweeps in (1, 3):
            _eq(CO.jacobi_smooth(u, f, g.hx, g.hy, 2.0 / 3.0, sweeps), G[f"{key}_jacobi_{sweeps}"], f"{key} jacobi")
            _eq(CO.jacobi_smooth(u, f, g.hx, g.hy, 4.0 / 5.0, sweeps), G[f"{key}_wjacobi_{sweeps}"], f"{key} wjacobi")
        for m in ("full_weighting", "injection", "half_weighting"):
            _eq(CO.restrict(u, m), G[f"{key}_restrict_{m}"], f"{key} restrict {m}")
        for m in ("bilinear", "injection"):
            _eq(CO.prolong(uc, m), G[f"{key}_prolong_{m}"], f"{key} prolong {m}")
        ref = float(G[f"{key}_l2"])
        assert abs(CO.l2_norm(f, g.hx, g.hy) - ref) <= (1e-14 if dt == "float64" else 1e-6) * ref


def test_c_oracle_solves_match_reference(solve_golden, golden_meta):
    for m in golden_meta["solves"]:
        if m["dtype"] != "float64" or m["precision"] is not None:
            continue
        s = _solver_for(m)
        s.ops = CO
        u, info = s.solve(O.mms_rhs(m["nx"], m["ny"]))
        assert info["iterations"] == m["iterations"], m["name"]
        # same arithmetic; np_oracle.l2_norm (NumPy pairwise sum) still drives the stopping tests
        assert np.array_equal(np.array(info["residual_history"]), solve_golden[f"{m['name']}_hist"]), m["name"]
        if f"{m['name']}_u" in solve_golden:
            assert np.array_equal(u, solve_golden[f"{m['name']}_u"]), m["name"]


def test_c_oracle_large_grid_matches_numpy_oracle():
    n = 513
    f = O.mms_rhs(n)
    a = O.OracleMultigrid(n, max_levels=8, max_iterations=2)
    b = O.OracleMultigrid(n, max_levels=8, max_iterations=2, ops=CO)
    ua, ia = a.solve(f)
    ub, ib = b.solve(f)
    assert np.array_equal(ua, ub) and ia["residual_history"] == ib["residual_history"]


def test_corrected_multigrid_oracle_matches_reference_runs():
    """oracle/corrected_oracle.py (the reference's SECONDARY solver, corrected_multigrid.py) against runs of the reference's
    own class: hierarchies, cycle counts, residual histories and solutions, bit for bit."""
    import json
    import os

    from oracle import corrected_oracle as CM
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "corrected_golden.npz"))
    meta = json.loads(str(g["meta"]))
    assert meta["cases"]
    for c in meta["cases"]:
        s = CM.OracleCorrectedMultigrid(max_levels=c["levels"], max_iterations=c["max_iterations"], tolerance=c["tolerance"])
        rhs = g[c["name"] + "_rhs"]
        r = s.solve(np.zeros_like(rhs), rhs)
        assert [list(x) for x in s.shapes] == c["hierarchy"], c["name"]
        assert r["iterations"] == c["iterations"] and r["converged"] == c["converged"], c["name"]
        assert np.array_equal(np.array(r["residual_history"]), g[c["name"] + "_hist"]), c["name"]
        assert np.array_equal(r["solution"], g[c["name"] + "_u"]), c["name"]
