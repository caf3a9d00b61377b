"""GPU parity of the basic kernel set (called through the C ABI) against
  (1) the golden outputs of the reference itself (tests/golden/ops_golden.npz), and
  (2) the NumPy oracle on fresh seeded inputs.
The basic kernels use strict IEEE arithmetic in the reference's operation order, so the bar is
BIT-EXACT for fp64 and fp32 (tolerance 0)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

mg = pytest.importorskip("mixed_precision_multigrid_solvers_for_pdes_b200")
from mixed_precision_multigrid_solvers_for_pdes_b200 import (GaussSeidelSmoother, Grid, JacobiSmoother,  # noqa: E402
                                                             LaplacianOperator, ProlongationOperator,
                                                             RestrictionOperator, SymmetricGaussSeidelSmoother,
                                                             WeightedJacobiSmoother, ops)
from mixed_precision_multigrid_solvers_for_pdes_b200.device import to_device, to_host  # noqa: E402


def _eq(a, b, what):
    assert a.dtype == b.dtype, (what, a.dtype, b.dtype)
    assert a.shape == b.shape, what
    if not np.array_equal(a, b):
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        raise AssertionError(f"{what}: {np.count_nonzero(d)} mismatches, max abs diff {d.max():.3e}")


def test_ops_match_reference_golden_bitwise(ops_golden, golden_meta):
    G = ops_golden
    for m in golden_meta["ops"]:
        key, dt = m["key"], np.dtype(m["dtype"]).type
        g = Grid(m["nx"], m["ny"], tuple(m["domain"]), dt)
        cg = g.coarsen()
        u, f, uc = G[f"{key}_u"], G[f"{key}_f"], G[f"{key}_uc"]
        for coeff in (1.0, -1.0, 2.5):
            op = LaplacianOperator(coeff)
            _eq(op.apply(g, u), G[f"{key}_apply_{coeff}"], f"{key} apply {coeff}")
            _eq(op.residual(g, u, f), G[f"{key}_residual_{coeff}"], f"{key} residual {coeff}")
        for omega in (1.0, 1.3):
            for sweeps in (1, 2, 3):
                s = GaussSeidelSmoother(relaxation_parameter=omega, red_black=True)
                _eq(s.smooth(g, None, u, f, sweeps), G[f"{key}_rbgs_{omega}_{sweeps}"], f"{key} rbgs {omega} {sweeps}")
            s = GaussSeidelSmoother(relaxation_parameter=omega)
            _eq(s.smooth(g, None, u, f, 2), G[f"{key}_lexgs_{omega}_2"], f"{key} lexgs {omega}")
        for sweeps in (1, 3):
            _eq(JacobiSmoother().smooth(g, None, u, f, sweeps), G[f"{key}_jacobi_{sweeps}"], f"{key} jacobi {sweeps}")
            _eq(WeightedJacobiSmoother().smooth(g, None, u, f, sweeps), G[f"{key}_wjacobi_{sweeps}"], f"{key} wjacobi")
        for meth in ("full_weighting", "injection", "half_weighting"):
            _eq(RestrictionOperator(meth).apply(g, u, cg), G[f"{key}_restrict_{meth}"], f"{key} restrict {meth}")
        for meth in ("bilinear", "injection"):
            _eq(ProlongationOperator(meth).apply(cg, uc, g), G[f"{key}_prolong_{meth}"], f"{key} prolong {meth}")
        # h-scaled L2 norm: fp64 accumulation on the device vs NumPy pairwise sum
        ref = float(G[f"{key}_l2"])
        got = g.l2_norm(to_device(f)[0])
        assert abs(got - ref) <= (1e-13 if dt is np.float64 else 2e-6) * ref


def test_inputs_untouched_and_types_preserved():
    g = Grid(17, 17)
    u = np.random.default_rng(1).uniform(-1, 1, (17, 17))
    f = np.ones((17, 17))
    u0 = u.copy()
    out = GaussSeidelSmoother(red_black=True).smooth(g, None, u, f, 2)
    assert isinstance(out, np.ndarray) and np.array_equal(u, u0) and not np.array_equal(out, u0)
    du, _ = to_device(u)
    dout = GaussSeidelSmoother(red_black=True).smooth(g, None, du, to_device(f)[0], 2)
    assert isinstance(dout, torch.Tensor) and dout.is_cuda and dout.data_ptr() != du.data_ptr()
    assert np.array_equal(to_host(dout), out) and np.array_equal(to_host(du), u0)
    # boundaries never move (reference tests/unit/test_iterative_solvers.py:62-72)
    assert np.array_equal(out[0], u0[0]) and np.array_equal(out[:, -1], u0[:, -1])


def test_restrict_dtype_follows_coarse_grid(ops_golden):
    u32 = ops_golden["probe_restrict_f32_in_f64_grid_u"]
    g = Grid(17, 17)
    out = RestrictionOperator().apply(g, u32, g.coarsen())
    _eq(out, ops_golden["probe_restrict_f32_in_f64_grid"], "fp32 field into fp64 coarse grid")


@pytest.mark.parametrize("n,m", [(33, 65), (129, 129), (257, 131), (3, 3), (5, 9)])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_ops_match_oracle_on_fresh_inputs(n, m, dt):
    rng = np.random.default_rng(n * 1000 + m)
    g = Grid(n, m, (0.0, 1.0, 0.0, 3.0), dt)
    u = rng.uniform(-1, 1, (n, m)).astype(dt)
    f = rng.uniform(-1, 1, (n, m)).astype(dt)
    _eq(LaplacianOperator(-1.0).residual(g, u, f), O.residual(u, f, g.hx, g.hy, -1.0), "residual")
    _eq(GaussSeidelSmoother(red_black=True).smooth(g, None, u, f, 4), O.rbgs_smooth(u, f, g.hx, g.hy, 1.0, 4), "rbgs")
    _eq(JacobiSmoother().smooth(g, None, u, f, 3), O.jacobi_smooth(u, f, g.hx, g.hy, 2 / 3, 3), "jacobi")
    _eq(GaussSeidelSmoother().smooth(g, None, u, f, 2), O.lexgs_smooth(u, f, g.hx, g.hy, 1.0, 2), "lexgs")
    if n >= 5 and m >= 5:
        cg = g.coarsen()
        _eq(RestrictionOperator().apply(g, u, cg), O.restrict(u), "restrict")
        uc = rng.uniform(-1, 1, cg.shape).astype(dt)
        _eq(ProlongationOperator().apply(cg, uc, g), O.prolong(uc), "prolong")


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n,m", [(1025, 513), (259, 1031), (2049, 2049), (35, 130), (66, 4), (3, 3), (34, 9)])
def test_lexicographic_gs_pipelined_over_warps_is_the_sequential_sweep(n, m, dt):
    """mg_smooth_lexgs runs the sweep as a skewed wavefront over many warps / CTAs (lexgs_pipe_kernel); the result is
    the sequential loop's (C oracle, smoothers.py:153-173) bit for bit: forward, backward and symmetric, any size."""
    from oracle import c_oracle as CO
    rng = np.random.default_rng(n + m)
    g = Grid(n, m, (0.0, 1.0, 0.0, 2.0), dt)
    u = rng.uniform(-1, 1, (n, m)).astype(dt)
    f = rng.uniform(-1, 1, (n, m)).astype(dt)
    fwd = CO.lexgs_smooth(u, f, g.hx, g.hy, 0.9, 2)
    du, df = to_device(u)[0], to_device(f)[0]
    _eq(to_host(ops.smooth_lexgs_(du, df, g.hx, g.hy, 0.9, 2, "forward")), fwd, "forward")
    flip = lambda a: np.ascontiguousarray(a[::-1, ::-1])  # noqa: E731
    bwd = flip(CO.lexgs_smooth(flip(u), flip(f), g.hx, g.hy, 0.9, 1))
    du = to_device(u)[0]
    _eq(to_host(ops.smooth_lexgs_(du, df, g.hx, g.hy, 0.9, 1, "backward")), bwd, "backward")
    sym = flip(CO.lexgs_smooth(flip(CO.lexgs_smooth(u, f, g.hx, g.hy, 0.9, 1)), flip(f), g.hx, g.hy, 0.9, 1))
    du = to_device(u)[0]
    _eq(to_host(ops.smooth_lexgs_(du, df, g.hx, g.hy, 0.9, 1, "symmetric")), sym, "symmetric")


def test_symmetric_gs_is_forward_then_backward():
    rng = np.random.default_rng(7)
    g = Grid(17, 33)
    u, f = rng.uniform(-1, 1, (17, 33)), rng.uniform(-1, 1, (17, 33))
    got = SymmetricGaussSeidelSmoother().smooth(g, None, u, f, 1)
    fwd = O.lexgs_smooth(u, f, g.hx, g.hy, 1.0, 1)
    # a backward lexicographic sweep is a forward sweep on the doubly flipped arrays
    exp = O.lexgs_smooth(fwd[::-1, ::-1].copy(), f[::-1, ::-1].copy(), g.hx, g.hy, 1.0, 1)[::-1, ::-1]
    _eq(got, np.ascontiguousarray(exp), "symmetric GS")


def test_mixed_precision_residual_and_axpy():
    rng = np.random.default_rng(3)
    n = 65
    g = Grid(n, n)
    u64, f64 = rng.uniform(-1, 1, (n, n)), rng.uniform(-1, 1, (n, n))
    du, df = to_device(u64)[0], to_device(f64)[0]
    r32 = ops.residual(du, df, g.hx, g.hy, -1.0, out_dtype=torch.float32)
    _eq(to_host(r32), O.residual(u64, f64, g.hx, g.hy, -1.0).astype(np.float32), "fp64 -> fp32 residual")
    u32, f32 = u64.astype(np.float32), f64.astype(np.float32)
    r64 = ops.residual(to_device(u32)[0], to_device(f32)[0], g.hx, g.hy, -1.0, out_dtype=torch.float64)
    _eq(to_host(r64), O.residual(u32.astype(np.float64), f32.astype(np.float64), g.hx, g.hy, -1.0), "fp32 -> fp64")
    e32 = to_device(rng.uniform(-1, 1, (n, n)).astype(np.float32))[0]
    y = du.clone()
    ops.axpy_(1.0, e32, y)
    _eq(to_host(y), u64 + to_host(e32).astype(np.float64), "fp64 += fp32")


def test_coarse_solve_matches_reference_loop():
    rng = np.random.default_rng(11)
    for n in (5, 9, 17):
        g = Grid(n, n)
        f = rng.uniform(-1, 1, (n, n))
        f[0, :] = f[-1, :] = f[:, 0] = f[:, -1] = 0.0  # otherwise the norm can never reach the tolerance
        s = O.OracleMultigrid(n, max_levels=1)
        s.rhs[0] = f
        exp = s._coarse_solve(0, np.zeros((n, n)))
        du, df = to_device(np.zeros((n, n)))[0], to_device(f)[0]
        info = torch.zeros(2, dtype=torch.float64, device="cuda")
        ops.coarse_solve_lexgs_(du, df, g.hx, g.hy, 1.0, -1.0, 1e-12, 1000, info=info)
        assert int(info[0].item()) == s.coarse_sweeps[-1]
        _eq(to_host(du), exp, f"coarse solve {n}")


def test_fill_and_maxerr():
    n = 129
    f = ops.fill_sinsin_(mg.device.empty_field(n, n, torch.float64), amplitude=2 * np.pi ** 2)
    ref = O.mms_rhs(n)
    np.testing.assert_allclose(to_host(f), ref, rtol=0, atol=1e-13)  # libm vs CUDA sin: ~1 ulp of 2 pi^2
    u = to_device(O.mms_exact(n) * 1.5)[0]
    assert abs(ops.maxerr_sinsin(u) - 0.5) < 1e-14
