"""variable-coefficient -div(a grad u) = f.  No reference operator exists (parity UNPINNED): the kernels are
checked bit-for-bit against the repo's own NumPy statement of the discretisation, the multigrid solve through the
unchanged MultigridSolver against a manufactured solution (O(h^2))."""
import numpy as np
import pytest

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

from mixed_precision_multigrid_solvers_for_pdes_b200 import (Grid, LaplacianOperator, MultigridSolver,  # noqa: E402
                                                             ProlongationOperator, RestrictionOperator,
                                                             VariableCoefficientOperator, VariableCoefficientSmoother)


def _coef(X, Y):
    return 1.0 + 0.5 * np.sin(2 * np.pi * X) * np.cos(np.pi * Y) + X * Y


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n,m", [(33, 33), (17, 65), (129, 129)])
def test_kernels_match_own_oracle_bitwise(n, m, dt):
    rng = np.random.default_rng(31)
    g = Grid(n, m, (0.0, 1.0, 0.0, 2.0), dt)
    a = _coef(g.X, g.Y).astype(dt)
    u, f = rng.uniform(-1, 1, (n, m)).astype(dt), rng.uniform(-1, 1, (n, m)).astype(dt)
    for shift in (0.0, 37.5):
        op = VariableCoefficientOperator(a, shift)
        np.testing.assert_array_equal(op.apply(g, u), O.varcoef_apply(u, a, g.hx, g.hy, shift))
        np.testing.assert_array_equal(op.residual(g, u, f), O.varcoef_residual(u, f, a, g.hx, g.hy, shift))
        for omega in (1.0, 1.2):
            sm = VariableCoefficientSmoother(op, relaxation_parameter=omega)
            np.testing.assert_array_equal(sm.smooth(g, op, u, f, 2), O.varcoef_rbgs_smooth(u, f, a, g.hx, g.hy, omega, 2, shift))


def test_constant_coefficient_reduces_to_the_laplacian():
    rng = np.random.default_rng(32)
    g = Grid(65, 65)
    u, f = rng.uniform(-1, 1, (65, 65)), rng.uniform(-1, 1, (65, 65))
    op = VariableCoefficientOperator(np.ones((65, 65)))
    ref = LaplacianOperator(-1.0).residual(g, u, f)
    assert np.max(np.abs(op.residual(g, u, f) - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_multigrid_solve_with_variable_coefficients_is_second_order():
    errs = []
    for n in (65, 129, 257):
        g = Grid(n, n)
        a = _coef(g.X, g.Y)
        # manufactured: u = sin(pi x) sin(pi y);  f = -div(a grad u) = -(a_x u_x + a_y u_y) - a lap u
        sx, cx, sy, cy = np.sin(np.pi * g.X), np.cos(np.pi * g.X), np.sin(np.pi * g.Y), np.cos(np.pi * g.Y)
        ax = np.pi * np.cos(2 * np.pi * g.X) * np.cos(np.pi * g.Y) + g.Y
        ay = -0.5 * np.pi * np.sin(2 * np.pi * g.X) * np.sin(np.pi * g.Y) + g.X
        f = -(ax * np.pi * cx * sy + ay * np.pi * sx * cy) + a * 2 * np.pi ** 2 * sx * sy
        # boundary values of f enter no equation but do enter the reference's all-points residual norm
        # (SURVEY appendix A): zero them so that the convergence test measures the interior residual
        f[0, :] = f[-1, :] = f[:, 0] = f[:, -1] = 0.0
        op = VariableCoefficientOperator(a)
        sm = VariableCoefficientSmoother(op)
        s = MultigridSolver(max_levels=8, max_iterations=40, tolerance=1e-9)
        s.setup(g, op, RestrictionOperator(), ProlongationOperator(), smoother=sm,
                coarse_solver=VariableCoefficientSmoother(op, max_iterations=400, tolerance=1e-12))
        u, info = s.solve(g, op, f)
        assert info["converged"] and info["iterations"] <= 14, info["iterations"]
        errs.append(np.max(np.abs(u - sx * sy)))
    assert abs(np.log2(errs[0] / errs[1]) - 2) < 0.15 and abs(np.log2(errs[1] / errs[2]) - 2) < 0.15, errs


# ------------------------------------------------------------------------------------------------------------------
# Fused / temporally blocked passes of the variable-coefficient operator (mg_vcv_*), against the strict kernels above
# (which equal the repo's NumPy statement bit for bit): bit-exact when the spacings are powers of two.
# ------------------------------------------------------------------------------------------------------------------
import torch  # noqa: E402

from mixed_precision_multigrid_solvers_for_pdes_b200 import MixedPrecisionMultigrid, PoissonProblem, ops  # noqa: E402
from mixed_precision_multigrid_solvers_for_pdes_b200.device import empty_field, to_device, to_host  # noqa: E402


def _case(n, m, dt, dom, seed=41):
    rng = np.random.default_rng(seed)
    g = Grid(n, m, dom, dt)
    a = _coef(g.X, g.Y).astype(dt)
    u, f = rng.uniform(-1, 1, (n, m)).astype(dt), rng.uniform(-1, 1, (n, m)).astype(dt)
    return g, a, u, f


def _close(got, exp, exact, what):
    got = to_host(got) if isinstance(got, torch.Tensor) else got
    assert got.dtype == exp.dtype and got.shape == exp.shape, what
    if exact:
        if not np.array_equal(got, exp):
            d = np.abs(got.astype(np.float64) - exp.astype(np.float64))
            raise AssertionError(f"{what}: {np.count_nonzero(d)} mismatches, max {d.max():.3e} at "
                                 f"{np.unravel_index(np.argmax(d), d.shape)}")
    else:
        tol = 1e-12 if exp.dtype == np.float64 else 5e-6
        assert np.max(np.abs(got.astype(np.float64) - exp.astype(np.float64))) <= tol * np.max(np.abs(exp)), what


# (nx, ny, domain, exact): exact = power-of-two spacings in both directions
VSHAPES = [(129, 129, (0.0, 1.0, 0.0, 1.0), True), (65, 257, (0.0, 1.0, 0.0, 4.0), True),
           (257, 129, (0.0, 1.0, 0.0, 2.0), True), (33, 17, (0.0, 1.0, 0.0, 1.0), True),
           (131, 77, (0.0, 1.3, -0.2, 0.9), False), (513, 513, (0.0, 1.0, 0.0, 1.0), True)]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,dom,exact", VSHAPES)
def test_fused_varcoef_smoothing_passes_equal_the_strict_kernels(n, m, dom, exact, dt):
    g, a, u, f = _case(n, m, dt, dom)
    du, df, da = to_device(u)[0], to_device(f)[0], to_device(a)[0]
    for shift in (0.0, 37.5):
        for omega in (1.0, 1.2):
            for sweeps in ((1,) if dt == np.float64 else (1, 2)):
                out = empty_field(n, m, dt)
                ops.vc_pass(du, out, df, g.hx, g.hy, sweeps=sweeps, omega=omega, shift=shift, a=da)
                exp = O.varcoef_rbgs_smooth(u, f, a, g.hx, g.hy, omega, sweeps, shift)
                _close(out, exp, exact, f"smooth {n}x{m} {dt.__name__} s={sweeps} w={omega} shift={shift}")
    assert np.array_equal(to_host(du), u)
    # from the zero iterate without reading it
    out = empty_field(n, m, dt)
    ops.vc_pass(du, out, df, g.hx, g.hy, sweeps=1, a=da, u_zero=True)
    _close(out, O.varcoef_rbgs_smooth(np.zeros_like(u), f, a, g.hx, g.hy, 1.0, 1), exact, "u_zero")
    if dt == np.float64:
        with pytest.raises(Exception, match="status"):
            ops.vc_pass(du, out, df, g.hx, g.hy, sweeps=2, a=da)   # fp64: one sweep per pass


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,dom,exact", [s for s in VSHAPES if s[0] % 2 == 1 and s[1] % 2 == 1])
def test_fused_varcoef_transfer_and_norm_passes(n, m, dom, exact, dt):
    g, a, u, f = _case(n, m, dt, dom, seed=43)
    du, df, da = to_device(u)[0], to_device(f)[0], to_device(a)[0]
    nc, mc = (n - 1) // 2 + 1, (m - 1) // 2 + 1
    shift = 12.25
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    for sweeps in ((0, 1) if dt == np.float64 else (0, 1, 2)):
        # down: smooth + residual + full weighting
        out, cf = empty_field(n, m, dt), empty_field(nc, mc, dt)
        ops.vc_pass(du, out if sweeps else None, df, g.hx, g.hy, sweeps=sweeps, shift=shift, a=da, coarse_out=cf)
        us = O.varcoef_rbgs_smooth(u, f, a, g.hx, g.hy, 1.0, sweeps, shift) if sweeps else u
        r = O.varcoef_residual(us, f, a, g.hx, g.hy, shift)
        if sweeps:
            _close(out, us, exact, f"down u s={sweeps}")
        _close(cf, O.restrict(r), exact, f"down restrict {n}x{m} {dt.__name__} s={sweeps}")
        # up: prolongation + correction + smooth (+ norm)
        ec = np.random.default_rng(7).uniform(-1, 1, (nc, mc)).astype(dt)
        dec = to_device(ec)[0]
        if sweeps == 0:
            out2 = empty_field(n, m, dt)
            ops.vc_pass(du, out2, df, g.hx, g.hy, sweeps=0, shift=shift, a=da, coarse_in=dec)
            _close(out2, u + O.prolong(ec), exact, "prolong only")
            ops.vc_pass(du, None, df, g.hx, g.hy, sweeps=0, shift=shift, a=da, sumsq_out=ss)
            want = float(np.sum(O.varcoef_residual(u, f, a, g.hx, g.hy, shift).astype(np.float64) ** 2))
            assert abs(ss.item() - want) <= (1e-12 if dt == np.float64 else 1e-5) * want
            continue
        out2 = empty_field(n, m, dt)
        ops.vc_pass(du, out2, df, g.hx, g.hy, sweeps=sweeps, shift=shift, a=da, coarse_in=dec, sumsq_out=ss)
        up = O.varcoef_rbgs_smooth(u + O.prolong(ec), f, a, g.hx, g.hy, 1.0, sweeps, shift)
        _close(out2, up, exact, f"up {n}x{m} {dt.__name__} s={sweeps}")
        want = float(np.sum(O.varcoef_residual(up, f, a, g.hx, g.hy, shift).astype(np.float64) ** 2))
        assert abs(ss.item() - want) <= (1e-12 if dt == np.float64 else 1e-5) * want


def test_fused_varcoef_defect_pass_and_coarse_solve():
    n, m, dom = 257, 129, (0.0, 2.0, 0.0, 1.0)
    g, a, u, f = _case(n, m, np.float64, dom, seed=44)
    e = np.random.default_rng(8).uniform(-1e-3, 1e-3, (n, m)).astype(np.float32)
    du, df, da, de = to_device(u)[0], to_device(f)[0], to_device(a)[0], to_device(e)[0]
    out, r32 = empty_field(n, m, np.float64), empty_field(n, m, np.float32)
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    ops.vc_defect_pass(du, out, df, g.hx, g.hy, e_in=de, r_out=r32, sumsq_out=ss, shift=3.5, a=da)
    unew = u + e.astype(np.float64)
    r = O.varcoef_residual(unew, f, a, g.hx, g.hy, 3.5)
    assert np.array_equal(to_host(out), unew) and np.array_equal(to_host(r32), r.astype(np.float32))
    assert abs(ss.item() - float(np.sum(r ** 2))) <= 1e-12 * float(np.sum(r ** 2))
    # coarsest-level solver: RB-GS sweeps to tolerance in one launch == the host-driven loop of the smoother class
    gc, ac, _, fc = _case(9, 9, np.float64, (0.0, 1.0, 0.0, 1.0), seed=45)
    fc[0, :] = fc[-1, :] = fc[:, 0] = fc[:, -1] = 0.0
    op = VariableCoefficientOperator(ac, 2.0)
    us, info = VariableCoefficientSmoother(op, max_iterations=500, tolerance=1e-12).solve(gc, op, fc)
    dev_u, dev_info = empty_field(9, 9, np.float64), torch.zeros(2, dtype=torch.float64, device="cuda")
    ops.varcoef_coarse_solve_(dev_u, to_device(fc)[0], to_device(ac)[0], gc.hx, gc.hy, 2.0, 1.0, 1e-12, 500, info=dev_info)
    assert np.array_equal(to_host(dev_u), us) and int(dev_info[0].item()) == info["iterations"]


def test_fused_varcoef_4097_equals_strict_bitwise():
    n = 4097
    g = Grid(n, n)
    dev = torch.device("cuda")
    x = torch.linspace(0, 1, n, dtype=torch.float64, device=dev)
    a = empty_field(n, n, np.float64)
    a.copy_(1.0 + 0.5 * torch.sin(2 * np.pi * x)[:, None] * torch.cos(np.pi * x)[None, :] + x[:, None] * x[None, :])
    gen = torch.Generator(device=dev).manual_seed(3)
    u, f = empty_field(n, n, np.float64), empty_field(n, n, np.float64)
    u.copy_(torch.rand((n, n), generator=gen, dtype=torch.float64, device=dev) * 2 - 1)
    f.copy_(torch.rand((n, n), generator=gen, dtype=torch.float64, device=dev) * 2 - 1)
    out = empty_field(n, n, np.float64)
    ops.vc_pass(u, out, f, g.hx, g.hy, sweeps=1, a=a, shift=5.0)
    op = VariableCoefficientOperator(a, 5.0)
    ref = VariableCoefficientSmoother(op).smooth(g, op, u, f, 1)
    assert torch.equal(out, ref)
    for dt in (torch.float32,):
        u32, f32, a32 = (empty_field(n, n, dt) for _ in range(3))
        u32.copy_(u), f32.copy_(f), a32.copy_(a)
        out32 = empty_field(n, n, dt)
        ops.vc_pass(u32, out32, f32, g.hx, g.hy, sweeps=2, a=a32)
        op32 = VariableCoefficientOperator(a32)
        assert torch.equal(out32, VariableCoefficientSmoother(op32).smooth(g, op32, u32, f32, 2))


def _mms(g):
    sx, cx, sy, cy = np.sin(np.pi * g.X), np.cos(np.pi * g.X), np.sin(np.pi * g.Y), np.cos(np.pi * g.Y)
    a = _coef(g.X, g.Y)
    ax = np.pi * np.cos(2 * np.pi * g.X) * np.cos(np.pi * g.Y) + g.Y
    ay = -0.5 * np.pi * np.sin(2 * np.pi * g.X) * np.sin(np.pi * g.Y) + g.X
    f = -(ax * np.pi * cx * sy + ay * np.pi * sx * cy) + a * 2 * np.pi ** 2 * sx * sy
    return f, sx * sy


@pytest.mark.parametrize("strategy", ["double", "adaptive"])
def test_facade_with_variable_coefficients_is_second_order_and_equals_the_unfused_path(strategy):
    errs = []
    for n in (129, 257, 513):
        g = Grid(n, n)
        f, exact = _mms(g)
        s = MixedPrecisionMultigrid(strategy, coefficient=_coef, tolerance=1e-9)
        u, info = s.solve(PoissonProblem(rhs=f, nx=n, ny=n))
        assert info["converged"] and info["iterations"] <= 12, info["residual_history"]
        errs.append(np.max(np.abs(u - exact)))
        if n == 129 and strategy == "double":
            # the same cycles through the strict one-launch-per-sweep kernels: identical counts and history
            ub, ib = MixedPrecisionMultigrid(strategy, coefficient=_coef, tolerance=1e-9, kernels="basic").solve(
                PoissonProblem(rhs=f, nx=n, ny=n))
            assert ib["iterations"] == info["iterations"]
            np.testing.assert_allclose(ib["residual_history"], info["residual_history"], rtol=1e-9)
            assert np.max(np.abs(ub - u)) <= 1e-12
    assert abs(np.log2(errs[0] / errs[1]) - 2) < 0.15 and abs(np.log2(errs[1] / errs[2]) - 2) < 0.15, errs
    # a coefficient given as an array, with a Helmholtz shift (one implicit heat step): converges as well
    g = Grid(257, 257)
    f, _ = _mms(g)
    _, info = MixedPrecisionMultigrid("adaptive", coefficient=_coef(g.X, g.Y), shift=400.0, tolerance=1e-9).solve(
        PoissonProblem(rhs=f, nx=257, ny=257))
    assert info["converged"] and info["iterations"] <= 8
