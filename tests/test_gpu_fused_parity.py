"""GPU parity of the fused / temporally blocked passes (mg_vc_pass through the C ABI) against
the strict basic kernels and the NumPy oracle.

Bar: BIT-EXACT whenever hx^2, hy^2 and 2/hx^2+2/hy^2 are powers of two (SQUARE n = 2^k+1 grids on
the unit square -- every BASELINE config), because all reciprocal multiplications are then exact;
<= 1e-13 relative to max|.| otherwise (reciprocal-multiply + FMA vs true division)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

from mixed_precision_multigrid_solvers_for_pdes_b200 import Grid, ops  # noqa: E402
from mixed_precision_multigrid_solvers_for_pdes_b200.device import empty_field, to_device, to_host  # noqa: E402

LOADERS = ["tma", "cp_async"]


def _fields(n, m, dt, seed, domain=(0.0, 1.0, 0.0, 1.0)):
    rng = np.random.default_rng(seed)
    g = Grid(n, m, domain, dt)
    u = rng.uniform(-1, 1, (n, m)).astype(dt)
    f = rng.uniform(-1, 1, (n, m)).astype(dt)
    return g, u, f


def _cmp(got, exp, exact, what, rtol=1e-13):
    got = to_host(got) if isinstance(got, torch.Tensor) else got
    exp = np.ascontiguousarray(exp)
    assert got.dtype == exp.dtype and got.shape == exp.shape, what
    if exact:
        if not np.array_equal(got, exp):
            d = np.abs(got.astype(np.float64) - exp.astype(np.float64))
            idx = np.unravel_index(np.argmax(d), d.shape)
            raise AssertionError(f"{what}: {np.count_nonzero(d)} mismatches, max {d.max():.3e} at {idx}")
    else:
        scale = np.max(np.abs(exp))
        tol = rtol if exp.dtype == np.float64 else 2e-6
        assert np.max(np.abs(got.astype(np.float64) - exp.astype(np.float64))) <= tol * scale, what


# (nx, ny, exact): exact = square power-of-two spacing
SHAPES = [(129, 129, True), (257, 513, False), (65, 1025, False), (33, 17, False), (5, 5, True), (9, 241, False),
          (131, 77, False), (1025, 129, False), (513, 513, True)]


@pytest.mark.parametrize("loader", LOADERS)
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,pow2", SHAPES)
def test_smooth_pass(n, m, pow2, dt, loader):
    dom = (0.0, 1.0, 0.0, 1.0) if (pow2 or n in (257, 65, 33, 1025)) else (0.0, 1.3, -0.2, 0.9)
    g, u, f = _fields(n, m, dt, 5)
    g = Grid(n, m, dom, dt)
    du, df = to_device(u)[0], to_device(f)[0]
    for sweeps in (1, 2):
        for omega in (1.0, 1.15):
            out = empty_field(n, m, dt)
            ops.vc_pass(du, out, df, g.hx, g.hy, sweeps=sweeps, omega=omega, loader=loader)
            exp = O.rbgs_smooth(u, f, g.hx, g.hy, omega, sweeps)
            _cmp(out, exp, pow2 and omega == 1.0, f"smooth {n}x{m} {dt.__name__} s={sweeps} w={omega} {loader}")
    assert np.array_equal(to_host(du), u)  # input untouched (out of place)


@pytest.mark.parametrize("loader", LOADERS)
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,pow2", [s for s in SHAPES if s[0] % 2 == 1 and s[1] % 2 == 1 and s[0] >= 5 and s[1] >= 5])
def test_smooth_residual_restrict_pass(n, m, pow2, dt, loader):
    dom = (0.0, 1.0, 0.0, 1.0) if (pow2 or n in (257, 65, 33, 1025)) else (0.0, 1.3, -0.2, 0.9)
    _, u, f = _fields(n, m, dt, 6)
    g = Grid(n, m, dom, dt)
    du, df = to_device(u)[0], to_device(f)[0]
    nc, mc = (n - 1) // 2 + 1, (m - 1) // 2 + 1
    for sweeps in (0, 1, 2):
        out = empty_field(n, m, dt)
        rc = empty_field(nc, mc, dt)
        rc.fill_(7.0)
        ops.vc_pass(du, out if sweeps else None, df, g.hx, g.hy, sweeps=sweeps, coefficient=-1.0, coarse_out=rc,
                    loader=loader)
        us = O.rbgs_smooth(u, f, g.hx, g.hy, 1.0, sweeps)
        exp_rc = O.restrict(O.residual(us, f, g.hx, g.hy, -1.0))
        if sweeps:
            _cmp(out, us, pow2, f"u after smooth+restrict {n}x{m} s={sweeps} {loader}")
        _cmp(rc, exp_rc, pow2, f"restricted residual {n}x{m} {dt.__name__} s={sweeps} {loader}")


@pytest.mark.parametrize("loader", LOADERS)
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,pow2", [s for s in SHAPES if s[0] % 2 == 1 and s[1] % 2 == 1 and s[0] >= 5 and s[1] >= 5])
def test_prolong_correct_smooth_norm_pass(n, m, pow2, dt, loader):
    dom = (0.0, 1.0, 0.0, 1.0) if (pow2 or n in (257, 65, 33, 1025)) else (0.0, 1.3, -0.2, 0.9)
    _, u, f = _fields(n, m, dt, 7)
    g = Grid(n, m, dom, dt)
    nc, mc = (n - 1) // 2 + 1, (m - 1) // 2 + 1
    ec = np.random.default_rng(8).uniform(-1, 1, (nc, mc)).astype(dt)  # non-zero boundary on purpose
    du, df, dec = to_device(u)[0], to_device(f)[0], to_device(ec)[0]
    for sweeps in (0, 1, 2):
        out = empty_field(n, m, dt)
        ss = torch.zeros(1, dtype=torch.float64, device="cuda")
        ops.vc_pass(du, out, df, g.hx, g.hy, sweeps=sweeps, coarse_in=dec, sumsq_out=ss, loader=loader)
        uc = u + O.prolong(ec)
        us = O.rbgs_smooth(uc, f, g.hx, g.hy, 1.0, sweeps)
        _cmp(out, us, pow2, f"prolong+correct+smooth {n}x{m} {dt.__name__} s={sweeps} {loader}")
        r = O.residual(us, f, g.hx, g.hy, -1.0)
        exp_ss = float(np.sum(r.astype(np.float64) ** 2))
        assert abs(ss.item() - exp_ss) <= (1e-12 if dt is np.float64 else 1e-5) * exp_ss
        out2 = empty_field(n, m, dt)
        ops.vc_pass(du, out2, df, g.hx, g.hy, sweeps=sweeps, coarse_in=dec, loader=loader)  # without the norm stage
        assert torch.equal(out, out2)


JAC_SHAPES = [(129, 129, True), (257, 513, False), (33, 17, False), (5, 5, True), (9, 241, False), (1025, 129, False),
              (513, 513, True)]


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n,m,pow2", JAC_SHAPES)
def test_jacobi_passes(n, m, pow2, dt):
    """Damped Jacobi in the streaming kernel (MG_VC_JACOBI): plain sweeps, sweeps + residual + restriction,
    prolongation + correction + sweeps + norm, and the zero-iterate flag, against the oracle (smoothers.py:41-86).
    The relaxation blend is not contracted, so square power-of-two grids are bit-exact for EVERY omega."""
    dom = (0.0, 1.0, 0.0, 1.0) if (pow2 or n in (257, 33, 1025)) else (0.0, 1.3, -0.2, 0.9)
    _, u, f = _fields(n, m, dt, 21)
    g = Grid(n, m, dom, dt)
    nc, mc = (n - 1) // 2 + 1, (m - 1) // 2 + 1
    ec = np.random.default_rng(22).uniform(-1, 1, (nc, mc)).astype(dt)
    du, df, dec = to_device(u)[0], to_device(f)[0], to_device(ec)[0]
    zeros = np.zeros_like(u)
    for sweeps in (1, 2):
        for omega in (2.0 / 3.0, 0.8, 1.0):
            what = f"jacobi {n}x{m} {dt.__name__} s={sweeps} w={omega:.3f}"
            out = empty_field(n, m, dt)
            ops.vc_pass(du, out, df, g.hx, g.hy, sweeps=sweeps, omega=omega, smoother="jacobi")
            us = O.jacobi_smooth(u, f, g.hx, g.hy, omega, sweeps)
            _cmp(out, us, pow2, what + " smooth")
            # down pass
            out = empty_field(n, m, dt)
            rc = empty_field(nc, mc, dt)
            rc.fill_(7.0)
            ops.vc_pass(du, out, df, g.hx, g.hy, sweeps=sweeps, omega=omega, coarse_out=rc, smoother="jacobi")
            _cmp(out, us, pow2, what + " u of the down pass")
            _cmp(rc, O.restrict(O.residual(us, f, g.hx, g.hy, -1.0)), pow2, what + " restricted residual")
            # down pass from the zero iterate (iterate buffer poisoned, must not be read)
            out.fill_(float("nan"))
            poison = torch.full_like(out, float("nan"))
            ops.vc_pass(poison, out, df, g.hx, g.hy, sweeps=sweeps, omega=omega, coarse_out=rc, smoother="jacobi",
                        u_zero=True)
            uz = O.jacobi_smooth(zeros, f, g.hx, g.hy, omega, sweeps)
            _cmp(out, uz, pow2, what + " u from zero")
            _cmp(rc, O.restrict(O.residual(uz, f, g.hx, g.hy, -1.0)), pow2, what + " restricted residual from zero")
            # up pass
            ss = torch.zeros(1, dtype=torch.float64, device="cuda")
            out = empty_field(n, m, dt)
            ops.vc_pass(du, out, df, g.hx, g.hy, sweeps=sweeps, omega=omega, coarse_in=dec, sumsq_out=ss,
                        smoother="jacobi")
            up = O.jacobi_smooth(u + O.prolong(ec), f, g.hx, g.hy, omega, sweeps)
            _cmp(out, up, pow2, what + " up pass")
            r = O.residual(up, f, g.hx, g.hy, -1.0)
            exp_ss = float(np.sum(r.astype(np.float64) ** 2))
            assert abs(ss.item() - exp_ss) <= (1e-12 if dt is np.float64 else 1e-5) * exp_ss, what
            out2 = empty_field(n, m, dt)
            ops.vc_pass(du, out2, df, g.hx, g.hy, sweeps=sweeps, omega=omega, coarse_in=dec, smoother="jacobi")
            assert torch.equal(out, out2), what
    assert np.array_equal(to_host(du), u)


def test_jacobi_large_grid_against_basic_kernels():
    """4097^2: fused Jacobi passes == the strict one-launch-per-sweep kernel, bit for bit, any tile height."""
    n = 4097
    g = Grid(n, n)
    gen = torch.Generator(device="cuda").manual_seed(1)
    for dt in (torch.float64, torch.float32):
        u, f = empty_field(n, n, dt), empty_field(n, n, dt)
        u.copy_(torch.rand((n, n), generator=gen, device="cuda", dtype=dt) * 2 - 1)
        f.copy_(torch.rand((n, n), generator=gen, device="cuda", dtype=dt) * 2 - 1)
        ref = u.clone()
        ops.smooth_jacobi_(ref, f, g.hx, g.hy, 2.0 / 3.0, 2)
        rref = ops.restrict(ops.residual(ref, f, g.hx, g.hy, -1.0))
        for rows in (0, 64):
            out = empty_field(n, n, dt)
            rc = empty_field(2049, 2049, dt)
            ops.vc_pass(u, out, f, g.hx, g.hy, sweeps=2, omega=2.0 / 3.0, coarse_out=rc, smoother="jacobi", rows=rows)
            assert torch.equal(out, ref)
            assert torch.equal(rc, rref)


def test_jacobi_needs_the_tma_loader():
    from mixed_precision_multigrid_solvers_for_pdes_b200 import MGLibraryError
    a, b = empty_field(33, 33, np.float64), empty_field(33, 33, np.float64)
    with pytest.raises(MGLibraryError):
        ops.vc_pass(a, b, a, 0.1, 0.1, smoother="jacobi", loader="cp_async")
    with pytest.raises(ValueError, match="smoother"):
        ops.vc_pass(a, b, a, 0.1, 0.1, smoother="sor")


@pytest.mark.parametrize("rows", [32, 64, 128])
def test_tile_height_does_not_change_results(rows):
    n = 513
    g, u, f = _fields(n, n, np.float64, 9)
    du, df = to_device(u)[0], to_device(f)[0]
    ref = empty_field(n, n, np.float64)
    ops.vc_pass(du, ref, df, g.hx, g.hy, sweeps=2)
    out = empty_field(n, n, np.float64)
    rc = empty_field(257, 257, np.float64)
    ops.vc_pass(du, out, df, g.hx, g.hy, sweeps=2, coarse_out=rc, rows=rows)
    assert torch.equal(out, ref)
    _cmp(rc, O.restrict(O.residual(to_host(ref), f, g.hx, g.hy, -1.0)), True, "restrict rows override")


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("case", ["shift", "non_dyadic", "iso_non_dyadic", "sor"])
def test_results_do_not_depend_on_tiling_when_scalings_are_inexact(case, dt):
    """With a Helmholtz shift, non-dyadic spacings or omega != 1 the scalings round, so every rounding of the point
    update must be pinned: the interior fast path and the masked edge path (which rows fall in which depends on the
    tile height / the slab) have to agree bit for bit.  (Found on 2 GPUs: nvcc fused `* 1/diag` into the next
    stage's add in the fast path only.)"""
    n = 513
    dom = {"shift": (0.0, 1.0, 0.0, 1.0), "non_dyadic": (0.0, 1.3, -0.2, 0.9), "iso_non_dyadic": (0.0, 1.1, 0.0, 1.1),
           "sor": (0.0, 1.0, 0.0, 1.0)}[case]
    shift = 777.7 if case == "shift" else 0.0
    omega = 1.15 if case == "sor" else 1.0
    _, u, f = _fields(n, n, dt, 31)
    g = Grid(n, n, dom, dt)
    ec = np.random.default_rng(32).uniform(-1, 1, (257, 257)).astype(dt)
    du, df, dec = to_device(u)[0], to_device(f)[0], to_device(ec)[0]
    for smoother, om in (("rbgs", omega), ("jacobi", 0.8)):
        ref = None
        for rows in (0, 8, 32, 128):
            down, rc = empty_field(n, n, dt), empty_field(257, 257, dt)
            ops.vc_pass(du, down, df, g.hx, g.hy, sweeps=2, omega=om, coarse_out=rc, rows=rows, shift=shift,
                        smoother=smoother)
            up = empty_field(n, n, dt)
            ops.vc_pass(du, up, df, g.hx, g.hy, sweeps=2, omega=om, coarse_in=dec, rows=rows, shift=shift,
                        smoother=smoother)
            if ref is None:
                ref = (down, rc, up)
            else:
                assert torch.equal(down, ref[0]) and torch.equal(rc, ref[1]) and torch.equal(up, ref[2]), (smoother, rows)


def test_alignment_is_checked():
    from mixed_precision_multigrid_solvers_for_pdes_b200 import MGLibraryError
    t = torch.zeros(33 * 40 + 1, dtype=torch.float64, device="cuda")[1:].view(33, 40)[:, :33]  # 8-byte offset base
    a = empty_field(33, 33, np.float64)
    with pytest.raises(MGLibraryError, match="align"):
        ops.vc_pass(t, a, a, 0.1, 0.1)
    with pytest.raises(MGLibraryError):
        ops.vc_pass(a, a, a, 0.1, 0.1)  # in place is not allowed


def test_large_grid_against_basic_kernels():
    """4097^2 (too slow for the loop reference): fused passes == strict basic kernels, bit for bit."""
    n = 4097
    g = Grid(n, n)
    gen = torch.Generator(device="cuda").manual_seed(0)
    for dt in (torch.float64, torch.float32):
        u = empty_field(n, n, dt)
        f = empty_field(n, n, dt)
        u.copy_(torch.rand((n, n), generator=gen, device="cuda", dtype=dt) * 2 - 1)
        f.copy_(torch.rand((n, n), generator=gen, device="cuda", dtype=dt) * 2 - 1)
        out = empty_field(n, n, dt)
        rc = empty_field(2049, 2049, dt)
        ops.vc_pass(u, out, f, g.hx, g.hy, sweeps=2, coarse_out=rc)
        ref = u.clone()
        ops.smooth_rbgs_(ref, f, g.hx, g.hy, 1.0, 2)
        assert torch.equal(out, ref)
        assert torch.equal(rc, ops.restrict(ops.residual(ref, f, g.hx, g.hy, -1.0)))


@pytest.mark.parametrize("loader", LOADERS)
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_zero_input_flag(dt, loader):
    """u_zero: the kernel must behave exactly as if it had read an all-zero iterate (and must not read it)."""
    n, m = 257, 129
    _, _, f = _fields(n, m, dt, 12)
    g = Grid(n, m, dtype=dt)
    df = to_device(f)[0]
    zeros = empty_field(n, m, dt)
    poison = empty_field(n, m, dt)
    poison.fill_(float("nan"))
    for sweeps in (1, 2):
        a, b = empty_field(n, m, dt), empty_field(n, m, dt)
        ra, rb = empty_field(129, 65, dt), empty_field(129, 65, dt)
        ops.vc_pass(zeros, a, df, g.hx, g.hy, sweeps=sweeps, coarse_out=ra, loader=loader)
        ops.vc_pass(poison, b, df, g.hx, g.hy, sweeps=sweeps, coarse_out=rb, loader=loader, u_zero=True)
        assert torch.equal(a, b) and torch.equal(ra, rb)


@pytest.mark.parametrize("loader", LOADERS)
@pytest.mark.parametrize("n,m", [(129, 129), (257, 65), (33, 1025), (9, 9)])
def test_defect_pass(n, m, loader):
    """u_out = u + e32 ; r32 = fp32(f - A u_out) ; sum r^2  -- against the oracle in fp64."""
    rng = np.random.default_rng(13)
    g = Grid(n, m)
    u, f = rng.uniform(-1, 1, (n, m)), rng.uniform(-1, 1, (n, m))
    e = rng.uniform(-1, 1, (n, m)).astype(np.float32)
    du, df, de = to_device(u)[0], to_device(f)[0], to_device(e)[0]
    uo = empty_field(n, m, np.float64)
    r32 = empty_field(n, m, np.float32)
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    ops.vc_defect_pass(du, uo, df, g.hx, g.hy, e_in=de, r_out=r32, sumsq_out=ss, loader=loader)
    un = u + e.astype(np.float64)
    assert np.array_equal(to_host(uo), un)
    r = O.residual(un, f, g.hx, g.hy, -1.0)
    exact = n == m
    _cmp(r32, r.astype(np.float32), exact, f"defect residual {n}x{m} {loader}", rtol=1e-13)
    assert abs(ss.item() - np.sum(r ** 2)) <= 1e-12 * np.sum(r ** 2)
    # residual only (no correction): nothing stored, same residual of the unchanged iterate
    r32b = empty_field(n, m, np.float32)
    ops.vc_defect_pass(du, None, df, g.hx, g.hy, r_out=r32b, sumsq_out=ss, loader=loader)
    _cmp(r32b, O.residual(u, f, g.hx, g.hy, -1.0).astype(np.float32), exact, "residual only")
    # update only
    uo2 = empty_field(n, m, np.float64)
    ops.vc_defect_pass(du, uo2, df, g.hx, g.hy, e_in=de, loader=loader)
    assert torch.equal(uo2, uo)


@pytest.mark.parametrize("shift", [0.0, 37.5])
@pytest.mark.parametrize("n,m", [(129, 129), (257, 65), (33, 1025), (9, 9), (17, 33), (513, 513), (1025, 257), (449, 129)])
def test_defect_down_pass(n, m, shift):
    """ONE pass (mg_stream_dd.cuh) = the defect pass followed by the first pass of the fp32 error cycle, bit for bit:
    u += e ; r32 ; ||r|| ; e' = 2 RB-GS sweeps from zero on A e = r32 ; f_c = R(r32 - A e')."""
    rng = np.random.default_rng(17)
    g = Grid(n, m)
    u, f = rng.uniform(-1, 1, (n, m)), rng.uniform(-1, 1, (n, m)) * 1e3
    e = rng.uniform(-1, 1, (n, m)).astype(np.float32)
    for a in (u, f, e):  # homogeneous Dirichlet ring, as in the solver
        a[0, :] = a[-1, :] = 0
        a[:, 0] = a[:, -1] = 0
    du, df, de = to_device(u)[0], to_device(f)[0], to_device(e)[0]
    nc, mc = (n + 1) // 2, (m + 1) // 2
    for with_e in (True, False):
        for u_zero in (False, True):
            if with_e and u_zero:
                uin = torch.full_like(du, float("nan"))  # U_ZERO: never read
            else:
                uin = du
            # two-launch reference
            uo = empty_field(n, m, np.float64)
            r32, eo, tmp = (empty_field(n, m, np.float32) for _ in range(3))
            co = empty_field(nc, mc, np.float32)
            ss = torch.zeros(1, dtype=torch.float64, device="cuda")
            ops.vc_defect_pass(uin, uo if with_e else None, df, g.hx, g.hy, e_in=de if with_e else None, r_out=r32,
                               sumsq_out=ss, shift=shift, u_zero=u_zero)
            ops.vc_pass(tmp, eo, r32, g.hx, g.hy, sweeps=2, coarse_out=co, u_zero=True, shift=shift)
            # fused
            uo2 = empty_field(n, m, np.float64)
            r32b, eo2 = empty_field(n, m, np.float32), empty_field(n, m, np.float32)
            co2 = empty_field(nc, mc, np.float32)
            ss2 = torch.zeros(1, dtype=torch.float64, device="cuda")
            for rows in (0, 8, 22, 40):  # tile heights: the row halo must cover the dependence cone exactly
                for t in (uo2, r32b, eo2, co2):
                    t.fill_(float("nan"))
                ops.vc_defect_down_pass(uin, uo2 if with_e else None, df, g.hx, g.hy, e_in=de if with_e else None,
                                        r_out=r32b, e_out=eo2, coarse_out=co2, sumsq_out=ss2, shift=shift,
                                        u_zero=u_zero, rows=rows)
                what = f"{n}x{m} e={with_e} u_zero={u_zero} shift={shift} rows={rows}"
                if with_e:
                    assert torch.equal(uo2, uo), "iterate " + what
                _cmp(r32b, to_host(r32), True, "residual " + what)
                _cmp(eo2[1:-1, 1:-1], to_host(eo)[1:-1, 1:-1], True, "pre-smoothed error " + what)
                _cmp(co2[1:-1, 1:-1], to_host(co)[1:-1, 1:-1], True, "restricted residual " + what)
                assert abs(ss2.item() - ss.item()) <= 1e-12 * ss.item(), "norm " + what
                assert torch.count_nonzero(eo2[0]) == 0 and torch.count_nonzero(eo2[:, 0]) == 0 and \
                    torch.count_nonzero(eo2[-1]) == 0 and torch.count_nonzero(eo2[:, -1]) == 0, "ring " + what


def test_defect_down_refinement_matches_two_pass_refinement():
    """The solver with the fused defect + down pass walks through the SAME residual history as with two launches."""
    from mixed_precision_multigrid_solvers_for_pdes_b200 import MixedPrecisionMultigrid, PoissonProblem
    hist = []
    for dd in (True, False):
        s = MixedPrecisionMultigrid(tolerance=1e-9, use_fused_defect_down="always" if dd else False)
        u, info = s.solve(PoissonProblem.manufactured(1025))
        assert s._dd_ok() == dd
        hist.append((info["residual_history"], u.copy()))
    # the two kernels sum the squared residual over different strip widths: same norm up to summation order
    assert len(hist[0][0]) >= 8 and len(hist[0][0]) == len(hist[1][0])
    np.testing.assert_allclose(hist[0][0], hist[1][0], rtol=1e-12)
    assert np.array_equal(hist[0][1], hist[1][1])


def test_last_cycle_hint_never_changes_results():
    """`last_hint` only chooses between the fused pass and its two launches (bit-identical): any sequence of hints,
    right or wrong, walks through the same iterates as the two-launch cycle."""
    from mixed_precision_multigrid_solvers_for_pdes_b200 import MixedPrecisionMultigrid
    n = 1025
    out = []
    for dd, hints in ((False, [False] * 6), (True, [False, True, True, False, True, False]), (True, [True] * 6)):
        s = MixedPrecisionMultigrid(precision_strategy="refinement", tolerance=1e-30,
                                    use_fused_defect_down="always" if dd else False,
                                    use_cuda_graphs=True)
        s.setup(n, n)
        b64 = s._engine.levels[0].bufs(torch.float64)
        ops.fill_sinsin_(b64.f, (0.0, 1.0, 0.0, 1.0), 2 * np.pi ** 2, 1.0, 1.0)
        ops.zero_ring_(b64.f)
        norms = [s._refinement_residual(u_zero=True)]
        for k, h in enumerate(hints):
            norms.append(s._cycle_refinement(u_zero=(k == 0), last_hint=h))
        out.append((norms, s._engine.levels[0].bufs(torch.float64).u.clone()))
    for norms, u in out[1:]:
        np.testing.assert_allclose(norms, out[0][0], rtol=1e-12)
        assert torch.equal(u, out[0][1])
