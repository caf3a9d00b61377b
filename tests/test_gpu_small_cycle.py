"""mg_small_cycle (csrc/mg_small.cu): the coarse end of a cycle in one launch -- whole block on the larger levels, warp 0
alone on grids up to 17 x 17, the register-resident single-thread solver on the 5 x 5 coarsest grid -- against the
oracle's recursion (reference solvers/multigrid.py:253-337, solvers/base.py:258-285) and against the strict one-launch
coarse solver.  Dyadic grids: bit for bit; otherwise to rounding."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

from mixed_precision_multigrid_solvers_for_pdes_b200 import ops  # noqa: E402
from mixed_precision_multigrid_solvers_for_pdes_b200.device import to_device, to_host  # noqa: E402


def _oracle_cycle(n, m, nlev, cycle, u0, f, dt, cdt, domain, shift=0.0, ctol=1e-12, cmax=1000):
    s = O.OracleMultigrid(n, m, max_levels=nlev, cycle_type=cycle, dtype=cdt, domain=domain, shift=shift,
                          coarse_tolerance=ctol, coarse_max_iterations=cmax)
    assert len(s.grids) == nlev
    s.level_dtypes = [dt] * (nlev - 1) + [cdt]
    s.rhs[0] = f.astype(dt if nlev > 1 else cdt).copy()
    u = s._cycle(u0.astype(dt if nlev > 1 else cdt).copy(), 0)
    return u, s.coarse_sweeps


def _device_cycle(u0, f, g, nlev, cycle, cdt, shift=0.0, u_zero=False, ctol=1e-12, cmax=1000):
    du, df = to_device(u0.copy())[0], to_device(f)[0]
    info = torch.zeros(2, dtype=torch.float64, device="cuda")
    ops.small_cycle_(du, df, g.hx, g.hy, nlev=nlev, cycle_type=cycle, shift=shift, coarse_dtype=cdt, u_zero=u_zero,
                     coarse_tolerance=ctol, coarse_max_iterations=cmax, info=info)
    return to_host(du), info.cpu().numpy()


@pytest.mark.parametrize("cycle", ["V", "W", "F"])
@pytest.mark.parametrize("n,nlev,dt,cdt", [(9, 2, np.float64, np.float64), (17, 3, np.float64, np.float64),
                                           (33, 4, np.float64, np.float64), (65, 5, np.float64, np.float64),
                                           (65, 5, np.float32, np.float64), (129, 6, np.float32, np.float64),
                                           (33, 4, np.float32, np.float32), (65, 3, np.float64, np.float64)])
def test_sub_cycle_on_dyadic_grids_is_bitwise_the_reference_recursion(n, nlev, dt, cdt, cycle):
    rng = np.random.default_rng(7 * n + nlev)
    g = O.OGrid(n, n)
    f = rng.uniform(-1, 1, (n, n)).astype(dt)
    f[0, :] = f[-1, :] = f[:, 0] = f[:, -1] = 0          # the facade's zero ring: the zero-boundary register solver
    u0 = np.zeros((n, n), dt)
    u0[1:-1, 1:-1] = rng.uniform(-1, 1, (n - 2, n - 2))
    cmax = 1000 if cdt == np.float64 else 60             # an fp32 coarsest level never meets 1e-12: bound the sweeps
    want, sweeps = _oracle_cycle(n, n, nlev, cycle, u0, f, dt, cdt, (0.0, 1.0, 0.0, 1.0), cmax=cmax)
    got, info = _device_cycle(u0, f, g, nlev, cycle, cdt, cmax=cmax)
    if dt != cdt:
        # fp32 levels under an fp64 coarsest level: the reference interpolates in the GRID dtype (fp64) and rounds the
        # sum once (transfer.py:236, multigrid.py:329) where the kernel interpolates in fp32 -- fp32 rounding apart
        # (a W or F cycle revisits the levels and compounds the differences)
        assert np.max(np.abs(got - want.astype(dt))) <= (2e-6 if cycle == "V" else 2e-4) * np.max(np.abs(want))
        return
    assert np.array_equal(got, want.astype(dt))
    assert int(info[0]) == sweeps[-1]                    # sweep count of the last coarsest solve: same stopping decisions
    # zero-iterate flag == an explicit zero iterate
    z, _ = _device_cycle(rng.uniform(-1, 1, (n, n)).astype(dt), f, g, nlev, cycle, cdt, u_zero=True, cmax=cmax)
    wz, _ = _oracle_cycle(n, n, nlev, cycle, np.zeros((n, n), dt), f, dt, cdt, (0.0, 1.0, 0.0, 1.0), cmax=cmax)
    assert np.array_equal(z, wz.astype(dt))


@pytest.mark.parametrize("n,m,nlev,domain,shift", [(17, 17, 3, (0.0, 1.0, 0.0, 1.0), 37.5),      # Helmholtz: not exact
                                                   (17, 33, 3, (0.0, 1.0, 0.0, 1.0), 0.0),       # 5 x 9 coarsest, hx != hy
                                                   (33, 33, 4, (-1.0, 2.0, 0.5, 1.7), 0.0),      # non-dyadic spacing
                                                   (9, 17, 2, (0.0, 1.0, 0.0, 2.0), 0.0)])
def test_sub_cycle_on_general_grids_matches_to_rounding(n, m, nlev, domain, shift):
    rng = np.random.default_rng(n + 3 * m)
    g = O.OGrid(n, m, domain)
    f = rng.uniform(-1, 1, (n, m))
    u0 = rng.uniform(-1, 1, (n, m))
    for cycle in ("V", "W"):
        want, sweeps = _oracle_cycle(n, m, nlev, cycle, u0, f, np.float64, np.float64, domain, shift, cmax=200)
        got, info = _device_cycle(u0, f, g, nlev, cycle, np.float64, shift, cmax=200)
        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
        assert abs(int(info[0]) - sweeps[-1]) <= 1


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("ctol,cmax", [(1e-12, 1000), (1e-3, 1000), (0.0, 7), (1e-12, 3)])
def test_register_resident_5x5_solver_equals_the_strict_coarse_solver(dt, ctol, cmax):
    """nlev = 1: the launch IS the coarsest solve.  Non-zero boundary values of u and f (general variant) and the zero
    ring (lean variant) against mg_coarse_solve_lexgs: same bits, same sweep count, same final norm."""
    rng = np.random.default_rng(55)
    g = O.OGrid(5, 5)
    for ring in (True, False):
        f = rng.uniform(-1, 1, (5, 5)).astype(dt)
        u0 = rng.uniform(-1, 1, (5, 5)).astype(dt)
        if not ring:
            f[0, :] = f[-1, :] = f[:, 0] = f[:, -1] = 0
            u0[0, :] = u0[-1, :] = u0[:, 0] = u0[:, -1] = 0
        got, info = _device_cycle(u0, f, g, 1, "V", dt, ctol=ctol, cmax=cmax)
        du, df = to_device(u0.copy())[0], to_device(f)[0]
        ref_info = torch.zeros(2, dtype=torch.float64, device="cuda")
        ops.coarse_solve_lexgs_(du, df, g.hx, g.hy, 1.0, -1.0, ctol, cmax, info=ref_info)
        assert np.array_equal(got, to_host(du))
        ri = ref_info.cpu().numpy()
        assert int(info[0]) == int(ri[0]) and info[1] == ri[1], (info, ri)
        # ... and against the reference's loop itself
        want, sweeps = _oracle_cycle(5, 5, 1, "V", u0, f, dt, dt, (0.0, 1.0, 0.0, 1.0), ctol=ctol, cmax=cmax)
        assert np.array_equal(got, want) and int(info[0]) == sweeps[-1]
