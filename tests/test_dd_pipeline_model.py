"""The row schedule of the fused defect + down pass (csrc/mg_stream_dd.cuh) replayed in NumPy, tile by tile, against
the oracle's separate operators (no GPU): the 6 + 6 row halo is exactly the dependence cone -- one row less on either
side and the results of a tile depend on rows it never loaded."""
import numpy as np
import pytest

from dd_pipeline_model import fused_defect_down
from oracle import np_oracle as O


def _reference(u, f, e, hx, hy, shift):
    un = u + e.astype(np.float64)
    r64 = O.residual(un, f, hx, hy, -1.0, shift)
    r32 = r64.astype(np.float32)
    ep = O.rbgs_smooth(np.zeros_like(r32), r32, np.float32(hx), np.float32(hy), np.float32(1.0), 2, np.float32(shift))
    fc = O.restrict(O.residual(ep, r32, np.float32(hx), np.float32(hy), np.float32(-1.0), np.float32(shift)))
    return un, r32, float(np.sum(r64 ** 2)), ep, fc


@pytest.mark.parametrize("shift", [0.0, 12.5])
@pytest.mark.parametrize("nx,ny,rows", [(33, 17, 8), (65, 33, 16), (65, 9, 64), (17, 17, 6), (129, 9, 22)])
def test_row_schedule_and_halo(nx, ny, rows, shift):
    rng = np.random.default_rng(nx + ny + rows)
    hx, hy = 1.0 / (nx - 1), 1.0 / (ny - 1)
    u, f = rng.uniform(-1, 1, (nx, ny)), rng.uniform(-1, 1, (nx, ny)) * 10
    e = rng.uniform(-1, 1, (nx, ny)).astype(np.float32)
    for a in (u, f, e):
        a[0], a[-1], a[:, 0], a[:, -1] = 0, 0, 0, 0
    un, r32, ss, ep, fc = _reference(u, f, e, hx, hy, shift)
    g_u, g_r, g_ss, g_e, g_c = fused_defect_down(u, f, e, hx, hy, rows, shift)
    assert np.array_equal(g_u, un)
    tol = dict(rtol=0, atol=2e-6 * max(1.0, float(np.abs(r32).max())))
    assert not np.isnan(g_r).any() and np.allclose(g_r, r32, **tol)
    assert abs(g_ss - ss) <= 1e-12 * ss
    assert not np.isnan(g_e[1:-1, 1:-1]).any() and np.allclose(g_e[1:-1, 1:-1], ep[1:-1, 1:-1], rtol=0, atol=2e-6 * float(np.abs(ep).max()))
    assert not np.isnan(g_c).any() and np.allclose(g_c[1:-1, 1:-1], fc[1:-1, 1:-1], rtol=0, atol=2e-5 * float(np.abs(fc).max()))


def test_one_halo_row_less_is_wrong():
    """With 5 lead rows a tile's restricted residual is off by percent of its scale, with 5 tail rows its last coarse
    row is never emitted: the 6 + 6 halo of the kernel is the dependence cone, not a safety margin."""
    nx, ny, rows = 65, 17, 16
    rng = np.random.default_rng(5)
    hx, hy = 1.0 / (nx - 1), 1.0 / (ny - 1)
    u, f = rng.uniform(-1, 1, (nx, ny)), rng.uniform(-1, 1, (nx, ny)) * 10
    e = rng.uniform(-1, 1, (nx, ny)).astype(np.float32)
    for a in (u, f, e):
        a[0], a[-1], a[:, 0], a[:, -1] = 0, 0, 0, 0
    _, _, _, ep, fc = _reference(u, f, e, hx, hy, 0.0)
    scale = float(np.abs(fc).max())
    for kw in (dict(lead=5), dict(tail=5, tail_last=8)):
        _, _, _, g_e, g_c = fused_defect_down(u, f, e, hx, hy, rows, 0.0, **kw)
        bad = np.nan_to_num(np.abs(g_c[1:-1, 1:-1] - fc[1:-1, 1:-1]), nan=np.inf).max()
        assert bad > 1e-3 * scale, kw
