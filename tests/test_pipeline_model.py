"""The temporal-blocking schedule of the streaming kernel, replayed in NumPy against the oracle (no GPU)."""
import numpy as np
import pytest

from oracle import np_oracle as O
from pipeline_model import model


@pytest.mark.parametrize("jacobi", [True, False])
@pytest.mark.parametrize("nx,ny", [(33, 17), (65, 33), (9, 9), (5, 5)])
def test_row_pipeline_schedule(nx, ny, jacobi):
    rng = np.random.default_rng(0)
    hx, hy = 1.0 / (nx - 1), 1.0 / (ny - 1)
    u, f = rng.uniform(-1, 1, (nx, ny)), rng.uniform(-1, 1, (nx, ny))
    om = 2 / 3 if jacobi else 1.0
    for nu in (1, 2):
        exp = O.jacobi_smooth(u, f, hx, hy, om, nu) if jacobi else O.rbgs_smooth(u, f, hx, hy, om, nu)
        er = O.residual(exp, f, hx, hy, -1.0)
        erc = O.restrict(er)
        for back in (False, True):
            for rows in (8, 16, 64):
                out, res, rc = model(u, f, hx, hy, om, nu, jacobi, back, rows)
                assert np.array_equal(out, exp), (nu, back, rows)
                if back:
                    assert not np.isnan(rc).any() and not np.isnan(res).any()
                    assert np.allclose(res, er, rtol=0, atol=1e-9 * np.abs(er).max())
                    assert np.allclose(rc, erc, rtol=0, atol=1e-9 * np.abs(erc).max())
