"""The skewed-wavefront schedule of the pipelined lexicographic Gauss-Seidel kernel, replayed in NumPy against the
oracle's sequential sweep (no GPU): forward and backward, grids that do not fill the last warp, single rows / columns."""
import numpy as np
import pytest

from lexgs_pipeline_model import sweep
from oracle import np_oracle as O


@pytest.mark.parametrize("nx,ny", [(37, 21), (70, 45), (5, 5), (34, 9), (35, 130), (3, 3), (66, 4)])
def test_skewed_wavefront_is_the_sequential_sweep(nx, ny):
    rng = np.random.default_rng(nx * 100 + ny)
    u, f = rng.uniform(-1, 1, (nx, ny)), rng.uniform(-1, 1, (nx, ny))
    hx, hy = 1 / (nx - 1), 1 / (ny - 1)
    assert np.array_equal(sweep(u, f, hx, hy, 0.9, True), O.lexgs_smooth(u.copy(), f, hx, hy, 0.9, 1))
    flip = lambda a: np.ascontiguousarray(a[::-1, ::-1])  # noqa: E731
    assert np.array_equal(sweep(u, f, hx, hy, 0.9, False), flip(O.lexgs_smooth(flip(u), flip(f), hx, hy, 0.9, 1)))
