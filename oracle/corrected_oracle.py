"""CPU restatement of the reference's SECONDARY solver, `CorrectedMultigridSolver`
(src/multigrid/solvers/corrected_multigrid.py:24-418) -- the V-cycle its validation modules and tutorials actually run
(validation/simple_validation.py, mms_validation.py, examples/01_basic_poisson_tutorial.ipynb).

TEST INFRASTRUCTURE ONLY (like everything under oracle/): a checker, never shipped or measured.
Pinned: tests/golden/corrected_golden.npz holds runs of the reference's own class (make_golden_corrected.py); this file
reproduces them bit for bit (tests/test_oracle_golden.py).

Differences from the primary path (np_oracle.py), all deliberate in the reference:
  * smoother: lexicographic GS written as 0.25*(W + E + S + N + h^2 f), summed left to right (corrected_multigrid.py:263-270),
    boundary re-zeroed after every sweep (:213-216);
  * residual = f - (-lap_h u) with the boundary set to ZERO (:284-290), norm = unscaled Frobenius norm of the interior (:306-313);
  * full weighting on coarse interior points only, coarse boundary = 0 (:315-332); textbook bilinear prolongation incl. the
    last row / column (:334-362);
  * hierarchy: n -> max(5, (n-1)//2+1), at most max_levels levels, stop at <= 5 (:80-97); coarsest: <= 100 GS sweeps with
    two stopping tests (:364-387)."""
from __future__ import annotations

from typing import Any, Dict, List

import numpy as np


def _bc(u: np.ndarray) -> None:
    u[0, :] = 0.0
    u[-1, :] = 0.0
    u[:, 0] = 0.0
    u[:, -1] = 0.0


def gs_sweep(u: np.ndarray, rhs: np.ndarray, h: float) -> np.ndarray:
    """One lexicographic sweep (:263-270) along anti-diagonals: same operands, same order, same results."""
    v = u.copy()
    nx, ny = v.shape
    h2 = h ** 2
    for d in range(2, nx + ny - 3):
        i = np.arange(max(1, d - (ny - 2)), min(nx - 2, d - 1) + 1)
        j = d - i
        v[i, j] = 0.25 * (v[i - 1, j] + v[i + 1, j] + v[i, j - 1] + v[i, j + 1] + h2 * rhs[i, j])
    return v


def apply_laplacian(u: np.ndarray, h: float) -> np.ndarray:
    out = np.zeros_like(u)
    out[1:-1, 1:-1] = -(u[:-2, 1:-1] + u[2:, 1:-1] + u[1:-1, :-2] + u[1:-1, 2:] - 4 * u[1:-1, 1:-1]) / h ** 2
    return out


def residual(u: np.ndarray, rhs: np.ndarray, h: float) -> np.ndarray:
    r = rhs - apply_laplacian(u, h)
    _bc(r)
    return r


def residual_norm(u: np.ndarray, rhs: np.ndarray, h: float) -> float:
    return float(np.linalg.norm(residual(u, rhs, h)[1:-1, 1:-1]))


def restrict(f: np.ndarray, nxc: int, nyc: int) -> np.ndarray:
    nxf, nyf = f.shape
    c = np.zeros((nxc, nyc))
    I = np.arange(1, nxc - 1)
    J = np.arange(1, nyc - 1)
    I = I[2 * I < nxf - 1]
    J = J[2 * J < nyf - 1]
    if I.size == 0 or J.size == 0:
        return c
    a, b = np.meshgrid(2 * I, 2 * J, indexing="ij")
    c[np.ix_(I, J)] = (f[a - 1, b - 1] + 2 * f[a - 1, b] + f[a - 1, b + 1] + 2 * f[a, b - 1] + 4 * f[a, b] + 2 * f[a, b + 1]
                       + f[a + 1, b - 1] + 2 * f[a + 1, b] + f[a + 1, b + 1]) / 16.0
    return c


def prolongate(c: np.ndarray, nxf: int, nyf: int) -> np.ndarray:
    nxc, nyc = c.shape
    f = np.zeros((nxf, nyf))
    for i in range(nxc):
        fi = 2 * i
        if fi >= nxf:
            continue
        J = np.arange(nyc)
        J0 = J[2 * J < nyf]
        f[fi, 2 * J0] = c[i, J0]
        J1 = J[(2 * J + 1 < nyf) & (J + 1 < nyc)]
        f[fi, 2 * J1 + 1] = 0.5 * (c[i, J1] + c[i, J1 + 1])
        if fi + 1 < nxf and i + 1 < nxc:
            f[fi + 1, 2 * J0] = 0.5 * (c[i, J0] + c[i + 1, J0])
            f[fi + 1, 2 * J1 + 1] = 0.25 * (c[i, J1] + c[i + 1, J1] + c[i, J1 + 1] + c[i + 1, J1 + 1])
    return f


class OracleCorrectedMultigrid:
    def __init__(self, max_levels: int = 4, max_iterations: int = 50, tolerance: float = 1e-8, pre: int = 2, post: int = 2,
                 coarse_tolerance: float = 1e-12, coarse_max_iterations: int = 100):
        self.max_levels, self.max_iterations, self.tolerance = max_levels, max_iterations, tolerance
        self.pre, self.post = pre, post
        self.coarse_tolerance, self.coarse_max_iterations = coarse_tolerance, coarse_max_iterations
        self.shapes: List = []
        self.h: List[float] = []

    def setup(self, nx: int, ny: int, domain=(0.0, 1.0, 0.0, 1.0)) -> None:
        self.shapes, self.h = [(nx, ny)], [(domain[1] - domain[0]) / (nx - 1)]
        for _ in range(1, self.max_levels):
            cx, cy = max(5, (nx - 1) // 2 + 1), max(5, (ny - 1) // 2 + 1)
            self.shapes.append((cx, cy))
            self.h.append((domain[1] - domain[0]) / (cx - 1))
            nx, ny = cx, cy
            if cx <= 5 or cy <= 5:
                break

    def _coarsest(self, u, rhs, level):
        h = self.h[level]
        for _ in range(self.coarse_max_iterations):
            old = u.copy()
            u = gs_sweep(u, rhs, h)
            _bc(u)
            if residual_norm(u, rhs, h) < self.coarse_tolerance:
                break
            if np.linalg.norm(u - old) < self.coarse_tolerance:
                break
        return u

    def _v_cycle(self, u, rhs, level):
        if level == len(self.shapes) - 1:
            return self._coarsest(u, rhs, level)
        h = self.h[level]
        for _ in range(self.pre):
            u = gs_sweep(u, rhs, h)
            _bc(u)
        r = residual(u, rhs, h)
        rc = restrict(r, *self.shapes[level + 1])
        ec = self._v_cycle(np.zeros_like(rc), rc, level + 1)
        u = u + prolongate(ec, *self.shapes[level])
        _bc(u)
        for _ in range(self.post):
            u = gs_sweep(u, rhs, h)
            _bc(u)
        return u

    def solve(self, initial_guess: np.ndarray, rhs: np.ndarray, domain=(0.0, 1.0, 0.0, 1.0)) -> Dict[str, Any]:
        if not self.shapes or self.shapes[0] != rhs.shape:
            self.setup(rhs.shape[0], rhs.shape[1], domain)
        u = initial_guess.copy()
        _bc(u)
        h = self.h[0]
        hist = [residual_norm(u, rhs, h)]
        converged, it, rn = False, 0, hist[0]
        for it in range(1, self.max_iterations + 1):
            u = self._v_cycle(u, rhs, 0)
            rn = residual_norm(u, rhs, h)
            hist.append(rn)
            if rn < self.tolerance:
                converged = True
                break
        return {"solution": u, "converged": converged, "iterations": it, "final_residual": rn, "residual_history": hist}
