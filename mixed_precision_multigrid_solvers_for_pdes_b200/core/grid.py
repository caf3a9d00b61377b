"""Uniform vertex-centred grid, API-compatible with the reference ``multigrid.core.grid.Grid``
(reference core/grid.py:10-218).

Differences that do not change results: the coordinate meshes ``X``/``Y`` and the ``values`` /
``residual`` scratch arrays are created lazily (the reference allocates four full arrays per
grid, 8.6 GB of host RAM at 16385^2), and norms accept CUDA tensors (reduced on the device by
``mg_sumsq``)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


class Grid:
    def __init__(self, nx: int, ny: int, domain: Tuple[float, float, float, float] = (0.0, 1.0, 0.0, 1.0),
                 dtype=np.float64):
        if nx < 3 or ny < 3:
            raise ValueError("Grid must have at least 3 points in each direction")
        self.nx, self.ny, self.domain, self.dtype = nx, ny, domain, dtype
        self.hx = (domain[1] - domain[0]) / (nx - 1)
        self.hy = (domain[3] - domain[2]) / (ny - 1)
        self.h = min(self.hx, self.hy)
        self._x = self._y = self._X = self._Y = self._values = self._residual = None

    # -- lazily materialised host arrays (reference allocates them eagerly, grid.py:47-56) ------
    @property
    def x(self) -> np.ndarray:
        if self._x is None:
            self._x = np.linspace(self.domain[0], self.domain[1], self.nx, dtype=self.dtype)
        return self._x

    @property
    def y(self) -> np.ndarray:
        if self._y is None:
            self._y = np.linspace(self.domain[2], self.domain[3], self.ny, dtype=self.dtype)
        return self._y

    def _mesh(self):
        if self._X is None:
            self._X, self._Y = np.meshgrid(self.x, self.y, indexing="ij")

    @property
    def X(self) -> np.ndarray:
        self._mesh()
        return self._X

    @property
    def Y(self) -> np.ndarray:
        self._mesh()
        return self._Y

    @property
    def values(self) -> np.ndarray:
        if self._values is None:
            self._values = np.zeros((self.nx, self.ny), dtype=self.dtype)
        return self._values

    @values.setter
    def values(self, v) -> None:
        self._values = v

    @property
    def residual(self):
        if self._residual is None:
            self._residual = np.zeros((self.nx, self.ny), dtype=self.dtype)
        return self._residual

    @residual.setter
    def residual(self, v) -> None:
        self._residual = v

    # -- geometry -----------------------------------------------------------------------------
    @property
    def shape(self) -> Tuple[int, int]:
        return (self.nx, self.ny)

    @property
    def size(self) -> int:
        return self.nx * self.ny

    def interior_slice(self):
        return (slice(1, -1), slice(1, -1))

    def boundary_slice(self, side: str):
        table = {"left": (slice(0, 1), slice(None)), "right": (slice(-1, None), slice(None)),
                 "bottom": (slice(None), slice(0, 1)), "top": (slice(None), slice(-1, None))}
        if side not in table:
            raise ValueError(f"Unknown boundary side: {side}")
        return table[side]

    def apply_dirichlet_bc(self, value, side: Optional[str] = None) -> None:
        sides = ["left", "right", "bottom", "top"] if side in (None, "all") else [side]
        for s in sides:
            self.values[self.boundary_slice(s)] = value

    def apply_neumann_bc(self, derivative, side: str) -> None:
        v = self.values
        if side == "left":
            v[0, :] = v[1, :] - self.hx * derivative
        elif side == "right":
            v[-1, :] = v[-2, :] + self.hx * derivative
        elif side == "bottom":
            v[:, 0] = v[:, 1] - self.hy * derivative
        elif side == "top":
            v[:, -1] = v[:, -2] + self.hy * derivative
        else:
            raise ValueError(f"Unknown boundary side: {side}")

    def coarsen(self) -> "Grid":
        if (self.nx - 1) % 2 != 0 or (self.ny - 1) % 2 != 0:
            raise ValueError("Cannot coarsen grid: need even number of interior points")
        return Grid((self.nx - 1) // 2 + 1, (self.ny - 1) // 2 + 1, self.domain, self.dtype)

    def refine(self) -> "Grid":
        return Grid(2 * (self.nx - 1) + 1, 2 * (self.ny - 1) + 1, self.domain, self.dtype)

    # -- norms --------------------------------------------------------------------------------
    def l2_norm(self, field=None) -> float:
        """sqrt(hx*hy*sum(field**2)) over all points incl. the boundary (reference grid.py:174-187).
        CUDA tensors are reduced on the device (deterministic two-stage tree)."""
        if field is None:
            field = self.values
        if isinstance(field, np.ndarray):
            return np.sqrt(self.hx * self.hy * np.sum(field ** 2))
        from ..ops import sumsq  # device path
        return float(np.sqrt(self.hx * self.hy * sumsq(field)))

    def max_norm(self, field=None) -> float:
        if field is None:
            field = self.values
        if isinstance(field, np.ndarray):
            return np.max(np.abs(field))
        return float(field.abs().max().item())

    def copy(self) -> "Grid":
        g = Grid(self.nx, self.ny, self.domain, self.dtype)
        if self._values is not None:
            g.values = self._values.copy()
        if self._residual is not None:
            g.residual = self._residual.copy() if isinstance(self._residual, np.ndarray) else self._residual.clone()
        return g

    def __str__(self) -> str:
        return f"Grid({self.nx}x{self.ny}, h=({self.hx:.6f}, {self.hy:.6f}), dtype={self.dtype})"

    def __repr__(self) -> str:
        return f"Grid(nx={self.nx}, ny={self.ny}, domain={self.domain}, dtype={self.dtype})"
