"""Drop-in ``RestrictionOperator`` / ``ProlongationOperator`` (reference operators/transfer.py)
on the GPU, including the reference's edge semantics: restriction injects on the coarse
boundary (transfer.py:109-113) and bilinear prolongation leaves the odd points of the last fine
row / column at zero (transfer.py:250-259).  Output dtype = the TARGET grid's dtype
(transfer.py:102,236)."""
from __future__ import annotations

from .. import ops
from ..device import like_input, to_device
from .base import BaseOperator


class RestrictionOperator(BaseOperator):
    def __init__(self, method: str = "full_weighting"):
        super().__init__(f"Restriction({method})")
        self.method = method
        if method not in ("injection", "full_weighting", "half_weighting"):
            raise ValueError(f"Unknown restriction method: {method}")

    def can_apply(self, fine_grid, coarse_grid) -> bool:
        return (coarse_grid.nx == (fine_grid.nx - 1) // 2 + 1 and coarse_grid.ny == (fine_grid.ny - 1) // 2 + 1)

    def apply(self, fine_grid, field, coarse_grid):
        if not self.can_apply(fine_grid, coarse_grid):
            raise ValueError(f"Cannot restrict from {fine_grid.shape} to {coarse_grid.shape}")
        if tuple(field.shape) != tuple(fine_grid.shape):
            raise ValueError(f"Field shape {tuple(field.shape)} doesn't match fine grid {fine_grid.shape}")
        if (fine_grid.nx - 1) % 2 or (fine_grid.ny - 1) % 2:
            raise ValueError(f"Cannot restrict from {fine_grid.shape}: need odd point counts")
        d, was_np = to_device(field)
        return like_input(ops.restrict(d, self.method, out_dtype=coarse_grid.dtype), was_np)


class ProlongationOperator(BaseOperator):
    def __init__(self, method: str = "bilinear"):
        super().__init__(f"Prolongation({method})")
        self.method = method
        if method not in ("injection", "bilinear"):
            raise ValueError(f"Unknown prolongation method: {method}")

    def can_apply(self, coarse_grid, fine_grid) -> bool:
        return (fine_grid.nx == 2 * (coarse_grid.nx - 1) + 1 and fine_grid.ny == 2 * (coarse_grid.ny - 1) + 1)

    def apply(self, coarse_grid, field, fine_grid):
        if not self.can_apply(coarse_grid, fine_grid):
            raise ValueError(f"Cannot prolongate from {coarse_grid.shape} to {fine_grid.shape}")
        if tuple(field.shape) != tuple(coarse_grid.shape):
            raise ValueError(f"Field shape {tuple(field.shape)} doesn't match coarse grid {coarse_grid.shape}")
        d, was_np = to_device(field)
        return like_input(ops.prolong(d, self.method, out_dtype=fine_grid.dtype), was_np)
