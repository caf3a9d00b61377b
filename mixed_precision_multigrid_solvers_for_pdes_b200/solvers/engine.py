"""Device-resident multigrid cycle engine.

Owns the grid hierarchy in HBM (pitched torch buffers per level and dtype) and issues the kernel
sequence of one V/W/F cycle exactly in the order of the reference recursion
(solvers/multigrid.py:253-337):

    pre-smooth -> residual -> restrict -> zero coarse guess -> 1/2/2^(L-l-2) recursive calls
    -> prolong + add -> post-smooth;      coarsest level: lexicographic-GS solve to tolerance.

Two kernel sets implement the same sequence:
  * ``basic``  -- one launch per reference method (strict arithmetic; also the on-GPU cross-check);
  * ``fused``  -- temporally blocked red-black GS with fused residual+restriction and fused
                 prolongation+correction+post-smooth (mg_vcycle.cu), used when available.
No host synchronisation happens inside a cycle, so a cycle can be captured in a CUDA graph."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from .. import ops
from ..device import empty_field, require_cuda, torch_dtype


class _Buffers:
    __slots__ = ("u", "tmp", "f")

    def __init__(self, nx, ny, dtype, dev):
        self.u = empty_field(nx, ny, dtype, dev)
        self.tmp = empty_field(nx, ny, dtype, dev)
        self.f = empty_field(nx, ny, dtype, dev)


class _Level:
    def __init__(self, grid, dev):
        self.grid, self.dev = grid, dev
        self._bufs: Dict[torch.dtype, _Buffers] = {}

    def bufs(self, dtype) -> _Buffers:
        dt = torch_dtype(dtype)
        b = self._bufs.get(dt)
        if b is None:
            b = self._bufs[dt] = _Buffers(self.grid.nx, self.grid.ny, dt, self.dev)
        return b


class CycleEngine:
    def __init__(self, grids: Sequence, *, smoother, coarse_solver, operators: Sequence, restriction_ops: Sequence,
                 prolongation_ops: Sequence, cycle_type: str = "V", pre: int = 2, post: int = 2,
                 kernels: str = "auto", device=None):
        self.dev = require_cuda(device)
        self.levels: List[_Level] = [_Level(g, self.dev) for g in grids]
        self.smoother, self.coarse_solver = smoother, coarse_solver
        self.operators, self.restriction_ops, self.prolongation_ops = list(operators), list(restriction_ops), list(prolongation_ops)
        self.cycle_type, self.pre, self.post = cycle_type, pre, post
        self.kernels = kernels
        self.coarse_info = torch.zeros(2, dtype=torch.float64, device=self.dev)

    # -- per-level building blocks ----------------------------------------------------------------
    @property
    def num_levels(self) -> int:
        return len(self.levels)

    def _smooth(self, lvl: int, b: _Buffers, sweeps: int) -> None:
        g = self.levels[lvl].grid
        sm = self.smoother
        kind = getattr(sm, "kind", "custom")
        if kind == "jacobi":
            ops.smooth_jacobi_(b.u, b.f, g.hx, g.hy, sm.omega, sweeps, tmp=b.tmp)
        elif kind in ("rbgs", "lexgs", "sgs"):
            sm._smooth_device_(g, b.u, b.f, sweeps)
        else:  # foreign smoother object: use its public protocol on device tensors
            b.u.copy_(sm.smooth(g, self.operators[lvl], b.u, b.f, sweeps))

    def _residual(self, lvl: int, u, f, out):
        g = self.levels[lvl].grid
        op = self.operators[lvl]
        coeff = getattr(op, "coefficient", None)
        if coeff is not None and type(op).__name__ == "LaplacianOperator":
            return ops.residual(u, f, g.hx, g.hy, coeff, out=out)
        out.copy_(op.residual(g, u, f))
        return out

    def _coarse_solve(self, lvl: int, b: _Buffers, precision_manager=None) -> None:
        g = self.levels[lvl].grid
        cs = self.coarse_solver
        op = self.operators[lvl]
        coeff = getattr(op, "coefficient", None)
        if getattr(cs, "kind", None) == "lexgs" and coeff is not None and type(op).__name__ == "LaplacianOperator":
            ops.coarse_solve_lexgs_(b.u, b.f, g.hx, g.hy, cs.omega, coeff, cs.tolerance, cs.max_iterations,
                                    info=self.coarse_info)
        else:  # any other IterativeSolver: its own solve loop (host-checked convergence)
            sol, _ = cs.solve(g, op, b.f, b.u, precision_manager)
            b.u.copy_(sol)

    # -- the recursion ------------------------------------------------------------------------------
    def cycle(self, level_dtypes: Sequence, lvl: int = 0, precision_manager=None) -> None:
        """One cycle on level `lvl`, in place on that level's ``u`` for ``level_dtypes[lvl]``."""
        L = self.num_levels
        b = self.levels[lvl].bufs(level_dtypes[lvl])
        if lvl == L - 1:
            self._coarse_solve(lvl, b, precision_manager)
            return
        if self.pre > 0:
            self._smooth(lvl, b, self.pre)
        c = self.levels[lvl + 1].bufs(level_dtypes[lvl + 1])
        r = self._residual(lvl, b.u, b.f, b.tmp)
        rop = self.restriction_ops[lvl]
        ops.restrict(r, getattr(rop, "method", "full_weighting"), out=c.f)
        c.u.zero_()
        if self.cycle_type == "V":
            reps = 1
        elif self.cycle_type == "W":
            reps = 2
        elif self.cycle_type == "F":
            reps = max(1, 2 ** (L - lvl - 2))
        else:
            reps = 0  # reference: unknown cycle strings fall through all branches (multigrid.py:309-319)
        for _ in range(reps):
            self.cycle(level_dtypes, lvl + 1, precision_manager)
        pop = self.prolongation_ops[lvl]
        ops.prolong(c.u, getattr(pop, "method", "bilinear"), out=b.u, add=True)
        if self.post > 0:
            self._smooth(lvl, b, self.post)

    def residual_sumsq_async(self, dtype, slot: int = 0) -> torch.Tensor:
        """Launch r = f - A u on level 0 and its sum of squares; returns a device scalar view."""
        b = self.levels[0].bufs(dtype)
        r = self._residual(0, b.u, b.f, b.tmp)
        return ops.sumsq_async(r, slot)
