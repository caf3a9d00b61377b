"""Drop-in ``MultigridSolver`` (reference solvers/multigrid.py:28-390) driving libmgb200.

Same constructor, ``setup`` and ``solve`` signatures, same hierarchy rule (stop before a level
with fewer than 5 points per side), same defaults (lexicographic GS smoother and coarse solver
when none is given), same convergence test (h-scaled L2 norm of f - A u over all points
< tolerance) and the same ``info`` keys.  The whole hierarchy lives in HBM; per iteration the
host reads back one double (the residual norm)."""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import ops
from ..core.precision import PrecisionLevel
from ..device import like_input, to_device
from .base import BaseSolver
from .engine import CycleEngine
from .smoothers import GaussSeidelSmoother


class MultigridCycle:
    V_CYCLE = "V"
    W_CYCLE = "W"
    F_CYCLE = "F"


_PREC_DTYPE = {PrecisionLevel.SINGLE: torch.float32, PrecisionLevel.DOUBLE: torch.float64}


class MultigridSolver(BaseSolver):
    def __init__(self, max_levels: int = 4, max_iterations: int = 50, tolerance: float = 1e-8,
                 cycle_type: str = MultigridCycle.V_CYCLE, pre_smooth_iterations: int = 2,
                 post_smooth_iterations: int = 2, coarse_tolerance: float = 1e-12,
                 coarse_max_iterations: int = 1000, verbose: bool = False, kernels: str = "auto", loader: str = "tma",
                 device=None, use_cuda_graphs: bool = True):
        super().__init__(max_iterations, tolerance, verbose, "Multigrid")
        self.max_levels, self.cycle_type = max_levels, cycle_type
        self.pre_smooth_iterations, self.post_smooth_iterations = pre_smooth_iterations, post_smooth_iterations
        self.coarse_tolerance, self.coarse_max_iterations = coarse_tolerance, coarse_max_iterations
        self.kernels, self.loader, self.device = kernels, loader, device
        # A cycle is ~3 launches per level, most of them microseconds long: without a precision manager (whose decisions
        # happen on the host between levels) each (dtype, buffer-role state) is captured once and replayed (graphs.py)
        self.use_cuda_graphs = use_cuda_graphs
        self._graphs = None
        self.grids: List = []
        self.operators: List = []
        self.restriction_ops: List = []
        self.prolongation_ops: List = []
        self.smoother = None
        self.coarse_solver = None
        self.level_stats: Dict[int, Dict[str, float]] = {}
        self.engine: Optional[CycleEngine] = None

    # -- setup (multigrid.py:91-182) -----------------------------------------------------------------
    def setup(self, fine_grid, operator, restriction_op, prolongation_op, smoother=None, coarse_solver=None) -> None:
        if smoother is None:
            smoother = GaussSeidelSmoother(max_iterations=max(self.pre_smooth_iterations, self.post_smooth_iterations),
                                           tolerance=self.tolerance * 0.1, verbose=False)
        if coarse_solver is None:
            coarse_solver = GaussSeidelSmoother(max_iterations=self.coarse_max_iterations,
                                                tolerance=self.coarse_tolerance, verbose=self.verbose)
        self.smoother, self.coarse_solver = smoother, coarse_solver
        self.grids, self.operators = [fine_grid], [operator]
        self.restriction_ops, self.prolongation_ops = [], []
        g = fine_grid
        for _ in range(1, self.max_levels):
            try:
                c = g.coarsen()
            except ValueError:
                break
            if c.nx < 5 or c.ny < 5:
                break
            self.grids.append(c)
            self.operators.append(operator)
            self.restriction_ops.append(restriction_op)
            self.prolongation_ops.append(prolongation_op)
            g = c
        self.level_stats = {l: {"smooth_time": 0.0, "restrict_time": 0.0, "prolong_time": 0.0}
                            for l in range(len(self.grids))}
        self.engine = CycleEngine(self.grids, smoother=smoother, coarse_solver=coarse_solver,
                                  operators=self.operators, restriction_ops=self.restriction_ops,
                                  prolongation_ops=self.prolongation_ops, cycle_type=self.cycle_type,
                                  pre=self.pre_smooth_iterations, post=self.post_smooth_iterations,
                                  kernels=self.kernels, loader=self.loader, device=self.device)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=self.engine.dev)
        from .graphs import GraphCache
        self._graphs = GraphCache(self.engine.buffer_state, self.engine.snapshot_roles, CycleEngine.restore_roles,
                                  self.use_cuda_graphs)

    # -- solve (multigrid.py:184-251) ------------------------------------------------------------------
    def _level_dtypes(self, base_dtype, precision_manager) -> List[torch.dtype]:
        L = len(self.grids)
        if precision_manager is None:
            return [base_dtype] * L
        out = []
        for l in range(L - 1):
            p = precision_manager.get_precision_for_level(l, L)
            out.append(_PREC_DTYPE.get(p, torch.float64))  # MIXED resolves per level; bare MIXED -> float64
        # the coarsest level is never converted (multigrid.py:270-272): it runs in whatever dtype the
        # restricted residual arrives in, which is the grid dtype (transfer.py:102)
        out.append(ops.torch_dtype(self.grids[-1].dtype))
        return out

    def solve(self, grid, operator, rhs, initial_guess=None, precision_manager=None) -> Tuple[Any, Dict[str, Any]]:
        if not self.grids or tuple(grid.shape) != tuple(self.grids[0].shape) or self.engine is None:
            raise ValueError("Multigrid not properly setup or grid mismatch")
        if tuple(rhs.shape) != tuple(grid.shape):
            raise ValueError("Multigrid not properly setup or grid mismatch")
        self.reset()
        eng = self.engine
        f_in, was_np = to_device(rhs, device=eng.dev)
        base = f_in.dtype
        u_in = None
        if initial_guess is not None:
            u_in, _ = to_device(initial_guess, device=eng.dev)
        cur_dtype = None
        residual_norm = float("inf")
        iteration = 0
        hxhy = grid.hx * grid.hy
        for iteration in range(1, self.max_iterations + 1):
            t0 = time.time()
            if precision_manager is not None:
                shapes = [(g.nx, g.ny) for g in self.grids]
                if cur_dtype is None:  # norm of the initial iterate, in the input precision
                    b0 = eng.levels[0].bufs(base)
                    b0.f.copy_(f_in)
                    b0.u.zero_() if u_in is None else b0.u.copy_(u_in)
                    cur_dtype = base
                cur = float(np.sqrt(hxhy * eng.residual_sumsq_async(cur_dtype).item()))
                precision_manager.update_precision(cur, shapes)
            dts = self._level_dtypes(base, precision_manager)
            if cur_dtype != dts[0]:  # (re)cast the level-0 iterate and rhs (multigrid.py:281-285)
                b_new = eng.levels[0].bufs(dts[0])
                if cur_dtype is None:
                    b_new.f.copy_(f_in)
                    b_new.u.zero_() if u_in is None else b_new.u.copy_(u_in)
                else:
                    b_old = eng.levels[0].bufs(cur_dtype)
                    ops.cast(b_old.u, dts[0], out=b_new.u)
                    b_new.f.copy_(f_in)
                cur_dtype = dts[0]
            if precision_manager is None and self.use_cuda_graphs and self._graphable(dts):
                # fused levels + native coarse solve: no host synchronisation inside the cycle -> CUDA-graph replay
                self._graphs.enabled = True
                self._graphs.run(f"cycle:{dts[0]}", lambda: eng.cycle(dts, 0, None, sumsq_out=self._sumsq))
                ss = self._sumsq
            else:
                fused_norm = eng.cycle(dts, 0, precision_manager, sumsq_out=self._sumsq)
                ss = self._sumsq if fused_norm else eng.residual_sumsq_async(cur_dtype)
            residual_norm = float(np.sqrt(hxhy * ss.item()))
            prec = precision_manager.current_precision.value if precision_manager else "double"
            self.history.record_iteration(residual_norm, time.time() - t0, prec, 0)
            self.log_iteration(iteration, residual_norm)
            if self.check_convergence(residual_norm, iteration):
                self.converged = True
                break
        self.iterations_performed = iteration
        self.final_residual = residual_norm
        u = eng.levels[0].bufs(cur_dtype).u
        out = like_input(u, was_np) if was_np else u.clone()
        return out, self.get_convergence_info()

    def _graphable(self, dts) -> bool:
        """A cycle can be captured when every level runs fused passes (or the small-cycle kernel) and the coarsest
        solve is the one-launch native one: nothing in it synchronises with the host."""
        eng = self.engine
        key = tuple(str(d) for d in dts)
        hit = getattr(self, "_graphable_cache", {}).get(key)
        if hit is not None:
            return hit
        L = eng.num_levels
        # level 0 must be a fused pass: it is the one that leaves the residual norm in `_sumsq`
        ok = getattr(eng.coarse_solver, "kind", None) in ("lexgs", "rbgs_var") and L >= 2 and not eng._small_ok(0, dts)
        for lvl in range(L - 1):
            if not ok or eng._small_ok(lvl, dts):
                break
            try:
                if not eng._fusable(lvl, dts):
                    ok = False
                    break
            except ValueError:
                ok = False
                break
        cache = dict(getattr(self, "_graphable_cache", {}))
        cache[key] = ok
        self._graphable_cache = cache
        return ok

    def apply_cycles(self, rhs, num_cycles: int = 1, initial_guess=None):
        """`num_cycles` cycles on A u = rhs from `initial_guess` (zero by default) with NO residual norms and no host
        synchronisation: the preconditioner fast path (reference multigrid_preconditioner.py:119-163 uses
        solve(tolerance=1e-16, max_iterations=num_cycles))."""
        if not self.grids or self.engine is None or tuple(rhs.shape) != tuple(self.grids[0].shape):
            raise ValueError("Multigrid not properly setup or grid mismatch")
        eng = self.engine
        f_in, was_np = to_device(rhs, device=eng.dev)
        dts = [f_in.dtype] * len(self.grids)
        b = eng.levels[0].bufs(f_in.dtype)
        b.f.copy_(f_in)
        if initial_guess is None:
            b.u.zero_()
        else:
            b.u.copy_(to_device(initial_guess, device=eng.dev, dtype=f_in.dtype)[0])
        for _ in range(num_cycles):
            eng.cycle(dts, 0, None)
        u = eng.levels[0].bufs(f_in.dtype).u
        return like_input(u, True) if was_np else u.clone()

    def get_convergence_info(self) -> Dict[str, Any]:
        info = super().get_convergence_info()
        info.update({
            "cycle_type": self.cycle_type,
            "num_levels": len(self.grids),
            "grid_hierarchy": [(g.nx, g.ny) for g in self.grids],
            "level_timings": {k: dict(v) for k, v in self.level_stats.items()},
            "pre_smooth_iterations": self.pre_smooth_iterations,
            "post_smooth_iterations": self.post_smooth_iterations,
        })
        return info
