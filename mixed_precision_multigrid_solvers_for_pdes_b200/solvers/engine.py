"""Device-resident multigrid cycle engine.

Owns the grid hierarchy in HBM (pitched torch buffers per level and dtype) and issues the kernel
sequence of one V/W/F cycle exactly in the order of the reference recursion
(solvers/multigrid.py:253-337):

    pre-smooth -> residual -> restrict -> zero coarse guess -> 1/2/2^(L-l-2) recursive calls
    -> prolong + add -> post-smooth;      coarsest level: lexicographic-GS solve to tolerance.

Two kernel sets implement the same sequence:
  * ``basic``  -- one launch per reference method (strict arithmetic; also the on-GPU cross-check);
  * ``fused``  -- temporally blocked red-black GS / damped Jacobi with fused residual+restriction and
                 fused prolongation+correction+post-smooth (mg_stream.cuh), used when available.
No host synchronisation happens inside a cycle, so a cycle can be captured in a CUDA graph."""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from .. import ops
from ..device import empty_field, require_cuda, torch_dtype


_NATIVE_OPERATORS = ("LaplacianOperator", "HelmholtzOperator")  # operators the kernels implement directly
_VARCOEF_OPERATOR = "VariableCoefficientOperator"               # -div(a grad u) + shift*u: the mg_vcv_* passes


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
                 kernels: str = "auto", loader: str = "tma", device=None, use_small_cycle: bool = True):
        self.dev = require_cuda(device)
        self.levels: List[_Level] = [_Level(g, self.dev) for g in grids]
        self.smoother, self.coarse_solver = smoother, coarse_solver
        self.operators, self.restriction_ops, self.prolongation_ops = list(operators), list(restriction_ops), list(prolongation_ops)
        self.cycle_type, self.pre, self.post = cycle_type, pre, post
        if kernels not in ("auto", "basic", "fused"):
            raise ValueError(f"kernels must be 'auto', 'basic' or 'fused', got {kernels!r}")
        self.kernels = kernels
        self.loader = loader
        self.use_small_cycle = use_small_cycle
        self._small_cache: Dict = {}
        self.coarse_info = torch.zeros(2, dtype=torch.float64, device=self.dev)
        # reduction scratch of the fused norm passes, owned by this engine and alive as long as the CUDA graphs
        # captured from it (a shared per-device scratch could be outgrown and freed under a captured graph)
        g0 = self.levels[0].grid
        self.workspace = torch.zeros(max(8, ops.vc_workspace_doubles(g0.nx, g0.ny)), dtype=torch.float64, device=self.dev)

    # -- buffer roles (u / tmp swap on every out-of-place pass) ---------------------------------------
    def buffer_state(self):
        """Which physical buffer currently plays `u` on every (level, dtype): the key of a captured graph."""
        return tuple((li, str(dt), b.u.data_ptr()) for li, lv in enumerate(self.levels)
                     for dt, b in sorted(lv._bufs.items(), key=lambda kv: str(kv[0])))

    def snapshot_roles(self):
        return [(b, b.u, b.tmp) for lv in self.levels for b in lv._bufs.values()]

    @staticmethod
    def restore_roles(snap) -> None:
        for b, u, tmp in snap:
            b.u, b.tmp = u, tmp

    # -- per-level building blocks ----------------------------------------------------------------
    @property
    def num_levels(self) -> int:
        return len(self.levels)

    def _smooth(self, lvl: int, b: _Buffers, sweeps: int) -> None:
        g = self.levels[lvl].grid
        sm = self.smoother
        kind = getattr(sm, "kind", "custom")
        shift = getattr(self.operators[lvl], "shift", 0.0)
        if shift and kind not in ("rbgs", "rbgs_var"):
            raise ValueError("a shifted (Helmholtz) operator is smoothed with red-black Gauss-Seidel only")
        if kind == "rbgs_var":  # strict per-colour kernels of the variable-coefficient smoother, in place
            sm._smooth_device_(g, b.u, b.f, sweeps)
        elif kind == "jacobi":
            ops.smooth_jacobi_(b.u, b.f, g.hx, g.hy, sm.omega, sweeps, tmp=b.tmp)
        elif kind == "rbgs":
            ops.smooth_rbgs_(b.u, b.f, g.hx, g.hy, sm.omega, sweeps, shift=shift)
        elif kind in ("lexgs", "sgs"):
            sm._smooth_device_(g, b.u, b.f, sweeps)
        else:  # foreign smoother object: use its public protocol on device tensors
            b.u.copy_(sm.smooth(g, self.operators[lvl], b.u, b.f, sweeps))

    def _residual(self, lvl: int, u, f, out):
        g = self.levels[lvl].grid
        op = self.operators[lvl]
        coeff = getattr(op, "coefficient", None)
        if coeff is not None and type(op).__name__ in _NATIVE_OPERATORS:
            return ops.residual(u, f, g.hx, g.hy, coeff, out=out, shift=getattr(op, "shift", 0.0))
        out.copy_(op.residual(g, u, f))
        return out

    def _coarse_solve(self, lvl: int, b: _Buffers, precision_manager=None) -> None:
        g = self.levels[lvl].grid
        cs = self.coarse_solver
        op = self.operators[lvl]
        coeff = getattr(op, "coefficient", None)
        if getattr(cs, "kind", None) == "lexgs" and coeff is not None and type(op).__name__ in _NATIVE_OPERATORS:
            ops.coarse_solve_lexgs_(b.u, b.f, g.hx, g.hy, cs.omega, coeff, cs.tolerance, cs.max_iterations,
                                    info=self.coarse_info, shift=getattr(op, "shift", 0.0))
        elif getattr(cs, "kind", None) == "rbgs_var" and type(op).__name__ == _VARCOEF_OPERATOR:
            ops.varcoef_coarse_solve_(b.u, b.f, op.coefficients(g.nx, g.ny, b.u.dtype), g.hx, g.hy, op.shift, cs.omega,
                                      cs.tolerance, cs.max_iterations, info=self.coarse_info)
        else:  # any other IterativeSolver: its own solve loop (host-checked convergence)
            sol, _ = cs.solve(g, op, b.f, b.u, precision_manager)
            b.u.copy_(sol)

    # -- fused path ---------------------------------------------------------------------------------
    def _fusable(self, lvl: int, level_dtypes: Sequence) -> bool:
        """The fused passes implement: red-black GS, LaplacianOperator, full weighting, bilinear,
        equal dtypes on both levels, 16-byte aligned pitched buffers (ours always are)."""
        if self.kernels == "basic":
            return False
        op = self.operators[lvl]
        kind = getattr(self.smoother, "kind", None)
        # Jacobi: TMA-staged instantiations only, and (like the strict kernel) no Helmholtz shift
        jac = kind == "jacobi" and self.loader == "tma" and not getattr(op, "shift", 0.0)
        var = kind == "rbgs_var" and type(op).__name__ == _VARCOEF_OPERATOR and self.loader == "tma"
        ok = (((kind == "rbgs" or jac) and type(op).__name__ in _NATIVE_OPERATORS or var)
              and getattr(self.restriction_ops[lvl], "method", None) == "full_weighting"
              and getattr(self.prolongation_ops[lvl], "method", None) == "bilinear"
              and torch_dtype(level_dtypes[lvl]) == torch_dtype(level_dtypes[lvl + 1]))
        if not ok and self.kernels == "fused":
            raise ValueError("kernels='fused' needs GaussSeidelSmoother(red_black=True) or a Jacobi smoother, LaplacianOperator, "
                             "full_weighting restriction, bilinear prolongation and one dtype per level pair")
        return ok

    # -- the coarse end of the hierarchy in one launch -------------------------------------------------------------
    def _small_ok(self, lvl: int, level_dtypes: Sequence) -> bool:
        """Can levels lvl..coarsest run as ONE mg_small_cycle launch (everything in shared memory)?"""
        if self.kernels == "basic" or not self.use_small_cycle or self.cycle_type not in ("V", "W", "F"):
            return False
        L = self.num_levels
        key = (lvl, tuple(str(torch_dtype(d)) for d in level_dtypes[lvl:]))
        hit = self._small_cache.get(key)
        if hit is not None:
            return hit
        ok = False
        cs, sm = self.coarse_solver, self.smoother
        dts = [torch_dtype(d) for d in level_dtypes[lvl:]]
        if (getattr(sm, "kind", None) == "rbgs" and getattr(cs, "kind", None) == "lexgs" and getattr(cs, "omega", 1.0) == 1.0
                and all(type(op).__name__ in _NATIVE_OPERATORS for op in self.operators[lvl:])
                and all(getattr(r, "method", None) == "full_weighting" for r in self.restriction_ops[lvl:])
                and all(getattr(p, "method", None) == "bilinear" for p in self.prolongation_ops[lvl:])
                and len(set(dts[:-1])) <= 1 and (len(dts) == 1 or dts[-1] in (dts[0], torch.float64))
                and len({(getattr(op, "coefficient", None), getattr(op, "shift", 0.0)) for op in self.operators[lvl:]}) == 1):
            g = self.levels[lvl].grid
            ok = ops.small_cycle_fits(g.nx, g.ny, L - lvl, dts[0], dts[-1])
        self._small_cache[key] = ok
        return ok

    def _small_cycle(self, lvl: int, level_dtypes: Sequence, u_zero: bool) -> None:
        g = self.levels[lvl].grid
        b = self.levels[lvl].bufs(level_dtypes[lvl])
        op, cs = self.operators[lvl], self.coarse_solver
        ops.small_cycle_(b.u, b.f, g.hx, g.hy, nlev=self.num_levels - lvl, cycle_type=self.cycle_type, pre=self.pre,
                         post=self.post, omega=self.smoother.omega, coefficient=op.coefficient,
                         shift=getattr(op, "shift", 0.0), coarse_tolerance=cs.tolerance,
                         coarse_max_iterations=cs.max_iterations, coarse_dtype=torch_dtype(level_dtypes[-1]),
                         u_zero=u_zero, info=self.coarse_info)

    def _reps(self, lvl: int) -> int:
        L = self.num_levels
        if self.cycle_type == "V":
            return 1
        if self.cycle_type == "W":
            return 2
        if self.cycle_type == "F":
            return max(1, 2 ** (L - lvl - 2))
        return 0  # reference: unknown cycle strings fall through all branches (multigrid.py:309-319)

    def _cycle_fused(self, level_dtypes, lvl, precision_manager, sumsq_out, u_zero, skip_down: bool = False) -> None:
        g = self.levels[lvl].grid
        b = self.levels[lvl].bufs(level_dtypes[lvl])
        c = self.levels[lvl + 1].bufs(level_dtypes[lvl + 1])
        op = self.operators[lvl]
        omega, ld = self.smoother.omega, self.loader
        sh = getattr(op, "shift", 0.0)
        kw = {"omega": omega, "loader": ld, "shift": sh}
        ms = 2  # sweeps per HBM pass
        if type(op).__name__ == _VARCOEF_OPERATOR:  # variable coefficients: the level's nodal field rides along
            kw["a"] = op.coefficients(g.nx, g.ny, b.u.dtype)
            if b.u.dtype == torch.float64:
                ms = 1
        else:
            kw["coefficient"] = op.coefficient
            kw["smoother"] = self.smoother.kind  # "rbgs" or "jacobi": same passes, the sweeps differ
        # down: pre-smooth (`ms` sweeps per HBM pass) with residual + restriction fused into the last pass
        # (skip_down: the caller's fused defect + down pass has already left the pre-smoothed iterate in b.u and the
        # restricted residual in c.f, ops.vc_defect_down_pass)
        n = 0 if skip_down else self.pre
        if skip_down:
            pass
        elif u_zero and n == 0:
            ops.zero_(b.u)  # nothing will overwrite the iterate before it is read
            u_zero = False
        while n > ms:
            ops.vc_pass(b.u, b.tmp, b.f, g.hx, g.hy, sweeps=ms, u_zero=u_zero, **kw)
            b.u, b.tmp = b.tmp, b.u
            n -= ms
            u_zero = False
        if n > 0:
            ops.vc_pass(b.u, b.tmp, b.f, g.hx, g.hy, sweeps=n, coarse_out=c.f, u_zero=u_zero, **kw)
            b.u, b.tmp = b.tmp, b.u
        elif not skip_down:
            ops.vc_pass(b.u, None, b.f, g.hx, g.hy, sweeps=0, coarse_out=c.f, **kw)
        # the coarse error equation starts from e = 0: the first coarse pass is told so instead of reading zeros
        for rep in range(self._reps(lvl)):
            self.cycle(level_dtypes, lvl + 1, precision_manager, u_zero=(rep == 0))
        # up: prolongation + correction fused into the first post-smoothing pass, norm into the last
        n = self.post
        first = min(n, ms)
        last = (n - first) == 0
        ops.vc_pass(b.u, b.tmp, b.f, g.hx, g.hy, sweeps=first, coarse_in=c.u, sumsq_out=sumsq_out if last else None,
                    workspace=self.workspace, **kw)
        b.u, b.tmp = b.tmp, b.u
        n -= first
        while n > 0:
            k = min(n, ms)
            n -= k
            ops.vc_pass(b.u, b.tmp, b.f, g.hx, g.hy, sweeps=k, sumsq_out=sumsq_out if n == 0 else None,
                        workspace=self.workspace, **kw)
            b.u, b.tmp = b.tmp, b.u

    # -- the recursion ------------------------------------------------------------------------------
    def cycle(self, level_dtypes: Sequence, lvl: int = 0, precision_manager=None, sumsq_out=None,
              u_zero: bool = False, skip_down: bool = False) -> bool:
        """One cycle on level `lvl`, updating that level's ``u`` for ``level_dtypes[lvl]``.
        With ``sumsq_out`` (1 float64 on the device) the fused path also leaves sum((f - A u)^2) of the
        final iterate there; returns True when it did.  ``u_zero``: the iterate of this level is to be
        taken as zero whatever its buffer holds (the zero initial guess of a coarse error equation)."""
        L = self.num_levels
        b = self.levels[lvl].bufs(level_dtypes[lvl])
        if skip_down:  # only the fused path can pick a cycle up after its down pass
            if not (lvl < L - 1 and self._fusable(lvl, level_dtypes)) or self._small_ok(lvl, level_dtypes):
                raise ValueError("skip_down needs a fused level")
            self._cycle_fused(level_dtypes, lvl, precision_manager, sumsq_out, u_zero, skip_down=True)
            return sumsq_out is not None
        if precision_manager is None and self._small_ok(lvl, level_dtypes):
            self._small_cycle(lvl, level_dtypes, u_zero)
            return False
        if lvl == L - 1:
            if u_zero:
                ops.zero_(b.u)
            self._coarse_solve(lvl, b, precision_manager)
            return False
        if self._fusable(lvl, level_dtypes):
            self._cycle_fused(level_dtypes, lvl, precision_manager, sumsq_out, u_zero)
            return sumsq_out is not None
        if u_zero:
            ops.zero_(b.u)
        if self.pre > 0:
            self._smooth(lvl, b, self.pre)
        c = self.levels[lvl + 1].bufs(level_dtypes[lvl + 1])
        r = self._residual(lvl, b.u, b.f, b.tmp)
        rop = self.restriction_ops[lvl]
        ops.restrict(r, getattr(rop, "method", "full_weighting"), out=c.f)
        for rep in range(self._reps(lvl)):
            self.cycle(level_dtypes, lvl + 1, precision_manager, u_zero=(rep == 0))
        pop = self.prolongation_ops[lvl]
        ops.prolong(c.u, getattr(pop, "method", "bilinear"), out=b.u, add=True)
        if self.post > 0:
            self._smooth(lvl, b, self.post)
        return False

    def residual_sumsq_async(self, dtype, slot: int = 0) -> torch.Tensor:
        """Launch r = f - A u on level 0 and its sum of squares; returns a device scalar view."""
        b = self.levels[0].bufs(dtype)
        r = self._residual(0, b.u, b.f, b.tmp)
        return ops.sumsq_async(r, slot)
