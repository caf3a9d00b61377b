"""Row-slab domain decomposition of the multigrid cycle over the GPUs of one box.

One process per GPU (torchrun).  The fine levels are partitioned into contiguous slabs of rows
(first index; rows are contiguous in memory, so halo rows are contiguous messages); every slab carries
GHOST = 8 ghost rows per interior side on EVERY distributed level.  The same fused kernels as on one GPU
run on the slab (`mg_vc_pass_slab`): a pass with 2 RB-GS sweeps and a residual/restriction stage has a
dependency cone of 6 rows, so after a pass the owned rows are exact while the outer ghost rows are stale;
they are refreshed by ONE grouped NCCL send/recv pair per neighbour per pass output (temporal blocking
means one exchange per 2 sweeps instead of one per colour half-sweep).  Ownership is aligned to powers
of two so every coarse row has a unique owner and the transfers need no extra communication.

Below `agglomerate_below` fine points per side the remaining levels are AGGLOMERATED: the owned rows of
the coarse right-hand side are all-gathered and every rank runs the remaining sub-cycle redundantly on the
full (small) grid — identical results on all ranks, no broadcast back, no idle GPUs — then copies its slab
(including ghosts) out of the full correction.  The residual norm is one NCCL all-reduce of one double per
cycle.  Reference counterpart: gpu/multi_gpu_solver.py:90-185, 244-383 (strip decomposition, peer copies,
host-side sum; no coarse levels) — a design sketch, not an algorithm.

Results on owned rows are bit-identical to the single-GPU run (RB-GS is order independent within a
colour); only the norm reduction order differs (1e-15 relative)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

GHOST = 8  # ghost rows per interior side on every distributed level (even; >= 6 = cone of a fused pass)
# bench.py's defaults (--ghost, --agg), measured on 8 B200 (profiles/r02_bench_8gpu*.json):
# ghost 8 / 16 / 32 -> 5.9 / 3.9 / 1.9 exchanges and 3.90 / 3.85 / 3.80 ms per cycle; agglomerating at 513 instead of 1025
# points per side (4097 x 513 instead of 8193 x 1025 solved redundantly) -> 3.71 ms
BENCH_GHOST = 32
BENCH_AGG = 513


# ======================================================================================================
# Partition (pure host logic)
# ======================================================================================================
@dataclass
class SlabLevel:
    level: int
    nx_glob: int
    ny: int
    hx: float
    hy: float
    own_lo: int      # global rows owned: [own_lo, own_hi)
    own_hi: int
    g_lo: int        # ghost rows present below / above (0 on a physical boundary)
    g_hi: int

    @property
    def row0(self) -> int:           # global row of local row 0
        return self.own_lo - self.g_lo

    @property
    def loc_nx(self) -> int:
        return (self.own_hi - self.own_lo) + self.g_lo + self.g_hi

    @property
    def own_local(self) -> Tuple[int, int]:
        return (self.g_lo, self.g_lo + (self.own_hi - self.own_lo))


class SlabPartition:
    """Ownership of global rows per level for `world` ranks; levels [0, dist_levels) are distributed."""

    def __init__(self, nx: int, ny: int, world: int, rank: int, num_levels: int, dist_levels: int,
                 domain=(0.0, 1.0, 0.0, 1.0), ghost: int = GHOST):
        if world < 1 or not (0 <= rank < world):
            raise ValueError("bad world/rank")
        if dist_levels < 1 or dist_levels > num_levels:
            raise ValueError("dist_levels must be in [1, num_levels]")
        if (nx - 1) % (world * 2 ** (dist_levels - 1)) != 0:
            raise ValueError(f"(nx-1)={nx - 1} must be a multiple of world*2^(dist_levels-1) = {world * 2 ** (dist_levels - 1)}")
        self.world, self.rank, self.ghost = world, rank, ghost
        self.num_levels, self.dist_levels = num_levels, dist_levels
        self.levels: List[SlabLevel] = []
        n, m = nx, ny
        hx = (domain[1] - domain[0]) / (nx - 1)
        hy = (domain[3] - domain[2]) / (ny - 1)
        for l in range(dist_levels):
            per = (n - 1) // world
            if world > 1 and per < ghost:
                raise ValueError(f"level {l}: {per} rows per rank < ghost depth {ghost}; lower dist_levels")
            lo, hi = rank * per, (rank + 1) * per + (1 if rank == world - 1 else 0)
            self.levels.append(SlabLevel(l, n, m, hx, hy, lo, hi, 0 if rank == 0 else ghost,
                                         0 if rank == world - 1 else ghost))
            n, m, hx, hy = (n - 1) // 2 + 1, (m - 1) // 2 + 1, hx * 2, hy * 2
        self.agg_shape = (n, m) if dist_levels < num_levels else None  # global grid of the first agglomerated level
        self.agg_h = (hx, hy)
        # rows of the agglomerated level this rank copies back into its level-(D-1)-coarse buffer
        if self.agg_shape is not None:
            per = (n - 1) // world
            lo, hi = rank * per, (rank + 1) * per + (1 if rank == world - 1 else 0)
            self.agg_slab = SlabLevel(dist_levels, n, m, hx, hy, lo, hi, 0 if rank == 0 else min(ghost, lo),
                                      0 if rank == world - 1 else min(ghost, n - hi))
        else:
            self.agg_slab = None

    def slab(self, l: int) -> SlabLevel:
        return self.levels[l] if l < self.dist_levels else self.agg_slab

    def coarse_view(self, l: int) -> Tuple[int, int]:
        """(offset, rows) of the window of level l+1's LOCAL array that lines up with level l's local rows
        (coarse local row ic <-> fine local row 2ic)."""
        f, c = self.slab(l), self.slab(l + 1)
        off = f.row0 // 2 - c.row0
        rows = (f.loc_nx - 1) // 2 + 1
        assert f.row0 % 2 == 0 and off >= 0 and off + rows <= c.loc_nx, (f, c, off, rows)
        return off, rows


def choose_dist_levels(nx: int, ny: int, world: int, num_levels: int, agglomerate_below: int = 1025,
                       ghost: int = GHOST) -> int:
    """Distribute level l while its grid has more than `agglomerate_below` points per side and every rank
    owns at least 4*ghost rows with power-of-two aligned slab boundaries."""
    d = 0
    n, m = nx, ny
    for l in range(num_levels):
        per = (n - 1) // world
        ok = (n - 1) % world == 0 and per >= 4 * ghost and (nx - 1) % (world * 2 ** l) == 0
        if l > 0 and min(n, m) <= agglomerate_below:
            ok = False
        if not ok:
            break
        d = l + 1
        n, m = (n - 1) // 2 + 1, (m - 1) // 2 + 1
    return max(1, min(d, num_levels))


# ======================================================================================================
# Local compute back ends
# ======================================================================================================
class DeviceBackend:
    """Production back end: libmgb200 kernels on pitched CUDA tensors."""

    def __init__(self, device, loader: str = "tma"):
        from . import ops
        from .device import empty_field
        self.ops, self._empty, self.device, self.loader = ops, empty_field, device, loader

    def empty(self, nx, ny, dtype):
        return self._empty(nx, ny, dtype, self.device)

    def scalar(self, n=1):
        return torch.zeros(n, dtype=torch.float64, device=self.device)

    def _workspace(self, f):
        """Reduction scratch of the norm passes, owned by this back end (= by one solver); grow-only, outgrown
        buffers stay alive because captured CUDA graphs reference them."""
        need = self.ops.vc_workspace_doubles(f.shape[0], f.shape[1])
        w = getattr(self, "_ws", None)
        if w is None or w.numel() < need:
            if w is not None:
                self._ws_retired = getattr(self, "_ws_retired", []) + [w]
            w = self._ws = torch.zeros(need, dtype=torch.float64, device=self.device)
        return w

    def vc_pass(self, u_in, u_out, f, *a, **k):
        if k.get("sumsq_out") is not None:
            k["workspace"] = self._workspace(f)
        return self.ops.vc_pass(u_in, u_out, f, *a, loader=self.loader, **k)

    def vc_defect_pass(self, u_in, u_out, f, *a, **k):
        if k.get("r_out") is not None:
            k["workspace"] = self._workspace(f)
        return self.ops.vc_defect_pass(u_in, u_out, f, *a, loader=self.loader, **k)

    def vc_defect_down_pass(self, u_in, u_out, f, *a, **k):
        k["workspace"] = self._workspace(f)
        return self.ops.vc_defect_down_pass(u_in, u_out, f, *a, **k)

    def zero_ring(self, t, first_row: bool, last_row: bool):
        self.ops.zero_ring_(t, first_row, last_row)

    def heat_rhs(self, u, rhs, hx, hy, **k) -> torch.Tensor:
        """One-kernel right-hand side of a theta-method step (mg_heat_rhs); returns sum(rhs^2) over `norm_rows`."""
        return self.ops.heat_rhs_(u, rhs, hx, hy, **k).clone()

    def sumsq(self, t) -> torch.Tensor:
        """Sum of squares of a (row window of a) pitched field as a 1-element device tensor (no sync)."""
        return self.ops.sumsq_async(t, slot=2).clone()

    def apply_laplacian(self, u, hx, hy):
        """lap_h(u) on the interior of the local array, 0 on its first/last rows and columns."""
        return self.ops.apply_laplacian(u, hx, hy, 1.0)

    def make_coarse_engine(self, nx, ny, domain, levels, cycle_type, pre, post, coarse_tol, coarse_max, shift=0.0,
                           coefficient=None):
        from .core.grid import Grid
        from .operators.laplacian import HelmholtzOperator, LaplacianOperator
        from .operators.transfer import ProlongationOperator, RestrictionOperator
        from .solvers.engine import CycleEngine
        from .solvers.smoothers import GaussSeidelSmoother
        grids = [Grid(nx, ny, domain)]
        for _ in range(1, levels):
            grids.append(grids[-1].coarsen())
        L = len(grids)
        if coefficient is not None:  # the full coarse coefficient field, identical on every rank
            from .operators.variable import VariableCoefficientOperator, VariableCoefficientSmoother
            with torch.cuda.device(self.device):
                op = VariableCoefficientOperator(np.ascontiguousarray(coefficient, dtype=np.float64), shift)
            smoother = VariableCoefficientSmoother(op)
            coarse = VariableCoefficientSmoother(op, max_iterations=coarse_max, tolerance=coarse_tol)
        else:
            op = HelmholtzOperator(-1.0, shift) if shift else LaplacianOperator(-1.0)
            smoother = GaussSeidelSmoother(red_black=True)
            coarse = GaussSeidelSmoother(max_iterations=coarse_max, tolerance=coarse_tol)
        eng = CycleEngine(grids, smoother=smoother, coarse_solver=coarse,
                          operators=[op] * L, restriction_ops=[RestrictionOperator()] * max(0, L - 1),
                          prolongation_ops=[ProlongationOperator()] * max(0, L - 1), cycle_type=cycle_type, pre=pre,
                          post=post, kernels="auto", loader=self.loader, device=self.device)
        return _DeviceCoarse(eng)


class _DeviceCoarse:
    def __init__(self, eng):
        self.eng = eng
        self.L = eng.num_levels

    def dtypes(self, dtype):
        # the coarsest level stays fp64 (reference: never converted, multigrid.py:270-272)
        return [dtype] * (self.L - 1) + [torch.float64] if self.L > 1 else [torch.float64]

    def bufs(self, dtype):
        return self.eng.levels[0].bufs(self.dtypes(dtype)[0])

    def cycle(self, dtype, u_zero: bool):
        self.eng.cycle(self.dtypes(dtype), 0, None, u_zero=u_zero)
        return self.bufs(dtype).u


# ======================================================================================================
# Distributed cycle engine
# ======================================================================================================
class _Bufs:
    __slots__ = ("u", "tmp", "f")


class DistributedCycleEngine:
    def __init__(self, nx: int, ny: int, *, domain=(0.0, 1.0, 0.0, 1.0), num_levels: Optional[int] = None,
                 cycle_type: str = "V", pre: int = 2, post: int = 2, agglomerate_below: int = 1025,
                 dist_levels: Optional[int] = None, coarse_tolerance: float = 1e-12, coarse_max_iterations: int = 1000,
                 shift: float = 0.0, backend=None, group=None, device=None, transport=None, coefficient=None,
                 ghost: Optional[int] = None):
        if not (1 <= pre <= 2 and 1 <= post <= 2):
            raise ValueError("the distributed engine runs 1 or 2 pre/post sweeps per pass")
        if not shift >= 0.0:
            raise ValueError("shift must be >= 0")
        self.shift = float(shift)  # Helmholtz term of the operator -lap_h + shift (0: Poisson); same on every level
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nx, self.ny, self.domain = nx, ny, tuple(domain)
        self.cycle_type, self.pre, self.post = cycle_type, pre, post
        if num_levels is None:
            num_levels, a, b = 1, nx, ny
            while (a - 1) % 2 == 0 and (b - 1) % 2 == 0 and (a - 1) // 2 + 1 >= 5 and (b - 1) // 2 + 1 >= 5:
                a, b, num_levels = (a - 1) // 2 + 1, (b - 1) // 2 + 1, num_levels + 1
        self.num_levels = num_levels
        # ghost rows per interior side (even, >= 6 = dependency cone of a fused 2-sweep pass with restriction).  Deeper
        # ghosts buy fewer exchanges: every pass spends 2 rows of validity per sweep, restriction halves what is left
        ghost = GHOST if ghost is None else int(ghost)
        if ghost < 6 or ghost % 2:
            raise ValueError("ghost must be even and >= 6")
        if dist_levels is None:
            dist_levels = choose_dist_levels(nx, ny, self.world, num_levels, agglomerate_below, ghost)
        if dist_levels >= num_levels:
            dist_levels = num_levels - 1  # the coarsest level is always solved on the gathered grid
        self.part = SlabPartition(nx, ny, self.world, self.rank, num_levels, max(1, dist_levels), domain, ghost)
        self.D = self.part.dist_levels
        self.be = backend if backend is not None else DeviceBackend(device)
        # optional one-sided halo transport (halo.py): slab arrays in symmetric memory, ghost rows pushed into the
        # neighbours over peer mappings instead of NCCL send/recv.  None = the NCCL exchange below.
        self.transport = transport
        self._bufs: Dict[Tuple[int, torch.dtype], _Bufs] = {}
        self.valid: Dict[int, int] = {}  # data_ptr -> number of ghost rows per side that currently hold exact values
        # variable coefficients: -div(a grad u) + shift*u with a callable a(X, Y) (or a full (nx, ny) array), SURVEY 8f-1.
        # Every level evaluates / slices its own slab rows (ghost rows included: the field is static, it never needs an
        # exchange); the agglomerated levels get the full coarse field on every rank.
        self.coefficient = coefficient
        self._coef: Dict[Tuple[int, torch.dtype], torch.Tensor] = {}
        an, am = self.part.agg_shape
        ckw = {"shift": self.shift} if self.shift else {}
        if coefficient is not None:
            ckw["coefficient"] = self._coef_rows(self.D, 0, an)
        self.coarse = self.be.make_coarse_engine(an, am, self.domain, num_levels - self.D, cycle_type, pre, post,
                                                 coarse_tolerance, coarse_max_iterations, **ckw)
        self._kw = {"shift": self.shift} if self.shift else {}  # forwarded to every slab pass
        self.exchanges = 0
        self.phase_events: Optional[List[Any]] = None  # a list while an eager run is being phase-timed (_phase)
        self._agg_flat: Dict[Any, torch.Tensor] = {}  # gather buffers of the agglomeration step, per (dtype, size)

    # -- buffer roles (for CUDA-graph replay) ---------------------------------------------------------------
    def _all_bufs(self):
        inner = getattr(self.coarse, "eng", None)
        mine = list(self._bufs.values())
        return mine + ([b for lv in inner.levels for b in lv._bufs.values()] if inner is not None else [])

    def buffer_state(self):
        # roles AND ghost validity: a captured graph bakes in which exchanges were (not) needed
        return tuple((b.u.data_ptr(), self.vdepth(b.u), self.vdepth(b.tmp), self.vdepth(b.f)) for b in self._all_bufs())

    def snapshot_roles(self):
        # buffer roles + ghost-validity bookkeeping: both must be re-applied after a graph replay, which runs no
        # Python and therefore updates neither
        return ([(b, b.u, b.tmp) for b in self._all_bufs()], dict(self.valid), self.exchanges)

    def restore_roles(self, snap) -> None:
        roles, valid, _ = snap
        for b, u, tmp in roles:
            b.u, b.tmp = u, tmp
        self.valid = dict(valid)

    # -- variable coefficients -----------------------------------------------------------------------------
    def _coef_rows(self, l: int, r0: int, r1: int) -> np.ndarray:
        """Nodal coefficient values of global rows [r0, r1) of level l (fp64 NumPy).  A callable is evaluated in blocks
        of 64 rows aligned to the global row index (identical arrays enter it on every decomposition, so the values
        do not depend on the number of ranks: see DistributedHeatSolver.evaluate); an array is sliced with the
        level's stride (injection, like VariableCoefficientOperator.coefficients)."""
        nxl, nyl = (self.nx - 1) // 2 ** l + 1, (self.ny - 1) // 2 ** l + 1
        c = self.coefficient
        if not callable(c):
            full = c.detach().cpu().numpy() if isinstance(c, torch.Tensor) else np.asarray(c, dtype=np.float64)
            st = 2 ** l
            return np.ascontiguousarray(full[::st, ::st][r0:r1], dtype=np.float64)
        hx = (self.domain[1] - self.domain[0]) / (self.nx - 1) * 2 ** l
        hy = (self.domain[3] - self.domain[2]) / (self.ny - 1) * 2 ** l
        y = self.domain[2] + np.arange(nyl) * hy
        out = np.empty((r1 - r0, nyl), dtype=np.float64)
        EB = 64
        for blk in range(r0 // EB, (r1 - 1) // EB + 1):
            a0, a1 = blk * EB, min((blk + 1) * EB, nxl)
            Xb, Yb = np.meshgrid(self.domain[0] + np.arange(a0, a1) * hx, y, indexing="ij")
            vb = np.broadcast_to(np.asarray(c(Xb, Yb), dtype=np.float64), Xb.shape)
            lo_, hi_ = max(a0, r0), min(a1, r1)
            out[lo_ - r0:hi_ - r0] = vb[lo_ - a0:hi_ - a0]
        if not (out > 0).all():
            raise ValueError("the diffusion coefficient must be positive")
        return out

    def coef(self, l: int, dtype) -> Optional[torch.Tensor]:
        """Coefficient slab of distributed level l (None for the constant-coefficient operator)."""
        if self.coefficient is None:
            return None
        t = self._coef.get((l, dtype))
        if t is None:
            s = self.part.slab(l)
            t = self.be.empty(s.loc_nx, s.ny, dtype)
            t.copy_(torch.from_numpy(self._coef_rows(l, s.row0, s.row0 + s.loc_nx)).to(t.device))
            self._coef[(l, dtype)] = t
        return t

    def _var_kw(self, l: int, dtype) -> Dict[str, Any]:
        a = self.coef(l, dtype)
        return dict(self._kw, a=a) if a is not None else dict(self._kw)

    # -- buffers ------------------------------------------------------------------------------------------
    def bufs(self, l: int, dtype) -> _Bufs:
        key = (l, dtype)
        b = self._bufs.get(key)
        if b is None:
            s = self.part.slab(l)
            b = _Bufs()
            if self.transport is not None:  # collective: every rank creates its buffers in the same order
                rows_max = (s.nx_glob - 1) // self.world + 2 * self.part.ghost + 1
                b.u, b.tmp, b.f = (self.transport.alloc(s.loc_nx, rows_max, s.ny, dtype) for _ in range(3))
            else:
                b.u, b.tmp, b.f = (self.be.empty(s.loc_nx, s.ny, dtype) for _ in range(3))
            self._bufs[key] = b
        return b

    # -- communication ------------------------------------------------------------------------------------
    # -- optional phase timing (eager runs only): CUDA events around the communication steps and the agglomerated tail
    def _phase(self, name: str):
        eng = self

        class _Ctx:
            def __enter__(self_inner):
                self_inner.on = eng.phase_events is not None
                if self_inner.on:
                    self_inner.a = torch.cuda.Event(enable_timing=True)
                    self_inner.a.record()

            def __exit__(self_inner, *exc):
                if self_inner.on:
                    b = torch.cuda.Event(enable_timing=True)
                    b.record()
                    eng.phase_events.append((name, self_inner.a, b))
                return False
        return _Ctx()

    def phase_summary(self) -> Dict[str, float]:
        """Total milliseconds per phase name of the events recorded since `phase_events` was set to a list."""
        torch.cuda.synchronize()
        out: Dict[str, float] = {}
        for name, a, b in self.phase_events or []:
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    def exchange(self, t: torch.Tensor, l: int) -> None:
        """Refresh the ghost rows of one level-l local array from the neighbouring slabs."""
        self.exchange_many([(t, l)])

    def exchange_many(self, items) -> None:
        """Refresh the ghost rows of several (array, level) pairs with ONE grouped send/recv launch."""
        for t, l in items:
            self.valid[t.data_ptr()] = self.part.ghost
        if self.world == 1 or not items:
            return
        with self._phase("halo_exchange"):
            self._exchange_now(items)

    def _exchange_now(self, items) -> None:
        if self.transport is not None:
            self.transport.exchange(self.part, items)
            self.exchanges += 1
            return
        G = self.part.ghost
        ops_, fix = [], []
        for t, l in items:
            s = self.part.slab(l)
            lo, hi = s.own_local
            contiguous = _rows_contiguous(t)
            for has, peer, send_rows, recv_rows in ((s.g_lo, self.rank - 1, (lo, lo + G), (0, s.g_lo)),
                                                    (s.g_hi, self.rank + 1, (hi - G, hi), (hi, hi + s.g_hi))):
                if not has:
                    continue
                send = _rows(t, *send_rows) if contiguous else t[send_rows[0]:send_rows[1]].contiguous()
                recv = _rows(t, *recv_rows) if contiguous else torch.empty_like(t[recv_rows[0]:recv_rows[1]])
                ops_ += [dist.P2POp(dist.isend, send, self._peer(peer), self.group),
                         dist.P2POp(dist.irecv, recv, self._peer(peer), self.group)]
                if not contiguous:
                    fix.append((t, recv, recv_rows))
        for w in dist.batch_isend_irecv(ops_):
            w.wait()
        for t, recv, (a, b) in fix:
            t[a:b].copy_(recv)
        self.exchanges += 1

    # -- ghost validity bookkeeping: exchange only when the next pass needs deeper valid ghosts than it has ------
    def vdepth(self, t: Optional[torch.Tensor]) -> int:
        return self.part.ghost if t is None else self.valid.get(t.data_ptr(), 0)

    def set_valid(self, t: torch.Tensor, depth: int) -> None:
        self.valid[t.data_ptr()] = max(0, min(self.part.ghost, depth))

    def ensure(self, need: int, items, effective=None) -> None:
        """`items`: (array, level) inputs of the next pass, `effective(depths) -> fine-row validity`.
        If the inputs do not provide `need` valid ghost rows, every input that is not fully valid is
        refreshed in one batch."""
        depths = [self.vdepth(t) for t, _ in items]
        have = effective(depths) if effective is not None else min(depths)
        if have < need:
            self.exchange_many([(t, l) for (t, l), d in zip(items, depths) if d < self.part.ghost])

    def _peer(self, r: int) -> int:
        return r if self.group is None else dist.get_global_rank(self.group, r)

    def allreduce_sum(self, x: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            with self._phase("norm_allreduce"):
                dist.all_reduce(x, op=dist.ReduceOp.SUM, group=self.group)
        return x

    # -- agglomerated levels --------------------------------------------------------------------------------
    def _agglomerated_cycle(self, dtype, u_zero: bool) -> None:
        """Level D: gather the owned rows of the local right-hand side into the full grid on every rank,
        run the remaining sub-cycle redundantly, copy the local slab (with ghosts) of the correction back."""
        s = self.part.agg_slab
        b = self.bufs(self.D, dtype)
        cb = self.coarse.bufs(dtype)
        lo, hi = s.own_local
        per = (s.nx_glob - 1) // self.world
        if u_zero:  # first visit in this cycle: the right-hand side is new
            if self.world == 1:
                cb.f.copy_(b.f)
            else:
                with self._phase("agglomeration_gather"):
                    # every rank contributes rows [own_lo, own_lo + per]: `per` owned rows plus one more (the
                    # boundary row on the last rank; elsewhere a ghost row that is ignored below).  Rows of a pitched
                    # field are contiguous including their padding, so the slab rows go out as they lie (no staging
                    # copy), ONE all_gather lands them in a flat buffer and ONE strided copy places the `per` owned
                    # rows of every rank in the full grid (round 2: was a staging copy, a list all_gather with its
                    # per-chunk copy-out and world + 1 slice copies -- 0.64 ms per cycle at 8 GPUs, eager).
                    ldp = b.f.stride(0)
                    mine = _rows(b.f, lo, lo + per + 1) if _rows_contiguous(b.f) else \
                        _pitched_rows(b.f, lo, lo + per + 1).contiguous().view(-1)
                    key = (b.f.dtype, mine.numel())
                    flat = self._agg_flat.get(key)
                    if flat is None:
                        flat = self._agg_flat[key] = torch.empty(self.world * mine.numel(), dtype=b.f.dtype,
                                                                 device=b.f.device)
                    dist.all_gather_into_tensor(flat, mine, group=self.group)
                    ny = s.ny
                    src = flat.view(self.world, per + 1, ldp)
                    ldc = cb.f.stride(0)
                    dst = torch.as_strided(cb.f, (self.world, per, ny), (per * ldc, ldc, 1), cb.f.storage_offset())
                    dst.copy_(src[:, :per, :ny])
                    cb.f[self.world * per].copy_(src[self.world - 1, per, :ny])
        with self._phase("agglomerated_sub_cycle"):
            full_u = self.coarse.cycle(dtype, u_zero)
            b.u.copy_(full_u[s.row0:s.row0 + s.loc_nx])
        self.set_valid(b.u, self.part.ghost)  # cut out of the full correction: every ghost row is exact

    # -- the recursion ----------------------------------------------------------------------------------------
    def _reps(self, l: int) -> int:
        if self.cycle_type == "V":
            return 1
        if self.cycle_type == "W":
            return 2
        if self.cycle_type == "F":
            return max(1, 2 ** (self.num_levels - l - 2))
        return 0

    def cycle(self, dtype, l: int = 0, u_zero: bool = False, sumsq_out: Optional[torch.Tensor] = None,
              skip_down: bool = False) -> None:
        """One cycle on distributed level l (all distributed levels in `dtype`).  ``skip_down`` (level 0 only): the
        caller's fused defect + down pass has already left the pre-smoothed iterate in b.u and the restricted
        residual in the next level's f (with their ghost validity recorded)."""
        if l == self.D:
            self._agglomerated_cycle(dtype, u_zero)
            return
        s = self.part.slab(l)
        b, c = self.bufs(l, dtype), self.bufs(l + 1, dtype)
        off, rows = self.part.coarse_view(l)
        G = self.part.ghost
        kw = self._var_kw(l, dtype)
        # sweeps per HBM pass: 2, except fp64 variable-coefficient passes (one sweep: register budget of the kernel)
        ms = 1 if (self.coefficient is not None and dtype == torch.float64) else 2
        # down: `pre` sweeps, the last pass with residual + restriction; dependency cone of a pass = 2 rows per sweep
        # (+2 for the owned coarse rows of the restriction)
        n, uz = (0 if skip_down else self.pre), u_zero
        while n > 0:
            k = min(n, ms)
            last = n - k == 0
            ins = [(b.f, l)] + ([] if uz else [(b.u, l)])
            self.ensure(2 * k + (2 if last else 0), ins)
            v = min(self.vdepth(t) for t, _ in ins)
            self.be.vc_pass(b.u, b.tmp, b.f, s.hx, s.hy, sweeps=k, coefficient=-1.0,
                            coarse_out=c.f[off:off + rows] if last else None, u_zero=uz, **kw)
            b.u, b.tmp = b.tmp, b.u
            self.set_valid(b.u, v - 2 * k)
            if last:
                self.set_valid(c.f, (v - (2 * k + 2)) // 2)
            n -= k
            uz = False
        for rep in range(self._reps(l)):
            self.cycle(dtype, l + 1, u_zero=(rep == 0))
        # up: prolongation in the first pass, `post` sweeps (+ norm in the last): cone 2 rows per sweep (+1); the
        # prolongation of a coarse field with v_c valid ghost rows is exact on 2*v_c - 1 fine ghost rows
        norm = sumsq_out is not None and l == 0
        n, first = self.post, True
        while n > 0:
            k = min(n, ms)
            last = n - k == 0
            items = [(b.u, l), (b.f, l)] + ([(c.u, l + 1)] if first else [])
            eff = (lambda d: min(d[0], d[1], 2 * d[2] - 1)) if first else None
            self.ensure(2 * k + (1 if (norm and last) else 0), items, effective=eff)
            v = min(self.vdepth(b.u), self.vdepth(b.f))
            if first:
                v = min(v, 2 * self.vdepth(c.u) - 1)
            self.be.vc_pass(b.u, b.tmp, b.f, s.hx, s.hy, sweeps=k, coefficient=-1.0,
                            coarse_in=c.u[off:off + rows] if first else None,
                            sumsq_out=sumsq_out if (norm and last) else None, norm_rows=s.own_local, **kw)
            b.u, b.tmp = b.tmp, b.u
            self.set_valid(b.u, v - 2 * k)
            n -= k
            first = False

    # -- data movement helpers -----------------------------------------------------------------------------------
    def scatter_rows(self, dst: torch.Tensor, full_rows_fn, l: int = 0) -> None:
        """Fill a level-l local array (owned + ghost rows) from a function global_row_range -> tensor."""
        s = self.part.slab(l)
        dst.copy_(full_rows_fn(s.row0, s.row0 + s.loc_nx))

    def gather_solution(self, local_u: torch.Tensor, l: int = 0) -> Optional[torch.Tensor]:
        """All ranks -> the full (nx_glob, ny) field on every rank (tests / small grids only)."""
        s = self.part.slab(l)
        lo, hi = s.own_local
        if self.world == 1:
            return local_u[lo:hi].clone()
        per = (s.nx_glob - 1) // self.world
        mine = local_u[lo:lo + per + 1].contiguous() if self.rank == self.world - 1 else torch.cat(
            [local_u[lo:lo + per], local_u[lo + per:lo + per + 1]]).contiguous()
        out = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(out, mine, group=self.group)
        return torch.cat([o[:per] for o in out] + [out[-1][per:per + 1]])


def _rows_contiguous(t: torch.Tensor) -> bool:
    # a pitched field: rows of `ld` elements back to back; a block of rows is one contiguous memory range
    return t.stride(1) == 1


def _rows(t: torch.Tensor, a: int, b: int) -> torch.Tensor:
    """Rows [a, b) of a pitched field INCLUDING the row padding, as one contiguous 1-D tensor (a single message)."""
    ldp = t.stride(0)
    return torch.as_strided(t, ((b - a) * ldp,), (1,), t.storage_offset() + a * ldp)


def _pitched_rows(t: torch.Tensor, a: int, b: int) -> torch.Tensor:
    ldp = t.stride(0)
    return torch.as_strided(t, (b - a, ldp), (ldp, 1), t.storage_offset() + a * ldp)


# ======================================================================================================
# Mixed-precision driver on the distributed engine (same phases as solvers/mixed_precision.py)
# ======================================================================================================
class DistributedMixedPrecisionSolver:
    """fp32-cycle / fp64-residual refinement with a switch to fp64 cycles, on row slabs.
    `step()` runs one cycle of the solve loop and returns the global h-scaled residual norm."""

    def __init__(self, nx: int, ny: int, *, domain=(0.0, 1.0, 0.0, 1.0), precision_strategy: str = "adaptive",
                 switch_threshold: float = 1e-6, tolerance: float = 1e-8, max_iterations: int = 50, backend=None,
                 device=None, use_cuda_graphs: bool = False, stagnation_ratio: float = 0.95,
                 stop_on_rounding_floor: bool = True, use_fused_defect_down: bool = True, **engine_kw):
        self.eng = DistributedCycleEngine(nx, ny, domain=domain, backend=backend, device=device, **engine_kw)
        self.mode = {"double": "fp64", "fp64": "fp64", "single": "fp32", "fp32": "fp32", "refinement": "refine"}.get(
            precision_strategy, "switch")
        self.switch_threshold, self.tolerance, self.max_iterations = switch_threshold, tolerance, max_iterations
        self.stagnation_ratio, self.stop_on_rounding_floor = stagnation_ratio, stop_on_rounding_floor
        s0 = self.eng.part.slab(0)
        self.s0 = s0
        self.hxhy = s0.hx * s0.hy
        self.ss = self.eng.be.scalar(2)
        self.phase = None
        # the fused pass pays off on large slabs only (see MixedPrecisionMultigrid._dd_ok); "always" forces it (tests)
        self.fused_defect_down_min_points = 0 if use_fused_defect_down == "always" else 6000 * 6000
        self.fused_defect_down = bool(use_fused_defect_down) and self._dd_ok()
        self._pre_smoothed = False
        self.precision_switches: List[Dict[str, Any]] = []
        for l in range(self.eng.D + 1):  # final shape of the buffer-role state before the first step
            for dt in ((torch.float64, torch.float32) if self.mode in ("switch", "refine") else
                       ((torch.float64,) if self.mode == "fp64" else (torch.float32,))):
                self.eng.bufs(l, dt)
        from .solvers.graphs import GraphCache
        self.graphs = GraphCache(self.eng.buffer_state, self.eng.snapshot_roles, self.eng.restore_roles,
                                 enabled=use_cuda_graphs)

    # -- right-hand side ---------------------------------------------------------------------------------------
    def set_rhs_from_global(self, f_global) -> None:
        """`f_global`: the full (nx, ny) float64 right-hand side (NumPy or tensor) present on every rank."""
        b = self.eng.bufs(0, torch.float64)
        s = self.s0
        src = torch.as_tensor(f_global)[s.row0:s.row0 + s.loc_nx]
        b.f.copy_(src)
        self.eng.set_valid(b.f, self.eng.part.ghost)

    def set_rhs_sinsin_device(self, amplitude: float = 2 * math.pi ** 2) -> None:
        """Manufactured f = amplitude*sin(pi x) sin(pi y) generated in HBM on the slab (global coordinates)."""
        from . import ops
        b = self.eng.bufs(0, torch.float64)
        s = self.s0
        x0, x1, y0, y1 = self.eng.domain
        xa = x0 + s.row0 * s.hx
        xb = x0 + (s.row0 + s.loc_nx - 1) * s.hx
        ops.fill_sinsin_(b.f, (xa, xb, y0, y1), amplitude, 1.0, 1.0)
        self.eng.set_valid(b.f, self.eng.part.ghost)  # generated on the ghost rows as well

    def zero_boundary_ring_of_rhs(self) -> None:
        b = self.eng.bufs(0, torch.float64)
        s = self.s0
        zr = getattr(self.eng.be, "zero_ring", None)
        if zr is not None:
            zr(b.f, s.own_lo == 0, s.own_hi == s.nx_glob)
            return
        b.f[:, 0] = 0
        b.f[:, -1] = 0
        if s.own_lo == 0:
            b.f[0, :] = 0
        if s.own_hi == s.nx_glob:
            b.f[-1, :] = 0

    # -- solve loop ----------------------------------------------------------------------------------------------
    def _iterate_norm(self, phase: str = "fp64") -> float:
        """Global h-scaled L2 norm of the current iterate (collective: every rank calls it at the same cycle, which
        the policy guarantees because all ranks see the same all-reduced residual norms)."""
        eng = self.eng
        dt = torch.float32 if phase == "fp32" else torch.float64
        lo, hi = self.s0.own_local
        ss = eng.allreduce_sum(eng.be.sumsq(eng.bufs(0, dt).u[lo:hi]))
        return float(np.sqrt(self.hxhy * float(ss.item())))

    def restart(self, keep_iterate: bool = False) -> float:
        """Begin a solve from u = 0 (the iterate is then neither memset nor read: the first passes carry the U_ZERO
        flag), or (keep_iterate) from whatever the fp64 level-0 iterate holds."""
        from .solvers.policy import CyclePolicy, weak_method
        eng = self.eng
        b64 = eng.bufs(0, torch.float64)
        self._fresh = not keep_iterate
        self.policy = CyclePolicy(self.mode, self.tolerance, self.switch_threshold, self.s0.hx, self.s0.hy, eng.shift,
                                  stagnation_ratio=self.stagnation_ratio, stop_on_floor=self.stop_on_rounding_floor,
                                  u_norm=weak_method(self._iterate_norm))
        self.phase = self.policy.phase
        self.history = self.policy.history
        self.last_action = None
        if self.phase == "refine":
            return self._defect(with_update=False, u_zero=self._fresh)
        if self.phase == "fp32":
            b32 = eng.bufs(0, torch.float32)
            b32.f.copy_(b64.f)
            eng.set_valid(b32.f, eng.vdepth(b64.f))
            if not self._fresh:
                b32.u.copy_(b64.u)
                eng.set_valid(b32.u, eng.vdepth(b64.u))
        return float("nan")

    def _norm(self, slot: int) -> float:
        return float(np.sqrt(self.hxhy * self.ss[slot].item()))

    def _defect(self, with_update: bool, u_zero: bool = False) -> float:
        # from the zero iterate the residual is f itself: two launches beat the fused pass there (see
        # MixedPrecisionMultigrid._refinement_residual)
        fuse = self.fused_defect_down and not (u_zero and not with_update)
        self.graphs.run(("defect_u" if with_update else "defect") + ("0" if u_zero else ""),
                        lambda: self._launch_defect(with_update, u_zero, fuse))
        self._pre_smoothed = fuse
        return self._norm(1)

    def _launch_defect(self, with_update: bool, u_zero: bool = False, fuse_next: bool = True) -> None:
        if self.fused_defect_down and fuse_next:
            self._launch_defect_down(with_update, u_zero)
            return
        eng, s = self.eng, self.s0
        b64, b32 = eng.bufs(0, torch.float64), eng.bufs(0, torch.float32)
        kw = eng._var_kw(0, torch.float64)
        if u_zero:
            kw["u_zero"] = True
        self.ss.zero_()
        G = eng.part.ghost
        if with_update:
            eng.ensure(1, [(b32.u, 0), (b64.f, 0)] + ([] if u_zero else [(b64.u, 0)]))
            v = min(G if u_zero else eng.vdepth(b64.u), eng.vdepth(b32.u))
            eng.be.vc_defect_pass(b64.u, b64.tmp, b64.f, s.hx, s.hy, e_in=b32.u, r_out=b32.f, sumsq_out=self.ss[1:2],
                                  norm_rows=s.own_local, **kw)
            b64.u, b64.tmp = b64.tmp, b64.u
            eng.set_valid(b64.u, v)
        else:
            eng.ensure(1, [(b64.f, 0)] + ([] if u_zero else [(b64.u, 0)]))
            v = G if u_zero else eng.vdepth(b64.u)
            eng.be.vc_defect_pass(b64.u, None, b64.f, s.hx, s.hy, r_out=b32.f, sumsq_out=self.ss[1:2],
                                  norm_rows=s.own_local, **kw)
        eng.set_valid(b32.f, min(v, eng.vdepth(b64.f)) - 1)
        eng.allreduce_sum(self.ss)

    def _launch_refine(self, u_zero: bool = False, pre_smoothed: bool = True, fuse_next: bool = True) -> None:
        """pre_smoothed: the down pass of this cycle ran inside the previous defect pass; fuse_next: this cycle's defect
        pass also runs the down pass of the next cycle (skipped when the policy expects this cycle to be the last)."""
        if self.fused_defect_down and pre_smoothed:
            self.eng.cycle(torch.float32, 0, u_zero=True, skip_down=True)
        else:
            self.eng.cycle(torch.float32, 0, u_zero=True)
        self._launch_defect(True, u_zero, fuse_next)

    def _dd_ok(self) -> bool:
        """The refinement cycle can use the fused defect + down pass (ops.vc_defect_down_pass) on the slabs: constant
        coefficients, two pre-smoothing sweeps, at least one distributed level below level 0, a back end that has it."""
        eng = self.eng
        return (self.mode in ("switch", "refine") and eng.coefficient is None and eng.pre == 2 and eng.D >= 1
                and hasattr(eng.be, "vc_defect_down_pass") and getattr(eng.be, "loader", "tma") == "tma"
                and self.s0.loc_nx * self.s0.ny >= self.fused_defect_down_min_points)

    def _launch_defect_down(self, with_update: bool, u_zero: bool) -> None:
        """u64 += e32 ; r32 ; ||r|| over the owned rows ; e' = 2 sweeps from zero on A e = r32 ; f_c = R(r32 - A e')
        in ONE pass over the slab.  Ghost validity: the update is pointwise, the residual reaches one row, two
        sweeps four more, the restriction two more (and halves)."""
        eng, s = self.eng, self.s0
        b64, b32, c32 = eng.bufs(0, torch.float64), eng.bufs(0, torch.float32), eng.bufs(1, torch.float32)
        off, rows = eng.part.coarse_view(0)
        G = eng.part.ghost
        self.ss.zero_()
        ins = [(b64.f, 0)] + ([] if u_zero else [(b64.u, 0)]) + ([(b32.u, 0)] if with_update else [])
        eng.ensure(7, ins)
        v = min(eng.vdepth(t) for t, _ in ins)
        eng.be.vc_defect_down_pass(b64.u, b64.tmp if with_update else None, b64.f, s.hx, s.hy,
                                   e_in=b32.u if with_update else None, r_out=b32.f, e_out=b32.tmp,
                                   coarse_out=c32.f[off:off + rows], sumsq_out=self.ss[1:2], u_zero=u_zero,
                                   norm_rows=s.own_local, shift=eng.shift)
        if with_update:
            b64.u, b64.tmp = b64.tmp, b64.u
            eng.set_valid(b64.u, min(G if u_zero else eng.vdepth(b64.tmp), eng.vdepth(b32.u)))
        b32.u, b32.tmp = b32.tmp, b32.u
        eng.set_valid(b32.f, v - 1)
        eng.set_valid(b32.u, v - 1 - 4)
        eng.set_valid(c32.f, (v - 1 - 6) // 2)
        eng.allreduce_sum(self.ss)

    def _launch_uniform(self, dt, u_zero: bool = False) -> None:
        self.ss.zero_()
        self.eng.cycle(dt, 0, u_zero=u_zero, sumsq_out=self.ss[0:1])
        self.eng.allreduce_sum(self.ss)

    def step(self) -> float:
        """One cycle of the solve loop; returns the global h-scaled residual norm.  `self.last_action` says what
        the policy (solvers/policy.py) made of it: 'continue', 'converged' or 'rounding_floor'."""
        first = self._fresh and not self.history
        if self.phase == "refine":
            if self.fused_defect_down:
                # every rank sees the same all-reduced norms, so every rank takes the same variant
                pre, fuse = self._pre_smoothed, not self.policy.likely_last()
                key = "refine" + ("_dd" if pre else "_full") + ("" if fuse else "_last") + ("0" if first else "")
                self.graphs.run(key, lambda: self._launch_refine(first, pre, fuse))
                self._pre_smoothed = fuse
            else:
                self.graphs.run("refine0" if first else "refine", lambda: self._launch_refine(first))
            norm = self._norm(1)
        else:
            dt = torch.float64 if self.phase == "fp64" else torch.float32
            self.graphs.run(self.phase + ("0" if first else ""), lambda: self._launch_uniform(dt, first))
            norm = self._norm(0)
        self.last_action = self.policy.observe(norm)
        self.phase = self.policy.phase
        self.precision_switches = self.policy.switches
        return norm

    def solve(self, keep_iterate: bool = False):
        from .solvers.policy import CONTINUE, CONVERGED
        self.precision_switches = []
        r0 = self.restart(keep_iterate)
        converged = False
        for _ in range(self.max_iterations):
            self.step()
            if self.last_action != CONTINUE:
                converged = self.last_action == CONVERGED  # the reference's meaning; see `stopped_on` otherwise
                break
        dt = torch.float32 if self.mode == "fp32" else torch.float64
        u = self.eng.bufs(0, dt).u
        return u, {"converged": converged, "iterations": len(self.history), "residual_history": list(self.history),
                   "final_residual": self.history[-1], "initial_residual": r0,
                   "stopped_on": self.policy.stopped_on, "attainable_residual": self.policy.floor_bound,
                   "switch_blocked": self.policy.switch_blocked,
                   "precision_switches": list(self.precision_switches), "dist_levels": self.eng.D,
                   "num_levels": self.eng.num_levels, "halo_exchanges": self.eng.exchanges}


    def solve_many(self, f_hosts: Sequence[torch.Tensor], u_hosts: Sequence[torch.Tensor]) -> List[Dict[str, Any]]:
        """A batch of solves on the slabs (many right-hand sides on one grid), software-pipelined over this rank's
        copy engines exactly like `MixedPrecisionMultigrid.solve_many`: while solve k cycles, the slab of right-hand
        side k+1 is uploaded on one side stream and the slab of solution k-1 downloaded on another.
        `f_hosts[k]` / `u_hosts[k]`: this rank's PINNED host slabs, (loc_nx, ny) float64 -- owned rows plus ghost rows,
        the layout of `eng.bufs(0, float64).f`.  Collective: every rank calls it with the same number of problems.
        Each solve starts from u = 0 and runs the same launches as `solve()`."""
        n = len(f_hosts)
        if n != len(u_hosts):
            raise ValueError("solve_many: one output slab per right-hand side")
        if n == 0:
            return []
        eng = self.eng
        b = eng.bufs(0, torch.float64)
        dev = b.f.device
        shape = (self.s0.loc_nx, self.s0.ny)
        for t in list(f_hosts) + list(u_hosts):
            if tuple(t.shape) != shape or t.dtype != torch.float64:
                raise ValueError(f"solve_many: host slabs must be float64 {shape}")
        if getattr(self, "_stage", None) is None:
            self._stage = ([eng.be.empty(*shape, torch.float64) for _ in range(2)],
                           [eng.be.empty(*shape, torch.float64) for _ in range(2)],
                           torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        f_stage, u_stage, s_in, s_out = self._stage
        cur = torch.cuda.current_stream(dev)
        consumed: List[Any] = [None, None]
        drained: List[Any] = [None, None]
        uploaded: List[Any] = [None] * n

        def upload(k):
            slot = k % 2
            with torch.cuda.stream(s_in):
                if consumed[slot] is not None:
                    s_in.wait_event(consumed[slot])
                f_stage[slot].copy_(f_hosts[k], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
            uploaded[k] = ev

        upload(0)
        infos = []
        for k in range(n):
            slot = k % 2
            if k + 1 < n:
                upload(k + 1)
            cur.wait_event(uploaded[k])
            b = eng.bufs(0, torch.float64)
            b.f.copy_(f_stage[slot])
            eng.set_valid(b.f, eng.part.ghost)  # the host slab carries its ghost rows
            ev = torch.cuda.Event()
            ev.record(cur)
            consumed[slot] = ev
            u, info = self.solve()
            if drained[slot] is not None:
                cur.wait_event(drained[slot])
            u_stage[slot].copy_(u)
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ready)
                u_hosts[k].copy_(u_stage[slot], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_out)
            drained[slot] = ev
            infos.append(info)
        s_out.synchronize()
        return infos

    def release_staging(self) -> None:
        """Free the double-buffered device staging slabs `solve_many` keeps between calls."""
        self._stage = None


# ======================================================================================================
# Implicit heat stepping on row slabs (BASELINE configs[4]: one distributed shifted multigrid solve per step)
# ======================================================================================================
class DistributedHeatSolver:
    """theta-method time stepping of u_t = alpha*lap(u) + f on row slabs, the multi-GPU twin of
    `applications.heat_solver.HeatSolver2D` (same linear system, docs/methodology.md:710 divided by theta*alpha*dt:
    (-lap_h + lambda) u^{n+1} = lambda*rhs, lambda = 1/(theta*alpha*dt); same result keys as the reference's
    applications/heat_solver.py:227-247).  The iterate never leaves the slabs: per step every rank forms its rows of
    the right-hand side, one all-reduce gives the rhs norm for the relative stopping test, and the shifted
    mixed-precision solve starts from u^n.  No reference oracle exists for this row: parity unpinned,
    validated against the single-GPU solver (bit-identical owned rows) and the analytical solutions."""

    def __init__(self, *, max_iterations: int = 50, tolerance: float = 1e-8, cycle_type: str = "V",
                 precision_strategy: str = "adaptive", backend=None, device=None, use_cuda_graphs: bool = False,
                 **engine_kw):
        if precision_strategy in ("single", "fp32"):
            raise ValueError("the heat driver keeps the fp64 iterate between steps: use 'adaptive', 'refinement' or 'double'")
        self.max_iterations, self.tolerance, self.cycle_type = max_iterations, tolerance, cycle_type
        self.precision_strategy, self.backend, self.device = precision_strategy, backend, device
        self.use_cuda_graphs, self.engine_kw = use_cuda_graphs, engine_kw
        self._solvers: Dict[Tuple, DistributedMixedPrecisionSolver] = {}
        self.time_history: List[Dict[str, Any]] = []

    @staticmethod
    def _theta(cfg) -> float:
        m = getattr(cfg.method, "value", cfg.method)
        if m == "backward_euler":
            return 1.0
        if m == "crank_nicolson":
            return 0.5
        if m == "theta_method":
            if not 0.0 < cfg.theta <= 1.0:
                raise ValueError("theta must be in (0, 1] for an implicit step")
            return cfg.theta
        raise ValueError(f"Unknown time stepping method: {cfg.method}")

    def _solver(self, nx, ny, domain, lam, coefficient=None) -> "DistributedMixedPrecisionSolver":
        key = (nx, ny, tuple(domain), lam, id(coefficient) if coefficient is not None else None)
        sol = self._solvers.get(key)
        if sol is None:
            kw = dict(self.engine_kw)
            if coefficient is not None:
                kw["coefficient"] = coefficient
            sol = DistributedMixedPrecisionSolver(nx, ny, domain=domain, precision_strategy=self.precision_strategy,
                                                  tolerance=self.tolerance, max_iterations=self.max_iterations,
                                                  cycle_type=self.cycle_type, shift=lam, backend=self.backend,
                                                  device=self.device, use_cuda_graphs=self.use_cuda_graphs,
                                                  **kw)
            self._solvers[key] = sol
        return sol

    def solve_heat_problem(self, problem, nx: int, ny: int, time_config, gather: bool = True) -> Dict[str, Any]:
        import time
        domain = tuple(float(v) for v in problem.domain)
        theta = self._theta(time_config)
        diff = problem.thermal_diffusivity
        # a number alpha: u_t = alpha lap u + f; a callable a(X, Y) / an (nx, ny) array: u_t = div(a grad u) + f
        coef = diff if (callable(diff) or isinstance(diff, (np.ndarray, torch.Tensor))) else None
        alpha = 1.0 if coef is not None else float(diff)
        dt, t_cur, step, total_mg = float(time_config.dt), 0.0, 0, 0
        sol = self._solver(nx, ny, domain, 1.0 / (theta * alpha * dt), coef)
        eng, s = sol.eng, sol.s0
        G = eng.part.ghost
        lo, hi = s.own_local
        x = domain[0] + (s.row0 + np.arange(s.loc_nx)) * s.hx
        y = domain[2] + np.arange(ny) * s.hy
        X, Y = np.meshgrid(x, y, indexing="ij")  # the slab's rows of Grid.X / Grid.Y (core/grid.py:50), ghosts included

        EB = 64  # rows per callback call

        def evaluate(fn, *args):
            """User callback on the slab, in blocks of 64 grid rows ALIGNED TO THE GLOBAL ROW INDEX: vectorised libm
            routines may round the same argument differently depending on its position in the array, so evaluating the
            slab as a whole would make the last bit of the data depend on the number of ranks.  Every rank evaluates
            the same global blocks [64 k, 64 k + 64) (whole blocks, clipped only at the end of the grid) and keeps its
            rows: identical arrays go into the callback on every decomposition.  (Round 1 made one call per row:
            8193 Python calls per field at configs[4].)"""
            out = np.empty(X.shape, dtype=np.float64)
            g0, g1 = s.row0, s.row0 + s.loc_nx
            for blk in range(g0 // EB, (g1 - 1) // EB + 1):
                a0, a1 = blk * EB, min((blk + 1) * EB, s.nx_glob)
                xb = domain[0] + np.arange(a0, a1) * s.hx
                Xb, Yb = np.meshgrid(xb, y, indexing="ij")
                vb = np.broadcast_to(np.asarray(fn(Xb, Yb, *args), dtype=np.float64), Xb.shape)
                lo_, hi_ = max(a0, g0), min(a1, g1)
                out[lo_ - g0:hi_ - g0] = vb[lo_ - a0:hi_ - a0]
            return out

        def to_slab(a, like):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(like.device)

        def source(t):
            if problem.source_function is None:
                return None
            f = evaluate(problem.source_function, t)
            return f if f.any() else None

        b = eng.bufs(0, torch.float64)
        b.u.copy_(to_slab(evaluate(problem.initial_condition), b.u))
        b.u[:, 0] = 0
        b.u[:, -1] = 0
        if s.own_lo == 0:
            b.u[0, :] = 0
        if s.own_hi == s.nx_glob:
            b.u[-1, :] = 0
        eng.set_valid(b.u, G)  # evaluated on the ghost rows as well
        f_old = source(0.0) if theta < 1.0 else None
        solver_time, t_start = 0.0, time.time()
        while t_cur < time_config.t_final - 1e-14:
            if t_cur + dt > time_config.t_final:
                dt = time_config.t_final - t_cur
            step += 1
            t_new = t_cur + dt
            lam = 1.0 / (theta * alpha * dt)
            nsol = self._solver(nx, ny, domain, lam, coef)
            if nsol is not sol:  # shortened last step: another shift, hence another solver; hand the iterate over
                nb = nsol.eng.bufs(0, torch.float64)
                nb.u.copy_(eng.bufs(0, torch.float64).u)
                nsol.eng.set_valid(nb.u, eng.vdepth(eng.bufs(0, torch.float64).u))
                sol, eng = nsol, nsol.eng
            t0 = time.time()
            b = eng.bufs(0, torch.float64)
            # rhs = lambda * (u + (1-theta) alpha dt lap_h u + dt (theta f_new + (1-theta) f_old)) on the local rows, its
            # boundary ring zeroed and the sum of squares of the owned rows: ONE kernel (mg_heat_rhs)
            v = eng.vdepth(b.u)
            if theta < 1.0:
                eng.ensure(1, [(b.u, 0)])
                v = eng.vdepth(b.u) - 1  # lap_h of a ghost row needs the next one
            f_new = source(t_new)
            ss = eng.be.heat_rhs(b.u, b.f, s.hx, s.hy, lam=lam, c_lap=(1.0 - theta) * alpha * dt,
                                 f1=to_slab(f_new, b.f) if f_new is not None else None, c_f1=dt * theta,
                                 f0=to_slab(f_old, b.f) if (theta < 1.0 and f_old is not None) else None,
                                 c_f0=dt * (1.0 - theta), zero_first_row=s.own_lo == 0,
                                 zero_last_row=s.own_hi == s.nx_glob, norm_rows=(lo, hi),
                                 **({"a": eng.coef(0, torch.float64)} if coef is not None else {}))
            eng.set_valid(b.f, v)
            # relative stopping test against the GLOBAL rhs norm: the right-hand side scales with lambda
            ss = eng.allreduce_sum(ss)
            scale = float(np.sqrt(s.hx * s.hy * float(ss.item())))
            sol.tolerance = self.tolerance * max(scale, 1e-300)
            sol.switch_threshold = max(1e-6 * scale, sol.tolerance)
            _, info = sol.solve(keep_iterate=True)
            solver_time += time.time() - t0
            total_mg += info["iterations"]
            f_old, t_cur = f_new, t_new
        total = time.time() - t_start
        u_loc = eng.bufs(0, torch.float64).u
        errors: Dict[str, Any] = {}
        results: Dict[str, Any] = {
            "problem_name": problem.name, "grid_size": (nx, ny),
            "time_config": {"method": getattr(time_config.method, "value", time_config.method),
                            "dt_initial": time_config.dt, "dt_final": dt, "t_final": time_config.t_final,
                            "adaptive_dt": time_config.adaptive_dt},
            "final_solution": eng.gather_solution(u_loc).cpu().numpy() if gather else None,
            "local_solution": u_loc[lo:hi], "local_rows": (s.own_lo, s.own_hi),
            "final_time": t_cur, "total_steps": step, "total_time": total, "total_solver_time": solver_time,
            "avg_mg_iterations": total_mg / step if step else 0, "total_mg_iterations": total_mg, "errors": errors,
            "solver_type": "multigrid", "use_gpu": u_loc.is_cuda, "n_gpus": eng.world,
            "halo_exchanges": sum(sv.eng.exchanges for sv in self._solvers.values()),
        }
        if problem.analytical_solution is not None:
            exact = to_slab(evaluate(problem.analytical_solution, t_cur), u_loc)[lo:hi]
            err = u_loc[lo:hi] - exact
            acc = torch.stack([(err ** 2).sum(), (exact ** 2).sum()]).to(torch.float64)
            eng.allreduce_sum(acc)
            mx = torch.stack([err.abs().max(), exact.abs().max()]).to(torch.float64)
            if eng.world > 1:
                dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=eng.group)
            l2e, l2n = float(np.sqrt(acc[0].item() * s.hx * s.hy)), float(np.sqrt(acc[1].item() * s.hx * s.hy))
            errors.update({"l2_error": l2e, "relative_l2_error": l2e / l2n if l2n > 0 else l2e,
                           "max_error": float(mx[0].item()),
                           "relative_max_error": float(mx[0].item() / mx[1].item()) if mx[1].item() > 0 else float(mx[0].item()),
                           "grid_spacing": (s.hx, s.hy)})
        self.time_history.append(results)
        return results


# ======================================================================================================
# bench.py leg for N > 1 GPUs (weak scaling: every GPU owns a (n-1) x n slab of a (N(n-1)+1) x n grid)
# ======================================================================================================
def run_distributed_bench(a, world: int, rank: int, dev, peak: float, peak_src: str, ClockSampler):
    import time

    from . import _lib, ops
    n = a.n
    strong = bool(getattr(a, "strong", False))
    if strong:  # BASELINE configs[3] style: one n x n grid on the unit square split over the GPUs
        nx, ny, domain = n, n, (0.0, 1.0, 0.0, 1.0)
    else:
        nx, ny = world * (n - 1) + 1, n
        domain = (0.0, float(world), 0.0, 1.0)  # square cells, h = 1/(n-1): same spacing as the 1-GPU workload
    tol = a.tolerance if a.tolerance is not None else 1e-8  # the reference's; unattainable sizes end on the rounding floor
    parity = _bench_parity_check(a, world, rank, dev)  # N-GPU == 1-GPU evidence in the SCALE line (untimed)
    sol = DistributedMixedPrecisionSolver(nx, ny, domain=domain, precision_strategy=a.strategy, switch_threshold=1e-6,
                                          tolerance=tol, cycle_type=a.cycle, backend=DeviceBackend(dev, a.loader),
                                          device=dev, use_cuda_graphs=not a.no_graphs,
                                          agglomerate_below=_bench_agg(a),
                                          use_fused_defect_down=not getattr(a, "no_dd", False),
                                          ghost=getattr(a, "ghost", None) or BENCH_GHOST, **_halo_kw(a, dev))
    sol.set_rhs_sinsin_device()
    sol.zero_boundary_ring_of_rhs()
    sol.eng.exchange(sol.eng.bufs(0, torch.float64).f, 0)
    state = {"solves": 0, "cycles": [], "last": None}

    def step():
        sol.step()
        if sol.last_action != "continue" or len(sol.history) >= 30:
            state["solves"] += 1
            state["stopped_on"] = sol.policy.stopped_on
            state["cycles"].append(len(sol.history))
            state["last"] = list(sol.history)
            sol.restart()

    sol.restart()
    from .solvers.graphs import prime
    primed = prime(step, sol.graphs, lambda: state["solves"])  # setup: capture every step graph (see bench.py)
    for _ in range(max(3, a.warmup)):
        step()
    # settle: the first process on a fresh multi-GPU box measured ~4 % slower steps than a second run of the same
    # command (NVLink / NCCL channels warming up); keep stepping, untimed, in batches of 10 until two successive batches
    # agree within 1 % on every rank (at most 30 batches)
    settle = {"batches": 0}
    prev = None
    for _ in range(30):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            step()
        torch.cuda.synchronize()
        tb = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        cur = float(tb.item())
        settle["batches"] += 1
        if prev is not None and abs(cur - prev) <= 0.01 * prev:
            break
        prev = cur
    settle["last_batch_ms_per_step"] = round(cur * 100.0, 4)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ex0 = sol.eng.exchanges
    with ClockSampler(dev.index) as clk:
        torch.cuda.synchronize()
        dist.barrier()
        if hasattr(clk, "mark"):
            clk.mark(0)
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        if hasattr(clk, "mark"):
            clk.mark(1)
        dist.barrier()
    ms_local = e0.elapsed_time(e1)
    ex_per_step = (sol.eng.exchanges - ex0) / max(1, a.steps)
    t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = nx * ny * a.steps / (ms * 1e-3)
    # per-kernel durations and launch counts: the same steps once more, eagerly, with events on rank 0
    sol.graphs.enabled = False
    ops.TIMER = ops.KernelTimer(min_points=sol.s0.loc_nx * ny // 2) if rank == 0 else None
    launches0 = _lib.call("mg_launch_count")
    ex0 = sol.eng.exchanges
    sol.eng.phase_events = []
    torch.cuda.synchronize()
    dist.barrier()
    t_eager = time.perf_counter()
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    t_eager = (time.perf_counter() - t_eager) * 1e3 / max(1, a.steps)
    phases = {k: round(v / max(1, a.steps), 4) for k, v in sol.eng.phase_summary().items()}
    phases["eager_step_wall"] = round(t_eager, 4)
    sol.eng.phase_events = None
    dist.barrier()
    launches = _lib.call("mg_launch_count") - launches0
    ex_per_step = (sol.eng.exchanges - ex0) / max(1, a.steps)
    kern = ops.TIMER.summary() if ops.TIMER is not None else {}
    ops.TIMER = None
    sol.graphs.enabled = not a.no_graphs

    kernels = {}
    roof = None
    s0 = sol.s0
    pts = s0.loc_nx * ny
    for tag, d in sorted(kern.items(), key=lambda kv: -kv[1]["total_ms"]):
        name, dtn, _ = tag.split("/")
        w = 8 if dtn == "f64" else 4
        fused = name.startswith("dd:")  # fused defect + down pass: the defect pass's traffic + e' and f_c written
        if fused:
            name = name[3:]
        if "resid32" in name or "update" in name:  # defect pass: read u64, f64, write r32 (+ read e32, write u64)
            b = (20 if "resid32" in name else 0) + (12 if "update" in name else 0) - (8 if name.startswith("Z+") else 0)
            b += 5 if fused else 0
        else:
            b = 3.0 * w - (w if name.startswith("Z+") else 0) + (0.25 * w if "P+" in name else 0) + (0.25 * w if "+R" in name else 0)
        ach = b * pts / (d["mean_ms"] * 1e-3) / 1e9
        kernels[tag] = {"launches": d["launches"], "mean_ms": round(d["mean_ms"], 4), "hbm_gbs": round(ach, 1),
                        "frac_of_peak": round(ach / peak, 4)}
        if roof is None:
            roof = {"bound": "hbm", "kernel": tag, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": b * pts, "rank": 0}

    # end to end: pinned host slab of f -> device, solve, device slab of u -> pinned host (every rank its slab)
    e2e = None
    if not a.no_e2e:
        b64 = sol.eng.bufs(0, torch.float64)
        numa = bind_to_gpu_numa_node(dev)  # before the pinned slabs are allocated (first touch decides the node)
        # ONE pinned output slab receives every solution of the batch (downloads are serialised on their stream): the
        # host footprint stays at two slabs per rank (34 GB of pinned memory over 8 ranks).  A rank that cannot pin its
        # slabs must not leave the others waiting in a collective: all ranks agree first.
        try:
            f_host = torch.empty((s0.loc_nx, ny), dtype=torch.float64, pin_memory=True)
            u_hosts = [torch.empty((s0.loc_nx, ny), dtype=torch.float64, pin_memory=True)]
            ok = 1.0
        except (RuntimeError, MemoryError) as exc:
            import sys
            print(f"[bench] rank {rank}: cannot pin host slabs for the end-to-end leg: {exc}", file=sys.stderr)
            f_host, u_hosts, ok = None, None, 0.0
        flag = torch.tensor([ok], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if not a.no_e2e and float(flag.item()) > 0:
        f_host.copy_(b64.f)
        torch.cuda.synchronize()
        # single solve, nothing overlapped (what one `solve()` call with host slabs costs)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        b64.f.copy_(f_host, non_blocking=True)
        u, info1 = sol.solve()
        u_hosts[0].copy_(u, non_blocking=True)
        torch.cuda.synchronize()
        dist.barrier()
        t1 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t1, op=dist.ReduceOp.MAX)
        # the batch API, the same method as the 1-GPU line: B solves, each with its own upload and download inside
        # the timed region, transfers of neighbouring solves overlapping the cycles
        B = 6
        sol.solve_many([f_host] * 2, u_hosts * 2)  # warm-up: staging slabs, side streams
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        infos = sol.solve_many([f_host] * B, u_hosts * B)
        torch.cuda.synchronize()
        dist.barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        tot_c = sum(i["iterations"] for i in infos)
        info = infos[-1]
        e2e = {"value": nx * ny * tot_c / float(tt.item()), "unit": "unknowns/s",
               "h2d_bytes_per_step": s0.loc_nx * ny * 8 * world, "d2h_bytes_per_step": s0.loc_nx * ny * 8 * world,
               "step": "one solve of a solve_many() batch of %d on the slabs: every rank uploads its pinned host slab of f "
                       "and downloads its slab of u; the transfers of neighbouring solves overlap the cycles" % B,
               "seconds_per_solve": float(tt.item()) / B, "iterations": info["iterations"],
               "final_residual": info["final_residual"], "numa": numa,
               "single_solve": {"value": nx * ny * info1["iterations"] / float(t1.item()),
                                "seconds_per_solve": float(t1.item()),
                                "step": "one distributed solve() with host slabs, nothing overlapped"}}
        sol.release_staging()
        del u_hosts, f_host
    clocks = clk.summary()
    return {
        "metric": _metric_name(), "value": value, "unit": "unknowns/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(3, a.warmup), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32 cycle / f64 iterate+residual" if sol.mode in ("switch", "refine") else sol.mode,
        "data": "synthetic",
        "config": {"workload": f"2D Poisson {nx}x{ny} ({(nx - 1) // world} rows per GPU) manufactured sin*sin on "
                               f"({domain[0]:g},{domain[1]:g})x(0,1), "
                               f"{a.cycle}(2,2) red-black GS, precision_strategy={a.strategy}, row slabs over {world} GPUs",
                   "levels": sol.eng.num_levels, "distributed_levels": sol.eng.D, "ghost_rows": sol.eng.part.ghost,
                   "agglomerated_grid": list(sol.eng.part.agg_shape), "tolerance": tol,
                   "halo_exchanges_per_step": ex_per_step, "halo": getattr(a, "halo", "nccl"),
                   # rank 0, eager replay of the timed steps with CUDA events around the communication steps and the
                   # agglomerated tail (ms per step; the level passes are in `kernels`)
                   "phases_ms_per_step_eager": phases,
                   "cuda_graphs": (not a.no_graphs),
                   "graphs_captured": sol.graphs.captured, "priming_solves": primed,
                   "settle_warmup": settle,
                   "l2": "slab arrays (>= 1 GB) exceed the 126 MB L2; no flush needed",
                   "cycles_per_solve": state["cycles"][-3:], "last_residual_history": state["last"],
                   "stopped_on": state.get("stopped_on"), "parity": parity},
        "roofline": roof, "kernels": kernels, "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clocks,
    }


def _bench_parity_check(a, world: int, rank: int, dev, n: int = 4097) -> Dict[str, Any]:
    """Driver-visible evidence that N GPUs compute what one GPU computes (SURVEY 8e acceptance): the same n x n problem
    (same strategy and cycle type as the benchmark) is solved on the row slabs of all ranks and by the single-GPU
    solver on rank 0; owned rows must agree bit for bit, cycle counts must be equal.  Runs before the timed region."""
    from . import ops
    from .problems import PoissonProblem
    from .solvers.mixed_precision import MixedPrecisionMultigrid
    from .device import empty_field
    while (n - 1) % (world * 64) != 0:
        n = 2 * (n - 1) + 1
    f = empty_field(n, n, torch.float64, dev)
    ops.fill_sinsin_(f, (0.0, 1.0, 0.0, 1.0), 2 * math.pi ** 2, 1.0, 1.0)
    ops.zero_ring_(f)
    dsol = DistributedMixedPrecisionSolver(n, n, precision_strategy=a.strategy, switch_threshold=1e-6, tolerance=1e-8,
                                           cycle_type=a.cycle, backend=DeviceBackend(dev, a.loader), device=dev,
                                           use_cuda_graphs=False, agglomerate_below=_bench_agg(a),
                                           # the fused defect + down pass of the timed run, forced on this smaller grid
                                           use_fused_defect_down=False if getattr(a, "no_dd", False) else "always",
                                           ghost=getattr(a, "ghost", None) or BENCH_GHOST, **_halo_kw(a, dev))
    dsol.set_rhs_from_global(f)
    u, info = dsol.solve()
    full = dsol.eng.gather_solution(u)
    out: Dict[str, Any] = {"grid": [n, n], "cycles_distributed": info["iterations"], "dist_levels": dsol.eng.D,
                           "ghost_rows": dsol.eng.part.ghost, "fused_defect_down": bool(dsol.fused_defect_down)}
    if rank == 0:
        single = MixedPrecisionMultigrid(precision_strategy=a.strategy, switch_threshold=1e-6, tolerance=1e-8,
                                         cycle_type=a.cycle, loader=a.loader, device=dev, strict_reference_norm=True)
        us, si = single.solve(PoissonProblem(rhs=f, nx=n, ny=n))
        hd, hs = info["residual_history"], si["residual_history"]
        out.update({"cycles_single": si["iterations"], "cycles_equal": si["iterations"] == info["iterations"],
                    "max_abs_diff": float((full - us).abs().max().item()),
                    "history_max_rel_diff": max(abs(x - y) / y for x, y in zip(hd, hs)) if len(hd) == len(hs) else None})
        del single, us
    del dsol, u, full, f
    torch.cuda.synchronize()
    dist.barrier()
    return out


def _bench_agg(a) -> int:
    """V-cycles agglomerate at BENCH_AGG; a W / F cycle visits level l 2^l times, so every extra distributed level
    multiplies its exchanges: those keep the engine's default (1025)."""
    return getattr(a, "agg", None) or (BENCH_AGG if getattr(a, "cycle", "V") == "V" else 1025)


def _halo_kw(a, dev) -> Dict[str, Any]:
    if getattr(a, "halo", "nccl") != "p2p":
        return {}
    from .halo import SymmMemTransport
    return {"transport": SymmMemTransport(dev, getattr(a, "ghost", None) or BENCH_GHOST)}


def bind_to_gpu_numa_node(dev) -> Dict[str, Any]:
    """Pin this process to the CPU cores NVML reports as local to its GPU, so that pinned host buffers allocated (first
    touched) afterwards live on the GPU's NUMA node.  With one process per GPU all sharing the default affinity, every
    rank's staging memory otherwise lands on one socket and the host <-> device copies of 8 ranks share its memory
    controllers and PCIe root (round 1: 15 GB/s per rank).  Best effort: returns what it did."""
    import os
    info: Dict[str, Any] = {"bound": False}
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if vis and all(v.strip().isdigit() for v in vis.split(",")):
            idx = int(vis.split(",")[idx])
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = sorted(local & allowed)
        info.update({"gpu_local_cpus": len(local), "allowed_cpus": len(allowed), "usable": len(use)})
        if use and len(use) < len(allowed):
            os.sched_setaffinity(0, use)
            info["bound"] = True
    except Exception as exc:  # no NVML, no permission, ...
        info["error"] = str(exc)[:120]
    return info


def _metric_name() -> str:
    import json
    import os
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        return json.load(open(os.path.join(root, "BASELINE.json")))["metric"]
    except Exception:
        return "V-cycle fine-grid unknowns/sec + smoother HBM GB/s vs peak at 1/2/4/8 B200"
