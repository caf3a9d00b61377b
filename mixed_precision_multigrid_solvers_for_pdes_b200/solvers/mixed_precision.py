"""``MixedPrecisionMultigrid``: the facade the reference documents but never ships
(README.md:73-92, docs/sphinx/index.rst:23-36, docs/TROUBLESHOOTING.md:144-253), built on the
B200 cycle engine.

    solver = MixedPrecisionMultigrid(precision_strategy='adaptive', switch_threshold=1e-6, use_gpu=True)
    solution, info = solver.solve(problem)

Precision driver (normative spec: docs/methodology.md:325-360; thresholds and names:
core/precision.py:155-302, applications/mixed_precision_analysis.py:72-105):

  * fp32 phase  --  "fp32 smoothing / fp64 residual" iterative refinement: the iterate u and the
    right-hand side stay in fp64 in HBM; each cycle computes r = f - A u in fp64, rounds it to fp32,
    runs one V/W-cycle on A e = r entirely in fp32 (e0 = 0) and adds e to the fp64 iterate.  In exact
    arithmetic this IS one multigrid cycle on A u = f, so cycle counts match the fp64 reference, and
    unlike an all-fp32 iterate it can reach discretisation accuracy below fp32 resolution
    (3e-9 at 16385^2).
  * switch      --  once ||r|| <= switch_threshold, or the residual stagnates (ratio > 0.95 over the last
    cycles, precision.py:189-246 / methodology.md:337), the driver continues with fp64 cycles.
  * strategies  --  'double' (fp64 only: the reference's convergent configuration), 'single' (fp32 only,
    floors at fp32 resolution), 'adaptive' / 'mixed' / 'conservative' / 'mixed_conservative' (switch at
    `switch_threshold`, default 1e-6), 'aggressive' / 'mixed_aggressive' (default 1e-4),
    'refinement' (fp32 cycles all the way; promotion only on stagnation).

The convergence test is the reference's: h-scaled L2 norm of f - A u over all points < tolerance
(core/grid.py:174-187, solvers/base.py:123-143)."""
from __future__ import annotations

import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import ops
from ..core.grid import Grid
from ..device import empty_field, like_input, require_cuda, to_device
from ..operators.laplacian import HelmholtzOperator, LaplacianOperator
from ..operators.transfer import ProlongationOperator, RestrictionOperator
from .engine import CycleEngine
from .graphs import GraphCache
from .smoothers import GaussSeidelSmoother, JacobiSmoother

_STRATEGIES = {
    "double": ("fp64", None), "fp64": ("fp64", None), "float64": ("fp64", None),
    "single": ("fp32", None), "fp32": ("fp32", None), "float32": ("fp32", None),
    "adaptive": ("switch", 1e-6), "mixed": ("switch", 1e-6), "conservative": ("switch", 1e-6),
    "mixed_conservative": ("switch", 1e-6), "aggressive": ("switch", 1e-4), "mixed_aggressive": ("switch", 1e-4),
    "refinement": ("refine", None),
}


class MixedPrecisionMultigrid:
    def __init__(self, precision_strategy: str = "adaptive", switch_threshold: Optional[float] = None,
                 use_gpu: bool = True, max_iterations: int = 50, tolerance: float = 1e-8,
                 max_levels: Optional[int] = None, cycle_type: str = "V", pre_smooth_iterations: int = 2,
                 post_smooth_iterations: int = 2, smoother: str = "red_black_gauss_seidel",
                 damping_factor: float = 1.0, coarse_tolerance: float = 1e-12, coarse_max_iterations: int = 1000,
                 stagnation_ratio: float = 0.95, max_grid_size: Optional[int] = None,
                 gpu_memory_fraction: Optional[float] = None, min_precision: Optional[str] = None,
                 strict_reference_norm: bool = False, kernels: str = "auto", loader: str = "tma", device=None,
                 use_cuda_graphs: bool = True, shift: float = 0.0, fmg: bool = False, verbose: bool = False):
        key = str(precision_strategy).lower()
        if key not in _STRATEGIES:
            raise ValueError(f"Unknown precision strategy: {precision_strategy}")
        if not use_gpu:
            raise ValueError("use_gpu=False is not available: this build has no CPU path (the reference's NumPy "
                             "solver is the CPU implementation)")
        self.precision_strategy = key
        self.mode, default_thr = _STRATEGIES[key]
        self.switch_threshold = switch_threshold if switch_threshold is not None else (default_thr or 1e-6)
        if min_precision is not None and str(min_precision).lower() in ("double", "fp64", "float64"):
            self.mode = "fp64"
        self.use_gpu, self.max_iterations, self.tolerance = use_gpu, max_iterations, tolerance
        self.max_levels, self.cycle_type = max_levels, cycle_type
        self.pre, self.post = pre_smooth_iterations, post_smooth_iterations
        self.smoother_name, self.damping_factor = smoother, damping_factor
        self.coarse_tolerance, self.coarse_max_iterations = coarse_tolerance, coarse_max_iterations
        self.stagnation_ratio = stagnation_ratio
        self.max_grid_size, self.gpu_memory_fraction = max_grid_size, gpu_memory_fraction
        self.kernels, self.loader, self.device, self.verbose = kernels, loader, device, verbose
        # The reference's residual equals f on the boundary ring and its norm sums over ALL points
        # (laplacian.py:64,117; grid.py:187), so a source that does not vanish on the boundary can never
        # meet the tolerance although those values enter no equation.  Unless the strict
        # behaviour is requested, the ring of f is zeroed so the test measures the interior residual.
        self.strict_reference_norm = strict_reference_norm
        # One multigrid cycle is ~3 launches per level, most of them microseconds long on the coarse levels:
        # each (phase, buffer-role state) is captured once into a CUDA graph and replayed afterwards.
        self.use_cuda_graphs = use_cuda_graphs
        self._graph_cache = None
        # Full-multigrid start (reference solvers/advanced_multigrid.py:626-683): restrict f to every
        # level, solve the coarsest, then prolong + one cycle per level upwards.  Costs ~1.4 cycles (fp32 in the mixed
        # strategies) and starts the iteration at discretisation-error level; off by default so that cycle counts
        # stay comparable with the reference's zero-start runs.
        self.fmg = fmg
        # Helmholtz shift: solve (-lap + shift) u = f (implicit heat steps); 0 = the Poisson problem
        if shift < 0:
            raise ValueError("shift must be >= 0")
        self.shift = float(shift)
        self.enable_precision_monitoring = False
        self.precision_switches: List[Dict[str, Any]] = []
        self._engine: Optional[CycleEngine] = None
        self._shape = None

    # -- setup ----------------------------------------------------------------------------------------
    def _make_smoother(self):
        n = self.smoother_name.lower()
        if n in ("red_black_gauss_seidel", "rbgs", "red_black", "gauss_seidel_rb"):
            return GaussSeidelSmoother(relaxation_parameter=self.damping_factor, red_black=True)
        if n in ("gauss_seidel", "lexicographic"):
            return GaussSeidelSmoother(relaxation_parameter=self.damping_factor)
        if n in ("jacobi", "weighted_jacobi"):
            return JacobiSmoother(relaxation_parameter=2.0 / 3.0 if self.damping_factor == 1.0 else self.damping_factor)
        raise ValueError(f"Unknown smoother: {self.smoother_name}")

    def _levels_for(self, nx: int, ny: int) -> int:
        """Default: coarsen down to the 5x5 floor of the reference hierarchy (multigrid.py:158-160)."""
        if self.max_levels is not None:
            return self.max_levels
        L, a, b = 1, nx, ny
        while (a - 1) % 2 == 0 and (b - 1) % 2 == 0 and (a - 1) // 2 + 1 >= 5 and (b - 1) // 2 + 1 >= 5:
            a, b, L = (a - 1) // 2 + 1, (b - 1) // 2 + 1, L + 1
        return L

    def setup(self, nx: int, ny: int, domain=(0.0, 1.0, 0.0, 1.0)) -> None:
        if self.max_grid_size is not None and max(nx, ny) > self.max_grid_size:
            raise ValueError(f"grid {nx}x{ny} exceeds max_grid_size={self.max_grid_size}")
        dev = require_cuda(self.device)
        g = Grid(nx, ny, tuple(domain))
        grids = [g]
        for _ in range(1, self._levels_for(nx, ny)):
            try:
                c = grids[-1].coarsen()
            except ValueError:
                break
            if c.nx < 5 or c.ny < 5:
                break
            grids.append(c)
        # coefficient -1: the convergent sign convention
        op = HelmholtzOperator(-1.0, self.shift) if self.shift else LaplacianOperator(-1.0)
        L = len(grids)
        self._engine = CycleEngine(
            grids, smoother=self._make_smoother(),
            coarse_solver=GaussSeidelSmoother(max_iterations=self.coarse_max_iterations, tolerance=self.coarse_tolerance),
            operators=[op] * L, restriction_ops=[RestrictionOperator("full_weighting")] * (L - 1),
            prolongation_ops=[ProlongationOperator("bilinear")] * (L - 1), cycle_type=self.cycle_type, pre=self.pre,
            post=self.post, kernels=self.kernels, loader=self.loader, device=dev)
        self._shape, self._domain, self._grid = (nx, ny), tuple(domain), g
        self._sumsq = torch.zeros(2, dtype=torch.float64, device=dev)
        self._pinned_out = None
        eng = self._engine
        # allocate every (level, dtype) buffer the chosen strategy touches now, so the buffer-role state that keys
        # the CUDA graphs has its final shape before the first step
        L = eng.num_levels
        plans = {"fp64": [[torch.float64] * L], "fp32": [[torch.float32] * L]}.get(
            self.mode, [[torch.float64] * L, [torch.float32] * (L - 1) + [torch.float64]])
        for plan in plans:
            for lv, dt in zip(eng.levels, plan):
                lv.bufs(dt)
        self._graph_cache = GraphCache(eng.buffer_state, eng.snapshot_roles, eng.restore_roles, self.use_cuda_graphs)
        # fp64 iterate / rhs of the refinement phase live in the engine's fp64 level-0 buffers

    # -- CUDA-graph replay of one step (solvers/graphs.py) ------------------------------------------------------
    def _graphed(self, name: str, launch) -> None:
        self._graph_cache.enabled = self.use_cuda_graphs
        self._graph_cache.run(name, launch)

    @property
    def _graphs(self):
        return self._graph_cache.entries

    def _norm_from(self, ss: torch.Tensor) -> float:
        return float(np.sqrt(self._grid.hx * self._grid.hy * ops.read_scalar(ss)))

    # -- one cycle in each precision phase -----------------------------------------------------------------
    def _launch_uniform_cycle(self, dtype) -> None:
        eng = self._engine
        fused = eng.cycle([dtype] * eng.num_levels, 0, None, sumsq_out=self._sumsq[0:1])
        if not fused:
            self._sumsq[0:1].copy_(eng.residual_sumsq_async(dtype))

    def _cycle_fp64(self) -> float:
        self._graphed("fp64", lambda: self._launch_uniform_cycle(torch.float64))
        return self._norm_from(self._sumsq[0:1])

    def _cycle_fp32_only(self) -> float:
        self._graphed("fp32", lambda: self._launch_uniform_cycle(torch.float32))
        return self._norm_from(self._sumsq[0:1])

    def _inner_dtypes(self):
        """fp32 on every level except the coarsest, which stays fp64 like the reference's per-level
        mixed mode (the coarsest level is never converted, multigrid.py:270-272; in fp32 its 1e-12
        stopping test is unreachable and every coarse solve would run all 1000 sweeps)."""
        L = self._engine.num_levels
        return [torch.float32] * (L - 1) + [torch.float64]

    def _refinement_residual(self, with_update: bool = False) -> float:
        """One HBM pass over the fp64 iterate: [u += e32] ; r32 = fp32(f - A u) ; fp64 h-scaled ||r||."""
        self._launch_refinement_residual(with_update)
        return self._norm_from(self._sumsq[1:2])

    def _launch_refinement_residual(self, with_update: bool) -> None:
        eng, g = self._engine, self._grid
        b64, b32 = eng.levels[0].bufs(torch.float64), eng.levels[0].bufs(torch.float32)
        ss = self._sumsq[1:2]
        if ops.vc_aligned(b64.u, b64.f, b64.tmp, b32.u, b32.f) and eng.kernels != "basic":
            if with_update:
                ops.vc_defect_pass(b64.u, b64.tmp, b64.f, g.hx, g.hy, e_in=b32.u, r_out=b32.f, sumsq_out=ss,
                                   loader=eng.loader, shift=self.shift)
                b64.u, b64.tmp = b64.tmp, b64.u
            else:
                ops.vc_defect_pass(b64.u, None, b64.f, g.hx, g.hy, r_out=b32.f, sumsq_out=ss, loader=eng.loader,
                                   shift=self.shift)
        else:  # strict basic kernels
            if with_update:
                ops.axpy_(1.0, b32.u, b64.u)
            ops.residual(b64.u, b64.f, g.hx, g.hy, -1.0, out=b64.tmp, shift=self.shift)
            ss.copy_(ops.sumsq_async(b64.tmp, slot=1))
            ops.cast(b64.tmp, torch.float32, out=b32.f)

    def _launch_refinement_cycle(self) -> None:
        self._engine.cycle(self._inner_dtypes(), 0, None, u_zero=True)
        self._launch_refinement_residual(True)

    def _cycle_refinement(self) -> float:
        """One fp32 cycle on A e = r32 (e0 = 0, never read), then u64 += e32 fused with the next residual."""
        self._graphed("refine", self._launch_refinement_cycle)
        return self._norm_from(self._sumsq[1:2])

    def _fmg_start(self, dtypes) -> None:
        """Level-0 iterate of `dtypes[0]` := full-multigrid approximation for the right-hand side in that level's f."""
        eng = self._engine
        L = eng.num_levels
        bufs = [eng.levels[l].bufs(dtypes[l]) for l in range(L)]
        for l in range(L - 1):
            ops.restrict(bufs[l].f, "full_weighting", out=bufs[l + 1].f)
        eng.cycle(dtypes, L - 1, None, u_zero=True)  # coarsest solve from zero
        for l in range(L - 2, -1, -1):
            bufs = [eng.levels[k].bufs(dtypes[k]) for k in range(L)]  # roles may have swapped
            ops.prolong(bufs[l + 1].u, "bilinear", out=bufs[l].u)
            eng.cycle(dtypes, l, None)

    # -- public API ------------------------------------------------------------------------------------------
    def _resolve_grid(self, problem, nx, ny):
        nx = nx or getattr(problem, "nx", None)
        ny = ny or getattr(problem, "ny", None) or nx
        rhs = getattr(problem, "rhs_array", None)
        if nx is None:
            if rhs is None:
                raise ValueError("grid size unknown: give PoissonProblem(..., nx=, ny=) or solve(problem, nx=, ny=)")
            nx, ny = rhs.shape
        domain = tuple(getattr(problem, "domain", (0.0, 1.0, 0.0, 1.0)))
        if self._engine is None or self._shape != (nx, ny) or self._domain != domain:
            self.setup(nx, ny, domain)
        return nx, ny, domain, rhs

    def _load_rhs(self, problem, rhs, domain, dst: torch.Tensor) -> bool:
        """Right-hand side of `problem` into the pitched fp64 device field `dst` (on the current stream).
        Returns True when the caller handed in host data (host in -> host out)."""
        eng, g = self._engine, self._grid
        was_np = True
        if rhs is not None:
            was_np = not (isinstance(rhs, torch.Tensor) and rhs.is_cuda)
            if isinstance(rhs, torch.Tensor):
                dst.copy_(rhs, non_blocking=True)  # pinned host tensors upload asynchronously
            else:
                dst.copy_(to_device(rhs, device=eng.dev)[0])
        elif getattr(problem, "device_mms", None) is not None:
            amp, kx, ky = problem.device_mms
            ops.fill_sinsin_(dst, domain, amp, kx, ky)
        else:
            f_host = np.asarray(problem.source_function(g.X, g.Y), dtype=np.float64)
            dst.copy_(to_device(f_host, device=eng.dev)[0])
        return was_np

    def _zero_ring(self, f: torch.Tensor) -> None:
        if not self.strict_reference_norm:
            f[0, :] = 0
            f[-1, :] = 0
            f[:, 0] = 0
            f[:, -1] = 0

    def _solve_device(self, have_guess: bool):
        """The cycle loop on the right-hand side / iterate already in the fp64 level-0 buffers.  Synchronises only
        the current stream (never the device), so copies on other streams keep flowing.
        Returns (device solution, partial info)."""
        eng = self._engine
        b64 = eng.levels[0].bufs(torch.float64)
        cur = torch.cuda.current_stream(eng.dev)
        history: List[float] = []
        precisions: List[str] = []
        self.precision_switches = []
        phase = {"fp64": "fp64", "fp32": "fp32"}.get(self.mode, "refine")
        if phase == "fp32":
            b32 = eng.levels[0].bufs(torch.float32)
            ops.cast(b64.f, torch.float32, out=b32.f)
            ops.cast(b64.u, torch.float32, out=b32.u)
        converged = False
        iteration = 0
        cur.synchronize()
        t_cycles = time.perf_counter()
        if self.fmg and not have_guess:
            if phase == "refine":  # fp32 FMG on the rounded right-hand side, result promoted to the fp64 iterate
                b32 = eng.levels[0].bufs(torch.float32)
                ops.cast(b64.f, torch.float32, out=b32.f)
                self._fmg_start(self._inner_dtypes())
                ops.cast(eng.levels[0].bufs(torch.float32).u, torch.float64, out=eng.levels[0].bufs(torch.float64).u)
            else:
                dt = torch.float64 if phase == "fp64" else torch.float32
                self._fmg_start([dt] * eng.num_levels)
        pending = self._refinement_residual() if phase == "refine" else None  # ||r(u_0)||
        for iteration in range(1, self.max_iterations + 1):
            if phase == "refine":
                norm = self._cycle_refinement()  # residual of the new iterate (also next cycle's rhs)
                precisions.append("mixed")
            elif phase == "fp64":
                norm = self._cycle_fp64()
                precisions.append("float64")
            else:
                norm = self._cycle_fp32_only()
                precisions.append("float32")
            history.append(norm)
            if self.verbose:
                print(f"cycle {iteration}: ||r|| = {norm:.3e} [{precisions[-1]}]")
            if norm < self.tolerance:
                converged = True
                break
            if phase == "refine":
                stagnating = (len(history) >= 3 and all(history[-k] > self.stagnation_ratio * history[-k - 1]
                                                        for k in (1, 2)))
                if (self.mode == "switch" and norm <= self.switch_threshold) or stagnating:
                    self.precision_switches.append({"iteration": iteration, "residual": norm, "from": "mixed",
                                                    "to": "float64",
                                                    "reason": "stagnation" if stagnating else "switch_threshold"})
                    phase = "fp64"
        cur.synchronize()
        t_solve = time.perf_counter() - t_cycles
        b64 = eng.levels[0].bufs(torch.float64)
        if phase == "fp32":
            u_dev = ops.cast(eng.levels[0].bufs(torch.float32).u, torch.float64, out=b64.tmp)
        else:
            u_dev = b64.u
        nx, ny = self._shape
        ratios = [history[k] / history[k - 1] for k in range(max(1, len(history) - 4), len(history))
                  if history[k - 1] > 0 and 0 < history[k] / history[k - 1] < 1]
        info = {
            "converged": converged, "iterations": iteration, "final_residual": history[-1] if history else pending,
            "residual": history[-1] if history else pending, "residual_history": history,
            "initial_residual": pending, "convergence_rate": float(np.mean(ratios)) if ratios else 0.0,
            "cycle_time": t_solve, "average_time_per_iteration": t_solve / max(1, iteration),
            "precision_history": precisions, "precision_levels_used": sorted(set(precisions)),
            "precision_switches": list(self.precision_switches), "precision_strategy": self.precision_strategy,
            "switch_threshold": self.switch_threshold, "fmg": self.fmg, "cycle_type": self.cycle_type,
            "num_levels": eng.num_levels, "grid_hierarchy": [(l.grid.nx, l.grid.ny) for l in eng.levels],
            "level_timings": {}, "pre_smooth_iterations": self.pre, "post_smooth_iterations": self.post,
            "unknowns_per_second": nx * ny * iteration / t_solve if t_solve > 0 else 0.0,
        }
        return u_dev, info

    def solve(self, problem, initial_guess=None, nx: Optional[int] = None, ny: Optional[int] = None
              ) -> Tuple[Any, Dict[str, Any]]:
        t_start = time.perf_counter()
        nx, ny, domain, rhs = self._resolve_grid(problem, nx, ny)
        eng = self._engine
        b64 = eng.levels[0].bufs(torch.float64)
        was_np = self._load_rhs(problem, rhs, domain, b64.f)
        self._zero_ring(b64.f)
        if initial_guess is None:
            b64.u.zero_()
        else:
            b64.u.copy_(to_device(initial_guess, device=eng.dev, dtype=torch.float64)[0])
        t_setup = time.perf_counter() - t_start
        u_dev, info = self._solve_device(initial_guess is not None)
        if was_np:
            # device -> pinned host staging (kept across solves; the returned array is a view of it and is
            # overwritten by the next solve of this solver object)
            if self._pinned_out is None or tuple(self._pinned_out.shape) != (nx, ny):
                self._pinned_out = torch.empty((nx, ny), dtype=torch.float64, pin_memory=True)
            self._pinned_out.copy_(u_dev, non_blocking=True)
            torch.cuda.current_stream(eng.dev).synchronize()
            solution = self._pinned_out.numpy()
        else:
            solution = u_dev.clone()
        total = time.perf_counter() - t_start
        info.update({"solve_time": total, "setup_time": t_setup, "total_time": total})
        return solution, info

    def solve_many(self, problems, outputs=None) -> Tuple[List[Any], List[Dict[str, Any]]]:
        """A batch of solves on ONE grid (many right-hand sides: time steps, parameter sweeps), software-pipelined
        over the copy engines: while solve k cycles, the right-hand side of solve k+1 is uploaded on its own stream
        and the solution of solve k-1 is downloaded on another, so a batch is bound by max(H2D, cycles, D2H) per
        solve instead of their sum (PCIe moves 2 x 8 bytes per unknown; a 16385^2 solve is ~27 ms of cycles between
        two ~40 ms transfers).  `problems`: PoissonProblem objects (or anything `solve` accepts) sharing grid size and
        domain; right-hand sides given as PINNED host tensors upload asynchronously.  `outputs`: optional list of
        host tensors (pinned for overlap) receiving the solutions; allocated pinned when omitted.  Each solve starts
        from u = 0 and is bit-identical to `solve(problem)`.  Returns ([solutions as NumPy views], [info])."""
        problems = list(problems)
        if not problems:
            return [], []
        t_start = time.perf_counter()
        nx, ny, domain, _ = self._resolve_grid(problems[0], None, None)
        eng = self._engine
        dev = eng.dev
        n = len(problems)
        if outputs is None:
            outputs = [torch.empty((nx, ny), dtype=torch.float64, pin_memory=True) for _ in range(n)]
        elif len(outputs) != n or any(tuple(o.shape) != (nx, ny) or o.dtype != torch.float64 for o in outputs):
            raise ValueError("solve_many: `outputs` must be one float64 (nx, ny) host tensor per problem")
        if getattr(self, "_stage", None) is None or self._stage[0][0].shape != (nx, ny):
            self._stage = ([empty_field(nx, ny, torch.float64, dev) for _ in range(2)],   # right-hand sides
                           [empty_field(nx, ny, torch.float64, dev) for _ in range(2)],   # solutions
                           torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        f_stage, u_stage, s_in, s_out = self._stage
        cur = torch.cuda.current_stream(dev)
        consumed = [None, None]   # event: the compute stream has copied f_stage[slot] into the engine
        drained = [None, None]    # event: u_stage[slot] has reached the host
        uploaded: List[Any] = [None] * n

        def upload(k):
            slot = k % 2
            pk = problems[k]
            kx, ky, kd, rhs = (getattr(pk, "nx", None) or nx, getattr(pk, "ny", None) or ny,
                               tuple(getattr(pk, "domain", domain)), getattr(pk, "rhs_array", None))
            if (kx, ky) != (nx, ny) or tuple(kd) != tuple(domain):
                raise ValueError("solve_many: every problem must share the grid size and domain of the first")
            with torch.cuda.stream(s_in):
                if consumed[slot] is not None:
                    s_in.wait_event(consumed[slot])
                self._load_rhs(pk, rhs, domain, f_stage[slot])
                ev = torch.cuda.Event()
                ev.record(s_in)
            uploaded[k] = ev

        upload(0)
        infos = []
        for k in range(n):
            slot = k % 2
            if k + 1 < n:
                upload(k + 1)
            t0 = time.perf_counter()
            b64 = eng.levels[0].bufs(torch.float64)
            cur.wait_event(uploaded[k])
            b64.f.copy_(f_stage[slot])
            ev = torch.cuda.Event()
            ev.record(cur)
            consumed[slot] = ev
            self._zero_ring(b64.f)
            b64.u.zero_()
            u_dev, info = self._solve_device(False)
            if drained[slot] is not None:
                cur.wait_event(drained[slot])
            u_stage[slot].copy_(u_dev)
            ready = torch.cuda.Event()
            ready.record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ready)
                outputs[k].copy_(u_stage[slot], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_out)
            drained[slot] = ev
            info.update({"solve_time": time.perf_counter() - t0, "setup_time": 0.0})
            infos.append(info)
        s_out.synchronize()
        total = time.perf_counter() - t_start
        for info in infos:
            info["total_time"] = total / n
            info["batch_time"] = total
        return [o.numpy() for o in outputs], infos

    def release_staging(self) -> None:
        """Free the double-buffered device staging fields and side streams `solve_many` keeps between calls
        (4 fields of the grid size: 8.6 GB at 16385^2)."""
        self._stage = None
        self._pinned_out = None


# doc-only aliases seen in the reference notebooks
MixedPrecisionMultigridSolver = MixedPrecisionMultigrid
