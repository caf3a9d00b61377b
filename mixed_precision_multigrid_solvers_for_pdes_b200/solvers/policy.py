"""Host-side decisions of the precision-switching cycle loop (pure Python: no device code, CPU-testable).

One object states, for every driver that runs the loop (``MixedPrecisionMultigrid``, the row-slab
``DistributedMixedPrecisionSolver``, ``bench.py``), what happens after each cycle's residual norm:

  * stop, the reference's test: ``norm < tolerance`` (h-scaled L2 norm of f - A u, solvers/base.py:123-143);
  * switch from fp32-cycle / fp64-residual refinement to fp64 cycles at ``norm <= switch_threshold``
    (docs/methodology.md:337-360) or when the refinement stagnates above the floor (two consecutive ratios above
    ``stagnation_ratio``: core/precision.py:189-246 states the same idea over a window of 5) -- unless the
    tolerance lies below the rounding floor (next item), in which case the solve stays in the refinement;
  * stop at the ROUNDING FLOOR of the residual evaluation.  Evaluating f - A u in a floating-point type T
    cannot return less than about  eps_T * (2/hx^2 + 2/hy^2 + shift) * ||u||  (h-scaled norms): 1.2e-7 at
    h = 1/16384 in fp64, above the reference's absolute tolerance 1e-8, which then nobody can meet.  The rule:
    once the norm is below that a-priori bound AND has not contracted (``norm > floor_ratio * previous``) for
    ``floor_confirmations`` consecutive cycles, the residual no longer measures the algebraic error and the solve
    ends with ``stopped_on = "rounding_floor"``.  The cycles spent on detection are not wasted: a residual on its
    floor is blind to the algebraic error, which keeps contracting underneath it -- in the REFINEMENT phase.  Measured
    at 16385^2 (profiles/r02_floor_study_16385.json; closed-form discretisation error 3.0639e-9): the
    refinement cycle that lands on the floor (8) leaves the MMS error 2.4 % short, the first stagnating cycle (9)
    1.0 %, the confirming one (10) 0.06 %.  fp64 cycles do not share this: once their residual is on the floor the
    smooth error contracts by 0.8 instead of 0.05 per cycle (same file; the C restatement of the reference shows
    the same at 4097^2), which is why the switch to fp64 is skipped for such solves.  The bound is a-priori and a
    few times above the measured floor, so a tolerance between the two (8193^2: floor 7e-9, tolerance 1e-8, bound
    3e-8) costs two cycles more than round 1 and gains an MMS error of 0.01 % instead of 0.6 %.  Where the tolerance
    is attainable (every grid up to 4097^2 at 1e-8) none of this fires and cycle counts are the reference's.  The
    fp32-only strategy is exempt: like the reference's all-fp32 runs it floors far above any useful
    tolerance and runs to ``max_iterations``.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional

EPS = {"fp64": 2.220446049250313e-16, "fp32": 1.1920928955078125e-07}

CONTINUE, CONVERGED, FLOOR = "continue", "converged", "rounding_floor"


def residual_floor_bound(hx: float, hy: float, shift: float, u_norm: float, precision: str = "fp64") -> float:
    """A-priori bound of the rounding floor of ||f - A u|| (h-scaled L2) when evaluated in `precision`."""
    return EPS[precision] * (2.0 / hx ** 2 + 2.0 / hy ** 2 + shift) * u_norm


def weak_method(bound_method) -> Callable:
    """`bound_method` without keeping its object alive (see CyclePolicy.u_norm)."""
    import weakref
    ref = weakref.WeakMethod(bound_method)

    def call(*a, **k):
        fn = ref()
        if fn is None:
            raise ReferenceError("the solver this policy belonged to is gone")
        return fn(*a, **k)
    return call


class CyclePolicy:
    """mode: 'fp64' | 'fp32' | 'switch' | 'refine' (solvers/mixed_precision.py strategies).
    ``u_norm(phase)``: callable returning the h-scaled L2 norm of the current iterate; only called when a stagnating
    residual has to be compared with the floor bound (costs one reduction).  Hand in a plain function or a
    ``weak_method`` of the driver: a closure over the policy itself would tie driver and policy into a reference
    cycle that only the cyclic garbage collector frees -- at an arbitrary moment, possibly while a CUDA graph of
    another solver is being captured, which the destruction of this solver's graphs would then invalidate."""

    def __init__(self, mode: str, tolerance: float, switch_threshold: float, hx: float, hy: float, shift: float = 0.0,
                 stagnation_ratio: float = 0.95, floor_ratio: float = 0.5, stop_on_floor: bool = True,
                 u_norm: Optional[Callable[[], float]] = None, floor_confirmations: int = 2):
        if mode not in ("fp64", "fp32", "switch", "refine"):
            raise ValueError(f"unknown mode {mode!r}")
        self.mode, self.tolerance, self.switch_threshold = mode, tolerance, switch_threshold
        self.hx, self.hy, self.shift = hx, hy, shift
        self.stagnation_ratio, self.floor_ratio, self.stop_on_floor = stagnation_ratio, floor_ratio, stop_on_floor
        self.u_norm = u_norm
        self.floor_confirmations = max(1, int(floor_confirmations))
        self.switch_guard = 1.0  # the switch to fp64 cycles is skipped when tolerance < switch_guard * floor bound
        self.start()

    def start(self) -> str:
        self.phase = {"fp64": "fp64", "fp32": "fp32"}.get(self.mode, "refine")
        self.history: List[float] = []
        self.switches: List[Dict[str, Any]] = []
        self.stopped_on: Optional[str] = None
        self.floor_bound: Optional[float] = None
        self._floor_hits = 0
        self.switch_blocked: Optional[Dict[str, Any]] = None
        return self.phase

    def _at_floor(self, norm: float) -> bool:
        """True once the fp64-evaluated residual has sat on its rounding floor for `floor_confirmations` cycles."""
        if not self.stop_on_floor or len(self.history) < 2 or self.u_norm is None or self.phase == "fp32":
            return False
        if not norm > self.floor_ratio * self.history[-2]:
            self._floor_hits = 0
            return False
        if norm > self._bound():
            self._floor_hits = 0
            return False
        self._floor_hits += 1
        return self._floor_hits >= self.floor_confirmations

    def likely_last(self) -> bool:
        """A HINT for the driver of the fused defect + down pass: the cycle about to run will probably end the solve
        (its norm meets the tolerance if it contracts like the last one did, or it is the confirming cycle on the
        rounding floor), so pre-smoothing the NEXT error equation inside its defect pass would be wasted work.
        Wrong guesses cost nothing but that one saving: the driver then runs the pre-smoothing as its own pass."""
        h = self.history
        if self.phase != "refine" or not h:
            return False
        if self._floor_hits > 0 and self._floor_hits >= self.floor_confirmations - 1:
            return True
        if self.stop_on_floor and self.floor_bound is not None and self.tolerance < self.floor_bound:
            return False  # this solve ends on the rounding floor, not on the tolerance: wait for the floor hits
        return len(h) >= 2 and h[-2] > 0 and h[-1] * min(1.0, h[-1] / h[-2]) < self.tolerance

    def _bound(self) -> float:
        if self.floor_bound is None:
            self.floor_bound = residual_floor_bound(self.hx, self.hy, self.shift, self.u_norm(self.phase), "fp64")
        return self.floor_bound

    def observe(self, norm: float) -> str:
        """Record the norm of the cycle that just ran in ``self.phase``; returns CONTINUE, CONVERGED or FLOOR and
        updates ``self.phase`` for the next cycle."""
        self.history.append(norm)
        h = self.history
        if norm < self.tolerance:
            self.stopped_on = "tolerance"
            return CONVERGED
        if self._at_floor(norm):
            self.stopped_on = FLOOR
            return FLOOR
        if self.phase == "refine":
            if self._floor_hits > 0:         # on the floor, awaiting confirmation: fp64 cycles would not lower it
                return CONTINUE
            stagnating = len(h) >= 3 and all(h[-k] > self.stagnation_ratio * h[-k - 1] for k in (1, 2))
            at_threshold = self.mode == "switch" and norm <= self.switch_threshold
            if at_threshold and not stagnating and self.stop_on_floor and self.u_norm is not None \
                    and self.tolerance < self.switch_guard * self._bound():
                # The tolerance lies below the rounding floor: this solve will end on the floor, and there the
                # refinement is the better iteration.  fp64 cycles smooth the O(1) iterate itself, whose last bit
                # (2.2e-16) is coarser than the updates a floor-level residual asks for, and their contraction of the
                # smooth error drops from 0.05 to 0.8 per cycle (measured, also for the reference's own arithmetic:
                # profiles/r02_floor_study_*.json); the refinement smooths the CORRECTION at its own scale.
                if not self.switch_blocked:
                    self.switch_blocked = {"iteration": len(h), "residual": norm, "floor_bound": self.floor_bound}
                return CONTINUE
            if at_threshold or stagnating:
                self.switches.append({"iteration": len(h), "residual": norm, "from": "mixed", "to": "float64",
                                      "reason": "stagnation" if stagnating else "switch_threshold"})
                self.phase = "fp64"
        return CONTINUE
