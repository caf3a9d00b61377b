"""Precision-switching rules, API-compatible with ``multigrid.core.precision`` of the reference
(core/precision.py:11-418).  Pure host logic: it decides WHICH dtype a level / an iteration
runs in; the casts themselves are ``mg_cast`` launches.

The thresholds are the reference's: DOUBLE->SINGLE while ``residual > 100*threshold`` or the
4-arrays-per-level memory estimate exceeds ``memory_threshold_gb`` (precision.py:155-187);
SINGLE->DOUBLE once ``residual < 10*threshold`` (precision.py:248-268); MIXED = finer half of
the levels fp64, coarser half fp32 (precision.py:337-357); stagnation promotion over the last
five residuals (precision.py:189-246)."""
from __future__ import annotations

from enum import Enum
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np


class PrecisionLevel(Enum):
    SINGLE = "float32"
    DOUBLE = "float64"
    MIXED = "mixed"


_NAMES = {"single": PrecisionLevel.SINGLE, "float32": PrecisionLevel.SINGLE, "double": PrecisionLevel.DOUBLE,
          "float64": PrecisionLevel.DOUBLE, "mixed": PrecisionLevel.MIXED}
_DTYPES = {PrecisionLevel.SINGLE: np.float32, PrecisionLevel.DOUBLE: np.float64, PrecisionLevel.MIXED: np.float64}


class PrecisionManager:
    def __init__(self, default_precision: Union[PrecisionLevel, str] = PrecisionLevel.DOUBLE, adaptive: bool = True,
                 convergence_threshold: float = 1e-6, memory_threshold_gb: float = 4.0):
        self.default_precision = self._parse_precision(default_precision)
        self.adaptive = adaptive
        self.convergence_threshold = convergence_threshold
        self.memory_threshold_bytes = memory_threshold_gb * 1024 ** 3
        self.precision_hierarchy = [PrecisionLevel.SINGLE, PrecisionLevel.DOUBLE]
        self.current_precision = self.default_precision
        self.precision_history: List[PrecisionLevel] = [self.current_precision]
        self.precision_stats = {p: {"operations": 0, "time": 0.0} for p in self.precision_hierarchy}

    @staticmethod
    def _parse_precision(p) -> PrecisionLevel:
        if isinstance(p, PrecisionLevel):
            return p
        if isinstance(p, str):
            try:
                return _NAMES[p.lower()]
            except KeyError:
                raise ValueError(f"Unknown precision level: {p}")
        raise TypeError(f"Precision must be PrecisionLevel or str, got {type(p)}")

    def get_dtype(self, precision: Optional[PrecisionLevel] = None):
        return _DTYPES[self.current_precision if precision is None else precision]

    def convert_array(self, array, target_precision: Optional[PrecisionLevel] = None):
        """astype to the target precision; NumPy arrays and torch tensors both supported
        (device tensors are converted by ``mg_cast``)."""
        target = self.get_dtype(target_precision)
        if isinstance(array, np.ndarray):
            return array if array.dtype == target else array.astype(target)
        from ..ops import cast
        return cast(array, target)

    def estimate_memory_usage(self, grid_shapes: Sequence) -> float:
        points = sum(nx * ny for nx, ny in grid_shapes)
        return points * np.dtype(self.get_dtype()).itemsize * 4  # 4 arrays per level

    def should_downgrade_precision(self, grid_shapes: Sequence, residual_norm: float) -> bool:
        if not self.adaptive:
            return False
        if self.estimate_memory_usage(grid_shapes) > self.memory_threshold_bytes:
            return True
        return self.current_precision == PrecisionLevel.DOUBLE and residual_norm > self.convergence_threshold * 100

    def should_upgrade_precision(self, residual_norm: float) -> bool:
        if not self.adaptive:
            return False
        return self.current_precision == PrecisionLevel.SINGLE and residual_norm < self.convergence_threshold * 10

    def should_promote_precision(self, convergence_history: Sequence[float], current_precision: PrecisionLevel) -> bool:
        if not self.adaptive or current_precision == PrecisionLevel.DOUBLE or len(convergence_history) < 5:
            return False
        r = list(convergence_history[-5:])
        ratios = [r[k] / r[k - 1] for k in range(1, 5) if r[k - 1] > 0]
        if ratios:
            if np.mean(ratios) > 0.9:  # stagnation
                return True
            rel = [abs(r[k] - r[k - 1]) / r[k - 1] for k in range(1, 5) if r[k - 1] > 0]
            if rel and np.mean(rel) < 1e-3:  # plateau
                return True
        return all(r[k] >= r[k - 1] * 0.99 for k in range(1, 5))  # growth

    def update_precision(self, residual_norm: float, grid_shapes: Optional[Sequence] = None) -> bool:
        if not self.adaptive:
            return False
        old = self.current_precision
        if grid_shapes and self.should_downgrade_precision(grid_shapes, residual_norm):
            if self.current_precision == PrecisionLevel.DOUBLE:
                self.current_precision = PrecisionLevel.SINGLE
        elif self.should_upgrade_precision(residual_norm):
            if self.current_precision == PrecisionLevel.SINGLE:
                self.current_precision = PrecisionLevel.DOUBLE
        if self.current_precision != old:
            self.precision_history.append(self.current_precision)
            return True
        return False

    def optimal_precision_per_level(self, grid_level: int, problem_size: int) -> PrecisionLevel:
        if not self.adaptive:
            return self.current_precision
        if grid_level == 0:
            return PrecisionLevel.DOUBLE
        if grid_level <= 2:
            return PrecisionLevel.SINGLE if problem_size > 500000 else PrecisionLevel.DOUBLE
        return PrecisionLevel.SINGLE

    def get_precision_for_level(self, level: int, max_levels: int) -> PrecisionLevel:
        if not self.adaptive or self.current_precision != PrecisionLevel.MIXED:
            return self.current_precision
        return PrecisionLevel.SINGLE if level >= max_levels // 2 else PrecisionLevel.DOUBLE

    def get_statistics(self) -> Dict[str, Any]:
        ops = sum(s["operations"] for s in self.precision_stats.values())
        return {
            "current_precision": self.current_precision.value,
            "precision_history": [p.value for p in self.precision_history],
            "total_operations": ops,
            "total_time": sum(s["time"] for s in self.precision_stats.values()),
            "precision_breakdown": {
                p.value: {"operations": s["operations"], "time": s["time"],
                          "percentage": (s["operations"] / ops * 100 if ops > 0 else 0)}
                for p, s in self.precision_stats.items()},
        }

    def record_operation(self, precision: PrecisionLevel, time_taken: float) -> None:
        if precision in self.precision_stats:
            self.precision_stats[precision]["operations"] += 1
            self.precision_stats[precision]["time"] += time_taken

    def reset_statistics(self) -> None:
        for s in self.precision_stats.values():
            s["operations"], s["time"] = 0, 0.0
        self.precision_history = [self.current_precision]

    def __repr__(self) -> str:
        return (f"PrecisionManager(default={self.default_precision.value}, current={self.current_precision.value}, "
                f"adaptive={self.adaptive}, threshold={self.convergence_threshold})")

    __str__ = __repr__
