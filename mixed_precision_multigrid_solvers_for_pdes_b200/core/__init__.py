from .grid import Grid
from .precision import PrecisionLevel, PrecisionManager

__all__ = ["Grid", "PrecisionManager", "PrecisionLevel"]
