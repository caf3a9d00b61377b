from .base import BaseOperator
from .laplacian import LaplacianOperator
from .transfer import ProlongationOperator, RestrictionOperator

__all__ = ["BaseOperator", "LaplacianOperator", "RestrictionOperator", "ProlongationOperator"]
