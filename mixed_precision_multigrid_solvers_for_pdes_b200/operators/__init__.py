from .base import BaseOperator
from .laplacian import HelmholtzOperator, LaplacianOperator
from .transfer import ProlongationOperator, RestrictionOperator

__all__ = ["BaseOperator", "LaplacianOperator", "HelmholtzOperator", "RestrictionOperator", "ProlongationOperator"]
from .variable import VariableCoefficientOperator, VariableCoefficientSmoother  # noqa: E402

__all__ += ["VariableCoefficientOperator", "VariableCoefficientSmoother"]
