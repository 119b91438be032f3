"""Modeling layer: parameters, the models of the accelerated path, and the optimiser front end."""
from .parameter import Parameter, Parameters  # noqa: F401
from .fit import Fit, FitResult  # noqa: F401
from .prior import GaussianPrior, Prior, UniformPrior  # noqa: F401
from . import io  # noqa: F401  (installs Models.write / read / to_yaml / from_yaml, *.to_dict / from_dict)
