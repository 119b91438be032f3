"""Priors on parameters (modeling/models/prior.py in the reference): ``-2 log pdf`` terms added to the joint
statistic by ``Datasets.stat_sum`` (datasets/core.py:264-277) and therefore seen by ``Fit``.

They are scalar functions of single parameters, evaluated on the host (vectorised over the parameter vectors of a
batched call); the likelihood itself stays on the GPU."""
import numpy as np

__all__ = ["Prior", "GaussianPrior", "UniformPrior"]


class Prior:
    """Base class: ``prior(parameter)`` returns ``weight * evaluate(parameter.value)`` (prior.py:120-123)."""

    tag = ["Prior"]

    def __init__(self, weight=1):
        self.weight = weight

    def evaluate(self, value):
        raise NotImplementedError

    def __call__(self, parameter):
        return self.weight * float(self.evaluate(np.float64(parameter.value)))

    def evaluate_many(self, values):
        """The same for an array of parameter values (one per parameter vector of a batched call)."""
        return self.weight * np.asarray(self.evaluate(np.asarray(values, dtype=float)), dtype=float)


class GaussianPrior(Prior):
    """``-2 log N(value; mu, sigma)`` (prior.py:156-183; scipy.stats.norm.logpdf written out)."""

    tag = ["GaussianPrior"]

    def __init__(self, mu=0.0, sigma=1.0, weight=1):
        super().__init__(weight)
        self.mu, self.sigma = float(mu), float(sigma)

    def evaluate(self, value):
        z = (value - self.mu) / self.sigma
        return z * z + 2.0 * np.log(self.sigma) + np.log(2.0 * np.pi)


class UniformPrior(Prior):
    """``2 log(max - min)`` inside ``[min, max]``, ``inf`` outside (prior.py:186-216)."""

    tag = ["UniformPrior"]

    def __init__(self, min=0.0, max=1.0, weight=1):
        super().__init__(weight)
        self.min, self.max = float(min), float(max)

    def evaluate(self, value):
        inside = (value >= self.min) & (value <= self.max)
        with np.errstate(divide="ignore"):
            return np.where(inside, 2.0 * np.log(self.max - self.min), np.inf)
