"""``Parameter`` / ``Parameters`` with the reference's value <-> factor contract (modeling/parameter.py).

``value = inverse_transform(factor) = scale_transform^-1(scale * factor)`` (:548-574), ``autoscale`` with
``scale10`` / ``factor1`` (:524-546), ``Parameters.set_parameter_factors`` (:850-859).  Units are labels
only: values are stored in the package's fixed unit system (deg, TeV, cm-2 s-1 TeV-1).
"""
import collections.abc
import copy
import itertools

import numpy as np

from ..utils import parse_quantity

__all__ = ["Parameter", "Parameters"]


def _scale_fn(name):
    if name == "lin":
        return (lambda v: v), (lambda v: v)
    if name == "log":
        tiny = np.finfo(float).tiny
        return (lambda v: np.log(np.clip(v, tiny, np.inf))), np.exp
    if name == "sqrt":
        return np.sqrt, (lambda v: np.power(v, 2))
    raise ValueError(f"Not a valid value scaling mode: '{name}'.")


def bind_mirror(parameters):
    """A float64 vector that the ``value`` / ``factor`` setters of ``parameters`` keep equal to the parameter values from
    now on: a likelihood engine reads its parameter vector from it instead of walking hundreds of Parameter objects on
    every evaluation of a fit (16 us of a 300 us step at 415 parameters)."""
    vec = np.array([p._value for p in parameters], dtype=float)
    for i, p in enumerate(parameters):
        p._mirrors = p._mirrors + ((vec, i),)
    return vec


def store_covariance(parameters, matrix):
    """Remember the covariance of ``parameters`` (the free parameters of a fit, value space) on the parameter objects:
    each keeps its row and weak references to the partners, so that any ``Models`` / ``Parameters`` view over them can
    assemble its matrix later (``covariance_of``), as the reference pushes sub-covariances down to the models
    (modeling/models/core.py ``DatasetModels.covariance`` setter, modeling/covariance.py)."""
    import weakref

    matrix = np.asarray(matrix, dtype=float)
    refs = tuple(weakref.ref(p) for p in parameters)
    for i, p in enumerate(parameters):
        p._cov_partners = refs
        p._cov_row = matrix[i].copy()


def covariance_of(parameters):
    """Covariance matrix over ``parameters`` (duplicates of a linked parameter get identical rows): the stored rows where
    both partners belong to the list, ``error ** 2`` on the diagonal otherwise, 0 for pairs never fitted together."""
    parameters = list(parameters)
    n = len(parameters)
    out = np.zeros((n, n))
    where = {}
    for j, q in enumerate(parameters):
        where.setdefault(id(q), []).append(j)
    for i, p in enumerate(parameters):
        partners, row = getattr(p, "_cov_partners", None), getattr(p, "_cov_row", None)
        filled = False
        if partners is not None:
            for ref, value in zip(partners, row):
                q = ref()
                if q is not None and id(q) in where:
                    for j in where[id(q)]:
                        out[i, j] = value
                    filled = filled or q is p
        if not filled:
            for j in where[id(p)]:
                out[i, j] = p.error ** 2
    return out


def release_mirror(parameters, vec):
    for p in parameters:
        p._mirrors = tuple(m for m in p._mirrors if m[0] is not vec)


class Parameter:
    """A model parameter (modeling/parameter.py:40-600)."""

    _mirrors = ()  # (vector, index) pairs of the engines this parameter is bound to (bind_mirror)
    _link_label_io = None  # serialisation label of a parameter shared between models (modeling/io.py)

    def __getstate__(self):  # copies and pickles are not bound to anybody's vector
        state = dict(self.__dict__)
        state.pop("_mirrors", None)
        state.pop("_cov_partners", None)  # (weak references: a copy keeps its error, not its correlations)
        state.pop("_cov_row", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)

    def __init__(self, name, value, unit="", scale=1, min=np.nan, max=np.nan, frozen=False, error=0,
                 scan_min=None, scan_max=None, scan_n_values=11, scan_n_sigma=2, scan_values=None,
                 scale_method="scale10", interp="lin", prior=None, scale_transform="lin"):
        if not isinstance(name, str):
            raise TypeError(f"Name must be string, got '{type(name)}' instead")
        self._name = name
        self._scale_method = scale_method
        self._scale_transform = scale_transform
        self.interp = interp
        self._scale = float(scale)
        self.frozen = frozen
        self._error = float(error)
        self._type = None
        if isinstance(value, str) or hasattr(value, "unit"):
            v, parsed_unit = parse_quantity(value)
            self.value = v
            self.unit = parsed_unit or unit
        else:
            self.value = float(value)
            self.unit = unit
        self.min = min
        self.max = max
        self.scan_min = scan_min
        self.scan_max = scan_max
        self.scan_values = scan_values
        self.scan_n_values = scan_n_values
        self.scan_n_sigma = int(scan_n_sigma)
        self.prior = prior

    # every assignment of a prior bumps a class-wide counter: consumers (the likelihood engines) cache "does any of
    # my parameters carry a prior" against it instead of scanning their parameters on every evaluation
    prior_epoch = 0

    @property
    def prior(self):
        return self._prior

    @prior.setter
    def prior(self, value):
        self._prior = value
        Parameter.prior_epoch += 1

    # descriptor protocol: class attributes on models are templates, instances hold their own copy
    def __get__(self, instance, owner):
        if instance is None:
            return self
        par = instance.__dict__[self.name]
        par._type = getattr(instance, "type", None)
        return par

    def __set__(self, instance, value):
        if isinstance(value, Parameter):
            instance.__dict__[self.name] = value
        else:
            par = instance.__dict__[self.name]
            raise TypeError(f"Cannot assign {value!r} to parameter {par!r}")

    def __set_name__(self, owner, name):
        if not self._name == name:
            raise ValueError(f"Expected parameter name '{name}', got {self._name}")

    @property
    def name(self):
        return self._name

    @property
    def type(self):
        return self._type

    @property
    def error(self):
        return self._error

    @error.setter
    def error(self, value):
        self._error = float(parse_quantity(value)[0])

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, val):
        # the reference computes the factor here (parameter.py:229-233); it is computed when first read instead, with the
        # same scale (the scale only changes through autoscale / reset_autoscale, which set the factor themselves): an
        # optimiser step writes hundreds of values whose factors nobody reads
        self._value = v = float(val)
        self._factor = None
        for vec, i in self._mirrors:
            vec[i] = v

    @property
    def quantity(self):
        return self._value

    @quantity.setter
    def quantity(self, val):
        self.value = parse_quantity(val)[0]

    @property
    def factor(self):
        f = self._factor
        if f is None:
            f = self._factor = self.transform(self._value)
        return f

    @factor.setter
    def factor(self, val):
        self._factor = float(val)
        self._value = v = float(self.inverse_transform(self._factor))
        for vec, i in self._mirrors:
            vec[i] = v

    @property
    def scale(self):
        return self._scale

    @property
    def scale_method(self):
        return self._scale_method

    @scale_method.setter
    def scale_method(self, val):
        if val not in ["scale10", "factor1"] and val is not None:
            raise ValueError(f"Invalid method: {val}")
        self.reset_autoscale()
        self._scale_method = val

    @property
    def scale_transform(self):
        return self._scale_transform

    @property
    def min(self):
        return self._min

    @min.setter
    def min(self, val):
        self._min = np.nan if val is None else float(parse_quantity(val)[0])

    @property
    def max(self):
        return self._max

    @max.setter
    def max(self, val):
        self._max = np.nan if val is None else float(parse_quantity(val)[0])

    @property
    def factor_min(self):
        return self.transform(self.min)

    @property
    def factor_max(self):
        return self.transform(self.max)

    @property
    def frozen(self):
        return self._frozen

    @frozen.setter
    def frozen(self, val):
        if not isinstance(val, (bool, np.bool_)):
            raise TypeError(f"Invalid type: {val}, {type(val)}")
        self._frozen = bool(val)

    @property
    def _step(self):
        return self.error if self.error > 0.0 else np.abs(self.value)

    @property
    def conf_min(self):
        """Lower end of the confidence search (parameter.py:385-397): the minimum if set, else a wide default."""
        if not np.isnan(self.min):
            return self.min
        min_ = self.value - self._step * self.scan_n_sigma
        large_step = np.maximum(self._step, np.abs(self.value))
        return np.minimum(min_, -large_step * 1e5)

    @property
    def conf_max(self):
        """Upper end of the confidence search (parameter.py:401-413)."""
        if not np.isnan(self.max):
            return self.max
        max_ = self.value + self._step * self.scan_n_sigma
        large_step = np.maximum(self._step, np.abs(self.value))
        return np.maximum(max_, large_step * 1e5)

    @property
    def scan_min(self):
        if self._scan_min is None:
            return self.value - self._step * self.scan_n_sigma
        return self._scan_min

    @scan_min.setter
    def scan_min(self, value):
        self._scan_min = value

    @property
    def scan_max(self):
        if self._scan_max is None:
            return self.value + self._step * self.scan_n_sigma
        return self._scan_max

    @scan_max.setter
    def scan_max(self, value):
        self._scan_max = value

    @property
    def scan_values(self):
        """modeling/parameter.py:449-458."""
        if self._scan_values is None:
            fwd, inv = _scale_fn(self.interp)
            parmin, parmax = fwd(np.array([self.scan_min, self.scan_max], dtype=float))
            return inv(np.linspace(parmin, parmax, self.scan_n_values))
        return self._scan_values

    @scan_values.setter
    def scan_values(self, values):
        self._scan_values = values

    def update_scale(self, value):
        """modeling/parameter.py:524-546."""
        if self.scale_method == "scale10":
            if value != 0:
                exponent = np.floor(np.log10(np.abs(value)))
                self._scale = float(np.power(10.0, exponent))
        elif self.scale_method == "factor1":
            self._scale = float(value)

    def transform(self, value, update_scale=False):
        """value -> factor (modeling/parameter.py:548-562)."""
        fwd, _ = _scale_fn(self._scale_transform)
        transformed = fwd(value)
        if update_scale:
            self.update_scale(transformed)
        return float(transformed / self.scale)

    def inverse_transform(self, factor):
        """factor -> value (modeling/parameter.py:564-574)."""
        _, inv = _scale_fn(self._scale_transform)
        return inv(self.scale * factor)

    def autoscale(self):
        self.factor = self.transform(self.value, update_scale=True)

    def reset_autoscale(self):
        self._factor = self._value
        self._scale = 1.0

    def check_limits(self):
        if not self.frozen:
            if (not np.isnan(self.min) and self.value < self.min) or (not np.isnan(self.max) and self.value > self.max):
                import logging

                logging.getLogger(__name__).warning(
                    f"Value {self.value} is outside bounds [{self.min}, {self.max}] for parameter '{self.name}'")

    def copy(self):
        return copy.deepcopy(self)

    def __repr__(self):
        return (f"Parameter(name={self.name!r}, value={self.value!r}, factor={self.factor!r}, scale={self.scale!r}, "
                f"unit={self.unit!r}, min={self.min!r}, max={self.max!r}, frozen={self.frozen!r})")


class Parameters(collections.abc.Sequence):
    """List of unique ``Parameter`` objects (modeling/parameter.py:600-900)."""

    def __init__(self, parameters=None):
        self._parameters = list(parameters) if parameters is not None else []

    @classmethod
    def from_stack(cls, parameters_list):
        pars = itertools.chain(*[_._parameters for _ in parameters_list])
        return cls(pars)

    @property
    def unique_parameters(self):
        seen, out = set(), []
        for p in self._parameters:
            if id(p) not in seen:
                seen.add(id(p))
                out.append(p)
        return Parameters(out)

    @property
    def free_parameters(self):
        return Parameters([p for p in self._parameters if not p.frozen])

    @property
    def free_unique_parameters(self):
        return self.unique_parameters.free_parameters

    @property
    def names(self):
        return [p.name for p in self._parameters]

    @property
    def value(self):
        return np.array([p.value for p in self._parameters], dtype=float)

    @value.setter
    def value(self, values):
        if len(self) != len(values):
            raise ValueError("Values must be of the same length as the parameters.")
        for p, v in zip(self._parameters, values):
            p.value = v

    @property
    def min(self):
        return np.array([p.min for p in self._parameters], dtype=float)

    @property
    def max(self):
        return np.array([p.max for p in self._parameters], dtype=float)

    def index(self, val):
        if isinstance(val, int):
            return val
        if isinstance(val, Parameter):
            for i, p in enumerate(self._parameters):
                if p is val:
                    return i
            raise IndexError(f"No parameter: {val!r}")
        if isinstance(val, str):
            for i, p in enumerate(self._parameters):
                if p.name == val:
                    return i
            raise IndexError(f"No parameter: {val!r}")
        raise TypeError(f"Invalid type: {type(val)!r}")

    def __getitem__(self, key):
        if isinstance(key, (np.ndarray, list)):
            arr = np.asarray(key)
            if arr.dtype == bool:  # boolean mask, as the reference's Parameters.__getitem__ on a numpy mask
                if arr.shape != (len(self._parameters),):
                    raise IndexError(f"boolean mask of length {arr.size} for {len(self._parameters)} parameters")
                return Parameters([p for p, keep in zip(self._parameters, arr) if keep])
            return Parameters([self[k if isinstance(k, str) else int(k)] for k in key])
        if isinstance(key, slice):
            return Parameters(self._parameters[key])
        return self._parameters[self.index(key)]

    def __len__(self):
        return len(self._parameters)

    def __add__(self, other):
        if isinstance(other, Parameters):
            return Parameters.from_stack([self, other])
        raise TypeError(f"Invalid type: {other!r}")

    def set_parameter_factors(self, factors):
        """modeling/parameter.py:850-859."""
        idx = 0
        for parameter in self._parameters:
            if not parameter.frozen:
                parameter.factor = factors[idx]
                idx += 1

    def autoscale(self):
        for par in self._parameters:
            par.autoscale()

    def check_limits(self):
        for par in self._parameters:
            par.check_limits()

    def prior_stat_sum(self):
        """Sum of the prior terms of the parameters that carry one (parameter.py:632-637)."""
        total = 0.0
        for par in self.unique_parameters:  # a linked parameter counts once
            if par.prior is not None:
                total += par.prior(par)
        return total

    def freeze_all(self):
        for p in self._parameters:
            p.frozen = True

    def unfreeze_all(self):
        for p in self._parameters:
            p.frozen = False

    def copy(self):
        return copy.deepcopy(self)

    def to_table(self):
        rows = [dict(name=p.name, value=p.value, unit=p.unit, error=p.error, min=p.min, max=p.max, frozen=p.frozen)
                for p in self._parameters]
        return rows

    def __repr__(self):
        return "Parameters([\n  " + ",\n  ".join(repr(p) for p in self._parameters) + "\n])"
