"""Sky models of the hot path: 3 spectral x 3 spatial models, ``SkyModel``, ``FoVBackgroundModel``.

Reference: modeling/models/core.py (ModelBase), spectral.py (PowerLaw :1002-1067, PowerLawNorm :1134-1162,
ExpCutoffPowerLaw :1520-1567, LogParabola :1865-1908), spatial.py (Point :558-631, Gaussian :639-701,
Disk :898-993), cube.py (SkyModel :33, FoVBackgroundModel :622-756).

Model objects only carry parameters; bin integrals (``SpectralModel.integral``) and pixel integrals
(``SpatialModel.integrate_geom``) run on the GPU through the C ABI.  ``evaluate`` is the plain
differential form, kept for user code that samples a model at a few energies / positions.
"""
import copy
import ctypes as C
import itertools

import numpy as np

from ..utils import DEG, SkyCoord, angular_separation, position_angle
from .parameter import Parameter, Parameters

__all__ = [
    "Model", "SpectralModel", "SpatialModel", "PowerLawSpectralModel", "PowerLawNormSpectralModel",
    "LogParabolaSpectralModel", "ExpCutoffPowerLawSpectralModel", "PointSpatialModel", "GaussianSpatialModel",
    "DiskSpatialModel", "SkyModel", "FoVBackgroundModel", "Models", "DatasetModels",
]

_name_counter = itertools.count()


def make_name(name=None):
    if name is None:
        return f"model-{next(_name_counter):04d}"
    return str(name)


class ModelBase:
    """Parameter bookkeeping of modeling/models/core.py:ModelBase."""

    _type = None
    tag = ["ModelBase"]

    def __init__(self, **kwargs):
        self.default_parameters = Parameters(
            [v for k, v in itertools.chain(*[vars(c).items() for c in reversed(type(self).__mro__)])
             if isinstance(v, Parameter)])
        names = []
        for par in self.default_parameters:
            if par.name in names:
                continue
            names.append(par.name)
            value = kwargs.pop(par.name, None)
            new = par.copy()
            if isinstance(value, Parameter):
                new = value
            elif value is not None:
                new.quantity = value
            self.__dict__[par.name] = new
        self._par_names = names
        if kwargs:
            raise TypeError(f"{type(self).__name__}: unexpected arguments {sorted(kwargs)}")

    @property
    def type(self):
        return self._type

    @property
    def parameters(self):
        return Parameters([self.__dict__[n] for n in self._par_names])

    def copy(self, **kwargs):
        return copy.deepcopy(self)

    def freeze(self):
        self.parameters.freeze_all()

    def unfreeze(self):
        for p, d in zip(self.parameters, [self.default_parameters[n] for n in self._par_names]):
            p.frozen = d.frozen

    def __repr__(self):
        pars = ", ".join(f"{p.name}={p.value!r}" for p in self.parameters)
        return f"{type(self).__name__}({pars})"


Model = ModelBase


# ------------------------------------------------------------------------------------------------
# spectral
# ------------------------------------------------------------------------------------------------
class SpectralModel(ModelBase):
    _type = "spectral"
    _gpy_type = None

    def __call__(self, energy):
        return self.evaluate(np.asarray(energy, dtype=float), *[p.value for p in self.parameters])

    def integral(self, energy_min, energy_max, **kwargs):
        """``SpectralModel.integral`` (spectral.py:353-371) per bin, computed by the CUDA kernel K0."""
        from .._lib import Context, check

        emin = np.atleast_1d(np.asarray(energy_min, dtype=float))
        emax = np.atleast_1d(np.asarray(energy_max, dtype=float))
        emin, emax = np.broadcast_arrays(emin, emax)
        out = np.empty(emin.shape, dtype=float)
        if self._gpy_type is None:
            raise NotImplementedError(f"{type(self).__name__}.integral is outside the accelerated path")
        ctx = Context.get()
        pars = np.array([p.value for p in self.parameters], dtype=float)
        for i, (lo, hi) in enumerate(zip(emin.ravel(), emax.ravel())):
            edges = np.array([lo, hi], dtype=float)
            nodes = integration_nodes(edges)
            res = np.empty(1)
            check(ctx.lib.gpy_spectral_integral(
                ctx.handle, self._gpy_type, pars.ctypes.data_as(C.c_void_p), edges.ctypes.data_as(C.c_void_p), 1,
                nodes.ctypes.data_as(C.c_void_p), nodes.shape[1], res.ctypes.data_as(C.c_void_p)))
            out.ravel()[i] = res[0]
        return out if np.ndim(energy_min) or np.ndim(energy_max) else float(out[0])

    def energy_flux(self, energy_min, energy_max, **kwargs):
        """``SpectralModel.energy_flux`` (spectral.py:423-445): the integral of E phi(E), analytic where the class has
        ``evaluate_energy_flux``, else the log-log trapezoid of the reference; host numpy."""
        emin, emax = np.asarray(energy_min, dtype=float), np.asarray(energy_max, dtype=float)
        if hasattr(self, "evaluate_energy_flux"):
            out = self.evaluate_energy_flux(emin, emax, *[p.value for p in self.parameters])
        else:
            out = integrate_loglog(lambda x: x * self(x), emin, emax, **kwargs)
        return out if np.ndim(out) else float(out)


def _pl_integral(energy_min, energy_max, index, amplitude, reference):
    """``PowerLawSpectralModel.evaluate_integral`` (spectral.py:1039-1067), numpy on the host."""
    val = -1 * index + 1
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        prefactor = amplitude * reference / val
        integral = prefactor * ((energy_max / reference) ** val - (energy_min / reference) ** val)
        flat = np.isclose(val, 0)
        if np.any(flat):
            integral = np.where(flat, amplitude * reference * np.log(energy_max / energy_min), integral)
    return integral


def integrate_loglog(func, energy_min, energy_max, ndecade=100):
    """``integrate_spectrum`` + ``trapz_loglog`` (spectral.py:103-164, utils/integrate.py:8-49) on the host: composite
    trapezoid in log-log space on ``ndecade`` geomspace nodes per decade.  (The per-bin integrals of the likelihood path
    are K0's; this serves the host-side conveniences such as ``energy_flux``.)"""
    energy_min, energy_max = np.asarray(energy_min, dtype=float), np.asarray(energy_max, dtype=float)
    num = np.maximum(np.max(ndecade * np.log10(energy_max / energy_min)), 2)
    x = np.moveaxis(np.geomspace(energy_min, energy_max, num=int(num), axis=-1), -1, 0)
    y = func(x)
    tiny = np.finfo(float).tiny
    with np.errstate(invalid="ignore", divide="ignore"):
        index = -np.log(np.clip(y[:-1] / y[1:], tiny, np.inf)) / np.log(np.clip(x[:-1] / x[1:], tiny, np.inf))
    index[np.isnan(index)] = np.inf
    return _pl_integral(x[:-1], x[1:], index, y[:-1], x[:-1]).sum(axis=0)


def integration_nodes(edges, ndecade=100):
    """geomspace nodes of ``integrate_spectrum`` (spectral.py:132-134): the node count sits on a float
    knife-edge, so it is evaluated here with the reference's numpy expression and passed to the kernel."""
    edges = np.asarray(edges, dtype=float)
    energy_min, energy_max = edges[:-1], edges[1:]
    num = np.maximum(np.max(ndecade * np.log10(energy_max / energy_min)), 2)
    return np.ascontiguousarray(np.geomspace(energy_min, energy_max, num=int(num), axis=-1))


class PowerLawSpectralModel(SpectralModel):
    tag = ["PowerLawSpectralModel", "pl"]
    _gpy_type = 0
    index = Parameter("index", 2.0)
    amplitude = Parameter("amplitude", "1e-12 cm-2 s-1 TeV-1", scale_method="scale10", interp="log")
    reference = Parameter("reference", "1 TeV", frozen=True)

    @staticmethod
    def evaluate_energy_flux(energy_min, energy_max, index, amplitude, reference):
        """spectral.py:1070-1100."""
        val = -1 * float(index) + 2
        if np.isclose(val, 0):  # index 2: the logarithmic case
            return amplitude * reference**2 * np.log(energy_max / energy_min)
        return amplitude * reference**2 / val * ((energy_max / reference) ** val - (energy_min / reference) ** val)

    @staticmethod
    def evaluate(energy, index, amplitude, reference):
        return amplitude * np.power((energy / reference), -index)


class PowerLawNormSpectralModel(SpectralModel):
    tag = ["PowerLawNormSpectralModel", "pl-norm"]
    tilt = Parameter("tilt", 0, frozen=True)
    norm = Parameter("norm", 1, unit="", interp="log")
    reference = Parameter("reference", "1 TeV", frozen=True)

    @staticmethod
    def evaluate_energy_flux(energy_min, energy_max, tilt, norm, reference):
        """spectral.py:1182-1200."""
        return PowerLawSpectralModel.evaluate_energy_flux(energy_min, energy_max, tilt, norm, reference)

    @staticmethod
    def evaluate(energy, tilt, norm, reference):
        return norm * np.power((energy / reference), -tilt)


class ExpCutoffPowerLawSpectralModel(SpectralModel):
    tag = ["ExpCutoffPowerLawSpectralModel", "ecpl"]
    _gpy_type = 2
    index = Parameter("index", 1.5)
    amplitude = Parameter("amplitude", "1e-12 cm-2 s-1 TeV-1", scale_method="scale10", interp="log")
    reference = Parameter("reference", "1 TeV", frozen=True)
    lambda_ = Parameter("lambda_", "0.1 TeV-1")
    alpha = Parameter("alpha", "1.0", frozen=True)

    @staticmethod
    def evaluate(energy, index, amplitude, reference, lambda_, alpha):
        pwl = amplitude * (energy / reference) ** (-index)
        cutoff = np.exp(-np.power(energy * lambda_, alpha))
        return pwl * cutoff


class LogParabolaSpectralModel(SpectralModel):
    tag = ["LogParabolaSpectralModel", "lp"]
    _gpy_type = 1
    amplitude = Parameter("amplitude", "1e-12 cm-2 s-1 TeV-1", scale_method="scale10", interp="log")
    reference = Parameter("reference", "10 TeV", frozen=True)
    alpha = Parameter("alpha", 2)
    beta = Parameter("beta", 1)

    @staticmethod
    def evaluate(energy, amplitude, reference, alpha, beta):
        xx = energy / reference
        exponent = -alpha - beta * np.log(xx)
        return amplitude * np.power(xx, exponent)


# ------------------------------------------------------------------------------------------------
# spatial
# ------------------------------------------------------------------------------------------------
class SpatialModel(ModelBase):
    _type = "spatial"
    _gpy_type = None

    def __init__(self, **kwargs):
        self.frame = kwargs.pop("frame", "icrs")
        super().__init__(**kwargs)

    @property
    def position(self):
        return SkyCoord(self.lon_0.value, self.lat_0.value, frame=self.frame)

    @property
    def position_lonlat(self):
        return self.lon_0.value * DEG, self.lat_0.value * DEG

    @property
    def is_energy_dependent(self):
        return False

    def __call__(self, lon, lat):
        return self.evaluate(np.asarray(lon, float), np.asarray(lat, float), *[p.value for p in self.parameters])

    def integrate_geom(self, geom, oversampling_factor=None):
        """``SpatialModel.integrate_geom`` (spatial.py:222-309) on an image geometry, CUDA kernel K1.

        Returns the float32 image the PSF convolution consumes (ndmap.py:984)."""
        from .._lib import Context, check, frame_code, geom_desc
        from ..maps import Map

        if oversampling_factor is not None:
            raise NotImplementedError("integrate_geom: the oversampling factor is chosen as the reference does")
        g = geom.to_image()
        desc = geom_desc(g)
        pars = np.array([p.value for p in self.parameters], dtype=float)
        out = np.empty(g.data_shape, dtype=np.float32)
        f = C.c_int32(0)
        ctx = Context.get()
        check(ctx.lib.gpy_integrate_geom(ctx.handle, C.byref(desc), self._gpy_type, frame_code(self.frame),
                                         pars.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.byref(f)))
        m = Map.from_geom(g, data=out, unit="" if self._gpy_type == 0 else "")
        m.meta["oversampling_factor"] = int(f.value)
        return m


class PointSpatialModel(SpatialModel):
    tag = ["PointSpatialModel", "point"]
    _gpy_type = 0
    lon_0 = Parameter("lon_0", "0 deg")
    lat_0 = Parameter("lat_0", "0 deg", min=-90, max=90)

    @property
    def evaluation_bin_size_min(self):
        return 0.0

    @property
    def evaluation_radius(self):
        return 0.0


class GaussianSpatialModel(SpatialModel):
    tag = ["GaussianSpatialModel", "gauss"]
    _gpy_type = 1
    lon_0 = Parameter("lon_0", "0 deg")
    lat_0 = Parameter("lat_0", "0 deg", min=-90, max=90)
    sigma = Parameter("sigma", "1 deg", min=0)
    e = Parameter("e", 0, min=0, max=1, frozen=True)
    phi = Parameter("phi", "0 deg", frozen=True, min=-180, max=180)

    @property
    def evaluation_bin_size_min(self):
        return self.sigma.value / 3.0

    @property
    def evaluation_radius(self):
        return 5 * self.sigma.value

    @staticmethod
    def evaluate(lon, lat, lon_0, lat_0, sigma, e, phi):
        """sr-1 (spatial.py:684-701); angles in deg."""
        lon, lat, lon_0, lat_0, sigma, phi = (v * DEG for v in (lon, lat, lon_0, lat_0, sigma, phi))
        sep = angular_separation(lon, lat, lon_0, lat_0)
        phi = phi % np.pi
        if e == 0:
            a = 1.0 - np.cos(sigma)
            norm = 1 / (4 * np.pi * a * (1.0 - np.exp(-1.0 / a)))
        else:
            minor_axis, sigma_eff = _sigma_eff(lon_0, lat_0, lon, lat, phi, sigma, e)
            a = 1.0 - np.cos(sigma_eff)
            norm = 1 / (2 * np.pi * sigma * minor_axis)
        return norm * np.exp(-0.5 * ((1 - np.cos(sep)) / a))


def _sigma_eff(lon_0, lat_0, lon, lat, phi, major_axis, e):
    phi_0 = position_angle(lon_0, lat_0, lon, lat)
    d_phi = phi - phi_0
    minor_axis = major_axis * np.sqrt(1 - e**2)
    a2 = (major_axis * np.sin(d_phi)) ** 2
    b2 = (minor_axis * np.cos(d_phi)) ** 2
    return minor_axis, major_axis * minor_axis / np.sqrt(a2 + b2)


class DiskSpatialModel(SpatialModel):
    tag = ["DiskSpatialModel", "disk"]
    _gpy_type = 2
    lon_0 = Parameter("lon_0", "0 deg")
    lat_0 = Parameter("lat_0", "0 deg", min=-90, max=90)
    r_0 = Parameter("r_0", "1 deg", min=0)
    e = Parameter("e", 0, min=0, max=1, frozen=True)
    phi = Parameter("phi", "0 deg", frozen=True, min=-180, max=180)
    edge_width = Parameter("edge_width", value=0.01, min=0, max=1, frozen=True)

    @property
    def evaluation_bin_size_min(self):
        return self.r_0.value * (1 - self.edge_width.value) / 10.0

    @property
    def evaluation_radius(self):
        return 1.1 * self.r_0.value * (1 + self.edge_width.value)


# ------------------------------------------------------------------------------------------------
# cube
# ------------------------------------------------------------------------------------------------
class SkyModel:
    """Spatial x spectral source model (modeling/models/cube.py:33)."""

    tag = ["SkyModel", "sky-model"]
    _apply_irf_default = {"exposure": True, "psf": True, "edisp": True}

    def __init__(self, spectral_model, spatial_model=None, temporal_model=None, name=None, apply_irf=None,
                 datasets_names=None):
        if temporal_model is not None:
            raise NotImplementedError("temporal models are outside the accelerated path")
        if spatial_model is None:
            raise NotImplementedError("SkyModel without a spatial model is outside the accelerated path")
        self.spatial_model = spatial_model
        self.spectral_model = spectral_model
        self.temporal_model = None
        self._name = make_name(name)
        self.apply_irf = dict(self._apply_irf_default if apply_irf is None else apply_irf)
        self.datasets_names = [datasets_names] if isinstance(datasets_names, str) else datasets_names

    @property
    def name(self):
        return self._name

    @property
    def parameters(self):
        return Parameters.from_stack([self.spatial_model.parameters, self.spectral_model.parameters])

    @property
    def position(self):
        return self.spatial_model.position

    @property
    def evaluation_radius(self):
        return self.spatial_model.evaluation_radius

    @property
    def frame(self):
        return self.spatial_model.frame

    def copy(self, name=None, **kwargs):
        new = copy.deepcopy(self)
        new._name = make_name(name)
        return new

    def freeze(self, model_type=None):
        """cube.py:320-330: all parameters, or those of the spatial / spectral component."""
        if model_type in (None, "spatial"):
            self.spatial_model.freeze()
        if model_type in (None, "spectral"):
            self.spectral_model.freeze()

    def unfreeze(self, model_type=None):
        """Frozen status back to the component classes' defaults (cube.py:332-342)."""
        if model_type in (None, "spatial"):
            self.spatial_model.unfreeze()
        if model_type in (None, "spectral"):
            self.spectral_model.unfreeze()

    @property
    def frozen(self):
        return all(p.frozen for p in self.parameters)

    def __repr__(self):
        return f"SkyModel(name={self.name!r}, spatial_model={self.spatial_model!r}, spectral_model={self.spectral_model!r})"


class FoVBackgroundModel:
    """Field-of-view background: norm * (E / E0)^-tilt on the background map (cube.py:622-756)."""

    tag = ["FoVBackgroundModel", "fov-bkg"]

    def __init__(self, dataset_name, spectral_model=None, spatial_model=None, name=None):
        if spatial_model is not None:
            raise NotImplementedError("FoVBackgroundModel spatial models are outside the accelerated path")
        if spectral_model is None:
            spectral_model = PowerLawNormSpectralModel()
        if not isinstance(spectral_model, PowerLawNormSpectralModel):
            raise NotImplementedError("FoVBackgroundModel: only PowerLawNormSpectralModel is accelerated")
        self.spectral_model = spectral_model
        self.spatial_model = None
        self.datasets_names = [dataset_name] if isinstance(dataset_name, str) else list(dataset_name)
        self._name = make_name(name if name is not None else f"{self.datasets_names[0]}-bkg")

    @property
    def name(self):
        return self._name

    @property
    def parameters(self):
        return self.spectral_model.parameters

    def copy(self, name=None, **kwargs):
        new = copy.deepcopy(self)
        new._name = make_name(name)
        return new

    def freeze(self, model_type=None):
        if model_type in (None, "spectral"):
            self.spectral_model.freeze()

    def unfreeze(self, model_type=None):
        if model_type in (None, "spectral"):
            self.spectral_model.unfreeze()

    @property
    def frozen(self):
        return all(p.frozen for p in self.parameters)

    def __repr__(self):
        return f"FoVBackgroundModel(name={self.name!r}, datasets_names={self.datasets_names}, spectral_model={self.spectral_model!r})"


class Models(list):
    """List of models with a joint ``Parameters`` view (modeling/models/core.py DatasetModels)."""

    def __init__(self, models=None):
        if models is None:
            models = []
        elif isinstance(models, (SkyModel, FoVBackgroundModel)):
            models = [models]
        super().__init__(models)
        names = [m.name for m in self]
        if len(set(names)) != len(names):
            raise ValueError("Model names must be unique")
        self._penalties = None

    @property
    def parameters(self):
        return Parameters.from_stack([m.parameters for m in self])

    @property
    def names(self):
        return [m.name for m in self]

    @property
    def parameters_unique_names(self):
        """``model_name.component.parameter`` of every parameter, model by model (core.py:531-540)."""
        out = []
        for m in self:
            for t in ("spectral", "spatial"):
                sub = getattr(m, f"{t}_model", None)
                out += [f"{m.name}.{t}.{p.name}" for p in (sub.parameters if sub is not None else [])]
        return out

    @property
    def covariance(self):
        """Covariance matrix over ``self.parameters`` (value space) as the last ``Fit.covariance`` left it on the parameter
        objects; parameters that were frozen or not part of that fit have ``error ** 2`` on the diagonal."""
        from .parameter import covariance_of

        return covariance_of(self.parameters)

    @covariance.setter
    def covariance(self, matrix):
        from .parameter import store_covariance

        pars = list(self.parameters)
        matrix = np.asarray(matrix, dtype=float)
        if matrix.shape != (len(pars), len(pars)):
            raise ValueError(f"covariance must have shape {(len(pars), len(pars))}, got {matrix.shape}")
        store_covariance(pars, matrix)
        for p, var in zip(pars, np.diag(matrix)):
            p.error = float(np.sqrt(var)) if var >= 0 else float("nan")

    def selection_mask(self, name_substring=None, datasets_names=None, tag=None, model_type=None, frozen=None):
        """Boolean mask of the models that meet all conditions (modeling/models/core.py:888-949)."""
        tags = None if not tag else ([tag] if isinstance(tag, str) else list(tag))
        wanted = None if not datasets_names else ([datasets_names] if isinstance(datasets_names, str) else list(datasets_names))
        mask = np.ones(len(self), dtype=bool)
        for i, m in enumerate(self):
            if name_substring:
                mask[i] &= name_substring in m.name
            if wanted:
                mask[i] &= m.datasets_names is None or any(n in m.datasets_names for n in wanted)
            if tags:
                sub = m if model_type is None else getattr(m, f"{model_type}_model", None)
                mask[i] &= bool(sub is not None and any(t in sub.tag for t in tags))
            if frozen is not None:
                all_frozen = all(p.frozen for p in m.parameters)
                mask[i] &= all_frozen if frozen else not all_frozen
        return mask

    def select(self, name_substring=None, datasets_names=None, tag=None, model_type=None, frozen=None):
        """Models that meet all conditions (modeling/models/core.py:855-886)."""
        mask = self.selection_mask(name_substring, datasets_names, tag, model_type, frozen)
        return Models([m for m, keep in zip(self, mask) if keep])

    # -- container operations with the reference's uniqueness rule (core.py:128-136, 776-783, 1358-1380) ----------
    def _check_name_unique(self, model, names=None):
        if model.name in (self.names if names is None else names):
            raise ValueError(f"Model names must be unique. Models named '{model.name}' are duplicated.")

    def append(self, model):
        self._check_name_unique(model)
        list.append(self, model)

    def insert(self, idx, model):
        self._check_name_unique(model)
        list.insert(self, idx, model)

    def extend(self, models):
        for m in models:
            self.append(m)

    def __add__(self, other):
        if isinstance(other, (SkyModel, FoVBackgroundModel)):
            self._check_name_unique(other)
            return Models([*self, other])
        if isinstance(other, list):
            return Models([*self, *other])
        raise TypeError(f"Invalid type: {other!r}")

    def __setitem__(self, key, model):
        idx = self.index(key)
        self._check_name_unique(model, [n for n in self.names if n != list.__getitem__(self, idx).name])
        list.__setitem__(self, idx, model)

    def __delitem__(self, key):
        list.__delitem__(self, self.index(key))

    def index(self, key, *args):
        if isinstance(key, str):
            return self.names.index(key)
        if isinstance(key, (int, np.integer)):
            return int(key)
        return list.index(self, key, *args)

    def remove(self, key):
        del self[key]

    def freeze(self, model_type=None):
        """core.py:1077-1087."""
        for m in self:
            m.freeze(model_type)

    def unfreeze(self, model_type=None):
        """Frozen status back to the model classes' defaults (core.py:1089-1099)."""
        for m in self:
            m.unfreeze(model_type)

    @property
    def frozen(self):
        return all(p.frozen for p in self.parameters)

    def set_parameters_bounds(self, tag, model_type, parameters_names=None, min=None, max=None, value=None):
        """Limits / start values for the named parameters of the selected model types (core.py:1046-1075)."""
        names = None if parameters_names is None else (
            [parameters_names] if isinstance(parameters_names, str) else list(parameters_names))
        for m in self.select(tag=tag, model_type=model_type):
            sub = getattr(m, f"{model_type}_model", None)
            for p in (sub.parameters if sub is not None else []):
                if names is not None and p.name not in names:
                    continue
                if min is not None:
                    p.min = min
                if max is not None:
                    p.max = max
                if value is not None:
                    p.value = value

    def reassign(self, dataset_name, new_dataset_name):
        """Copies of the models with ``dataset_name`` replaced in their ``datasets_names`` (core.py:410-447, 1106-1119)."""
        old = dataset_name if isinstance(dataset_name, list) else [dataset_name]
        new = new_dataset_name if isinstance(new_dataset_name, list) else [new_dataset_name]
        out = []
        for m in self:
            c = m.copy(name=m.name)
            if c.datasets_names:
                for a, b in zip(old, new):
                    c.datasets_names = [n.replace(a, b) for n in c.datasets_names]
            out.append(c)
        return Models(out)

    def to_parameters_table(self):
        """Rows ``{model, type, name, value, unit, error, min, max, frozen}`` of every parameter (core.py:702-708,
        parameter.py:816-831; a list of dicts instead of an astropy Table)."""
        rows = []
        for m in self:
            comps = [(t, getattr(m, f"{t}_model", None)) for t in ("spectral", "spatial")]
            for t, sub in comps:
                for p in (sub.parameters if sub is not None else []):
                    rows.append(dict(model=m.name, type=t, name=p.name, value=p.value, unit=p.unit, error=p.error,
                                     min=p.min, max=p.max, frozen=p.frozen))
        return rows

    def __str__(self):
        lines = [f"{type(self).__name__}", ""]
        for i, m in enumerate(self):
            lines.append(f"Component {i}: {type(m).__name__}")
            lines.append(f"  Name                      : {m.name}")
            lines.append(f"  Datasets names            : {m.datasets_names}")
            lines.append(f"  Spectral model type       : {type(m.spectral_model).__name__}")
            sp = getattr(m, "spatial_model", None)
            lines.append(f"  Spatial  model type       : {type(sp).__name__ if sp is not None else ''}")
            lines.append("  Parameters:")
            for row in Models([m]).to_parameters_table():
                frozen = " (frozen)" if row["frozen"] else ""
                err = "" if not row["error"] else f" +/- {row['error']:.2g}"
                lines.append(f"    {row['name']:<24}{frozen:<9}: {row['value']:.4g}{err} {row['unit']}".rstrip())
            lines.append("")
        return "\n".join(lines)

    def __getitem__(self, key):
        if isinstance(key, str):
            for m in self:
                if m.name == key:
                    return m
            raise KeyError(f"No model named {key!r}; available: {self.names}")
        if isinstance(key, slice):
            return Models(list.__getitem__(self, key))
        if isinstance(key, np.ndarray) and key.dtype == bool:  # core.py:785-790: boolean mask
            return Models([m for m, keep in zip(self, key) if keep])
        return list.__getitem__(self, key)

    def copy(self):
        return Models([m.copy(name=m.name) for m in self])


DatasetModels = Models
