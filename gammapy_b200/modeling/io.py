"""Model serialisation: the YAML / dict layout of the reference's ``Models.write`` / ``Models.read``.

Reference: modeling/models/core.py (``ModelBase.to_dict`` :330-364, ``from_dict`` :366-382,
``_build_parameters_from_dict`` :107-125, ``_set_model_link`` :54-81, ``DatasetModels.read / from_yaml / from_dict /
write / to_yaml / to_dict`` :545-700), modeling/models/cube.py (``SkyModel.to_dict`` :463-483, ``from_dict`` :485-522,
``FoVBackgroundModel.to_dict`` :784-791, ``from_dict`` :793-835), modeling/models/spatial.py (``frame`` in the spatial
block, :311-333), modeling/parameter.py (``Parameter.to_dict`` :503-522, ``Parameters.update_link_label`` :803-814),
modeling/models/prior.py (``Prior.to_dict`` :110-135), utils/scripts.py (``YAML_FORMAT`` :32, ``to_yaml`` :105-121).

The layout, for a file written by either side::

    components:
    -   name: crab
        type: SkyModel
        spectral:
            type: PowerLawSpectralModel
            parameters:
            -   name: index
                value: 2.0
            ...
        spatial:
            type: PointSpatialModel
            frame: icrs
            parameters: [...]
    -   type: FoVBackgroundModel
        datasets_names: [obs-1]
        spectral: {type: PowerLawNormSpectralModel, parameters: [...]}

Only the model classes of the accelerated path exist here (the registries below); a component of any other type raises
``NotImplementedError`` naming it.  The covariance goes to the side file of the reference, ``<stem>_covariance.dat``
(``_write_models``, core.py:169-204; ``read_covariance`` / ``write_covariance``, :716-757): the text layout of
``astropy.table`` ``ascii.fixed_width`` with ``|`` as the delimiter -- a header row ``Parameters | 0 | 1 | ...`` and one
row per parameter ``model.component.name | values`` -- written and parsed here without astropy.
"""
import logging
import math
import os
import warnings

import numpy as np
import yaml

from .models import (
    DiskSpatialModel,
    ExpCutoffPowerLawSpectralModel,
    FoVBackgroundModel,
    GaussianSpatialModel,
    LogParabolaSpectralModel,
    Models,
    PointSpatialModel,
    PowerLawNormSpectralModel,
    PowerLawSpectralModel,
    SkyModel,
    make_name,
)
from .parameter import Parameter
from .prior import GaussianPrior, UniformPrior

log = logging.getLogger(__name__)
YAML_FORMAT = dict(sort_keys=False, indent=4, width=80, default_flow_style=False)  # utils/scripts.py:32

SPECTRAL_MODELS = [PowerLawSpectralModel, PowerLawNormSpectralModel, ExpCutoffPowerLawSpectralModel, LogParabolaSpectralModel]
SPATIAL_MODELS = [PointSpatialModel, GaussianSpatialModel, DiskSpatialModel]
PRIORS = [GaussianPrior, UniformPrior]
_PARAMETER_ITEMS = ["min", "max", "error", "interp", "scale_method", "scale_transform"]


def _get_cls(registry, tag, kind):
    for cls in registry:
        if tag in cls.tag:
            return cls
    raise NotImplementedError(f"{kind} {tag!r} is outside the accelerated path "
                              f"(available: {[c.tag[0] for c in registry]})")


def _same(a, b):
    if isinstance(a, float) and isinstance(b, float) and math.isnan(a) and math.isnan(b):
        return True
    return a == b


# -- priors ------------------------------------------------------------------------------------------
def _prior_parameters(prior):
    names = ("mu", "sigma") if isinstance(prior, GaussianPrior) else ("min", "max")
    return names, [float(getattr(prior, n)) for n in names]


def prior_to_dict(prior, full_output=False):
    """``Prior.to_dict()["prior"]`` (prior.py:110-135): prior parameters are unitless ``PriorParameter`` entries."""
    names, values = _prior_parameters(prior)
    params = []
    for n, v in zip(names, values):
        entry = {"name": n, "value": v, "unit": ""}
        if full_output:
            entry.update({"error": 0.0, "min": float("nan"), "max": float("nan")})
        params.append(entry)
    return {"type": prior.tag[0], "parameters": params, "weight": prior.weight}


def prior_from_dict(data):
    data = data.get("prior", data)
    cls = _get_cls(PRIORS, data["type"], "prior")
    values = {p["name"]: float(p["value"]) for p in data.get("parameters", [])}
    return cls(weight=data.get("weight", 1), **values)


# -- parameters ----------------------------------------------------------------------------------------
def parameter_to_dict(par):
    """``Parameter.to_dict`` (parameter.py:503-522)."""
    out = {
        "name": par.name,
        "value": float(par.value),
        "unit": str(par.unit or ""),
        "error": float(par.error),
        "min": float(par.min),
        "max": float(par.max),
        "frozen": bool(par.frozen),
        "interp": par.interp,
        "scale_method": par.scale_method,
        "scale_transform": par.scale_transform,
    }
    label = getattr(par, "_link_label_io", None)
    if label is not None:
        out["link"] = label
    if par.prior is not None:
        out["prior"] = prior_to_dict(par.prior)
    return out


def parameter_from_dict(data):
    data = dict(data)
    prior = data.pop("prior", None)
    link = data.pop("link", None)
    data.pop("scale", None)  # files of older versions carry the scale; it is recomputed by autoscale
    par = Parameter(
        data["name"], data["value"], unit=data.get("unit", "") or "", min=data.get("min", float("nan")),
        max=data.get("max", float("nan")), frozen=bool(data.get("frozen", False)), error=data.get("error", 0) or 0,
        interp=data.get("interp", "lin"), scale_method=data.get("scale_method", "scale10"),
        scale_transform=data.get("scale_transform", "lin"),
        prior=prior_from_dict(prior) if prior not in (None, "") else None)
    par._link_label_io = link
    return par


def update_link_label(parameters):
    """``Parameters.update_link_label`` (parameter.py:803-814): a parameter object that occurs more than once gets a
    label ``name@unique`` which the reader uses to share it again."""
    seen, shared = [], []
    for p in parameters:
        if not any(p is q for q in seen):
            seen.append(p)
        elif not any(p is q for q in shared):
            shared.append(p)
    for p in seen:
        if not any(p is q for q in shared):
            p._link_label_io = None
    for p in shared:
        if getattr(p, "_link_label_io", None) is None:
            p._link_label_io = p.name + "@" + make_name()


# -- component models ------------------------------------------------------------------------------------
def component_to_dict(model, full_output=False):
    """``ModelBase.to_dict`` (core.py:330-364) for a spectral or spatial model: ``{type: {type, [frame], parameters}}``."""
    params = [parameter_to_dict(p) for p in model.parameters]
    if not full_output:
        defaults = {d.name: parameter_to_dict(d) for d in model.default_parameters}
        for par in params:
            init = defaults[par["name"]]
            for item in _PARAMETER_ITEMS:
                if _same(par[item], init[item]):
                    del par[item]
            if par["frozen"] == init["frozen"]:
                del par["frozen"]
            if init["unit"] == "":
                del par["unit"]
    data = {"type": model.tag[0]}
    if model.type == "spatial":
        data["frame"] = model.frame
    data["parameters"] = params
    return {model.type: data}


def component_from_dict(data, registry, kind):
    key0 = next(iter(data))
    if key0 in ("spatial", "spectral", "temporal"):
        data = data[key0]
    cls = _get_cls(registry, data["type"], kind)
    given = {p["name"]: p for p in data.get("parameters", [])}
    probe = cls()
    kwargs = {}
    for default in probe.default_parameters:
        if default.name in kwargs:
            continue
        entry = parameter_to_dict(default)
        if default.name in given:
            entry.update(given[default.name])
        else:
            log.warning(f"Parameter '{default.name}' not defined in YAML file. Using default value: "
                        f"{entry['value']} {entry['unit']}")
        kwargs[default.name] = parameter_from_dict(entry)
    if kind == "spatial model" and "frame" in data:
        kwargs["frame"] = data["frame"]
    return cls(**kwargs)


# -- sky models ------------------------------------------------------------------------------------------
def model_to_dict(model, full_output=False):
    if isinstance(model, FoVBackgroundModel):  # cube.py:784-791
        data = {"type": model.tag[0], "datasets_names": list(model.datasets_names)}
        data.update(component_to_dict(model.spectral_model, full_output))
        return data
    data = {"name": model.name, "type": model.tag[0]}  # cube.py:463-483
    if model.apply_irf != model._apply_irf_default:
        data["apply_irf"] = dict(model.apply_irf)
    if model.datasets_names is not None:
        data["datasets_names"] = list(model.datasets_names)
    data.update(component_to_dict(model.spectral_model, full_output))
    if model.spatial_model is not None:
        data.update(component_to_dict(model.spatial_model, full_output))
    return data


def model_from_dict(data):
    tag = data.get("type")
    if tag in FoVBackgroundModel.tag:
        names = data.get("datasets_names")
        if names is None:
            raise ValueError("FoVBackgroundModel must define a dataset name")
        if data.get("spatial") is not None:
            raise NotImplementedError("FoVBackgroundModel spatial models are outside the accelerated path")
        spectral = component_from_dict({"spectral": data["spectral"]}, SPECTRAL_MODELS, "spectral model") \
            if data.get("spectral") is not None else None
        return FoVBackgroundModel(dataset_name=names, spectral_model=spectral, name=data.get("name"))
    if tag in SkyModel.tag:
        if data.get("temporal") is not None:
            raise NotImplementedError("temporal models are outside the accelerated path")
        spectral = component_from_dict({"spectral": data["spectral"]}, SPECTRAL_MODELS, "spectral model")
        spatial = component_from_dict({"spatial": data["spatial"]}, SPATIAL_MODELS, "spatial model") \
            if data.get("spatial") is not None else None
        return SkyModel(spectral_model=spectral, spatial_model=spatial, name=data["name"],
                        apply_irf=data.get("apply_irf"), datasets_names=data.get("datasets_names"))
    raise NotImplementedError(f"model type {tag!r} is outside the accelerated path (SkyModel, FoVBackgroundModel)")


def _components(model):
    return [m for m in (model.spectral_model, getattr(model, "spatial_model", None)) if m is not None]


def _set_links(models):
    """``_set_models_link`` (core.py:54-81): parameters that carry the same link label become one object."""
    register = {}
    for model in models:
        for comp in _components(model):
            for name in comp._par_names:
                par = comp.__dict__[name]
                label = getattr(par, "_link_label_io", None)
                if label is None:
                    continue
                if label in register:
                    comp.__dict__[name] = register[label]
                else:
                    register[label] = par


# -- Models ------------------------------------------------------------------------------------------------
def models_to_dict(models, full_output=False, covariance_file=None):
    update_link_label(models.parameters)
    data = {"components": [model_to_dict(m, full_output) for m in models]}
    if covariance_file is not None:
        data["covariance"] = str(covariance_file)
    return data


def models_from_dict(data, path=""):
    models = Models([model_from_dict(c) for c in data["components"]])
    _set_links(models)
    if "covariance" in data:
        filename = os.path.join(str(path), str(data["covariance"]))
        try:
            read_covariance(models, filename)
        except (OSError, ValueError) as exc:  # core.py:604-615: a warning, the errors of the YAML entries remain
            log.warning(f"Impossible to read the covariance correctly: {exc}. Covariance will be created from errors "
                        "and non-diagonal terms ignored.")
    return models


def write_covariance(models, filename):
    """``ascii.fixed_width`` with ``|`` delimiters (right-justified cells, as astropy writes them)."""
    names = models.parameters_unique_names
    matrix = models.covariance
    header = ["Parameters"] + [str(i) for i in range(len(names))]
    rows = [[name] + [repr(float(v)) for v in row] for name, row in zip(names, matrix)]
    widths = [max(len(r[k]) for r in [header] + rows) for k in range(len(header))]
    with open(str(filename), "w") as fh:
        for r in [header] + rows:
            fh.write("| " + " | ".join(cell.rjust(w) for cell, w in zip(r, widths)) + " |\n")


def read_covariance(models, filename):
    with open(str(filename)) as fh:
        lines = [ln.strip() for ln in fh if ln.strip()]
    cells = [[c.strip() for c in ln.strip("|").split("|")] for ln in lines]
    if not cells or cells[0][0] != "Parameters":
        raise ValueError(f"{filename}: not a covariance table (first column must be 'Parameters')")
    matrix = np.array([[float(c) for c in row[1:]] for row in cells[1:]], dtype=float)
    n = len(models.parameters)
    if matrix.shape != (n, n):
        raise ValueError(f"{filename}: covariance of shape {matrix.shape} for {n} parameters")
    errors = [p.error for p in models.parameters]
    models.covariance = matrix
    for p, e in zip(models.parameters, errors):  # the YAML entries stay authoritative for the errors
        p.error = e


def models_to_yaml(models, full_output=False):
    return yaml.safe_dump(models_to_dict(models, full_output), **YAML_FORMAT)


def models_from_yaml(text, path=""):
    data = yaml.safe_load(text)
    data.pop("checksum", None)
    data.pop("metadata", None)
    return models_from_dict(data, path=path)


def _has_covariance(models):
    return any(getattr(p, "_cov_partners", None) is not None for p in models.parameters)


def write_models(models, path, overwrite=False, full_output=False, write_covariance_file=True):
    path = str(path)
    if os.path.exists(path) and not overwrite:
        raise OSError(f"File exists already: {path}")
    folder = os.path.dirname(os.path.abspath(path))
    os.makedirs(folder, exist_ok=True)
    covariance_file = None
    if write_covariance_file and len(models.parameters) and _has_covariance(models):
        covariance_file = os.path.splitext(os.path.basename(path))[0] + "_covariance.dat"
        write_covariance(models, os.path.join(folder, covariance_file))
    with open(path, "w") as fh:
        fh.write(yaml.safe_dump(models_to_dict(models, full_output, covariance_file), **YAML_FORMAT))


def read_models(path):
    with open(str(path)) as fh:
        return models_from_yaml(fh.read(), path=os.path.dirname(os.path.abspath(str(path))))


def _install():
    """The reference's method names on the model classes."""
    Models.to_dict = lambda self, full_output=False: models_to_dict(self, full_output)
    Models.from_dict = staticmethod(lambda data, path="": models_from_dict(data, path=path))
    Models.to_yaml = lambda self, full_output=False: models_to_yaml(self, full_output)
    Models.from_yaml = staticmethod(lambda text, path="": models_from_yaml(text, path=path))
    Models.write = lambda self, path, overwrite=False, full_output=False, write_covariance=True, **kwargs: write_models(
        self, path, overwrite, full_output, write_covariance)
    Models.write_covariance = lambda self, filename, **kwargs: write_covariance(self, filename)
    Models.read_covariance = lambda self, path, filename="_covariance.dat", **kwargs: read_covariance(
        self, os.path.join(str(path), filename))
    Models.read = staticmethod(read_models)
    SkyModel.to_dict = lambda self, full_output=False: model_to_dict(self, full_output)
    SkyModel.from_dict = staticmethod(model_from_dict)
    FoVBackgroundModel.to_dict = lambda self, full_output=False: model_to_dict(self, full_output)
    FoVBackgroundModel.from_dict = staticmethod(model_from_dict)
    for cls in SPECTRAL_MODELS:
        cls.to_dict = lambda self, full_output=False: component_to_dict(self, full_output)
        cls.from_dict = classmethod(lambda c, data, **kw: component_from_dict(data, [c], "spectral model"))
    for cls in SPATIAL_MODELS:
        cls.to_dict = lambda self, full_output=False: component_to_dict(self, full_output)
        cls.from_dict = classmethod(lambda c, data, **kw: component_from_dict(data, [c], "spatial model"))
    Parameter.to_dict = parameter_to_dict


_install()
