"""Oracle: spectral models and their energy-bin integrals.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (paths relative to /root/reference/gammapy):

* ``PowerLawSpectralModel.evaluate / evaluate_integral``   modeling/models/spectral.py:1034-1067
* ``PowerLawNormSpectralModel.evaluate``                   modeling/models/spectral.py:1159-1162
* ``ExpCutoffPowerLawSpectralModel.evaluate``              modeling/models/spectral.py:1561-1567
* ``LogParabolaSpectralModel.evaluate``                    modeling/models/spectral.py:1903-1908
* ``integrate_spectrum``                                   modeling/models/spectral.py:103-164
* ``trapz_loglog``                                         utils/integrate.py:8-49
* ``LogScale._scale`` (zero clipping)                      utils/interpolation.py:196-210
* ``SpectralModel.integral``                               modeling/models/spectral.py:353-371

Units by convention: energies TeV, amplitudes cm-2 s-1 TeV-1, lambda_ TeV-1.
A spectral model is a dict ``{"type": "pl"|"lp"|"ecpl"|"pl-norm", <parameter values>}``.
"""
import numpy as np

_TINY64 = np.finfo(np.float64).tiny


def pl_evaluate(energy, index, amplitude, reference):
    return amplitude * np.power((energy / reference), -index)


def pl_evaluate_integral(energy_min, energy_max, index, amplitude, reference):
    """spectral.py:1039-1067 (arrays broadcast; ``index`` may be an array, as trapz_loglog passes it)."""
    val = np.asarray(-1 * index + 1, dtype=float)  # Quantity arithmetic in the reference: numpy semantics
    with np.errstate(all="ignore"):
        prefactor = amplitude * reference / val
        upper = np.power((energy_max / reference), val)
        lower = np.power((energy_min / reference), val)
        integral = prefactor * (upper - lower)
        mask = np.isclose(val, 0)
        if np.any(mask):
            integral = np.array(np.broadcast_to(integral, np.broadcast(integral, mask).shape), dtype=float)
            alt = amplitude * reference * np.log(energy_max / energy_min)
            alt = np.broadcast_to(alt, integral.shape)
            mask = np.broadcast_to(mask, integral.shape)
            integral[mask] = alt[mask]
    return integral


def ecpl_evaluate(energy, index, amplitude, reference, lambda_, alpha):
    pwl = amplitude * (energy / reference) ** (-index)
    cutoff = np.exp(-np.power(energy * lambda_, alpha))
    return pwl * cutoff


def lp_evaluate(energy, amplitude, reference, alpha, beta):
    xx = energy / reference
    exponent = -alpha - beta * np.log(xx)
    return amplitude * np.power(xx, exponent)


def plnorm_evaluate(energy, tilt, norm, reference):
    return norm * np.power((energy / reference), -tilt)


def evaluate(model, energy):
    t = model["type"]
    if t == "pl":
        return pl_evaluate(energy, model["index"], model["amplitude"], model["reference"])
    if t == "ecpl":
        return ecpl_evaluate(energy, model["index"], model["amplitude"], model["reference"],
                             model["lambda_"], model.get("alpha", 1.0))
    if t == "lp":
        return lp_evaluate(energy, model["amplitude"], model["reference"], model["alpha"], model["beta"])
    if t == "pl-norm":
        return plnorm_evaluate(energy, model.get("tilt", 0.0), model["norm"], model.get("reference", 1.0))
    raise ValueError(f"oracle: unknown spectral model {t!r}")


def _log_scale(values):
    """utils/interpolation.py:207-210."""
    values = np.clip(values, _TINY64, np.inf)
    return np.log(values)


def trapz_loglog(y, x, axis=-1):
    """utils/integrate.py:8-49."""
    x, y = np.moveaxis(x, axis, 0), np.moveaxis(y, axis, 0)
    energy_min, energy_max = x[:-1], x[1:]
    vals_energy_min, vals_energy_max = y[:-1], y[1:]
    with np.errstate(invalid="ignore", divide="ignore"):
        index = -_log_scale(vals_energy_min / vals_energy_max) / _log_scale(energy_min / energy_max)
    index[np.isnan(index)] = np.inf
    return pl_evaluate_integral(
        energy_min=energy_min, energy_max=energy_max, index=index,
        reference=energy_min, amplitude=vals_energy_min,
    )


def integrate_num(energy_min, energy_max, ndecade=100):
    """The node count of integrate_spectrum (spectral.py:132-134) — a float knife-edge, keep the numpy expression."""
    num = np.maximum(np.max(ndecade * np.log10(energy_max / energy_min)), 2)
    return int(num)


def integrate_nodes(energy_min, energy_max, ndecade=100):
    """geomspace nodes, shape (n_energy, num) (spectral.py:132-134)."""
    num = integrate_num(energy_min, energy_max, ndecade)
    return np.geomspace(energy_min, energy_max, num=num, axis=-1)


def integrate_spectrum(func, energy_min, energy_max, ndecade=100):
    """spectral.py:103-164 (no parameter samples, no energy flux)."""
    energy = integrate_nodes(np.asarray(energy_min, float), np.asarray(energy_max, float), ndecade)
    values = func(energy)
    values = values.reshape(energy.shape)
    integral = trapz_loglog(values, energy, axis=-1)
    return integral.sum(axis=0)


def integral(model, energy_min, energy_max):
    """SpectralModel.integral (spectral.py:353-371): analytic for power laws, log-log trapezoid otherwise."""
    energy_min = np.asarray(energy_min, dtype=float)
    energy_max = np.asarray(energy_max, dtype=float)
    if model["type"] == "pl":
        return pl_evaluate_integral(energy_min, energy_max, model["index"], model["amplitude"],
                                    model["reference"])
    return integrate_spectrum(lambda e: evaluate(model, e), energy_min, energy_max)
