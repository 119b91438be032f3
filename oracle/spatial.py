"""Oracle: spatial models and ``SpatialModel.integrate_geom``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows (paths relative to /root/reference/gammapy):

* ``compute_sigma_eff``                         modeling/models/spatial.py:57-67
* ``SpatialModel.integrate_geom``               modeling/models/spatial.py:222-309
* ``PointSpatialModel`` (_grid_weights, integrate_geom)   modeling/models/spatial.py:558-631
* ``GaussianSpatialModel`` (evaluate, radii)    modeling/models/spatial.py:639-701
* ``DiskSpatialModel`` (evaluate, norm, edge)   modeling/models/spatial.py:898-993
* ``WcsNDMap.downsample`` (block sum)           maps/wcs/ndmap.py:412-441
* ``WcsNDMap.stack`` (cast to the float32 map)  maps/wcs/ndmap.py:1138-1181
* ``Map.from_geom`` default dtype float32       maps/core.py:269-305
* ``Map._arithmetics`` (array op, numpy promotion)  maps/core.py:1902-1921

A spatial model is a dict ``{"type": "point"|"gauss"|"disk", "lon_0", "lat_0", ...}``, angles in deg.
"""
import numpy as np
import scipy.integrate
import scipy.special
from . import frames as _frames
from .wcs import DEG, NoOverlapError, angular_separation, position_angle

MAX_OVERSAMPLING = 200  # spatial.py:54


def wrap_at(value, min_value, max_value):
    """utils/units.py:14-35."""
    range_size = max_value - min_value
    return ((value - min_value) % range_size) + min_value


def compute_sigma_eff(lon_0, lat_0, lon, lat, phi, major_axis, e):
    """spatial.py:57-67; all angles in radians."""
    phi_0 = position_angle(lon_0, lat_0, lon, lat)
    d_phi = phi - phi_0
    minor_axis = major_axis * np.sqrt(1 - e**2)
    a2 = (major_axis * np.sin(d_phi)) ** 2
    b2 = (minor_axis * np.cos(d_phi)) ** 2
    denominator = np.sqrt(a2 + b2)
    sigma_eff = major_axis * minor_axis / denominator
    return minor_axis, sigma_eff


def evaluation_radius(model):
    """deg; point :573-579, gauss :672-678, disk :941-947."""
    t = model["type"]
    if t == "point":
        return 0.0
    if t == "gauss":
        return 5 * model["sigma"]
    if t == "disk":
        return 1.1 * model["r_0"] * (1 + model.get("edge_width", 0.01))
    raise ValueError(t)


def evaluation_bin_size_min(model):
    """deg; point :568-571, gauss :667-670, disk :933-939."""
    t = model["type"]
    if t == "point":
        return 0.0
    if t == "gauss":
        return model["sigma"] / 3.0
    if t == "disk":
        return model["r_0"] * (1 - model.get("edge_width", 0.01)) / 10.0
    raise ValueError(t)


def gauss_evaluate(lon, lat, lon_0, lat_0, sigma, e=0.0, phi=0.0):
    """spatial.py:684-701; inputs deg, output sr-1."""
    lon, lat, lon_0, lat_0, sigma, phi = (np.asarray(v, float) * DEG for v in (lon, lat, lon_0, lat_0, sigma, phi))
    sep = angular_separation(lon, lat, lon_0, lat_0)
    phi = wrap_at(phi, 0.0, np.pi)
    if e == 0:
        a = 1.0 - np.cos(sigma)
        norm = 1 / (4 * np.pi * a * (1.0 - np.exp(-1.0 / a)))
    else:
        minor_axis, sigma_eff = compute_sigma_eff(lon_0, lat_0, lon, lat, phi, sigma, e)
        a = 1.0 - np.cos(sigma_eff)
        norm = 1 / (2 * np.pi * sigma * minor_axis)
    exponent = -0.5 * ((1 - np.cos(sep)) / a)
    return norm * np.exp(exponent)


def disk_norm_factor(r_0, e):
    """spatial.py:951-969; r_0 in radians."""
    semi_minor = r_0 * np.sqrt(1 - e**2)

    def integral_fcn(x, a, b):
        A = 1 / np.sin(a) ** 2
        B = 1 / np.sin(b) ** 2
        C = A - B
        cs2 = np.cos(x) ** 2
        return 1 - np.sqrt(1 - 1 / (B + C * cs2))

    return (2 * scipy.integrate.quad(lambda x: integral_fcn(x, r_0, semi_minor), 0, np.pi)[0]) ** -1


def disk_evaluate(lon, lat, lon_0, lat_0, r_0, e=0.0, phi=0.0, edge_width=0.01):
    """spatial.py:977-993; inputs deg, output sr-1."""
    lon, lat, lon_0, lat_0, r_0, phi = (np.asarray(v, float) * DEG for v in (lon, lat, lon_0, lat_0, r_0, phi))
    sep = angular_separation(lon, lat, lon_0, lat_0)
    phi = wrap_at(phi, 0.0, np.pi)
    if e == 0:
        sigma_eff = r_0
    else:
        sigma_eff = compute_sigma_eff(lon_0, lat_0, lon, lat, phi, r_0, e)[1]
    norm = disk_norm_factor(float(r_0), e)
    value = (sep - sigma_eff) / (sigma_eff * edge_width)
    edge_width_95 = 2.326174307353347
    in_ellipse = 0.5 * (1 - scipy.special.erf(value * edge_width_95))
    return norm * in_ellipse


def map_position(model):
    """``model.position`` in the map's frame (skycoord_to_pixel, maps/wcs/geom.py:886-930): the dict may carry the model's
    ``frame`` and the ``map_frame`` of the geometry it is evaluated on (both None / absent: same frame)."""
    lon, lat = _frames.transform(model["lon_0"], model["lat_0"], model.get("frame"), model.get("map_frame"))
    return float(lon), float(lat)


def model_coords(model, lon, lat):
    """Map coordinates in the model's frame (``geom.get_coord(frame=self.frame)``, spatial.py:196-220)."""
    return _frames.transform(lon, lat, model.get("map_frame"), model.get("frame"))


def evaluate(model, lon, lat):
    t = model["type"]
    if t == "gauss":
        return gauss_evaluate(lon, lat, model["lon_0"], model["lat_0"], model["sigma"],
                              model.get("e", 0.0), model.get("phi", 0.0))
    if t == "disk":
        return disk_evaluate(lon, lat, model["lon_0"], model["lat_0"], model["r_0"],
                             model.get("e", 0.0), model.get("phi", 0.0), model.get("edge_width", 0.01))
    raise ValueError(f"oracle: evaluate not defined for {t!r}")


def block_sum(data, factor):
    """astropy block_reduce(func=np.nansum) with block (factor, factor) (maps/wcs/ndmap.py:430)."""
    ny, nx = data.shape
    return np.nansum(data.reshape(ny // factor, factor, nx // factor, factor), axis=(1, 3))


def point_integrate_geom(model, geom):
    """spatial.py:589-631: bilinear 4-pixel weights in pixel coordinates, unit "", float64."""
    x = np.arange(geom.nx, dtype=float)[np.newaxis, :]
    y = np.arange(geom.ny, dtype=float)[:, np.newaxis]
    x0, y0 = geom.world_to_pix(*map_position(model))
    dx = np.abs(x - x0)
    dx = np.where(dx < 1, 1 - dx, 0)
    dy = np.abs(y - y0)
    dy = np.where(dy < 1, 1 - dy, 0)
    return dx * dy


def integrate_geom(model, geom, oversampling_factor=None):
    """spatial.py:222-309 on an image geometry -> (ny, nx) array with the reference's dtype.

    float32 when the oversampled branch stacks into the default float32 map and float64
    otherwise; both are then multiplied by the float64 solid angle (-> float64).
    """
    if model["type"] == "point":
        return point_integrate_geom(model, geom)

    result = np.zeros(geom.data_shape, dtype=np.float32)  # Map.from_geom default dtype
    pix_scale = geom.pixel_scale
    radius = evaluation_radius(model)
    geom_cut, parent_slices = geom, None
    try:
        width = 2 * np.maximum(radius, pix_scale)
        geom_cut, parent_slices = geom.cutout_info(map_position(model), width)
    except (NoOverlapError, ValueError):
        oversampling_factor = 1

    if oversampling_factor is None:
        res_scale = evaluation_bin_size_min(model)
        if res_scale > 0:
            oversampling_factor = np.minimum(int(np.ceil(pix_scale / res_scale)), MAX_OVERSAMPLING)
        else:
            oversampling_factor = MAX_OVERSAMPLING

    if oversampling_factor > 1:
        f = int(oversampling_factor)
        up = geom_cut.upsample(f)
        lon, lat = model_coords(model, *up.get_coord())
        values = evaluate(model, lon, lat) / f**2
        upsampled = np.zeros(up.data_shape, dtype=np.float32) + values  # float32 map += float64 array -> float64
        integrated = block_sum(upsampled, f)
        # result.stack(integrated): cast to result dtype (float32), non-finite -> 0, in-place add
        data = integrated.astype(result.dtype)
        not_finite = ~np.isfinite(data)
        if np.any(not_finite):
            data = data.copy()
            data[not_finite] = 0
        # cutout slices in the parent are recomputed from the cutout centre (mode="partial")
        sl = geom_cut.cutout_slices(geom)
        result[sl["parent-slices"]] += data[sl["cutout-slices"]]
    else:
        lon, lat = model_coords(model, *geom.get_coord())
        values = evaluate(model, lon, lat)
        result = result + values  # float32 zeros + float64 -> float64
    return result * geom.solid_angle()
