"""Oracle: IRF kernel extraction at a source position, restated on plain numpy arrays.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **Parity unpinned** beyond the position-independent limit: the
reference's own tests of position-dependent PSF / edisp maps need ``$GAMMAPY_DATA``; what pins this file is (1) the
position-independent case, where it must reproduce the kernel ``test_npred_psf_after_edisp`` pins through its npred
total, and (2) reading the reference source.

Follows (paths relative to /root/reference/gammapy):

* ``PSFMap.get_psf_kernel``            irf/psf/map.py:249-337   (kernel geometry centred on the exposure geometry's
  centre, per-energy cut to the 0.999 containment radius, <= 4x oversampling, block sum, plane normalisation)
* ``_psf_upsampling_factor``           irf/psf/map.py:23-37
* ``PSFMap.containment_radius``        irf/psf/map.py:198-222 -> ``PSF.containment_radius`` irf/psf/core.py:39-87
  (argmin of |containment - fraction| on the 20x upsampled rad axis) -> ``IRF.integral / cumsum`` irf/core.py:358-411
  (cumulative sum of psf * bin width * 2 pi rad at the upper bin edges, linear interpolation, 0 outside the nodes)
* ``PSFMap._get_irf_coords``           irf/psf/map.py:154-170   (position -> fractional pixel of the IRF geometry,
  multilinear interpolation in (lat_idx, lon_idx))
* ``Map.interp_by_coord(method="linear")`` for the kernel image: linear in rad between the rad bin centres with
  extrapolation (fill_value=None), bilinear in the IRF pixel grid
* ``EDispKernelMap.get_edisp_kernel``  irf/edisp/map.py:369-401 -> ``to_region_nd_map(region=position)``
  maps/wcs/ndmap.py:661-728: ``interp_by_coord(method="nearest")``, i.e. the matrix of the nearest IRF pixel
* ``PSFKernel.normalize``              irf/psf/kernel.py:70-80
* ``round_up_to_odd``                  utils/array.py:110; ``WcsGeom.to_odd_npix`` maps/wcs/geom.py:1106-1138
"""
import numpy as np

from .wcs import DEG, ImageGeom, angular_separation

PSF_MAX_OVERSAMPLING = 4  # irf/psf/map.py:20


class IrfGeom(ImageGeom):
    """Image grid of an IRF map: an ``ImageGeom`` (CAR or TAN about its reference point; about the equator pixel (i, j)
    has its centre at lon = crval_lon - binsz (i + 1 - crpix_x), lat = binsz (j + 1 - crpix_y))."""

    def coord_to_pix(self, lon, lat):
        px, py = self.world_to_pix(lon, lat)
        return float(px), float(py)


def _clamped_bilinear(data, px, py):
    """data[..., ny, nx] at fractional pixel (px, py), linear between pixel centres, clamped to the grid
    (``_get_nearest_valid_position`` keeps the lookup inside the map)."""
    ny, nx = data.shape[-2:]
    px = float(np.clip(px, 0.0, nx - 1.0))
    py = float(np.clip(py, 0.0, ny - 1.0))
    x0 = int(min(np.floor(px), max(nx - 2, 0)))
    y0 = int(min(np.floor(py), max(ny - 2, 0)))
    x1, y1 = min(x0 + 1, nx - 1), min(y0 + 1, ny - 1)
    tx, ty = px - x0, py - y0
    return ((1 - ty) * ((1 - tx) * data[..., y0, x0] + tx * data[..., y0, x1])
            + ty * ((1 - tx) * data[..., y1, x0] + tx * data[..., y1, x1]))


def psf_profile_at(psf_map, irf_geom, lon, lat):
    """(nT, nrad) radial profile at (lon, lat) [deg]; ``psf_map`` is (nT, nrad) or (nT, nrad, ny, nx)."""
    psf_map = np.asarray(psf_map, dtype=float)
    if psf_map.ndim == 2:
        return psf_map
    px, py = irf_geom.coord_to_pix(lon, lat)
    return _clamped_bilinear(psf_map, px, py)


def containment_radius(profile, rad_edges, fraction, factor=20):
    """One radius [deg] per energy: irf/psf/core.py:39-87 on the profile of one position."""
    edges = np.asarray(rad_edges, dtype=float)
    centers = 0.5 * (edges[:-1] + edges[1:])
    width = np.diff(edges)
    # IRF.cumsum (irf/core.py:372-390): values * bin_width * 2 pi rad_centre, nodes at the upper edges
    cum = np.cumsum(profile * (width * DEG) * (2 * np.pi * centers * DEG), axis=1)
    nodes = edges[1:]
    # MapAxis.upsample(factor): factor x more bins between the same edges (linear axis); centres of those bins
    n_up = (len(edges) - 1) * factor
    up_edges = np.linspace(edges[0], edges[-1], n_up + 1)
    rad = 0.5 * (up_edges[:-1] + up_edges[1:])
    out = np.empty(profile.shape[0])
    for t in range(profile.shape[0]):
        cont = np.interp(rad, nodes, cum[t], left=0.0, right=0.0)  # bounds_error=False, fill_value=0
        out[t] = rad[np.argmin(np.abs(cont - fraction))]
    return out


def round_up_to_odd(f):
    return int(np.floor(np.ceil(f) / 2.0) * 2.0 + 1.0)


def _interp_rad_linear(profile_t, centers, rad):
    """Linear interpolation on the rad bin centres, extrapolating with the edge segments."""
    idx = np.clip(np.searchsorted(centers, rad) - 1, 0, len(centers) - 2)
    t = (rad - centers[idx]) / (centers[idx + 1] - centers[idx])
    return profile_t[idx] * (1 - t) + profile_t[idx + 1] * t


def psf_kernel_at(psf_map, rad_edges, irf_geom, position, binsz, center, containment=0.999, precision_factor=12,
                  proj="CAR"):
    """(nT, K, K) kernel at ``position`` (lon, lat) [deg] for a map of pixel size ``binsz`` and projection ``proj`` whose
    centre is ``center``: every plane on ``geom.to_odd_npix(max_radius=radii[t]).upsample(factor)``, i.e. a geometry
    created about ``center`` in the map's projection (maps/wcs/geom.py:1106-1138), rad = separation of its pixel
    centres from ``center`` (irf/psf/map.py:310-318)."""
    profile = psf_profile_at(psf_map, irf_geom, position[0], position[1])
    edges = np.asarray(rad_edges, dtype=float)
    centers = 0.5 * (edges[:-1] + edges[1:])
    radii = containment_radius(profile, edges, containment)
    r68 = containment_radius(profile, edges, 0.68)
    max_radius = radii.max()
    K = round_up_to_odd(np.float64(2 * max_radius / binsz))
    nT = profile.shape[0]
    data = np.zeros((nT, K, K))
    lon_c, lat_c = center
    for t in range(nT):
        base = 2 * r68[t] / binsz
        with np.errstate(divide="ignore"):
            factor = int(min(int(np.ceil(precision_factor / base)) if base > 0 else PSF_MAX_OVERSAMPLING, PSF_MAX_OVERSAMPLING))
        n = round_up_to_odd(np.float64(2 * radii[t] / binsz))
        cut = ImageGeom.create(skydir=(lon_c, lat_c), binsz=binsz, npix=(n, n), proj=proj).upsample(factor)
        lon, lat = cut.get_coord()
        rad = angular_separation(lon * DEG, lat * DEG, lon_c * DEG, lat_c * DEG) / DEG
        vals = np.clip(_interp_rad_linear(profile[t], centers, rad.ravel()).reshape(rad.shape), 0, np.inf)
        img = vals.reshape(n, factor, n, factor).sum(axis=(1, 3))
        o = (K - n) // 2
        data[t, o:o + n, o:o + n] = img
    with np.errstate(divide="ignore", invalid="ignore"):
        data = np.nan_to_num(data / data.sum(axis=(1, 2), keepdims=True))
    return data


def edisp_kernel_at(edisp_map, irf_geom, position):
    """(nT, nR) matrix at ``position``: the nearest IRF pixel (clamped to the grid)."""
    edisp_map = np.asarray(edisp_map, dtype=float)
    if edisp_map.ndim == 2:
        return edisp_map
    px, py = irf_geom.coord_to_pix(position[0], position[1])
    ny, nx = edisp_map.shape[-2:]
    i = int(np.clip(np.floor(px + 0.5), 0, nx - 1))
    j = int(np.clip(np.floor(py + 0.5), 0, ny - 1))
    return edisp_map[:, :, j, i]
