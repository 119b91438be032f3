"""Synthetic benchmark datasets of BASELINE.json ``configs``: the value distributions and
seeds are the contract; everything is built through the public API (``MapDataset``, ``SkyModel`` ...)."""
import numpy as np

from .datasets import Datasets, MapDataset
from .irf import EDispKernelMap, PSFMap
from .maps import Map, MapAxis, WcsGeom
from .modeling.models import (DiskSpatialModel, ExpCutoffPowerLawSpectralModel, FoVBackgroundModel,
                              GaussianSpatialModel, LogParabolaSpectralModel, PointSpatialModel,
                              PowerLawSpectralModel, SkyModel)

FRAME = "galactic"


def make_dataset(width=(4.0, 4.0), binsz=0.02, n_reco=10, n_true=20, e_reco=(0.1, 10.0), e_true=(0.03, 30.0),
                 name="c1", exposure_scale=1.0, pointing_offset=(0.0, 0.0), mask_radius=1.8, skydir=(0.0, 0.0),
                 falloff="radial", proj="CAR"):
    """One MapDataset with smooth synthetic IRFs (no models, no counts).

    ``falloff``: "radial" = exp(-(theta / 2 deg)^2 / 2) about the pointing (one observation);
    "latitude" = the same profile in Galactic latitude only, flat along the plane (a survey mosaic: every
    low-energy bin of the map holds counts)."""
    energy_axis = MapAxis.from_energy_bounds(e_reco[0], e_reco[1], nbin=n_reco, unit="TeV")
    energy_axis_true = MapAxis.from_energy_bounds(e_true[0], e_true[1], nbin=n_true, unit="TeV", name="energy_true")
    geom = WcsGeom.create(skydir=skydir, binsz=binsz, width=width, frame=FRAME, proj=proj, axes=[energy_axis])
    ds = MapDataset.create(geom=geom, energy_axis_true=energy_axis_true, name=name)
    center = (skydir[0] + pointing_offset[0], skydir[1] + pointing_offset[1])
    if falloff == "latitude":
        c = geom.to_image().get_coord()
        theta = np.abs(c["lat"] - center[1]) + 0.0 * c["lon"]
    elif falloff == "radial":
        theta = geom.separation(center)
    else:
        raise ValueError(f"falloff must be 'radial' or 'latitude', not {falloff!r}")
    falloff = np.exp(-0.5 * (theta / 2.0) ** 2)
    e_t = energy_axis_true.center[:, None, None]
    e_r = energy_axis.center[:, None, None]
    ds.exposure = Map.from_geom(ds.exposure.geom, unit="m2 s",
                                data=(exposure_scale * 1e10 * falloff[None] * (e_t / 1.0) ** 0.3).astype(np.float32))
    ds.background = Map.from_geom(geom, data=(0.5 * (e_r / 1.0) ** -2.7 * falloff[None]).astype(np.float32))
    sigma = np.exp(np.interp(np.log(energy_axis_true.center), np.log([e_true[0], e_true[1]]),
                             np.log([0.12, 0.04])))
    ds.psf = PSFMap.from_gauss(energy_axis_true=energy_axis_true, sigma=sigma)
    ds.edisp = EDispKernelMap.from_gauss(energy_axis=energy_axis, energy_axis_true=energy_axis_true, sigma=0.15, bias=0)
    ds.mask_safe = Map.from_geom(geom, data=np.ones(geom.data_shape, dtype=bool), dtype=bool)
    if mask_radius is not None:
        circle = geom.separation(skydir) < mask_radius
        ds.mask_fit = Map.from_geom(geom, data=np.broadcast_to(circle, geom.data_shape).copy(), dtype=bool)
    return ds


def zoned_irfs(ds, binsz_irf=1.0, psf_scale=(1.0, 1.8), edisp_sigma=(0.10, 0.22), edisp_bias=(0.0, 0.06)):
    """Replace the IRFs of ``ds`` by maps on a coarse spatial grid with two zones split at the map's central meridian
    (lon > centre: zone 0, else zone 1): the PSF width and the edisp resolution / bias differ between the zones, with
    one transition column of IRF pixels in between for the (bilinear) PSF lookup.  Position-dependent IRFs as
    ``PSFMap`` / ``EDispKernelMap`` of a real DL4 dataset carry them."""
    geom = ds._geom
    e_true = ds.exposure.geom.axes["energy_true"]
    e_reco = geom.axes["energy"]
    c = geom.center_skydir
    # (as MapDatasetMaker builds them: the dataset's projection about the dataset's centre, coarser pixels)
    ig = WcsGeom.create(skydir=(c.lon, c.lat), binsz=binsz_irf, width=tuple(geom.width), frame=geom.frame,
                        proj=geom.projection)
    coords = ig.get_coord()
    dlon = (coords["lon"] - c.lon + 180.0) % 360.0 - 180.0
    zone = np.where(dlon > 0, 0, 1)  # (ny, nx)
    base = ds.psf  # position-independent Gaussian PSF of make_dataset
    psf = np.empty(base.data.shape + zone.shape)
    sig0 = np.exp(np.interp(np.log(e_true.center), np.log([e_true.edges[0], e_true.edges[-1]]), np.log([0.12, 0.04])))
    rad_axis = base.rad_axis
    for z in (0, 1):
        prof = PSFMap.from_gauss(energy_axis_true=e_true, rad_axis=rad_axis, sigma=sig0 * psf_scale[z]).data
        psf[..., zone == z] = prof[..., None]
    ds.psf = PSFMap(e_true, rad_axis, psf, geom=ig)
    ed = np.empty((e_true.nbin, e_reco.nbin) + zone.shape)
    for z in (0, 1):
        k = EDispKernelMap.from_gauss(energy_axis=e_reco, energy_axis_true=e_true, sigma=edisp_sigma[z],
                                      bias=edisp_bias[z]).kernel.pdf_matrix
        ed[..., zone == z] = k[..., None]
    ds.edisp = EDispKernelMap(data=ed, geom=ig, axes=[e_true, e_reco])
    return ds


def graded_irfs(ds, binsz_irf=0.5, psf_scale=(1.0, 1.6), edisp_sigma=(0.10, 0.20), edisp_bias=(0.0, 0.05)):
    """Replace the IRFs of ``ds`` by maps on a spatial grid of ``binsz_irf`` pixels whose PSF width and energy resolution /
    bias grow smoothly with the offset from the map centre, as they do across the field of view of a real observation
    (``MapDatasetMaker`` fills ``PSFMap`` / ``EDispKernelMap`` per IRF pixel): EVERY IRF pixel has its own edisp matrix, so
    every evaluator gets a matrix of its own (irf/edisp/map.py:369-401, nearest pixel) and a PSF kernel interpolated
    between its four neighbours."""
    geom = ds._geom
    e_true = ds.exposure.geom.axes["energy_true"]
    e_reco = geom.axes["energy"]
    c = geom.center_skydir
    ig = WcsGeom.create(skydir=(c.lon, c.lat), binsz=binsz_irf, width=tuple(geom.width), frame=geom.frame,
                        proj=geom.projection)
    offset = ig.separation((c.lon, c.lat))  # (ny, nx) deg
    u = offset / max(float(offset.max()), 1e-12)
    # break the circular symmetry so that no two IRF pixels share their numbers
    yy, xx = np.mgrid[0:ig.npix[1], 0:ig.npix[0]]
    u = np.clip(u + 0.01 * ((3 * xx + 7 * yy) % 11) / 11.0, 0.0, 1.0)
    rad_axis = ds.psf.rad_axis
    sig0 = np.exp(np.interp(np.log(e_true.center), np.log([e_true.edges[0], e_true.edges[-1]]), np.log([0.12, 0.04])))
    psf = np.empty(ds.psf.data.shape + u.shape)
    ed = np.empty((e_true.nbin, e_reco.nbin) + u.shape)
    for j in range(u.shape[0]):
        for i in range(u.shape[1]):
            w = float(u[j, i])
            scale = psf_scale[0] + w * (psf_scale[1] - psf_scale[0])
            psf[..., j, i] = PSFMap.from_gauss(energy_axis_true=e_true, rad_axis=rad_axis, sigma=sig0 * scale).data
            ed[..., j, i] = EDispKernelMap.from_gauss(
                energy_axis=e_reco, energy_axis_true=e_true, sigma=edisp_sigma[0] + w * (edisp_sigma[1] - edisp_sigma[0]),
                bias=edisp_bias[0] + w * (edisp_bias[1] - edisp_bias[0])).kernel.pdf_matrix
    ds.psf = PSFMap(e_true, rad_axis, psf, geom=ig)
    ds.edisp = EDispKernelMap(data=ed, geom=ig, axes=[e_true, e_reco])
    return ds


def c1_models(dataset_name="c1"):
    """analysis_3d tutorial shape: one Gaussian x PowerLaw + FoV background (simulate_3d.py:102-134)."""
    spatial = GaussianSpatialModel(lon_0="0.2 deg", lat_0="0.1 deg", sigma="0.3 deg", frame=FRAME)
    spectral = PowerLawSpectralModel(index=3, amplitude="1e-11 cm-2 s-1 TeV-1", reference="1 TeV")
    src = SkyModel(spatial_model=spatial, spectral_model=spectral, name="model-simu")
    bkg = FoVBackgroundModel(dataset_name=dataset_name)
    return [src, bkg]


def make_c1(name="c1", **kwargs):
    ds = make_dataset(name=name, **kwargs)
    ds.models = c1_models(name)
    return ds


def c2_models(dataset_name="c2"):
    """Five sources on a ring of 1 deg: Point x PL, Gauss(0.2) x LP, Disk(0.3) x ECPL, Gauss(0.05) x PL, Point x ECPL."""
    ang = np.deg2rad(np.arange(5) * 72.0 + 18.0)
    pos = [(float(np.round(np.cos(a), 4)), float(np.round(np.sin(a), 4))) for a in ang]
    spatial = [
        PointSpatialModel(lon_0=pos[0][0], lat_0=pos[0][1], frame=FRAME),
        GaussianSpatialModel(lon_0=pos[1][0], lat_0=pos[1][1], sigma=0.2, frame=FRAME),
        DiskSpatialModel(lon_0=pos[2][0], lat_0=pos[2][1], r_0=0.3, frame=FRAME),
        GaussianSpatialModel(lon_0=pos[3][0], lat_0=pos[3][1], sigma=0.05, frame=FRAME),
        PointSpatialModel(lon_0=pos[4][0], lat_0=pos[4][1], frame=FRAME),
    ]
    spectral = [
        PowerLawSpectralModel(index=2.3, amplitude=3e-12, reference=1),
        LogParabolaSpectralModel(amplitude=5e-12, reference=1, alpha=2.2, beta=0.2),
        ExpCutoffPowerLawSpectralModel(index=1.8, amplitude=8e-12, reference=1, lambda_=0.2),
        PowerLawSpectralModel(index=2.0, amplitude=2e-12, reference=1),
        ExpCutoffPowerLawSpectralModel(index=2.1, amplitude=4e-12, reference=1, lambda_=0.05),
    ]
    models = [SkyModel(spatial_model=s, spectral_model=p, name=f"src-{i}") for i, (s, p) in enumerate(zip(spatial, spectral))]
    models.append(FoVBackgroundModel(dataset_name=dataset_name))
    return models


def make_c2(name="c2", **kwargs):
    ds = make_dataset(name=name, **kwargs)
    ds.models = c2_models(name)
    return ds


def survey_models(n_sources=50, lon_range=(-19.0, 19.0), lat_range=(-3.0, 3.0), seed=2, dataset_name="c3"):
    """Galactic-plane population: types cycled over the 3 x 3 grid, positions uniform, extensions
    log-uniform 0.05-0.5 deg, amplitudes log-uniform 1e-13-1e-11 (rng seed 2)."""
    rng = np.random.RandomState(seed)
    models = []
    for i in range(n_sources):
        lon = float(rng.uniform(*lon_range))
        lat = float(rng.uniform(*lat_range))
        ext = float(np.exp(rng.uniform(np.log(0.05), np.log(0.5))))
        amp = float(np.exp(rng.uniform(np.log(1e-13), np.log(1e-11))))
        index = float(rng.uniform(1.8, 2.8))
        kind_sp, kind_sc = i % 3, (i // 3) % 3
        if kind_sp == 0:
            spatial = PointSpatialModel(lon_0=lon, lat_0=lat, frame=FRAME)
        elif kind_sp == 1:
            spatial = GaussianSpatialModel(lon_0=lon, lat_0=lat, sigma=ext, frame=FRAME)
        else:
            spatial = DiskSpatialModel(lon_0=lon, lat_0=lat, r_0=ext, frame=FRAME)
        if kind_sc == 0:
            spectral = PowerLawSpectralModel(index=index, amplitude=amp, reference=1)
        elif kind_sc == 1:
            spectral = LogParabolaSpectralModel(amplitude=amp, reference=1, alpha=index, beta=0.15)
        else:
            spectral = ExpCutoffPowerLawSpectralModel(index=index - 0.4, amplitude=amp, reference=1, lambda_=0.05)
        models.append(SkyModel(spatial_model=spatial, spectral_model=spectral, name=f"src-{i:02d}"))
    models.append(FoVBackgroundModel(dataset_name=dataset_name))
    return models


def make_c3(name="c3", width=(40.0, 10.0), n_sources=50, n_reco=30, n_true=60, **kwargs):
    """Galactic-plane survey dataset 2000 x 500 px, 30 reco / 60 true bins, 50 sources."""
    ds = make_dataset(width=width, n_reco=n_reco, n_true=n_true, e_reco=(0.1, 100.0), e_true=(0.03, 300.0),
                      name=name, mask_radius=None, **kwargs)
    half_w, half_h = width[0] / 2 - 1.0, width[1] / 2 - 2.0
    ds.models = survey_models(n_sources, lon_range=(-half_w, half_w), lat_range=(-max(half_h, 0.5), max(half_h, 0.5)),
                              dataset_name=name)
    return ds


def make_c4(n_obs=64, seed=100, indices=None, return_models=False, **kwargs):
    """Joint fit: ``n_obs`` observations of the C1 field, shared source model, per-dataset background norm,
    exposure scale U(0.5, 1.5), pointing offsets N(0, 0.5 deg).

    ``indices``: build only these observations (the shard of one rank); the random draws of all ``n_obs`` are
    made regardless, so observation ``i`` is the same dataset whichever rank builds it.  ``return_models``: also
    return the models of ALL observations in the global order (shared source, then every background model; for
    observations not built here a model object without data) -- what ``ShardedDatasets(global_models=...)`` needs."""
    rng = np.random.RandomState(seed)
    shared = c1_models("unused")[0]
    draws = [(float(rng.uniform(0.5, 1.5)), tuple(rng.normal(0, 0.5, size=2))) for _ in range(n_obs)]
    keep = set(range(n_obs)) if indices is None else set(int(i) for i in indices)
    datasets, backgrounds = [], []
    for i, (scale, offset) in enumerate(draws):
        name = f"obs-{i:03d}"
        bkg = FoVBackgroundModel(dataset_name=name)
        backgrounds.append(bkg)
        if i in keep:
            ds = make_dataset(name=name, exposure_scale=scale, pointing_offset=offset, **kwargs)
            ds.models = [shared, bkg]
            datasets.append(ds)
    out = Datasets(datasets)
    if return_models:
        return out, [shared] + backgrounds
    return out
