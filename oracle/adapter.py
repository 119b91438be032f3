"""Oracle adapter (TEST INFRASTRUCTURE ONLY): turn the product's host-side objects (numpy maps, model
parameter values) into the plain arrays / dicts the oracle restatement consumes.  Only reads attributes;
the oracle recomputes every geometry / model quantity itself."""
import numpy as np

from .evaluator import Dataset, DatasetEvaluator
from .wcs import ImageGeom

_SPATIAL = {"PointSpatialModel": "point", "GaussianSpatialModel": "gauss", "DiskSpatialModel": "disk"}
_SPECTRAL = {"PowerLawSpectralModel": "pl", "LogParabolaSpectralModel": "lp", "ExpCutoffPowerLawSpectralModel": "ecpl",
             "PowerLawNormSpectralModel": "pl-norm"}


def geom_to_oracle(geom):
    return ImageGeom(geom.npix[0], geom.npix[1], geom.binsz, geom.crval[0], geom.crpix[0], geom.crpix[1], geom.crval[1],
                     proj=geom.projection)


def model_to_oracle(model, map_frame=None):
    if type(model).__name__ == "FoVBackgroundModel":
        spec = {"type": "pl-norm", **{p.name: p.value for p in model.spectral_model.parameters}}
        return {"type": "fov-bkg", "name": model.name, "spectral": spec}
    spatial = {"type": _SPATIAL[type(model.spatial_model).__name__],
               **{p.name: p.value for p in model.spatial_model.parameters},
               "frame": model.spatial_model.frame, "map_frame": map_frame}
    spectral = {"type": _SPECTRAL[type(model.spectral_model).__name__],
                **{p.name: p.value for p in model.spectral_model.parameters}}
    return {"name": model.name, "spatial": spatial, "spectral": spectral, "apply_irf": dict(model.apply_irf)}


def dataset_to_oracle(ds):
    """gammapy_b200.MapDataset -> oracle.evaluator.Dataset (the IRF kernels are the identical static arrays)."""
    geom = ds._geom
    e_true = ds.exposure.geom.axes["energy_true"]
    e_reco = geom.axes["energy"]
    psf_by_position = ds.psf is not None and not getattr(ds.psf, "has_single_spatial_bin", True)
    edisp_by_position = ds.edisp is not None and not getattr(ds.edisp, "has_single_spatial_bin", True)
    psf = None if psf_by_position else ds._psf_kernel_for_evaluators()
    irf = {}
    if psf_by_position or edisp_by_position:
        from .irf import IrfGeom

        ig = (ds.psf if psf_by_position else ds.edisp).geom
        irf["irf_geom"] = IrfGeom(ig.npix[0], ig.npix[1], ig.binsz, ig.crval[0], ig.crpix[0], ig.crpix[1],
                                  crval_lat=ig.crval[1], proj=ig.projection)
        if psf_by_position:
            # the library holds the map in float32 (gpy_irf_maps_desc.psf_map): same numbers on both sides
            irf["psf_map"] = np.asarray(ds.psf.data, dtype=np.float32).astype(float)
            irf["rad_edges"] = np.asarray(ds.psf.rad_axis.edges, dtype=float)
        if edisp_by_position:
            irf["edisp_map"] = np.asarray(ds.edisp.data, dtype=float)
    if ds.edisp is not None:
        edisp = np.asarray(ds.edisp.get_edisp_kernel(energy_axis=e_reco).pdf_matrix, dtype=float)
    else:
        from gammapy_b200.irf import EDispKernel

        edisp = EDispKernel.from_diagonal_response(e_true, e_reco).pdf_matrix
    return Dataset(
        geom=geom_to_oracle(geom), energy_edges=e_reco.edges, energy_true_edges=e_true.edges,
        energy_center=e_reco.center, exposure=np.asarray(ds.exposure.data),
        background=None if ds.background is None else np.asarray(ds.background.data),
        counts=None if ds.counts is None else np.asarray(ds.counts.data),
        mask_safe=None if ds.mask_safe is None else np.asarray(ds.mask_safe.data),
        mask_fit=None if ds.mask_fit is None else np.asarray(ds.mask_fit.data),
        psf_kernel=None if psf is None else np.asarray(psf.data, dtype=float), edisp=edisp, name=ds.name, **irf,
    )


def evaluator_for(ds, conv="fft32"):
    """DatasetEvaluator with the dataset's current models / parameter values."""
    return DatasetEvaluator(dataset_to_oracle(ds), [model_to_oracle(m, ds._geom.frame) for m in (ds.models or [])], conv=conv)


def refresh_models(de, ds):
    """Copy the current parameter values into an existing oracle evaluator (keeps its cache state)."""
    for m in (ds.models or []):
        new = model_to_oracle(m, ds._geom.frame)
        if new.get("type") == "fov-bkg":
            de.bkg_model["spectral"].update(new["spectral"])
        else:
            ev = de.evaluators[new["name"]]
            ev.model["spatial"].update(new["spatial"])
            ev.model["spectral"].update(new["spectral"])
