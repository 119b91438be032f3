"""Host-side driver of libgpy_b200 for a list of MapDatasets sharing one parameter vector.

Plays the part of ``DatasetsActor`` (datasets/actors.py:16-87) in the reference: the free parameter
*values* go out, the per-dataset Cash sums come back.  Cubes are mirrored once into HBM as torch tensors
(device memory + stream plumbing only); every likelihood evaluation is one C-ABI call.
"""
import ctypes as C

import numpy as np

from . import _lib
from .modeling.parameter import bind_mirror, release_mirror
from .modeling.models import (DiskSpatialModel, ExpCutoffPowerLawSpectralModel, FoVBackgroundModel,
                              GaussianSpatialModel, LogParabolaSpectralModel, PointSpatialModel,
                              PowerLawSpectralModel, SkyModel, integration_nodes)

_SPATIAL_ORDER = {
    PointSpatialModel: (0, ["lon_0", "lat_0"]),
    GaussianSpatialModel: (1, ["lon_0", "lat_0", "sigma", "e", "phi"]),
    DiskSpatialModel: (2, ["lon_0", "lat_0", "r_0", "e", "phi", "edge_width"]),
}
_SPECTRAL_ORDER = {
    PowerLawSpectralModel: (0, ["index", "amplitude", "reference"]),
    LogParabolaSpectralModel: (1, ["amplitude", "reference", "alpha", "beta"]),
    ExpCutoffPowerLawSpectralModel: (2, ["index", "amplitude", "reference", "lambda_", "alpha"]),
}


def _ptr(arr):
    return C.c_void_p(arr.__array_interface__["data"][0])


class DeviceDataset:
    """HBM mirror + C handle of one MapDataset (exposure / background / counts float32, mask uint8)."""

    def __init__(self, dataset, ctx):
        import torch

        self.ctx = ctx
        self.dataset = dataset
        self.handle = None
        self.models_owner = None
        self.torch = torch
        self._keep = {}
        self._uploaded = {}
        self._held = ()
        self._build()

    _shared = {}  # (device, id of the source array) -> (weakref to the array, tensor): one upload per shared map

    def _to_device(self, name, arr, dtype, share=False):
        """HBM copy of ``arr``.  ``share``: datasets holding the very same numpy array (one exposure map for the
        observations of one pointing grid) get the same tensor."""
        import weakref

        key = (self.ctx.device, id(arr), np.dtype(dtype).str)
        if share:
            hit = DeviceDataset._shared.get(key)
            if hit is not None and hit[0]() is arr:
                self._keep[name] = hit[1]
                return hit[1]
        t = self.torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(self.ctx.torch_device)
        self._keep[name] = t
        if share:
            for k in [k for k, v in DeviceDataset._shared.items() if v[0]() is None]:
                del DeviceDataset._shared[k]
            try:
                DeviceDataset._shared[key] = (weakref.ref(arr), t)
            except TypeError:
                pass
        return t

    def _build(self):
        ds = self.dataset
        geom = ds._geom
        if geom.frame is None:
            raise ValueError("MapDataset geometry without a frame")
        e_true = ds.exposure.geom.axes["energy_true"]
        e_reco = geom.axes["energy"]
        g = _lib.geom_desc(geom, e_true.nbin, e_reco.nbin)
        if ds.exposure.geom.to_image() != geom.to_image():
            raise NotImplementedError("exposure and counts must share the image geometry")
        desc = _lib.DatasetDesc()
        desc.geom = g
        host = {}
        host["edges"] = np.ascontiguousarray(e_true.edges, dtype=float)
        host["ecenter"] = np.ascontiguousarray(e_reco.center, dtype=float)
        host["nodes"] = integration_nodes(host["edges"])
        desc.energy_true_edges = _ptr(host["edges"])
        desc.energy_reco_center = _ptr(host["ecenter"])
        desc.nodes = _ptr(host["nodes"])
        desc.n_nodes = host["nodes"].shape[1]
        desc.exposure = self._to_device("exposure", ds.exposure.data, np.float32, share=True).data_ptr()
        if ds.background is not None:
            desc.background = self._to_device("background", ds.background.data, np.float32).data_ptr()
        if ds.counts is not None:
            desc.counts = self._upload_counts().data_ptr()
        mask = ds.mask
        if mask is not None:
            mdata = np.asarray(mask.data)
            if mdata.dtype != bool:
                # WeightedCashFitStatistic (stats/fit_statistics.py:304-321): mask = ~(mask == False), weights = mask[mask]
                if ds.stat_type != "cash_weighted":
                    raise ValueError("float masks need stat_type='cash_weighted'")
                desc.weight = self._to_device("weight", mdata, np.float32).data_ptr()
                mdata = ~(mdata == False)  # noqa: E712
            elif ds.stat_type == "cash_weighted":
                desc.weight = self._to_device("weight", mdata, np.float32).data_ptr()
            if ds.storage == "compact":
                # one bit per bin, bit (p & 7) of byte (p >> 3) for pixel p = y nx + x, rows padded to 16 bytes
                P = mdata.shape[1] * mdata.shape[2]
                row_bytes = (P + 127) // 128 * 16
                bits = np.zeros((mdata.shape[0], row_bytes), dtype=np.uint8)
                packed = np.packbits(mdata.reshape(mdata.shape[0], P).astype(bool), axis=1, bitorder="little")
                bits[:, :packed.shape[1]] = packed
                desc.mask = self._to_device("mask", bits, np.uint8).data_ptr()
                desc.mask_row_bytes = row_bytes
            else:
                desc.mask = self._to_device("mask", mdata, np.uint8).data_ptr()
            host["mask_image"] = np.ascontiguousarray(np.logical_or.reduce(mdata, axis=0), dtype=np.uint8)
            desc.mask_image = _ptr(host["mask_image"])
        desc.compact = int(ds.storage == "compact")
        psf_by_position = ds.psf is not None and not getattr(ds.psf, "has_single_spatial_bin", True)
        edisp_by_position = ds.edisp is not None and not getattr(ds.edisp, "has_single_spatial_bin", True)
        psf_kernel = None if psf_by_position else ds._psf_kernel_for_evaluators()
        if psf_kernel is not None:
            kdata = np.ascontiguousarray(psf_kernel.data, dtype=np.float32)  # ndmap.py:961
            if kdata.ndim != 3 or kdata.shape[0] != e_true.nbin or kdata.shape[1] != kdata.shape[2]:
                raise ValueError("PSF kernel must have shape (n_energy_true, K, K)")
            host["psf"] = kdata
            desc.psf_kernel = _ptr(kdata)
            desc.psf_k = kdata.shape[-1]
        from .irf import EDispKernel

        diag = EDispKernel.from_diagonal_response(e_true, e_reco).pdf_matrix
        host["diag"] = np.ascontiguousarray(diag, dtype=float)
        if ds.edisp is not None:
            kernel = ds.edisp.get_edisp_kernel(energy_axis=e_reco)
            host["edisp"] = np.ascontiguousarray(kernel.pdf_matrix, dtype=float)
            desc.edisp_diagonal = _ptr(host["diag"])
        else:
            host["edisp"] = host["diag"]
        if host["edisp"].shape != (e_true.nbin, e_reco.nbin):
            raise ValueError("edisp matrix shape does not match the energy axes")
        desc.edisp = _ptr(host["edisp"])
        self.torch.cuda.synchronize(self.ctx.torch_device)  # uploads ran on torch's stream
        handle = C.c_void_p()
        _lib.check(self.ctx.lib.gpy_dataset_create(self.ctx.handle, C.byref(desc), C.byref(handle)))
        self.handle = handle
        if psf_by_position or edisp_by_position:
            self._set_irf_maps(ds, psf_by_position, edisp_by_position, e_true, e_reco)
        self._uploaded = ds._data_fingerprint()
        self._held = ds._data_arrays()

    def _set_irf_maps(self, ds, psf_by_position, edisp_by_position, e_true, e_reco):
        """PSFMap / EDispKernelMap with a spatial grid: the evaluators look their IRFs up at the source position
        (``gpy_dataset_set_irf_maps``; MapEvaluator.update, datasets/evaluator.py:196-227)."""
        grids = [m.geom for m, on in ((ds.psf, psf_by_position), (ds.edisp, edisp_by_position)) if on]
        ig = grids[0]
        if len(grids) == 2 and grids[1] != ig:
            raise NotImplementedError("PSF map and edisp map must share their image geometry")
        if ig.frame != ds._geom.frame:
            raise NotImplementedError("IRF maps must be in the frame of the dataset geometry")
        m = _lib.IrfMapsDesc()
        m.nx, m.ny, m.binsz = ig.npix[0], ig.npix[1], ig.binsz
        m.crval_lon, m.crpix_x, m.crpix_y = ig.crval[0], ig.crpix[0], ig.crpix[1]
        m.crval_lat, m.proj = ig.crval[1], _lib.PROJECTIONS[ig.projection]
        keep = {}
        if psf_by_position:
            if ds.psf.energy_axis_true.nbin != e_true.nbin:
                raise ValueError("PSF map energy axis does not match the exposure")
            keep["rad"] = np.ascontiguousarray(ds.psf.rad_axis.edges, dtype=float)
            keep["psf"] = np.ascontiguousarray(ds.psf.data, dtype=np.float32)
            m.n_rad, m.rad_edges, m.psf_map = ds.psf.rad_axis.nbin, _ptr(keep["rad"]), _ptr(keep["psf"])
        if edisp_by_position:
            if ds.edisp.data.shape[:2] != (e_true.nbin, e_reco.nbin):
                raise ValueError("edisp map axes do not match the dataset")
            keep["edisp"] = np.ascontiguousarray(ds.edisp.data, dtype=float)
            m.edisp_map = _ptr(keep["edisp"])
        _lib.check(self.ctx.lib.gpy_dataset_set_irf_maps(self.handle, C.byref(m)))

    def source_irf(self, source, with_kernel=True):
        """(PSF kernel (nT, K, K) float32 or None, K, IRF pixel of the edisp matrix, number of IRF look-ups) of one
        evaluator of a dataset with position-dependent IRFs (``gpy_dataset_source_psf_kernel``)."""
        k, pix, n = C.c_int32(0), C.c_int32(-1), C.c_int32(0)
        _lib.check(self.ctx.lib.gpy_dataset_source_psf_kernel(self.handle, source, None, 0, C.byref(k), C.byref(pix),
                                                              C.byref(n)))
        kernel = None
        if with_kernel and k.value > 0 and n.value > 0:
            nT = self.dataset.exposure.geom.axes["energy_true"].nbin
            kernel = np.zeros((nT, k.value, k.value), dtype=np.float32)
            _lib.check(self.ctx.lib.gpy_dataset_source_psf_kernel(self.handle, source, _ptr(kernel), kernel.size,
                                                                  C.byref(k), C.byref(pix), C.byref(n)))
        return kernel, k.value, pix.value, n.value

    def _upload_counts(self):
        counts = np.asarray(self.dataset.counts.data)
        if self.dataset.storage == "compact":
            c16 = counts.astype(np.uint16)
            if not np.array_equal(c16, counts):
                raise ValueError("compact storage holds integer counts in [0, 65535] only")
            return self._to_device("counts", c16, np.uint16)
        c32 = counts.astype(np.float32)
        if not np.array_equal(c32, counts, equal_nan=True):
            raise NotImplementedError("counts are not exactly representable in float32")
        return self._to_device("counts", c32, np.float32)

    def refresh_counts(self):
        t = self._upload_counts()
        self.torch.cuda.synchronize(self.ctx.torch_device)
        _lib.check(self.ctx.lib.gpy_dataset_set_counts(self.handle, C.c_void_p(t.data_ptr())))
        self._uploaded = self.dataset._data_fingerprint()
        self._held = self.dataset._data_arrays()

    def close(self):
        if self.handle is not None:
            self.ctx.lib.gpy_dataset_destroy(self.handle)
            self.handle = None
        self._keep.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LikelihoodEngine:
    """All datasets of a ``Datasets`` object on one GPU, evaluated through one parameter vector."""

    def __init__(self, datasets, device=None):
        self.ctx = _lib.Context.get(device)
        self.datasets = list(datasets)
        self.parameters = []   # unique Parameter objects, column order of the parameter vector
        self._col = {}
        for ds in self.datasets:
            for model in (ds.models or []):
                for par in model.parameters:
                    if id(par) not in self._col:
                        self._col[id(par)] = len(self.parameters)
                        self.parameters.append(par)
        self.n_par = max(len(self.parameters), 1)
        self._vec = bind_mirror(self.parameters)  # kept equal to the parameter values by their setters
        self._vec_ptr = _ptr(self._vec)
        self._stat1 = np.empty(1, dtype=float)
        self._stat1_ptr = _ptr(self._stat1)
        for ds in self.datasets:
            dev = ds._device_dataset(self.ctx)
            self._set_models(ds, dev)
        self._handles = (C.c_void_p * len(self.datasets))(*[ds._device.handle for ds in self.datasets])
        self._signature = self.signature(self.datasets)

    @staticmethod
    def signature(datasets):
        return tuple((id(ds), ds._models_version, id(ds._device)) for ds in datasets)

    def _set_models(self, ds, dev):
        srcs, names = [], []
        bkg = _lib.BackgroundDesc(0, -1, -1, -1)
        for model in (ds.models or []):
            if isinstance(model, FoVBackgroundModel):
                sm = model.spectral_model
                bkg = _lib.BackgroundDesc(1, self._col[id(sm.norm)], self._col[id(sm.tilt)], self._col[id(sm.reference)])
                continue
            if not isinstance(model, SkyModel):
                raise NotImplementedError(f"{type(model).__name__} is outside the accelerated path")
            if not model.apply_irf.get("exposure", True):
                raise NotImplementedError("apply_irf['exposure']=False is outside the accelerated path")
            try:
                stype, snames = _SPATIAL_ORDER[type(model.spatial_model)]
                ptype, pnames = _SPECTRAL_ORDER[type(model.spectral_model)]
            except KeyError as exc:
                raise NotImplementedError(f"model class {exc.args[0].__name__} is outside the accelerated path") from None
            d = _lib.SourceDesc()
            d.spatial_type, d.spectral_type = stype, ptype
            for k in range(6):
                d.spatial_par[k] = self._col[id(getattr(model.spatial_model, snames[k]))] if k < len(snames) else -1
            for k in range(5):
                d.spectral_par[k] = self._col[id(getattr(model.spectral_model, pnames[k]))] if k < len(pnames) else -1
            d.apply_psf = int(bool(model.apply_irf.get("psf", True)))
            d.apply_edisp = int(bool(model.apply_irf.get("edisp", True)))
            d.frame = _lib.frame_code(model.frame)  # (a frame other than the map's: coordinates are rotated on the device)
            srcs.append(d)
            names.append(model.name)
        arr = (_lib.SourceDesc * max(len(srcs), 1))(*srcs)
        _lib.check(self.ctx.lib.gpy_dataset_set_models(dev.handle, arr, len(srcs), C.byref(bkg)))
        ds._source_names = names
        dev.models_owner = self

    def _claim(self):
        """The C handle of a dataset holds ONE model table: (re)install this engine's parameter columns when
        another engine (e.g. ``MapDataset.npred`` vs ``Datasets.stat_sum``) configured it last."""
        for ds in self.datasets:
            if ds._device is None:
                raise RuntimeError("stale LikelihoodEngine: the dataset's maps or storage format changed after the engine "
                                   "was built; get a new one (Datasets._get_engine / MapDataset._get_engine)")
            if ds._device.models_owner is not self:
                self._set_models(ds, ds._device)

    # -- evaluation ----------------------------------------------------------------------------------
    def values(self):
        """Current parameter values as the [n_par] vector."""
        if not self.parameters:
            return np.zeros(self.n_par, dtype=float)
        return self._vec.copy()

    def __del__(self):
        try:
            release_mirror(self.parameters, self._vec)
        except Exception:
            pass

    def stat_sum(self, values=None, per_dataset=False):
        """Datasets.stat_sum for one ([n_par]) or many ([B, n_par]) parameter vectors."""
        if values is None and not per_dataset and self.parameters:
            # the step of a fit: the bound vector goes out as it is (the library copies it before it returns), the
            # result comes back into a one-element array kept for the purpose
            self._claim()
            _lib.check(self.ctx.lib.gpy_stat_sum(self.ctx.handle, self._handles, len(self.datasets), self._vec_ptr, 1,
                                                 self.n_par, self._stat1_ptr, None))
            return float(self._stat1[0])
        if values is None:
            values = self.values()
        values = np.ascontiguousarray(values, dtype=float)
        single = values.ndim == 1
        v2 = values.reshape(1, -1) if single else values
        if v2.shape[1] != self.n_par:
            raise ValueError(f"expected {self.n_par} parameter values per vector, got {v2.shape[1]}")
        B, D = v2.shape[0], len(self.datasets)
        self._claim()
        stat = np.empty(B, dtype=float)
        per = np.empty((B, D), dtype=float) if per_dataset else None
        _lib.check(self.ctx.lib.gpy_stat_sum(self.ctx.handle, self._handles, D, _ptr(v2), B, self.n_par, _ptr(stat),
                                             _ptr(per) if per_dataset else None))
        if per_dataset:
            return (stat[0], per[0]) if single else (stat, per)
        return float(stat[0]) if single else stat

    def stat_sum_device(self, values, out):
        """Asynchronous variant: per-(vector, dataset) stats into the CUDA tensor ``out`` [B, D] float64."""
        v2 = np.ascontiguousarray(values, dtype=float).reshape(-1, self.n_par)
        self._claim()
        _lib.check(self.ctx.lib.gpy_stat_sum_device(self.ctx.handle, self._handles, len(self.datasets), _ptr(v2),
                                                    v2.shape[0], self.n_par, C.c_void_p(out.data_ptr())))

    def stat_sum_joint_device(self, values, out):
        """Asynchronous joint statistic over the ranks of a sharded fit into the CUDA tensor ``out`` [B] float64: this
        rank's datasets + the one-shot all-reduce over NVLink peer memory (``gpy_stat_sum_joint_device``)."""
        v2 = np.ascontiguousarray(values, dtype=float).reshape(-1, self.n_par)
        self._claim()
        _lib.check(self.ctx.lib.gpy_stat_sum_joint_device(self.ctx.handle, self._handles, len(self.datasets), _ptr(v2),
                                                          v2.shape[0], self.n_par, C.c_void_p(out.data_ptr())))

    def stat_sum_joint(self, values):
        """Joint statistic over the ranks of a sharded fit as a numpy array [B] (``gpy_stat_sum_joint``: evaluation,
        one-shot all-reduce over NVLink peer memory, copy to the host, one synchronisation -- all inside the library)."""
        v2 = np.ascontiguousarray(values, dtype=float).reshape(-1, self.n_par)
        self._claim()
        out = np.empty(v2.shape[0], dtype=float)
        _lib.check(self.ctx.lib.gpy_stat_sum_joint(self.ctx.handle, self._handles, len(self.datasets), _ptr(v2),
                                                   v2.shape[0], self.n_par, _ptr(out)))
        return out

    def npred(self, index, values=None, source_enable=None, include_background=True):
        """npred cube of dataset ``index`` as a float64 numpy array."""
        return self.npred_device(index, values, source_enable, include_background).cpu().numpy()

    def npred_device(self, index, values=None, source_enable=None, include_background=True):
        """npred cube of dataset ``index`` as a float64 CUDA tensor (complete when the call returns)."""
        import torch

        if values is None:
            values = self.values()
        values = np.ascontiguousarray(values, dtype=float)
        ds = self.datasets[index]
        self._claim()
        out = torch.empty(ds.data_shape, dtype=torch.float64, device=self.ctx.torch_device)
        en = None
        if source_enable is not None:
            en = np.ascontiguousarray(source_enable, dtype=np.uint8)
        _lib.check(self.ctx.lib.gpy_npred(self.ctx.handle, ds._device.handle, _ptr(values), self.n_par,
                                          _ptr(en) if en is not None else None, int(include_background),
                                          C.c_void_p(out.data_ptr())))
        self.ctx.stream.synchronize()
        return out

    def source_state(self, index, source):
        st = _lib.SourceState()
        self._claim()
        _lib.check(self.ctx.lib.gpy_dataset_source_state(self.datasets[index]._device.handle, source, C.byref(st)))
        return {f: getattr(st, f) for f, _ in st._fields_}
