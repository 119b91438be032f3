"""ctypes binding of ``libgpy_b200.so`` (include/gpy_b200.h).  No CPU fallback: every compute entry
raises when the CUDA library or a CUDA device is missing."""
import ctypes as C
import os

from . import _build

_LIB = None


class GpyError(RuntimeError):
    """Failure reported by libgpy_b200 (CUDA error or unsupported configuration)."""


class GeomDesc(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("n_true", C.c_int32), ("n_reco", C.c_int32),
        ("binsz", C.c_double), ("crval_lon", C.c_double), ("crpix_x", C.c_double), ("crpix_y", C.c_double),
        ("crval_lat", C.c_double), ("proj", C.c_int32), ("frame", C.c_int32),
    ]


PROJECTIONS = {"CAR": 0, "TAN": 1}
FRAMES = {"galactic": 0, "icrs": 1}


def frame_code(frame):
    try:
        return FRAMES[frame]
    except KeyError:
        raise NotImplementedError(f"celestial frame {frame!r} (implemented: {sorted(FRAMES)})") from None


def geom_desc(geom, n_true=1, n_reco=1):
    """``gpy_geom_desc`` of a WcsGeom."""
    g = geom.to_image() if len(geom.axes) else geom
    return GeomDesc(g.npix[0], g.npix[1], n_true, n_reco, g.binsz, g.crval[0], g.crpix[0], g.crpix[1], g.crval[1],
                    PROJECTIONS[g.projection], frame_code(g.frame))


class DatasetDesc(C.Structure):
    _fields_ = [
        ("geom", GeomDesc),
        ("energy_true_edges", C.c_void_p),
        ("energy_reco_center", C.c_void_p),
        ("exposure", C.c_void_p),
        ("background", C.c_void_p),
        ("counts", C.c_void_p),
        ("mask", C.c_void_p),
        ("mask_image", C.c_void_p),
        ("psf_kernel", C.c_void_p),
        ("psf_k", C.c_int32),
        ("edisp", C.c_void_p),
        ("edisp_diagonal", C.c_void_p),
        ("nodes", C.c_void_p),
        ("n_nodes", C.c_int32),
        ("weight", C.c_void_p),
        ("compact", C.c_int32),
        ("mask_row_bytes", C.c_int64),
    ]


class SourceDesc(C.Structure):
    _fields_ = [
        ("spatial_type", C.c_int32), ("spectral_type", C.c_int32),
        ("spatial_par", C.c_int32 * 6), ("spectral_par", C.c_int32 * 5),
        ("apply_psf", C.c_int32), ("apply_edisp", C.c_int32), ("frame", C.c_int32),
    ]


class BackgroundDesc(C.Structure):
    _fields_ = [("enabled", C.c_int32), ("par_norm", C.c_int32), ("par_tilt", C.c_int32),
                ("par_reference", C.c_int32)]


class SourceState(C.Structure):
    _fields_ = [
        ("initialised", C.c_int32), ("contributes", C.c_int32),
        ("x0", C.c_int32), ("y0", C.c_int32), ("cx", C.c_int32), ("cy", C.c_int32),
        ("oversampling", C.c_int32), ("n_spatial_evals", C.c_int32),
    ]


class IrfMapsDesc(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32),
        ("binsz", C.c_double), ("crval_lon", C.c_double), ("crpix_x", C.c_double), ("crpix_y", C.c_double),
        ("crval_lat", C.c_double), ("proj", C.c_int32),
        ("n_rad", C.c_int32), ("rad_edges", C.c_void_p), ("psf_map", C.c_void_p), ("edisp_map", C.c_void_p),
    ]


# every symbol include/gpy_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "gpy_last_error": (C.c_char_p, []),
    "gpy_abi_version": (C.c_int, []),
    "gpy_init": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "gpy_finalize": (C.c_int, [_P]),
    "gpy_launch_count": (C.c_int64, [_P]),
    "gpy_kernel_timing": (C.c_int, [_P, C.c_int32]),
    "gpy_kernel_time": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "gpy_dataset_create": (C.c_int, [_P, C.POINTER(DatasetDesc), C.POINTER(_P)]),
    "gpy_dataset_destroy": (C.c_int, [_P]),
    "gpy_dataset_set_counts": (C.c_int, [_P, _P]),
    "gpy_dataset_set_models": (C.c_int, [_P, C.POINTER(SourceDesc), C.c_int32, C.POINTER(BackgroundDesc)]),
    "gpy_dataset_source_state": (C.c_int, [_P, C.c_int32, C.POINTER(SourceState)]),
    "gpy_dataset_set_irf_maps": (C.c_int, [_P, C.POINTER(IrfMapsDesc)]),
    "gpy_dataset_source_psf_kernel": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                                C.POINTER(C.c_int32)]),
    "gpy_stat_sum": (C.c_int, [_P, C.POINTER(_P), C.c_int32, _P, C.c_int32, C.c_int32, _P, _P]),
    "gpy_stat_sum_device": (C.c_int, [_P, C.POINTER(_P), C.c_int32, _P, C.c_int32, C.c_int32, _P]),
    "gpy_peer_reduce_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "gpy_peer_reduce_init": (C.c_int, [_P, C.POINTER(_P), C.c_int32, C.c_int32, C.c_int32]),
    "gpy_stat_sum_joint_device": (C.c_int, [_P, C.POINTER(_P), C.c_int32, _P, C.c_int32, C.c_int32, _P]),
    "gpy_stat_sum_joint": (C.c_int, [_P, C.POINTER(_P), C.c_int32, _P, C.c_int32, C.c_int32, _P]),
    "gpy_npred": (C.c_int, [_P, _P, _P, C.c_int32, _P, C.c_int32, _P]),
    "gpy_cash_sum": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "gpy_weighted_cash_sum": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P]),
    "gpy_wstat_sum_device": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, _P, C.c_int64, _P, _P]),
    "gpy_integrate_geom": (C.c_int, [_P, C.POINTER(GeomDesc), C.c_int32, C.c_int32, _P, _P, C.POINTER(C.c_int32)]),
    "gpy_psf_convolve": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32, _P]),
    "gpy_spectral_integral": (C.c_int, [_P, C.c_int32, _P, _P, C.c_int32, _P, C.c_int32, _P]),
    "gpy_solid_angle": (C.c_int, [_P, C.POINTER(GeomDesc), _P]),
    "gpy_f_cash_root": (C.c_int, [_P, C.c_double, _P, _P, _P, C.c_int64, _P]),
    "gpy_norm_bounds": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P]),
    "gpy_ts_map": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32, _P,
                             C.c_double, C.c_int32, C.c_double, _P]),
}


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """dlopen libgpy_b200.so and type every exported symbol.  Raises if the library cannot be had."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise GpyError(f"{path} is missing: build it with `python -m gammapy_b200._build`")
        _build.build()
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.gpy_abi_version() != 2:
        raise GpyError("libgpy_b200.so ABI version mismatch")
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().gpy_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        if rc == -3:
            raise NotImplementedError(msg)
        raise GpyError(msg)


class Context:
    """One libgpy_b200 context per (process, device) with its own torch-visible CUDA stream."""

    _instances = {}

    def __init__(self, device):
        import torch

        if not torch.cuda.is_available():
            raise GpyError("gammapy_b200 needs a CUDA device: there is no CPU fallback for the likelihood path")
        self.lib = load()
        self.device = int(device)
        self.torch_device = torch.device("cuda", self.device)
        with torch.cuda.device(self.device):
            # a dedicated stream: kernels, copies and the CUDA events that time them all live on it
            self.stream = torch.cuda.Stream(device=self.torch_device)
            handle = _P()
            check(self.lib.gpy_init(self.device, _P(self.stream.cuda_stream), C.byref(handle)))
        self.handle = handle

    @classmethod
    def get(cls, device=None):
        import torch

        if device is None:
            if not torch.cuda.is_available():
                raise GpyError("gammapy_b200 needs a CUDA device: there is no CPU fallback for the likelihood path")
            device = torch.cuda.current_device()
        device = int(device)
        if device not in cls._instances:
            cls._instances[device] = cls(device)
        return cls._instances[device]

    @property
    def launch_count(self):
        return int(self.lib.gpy_launch_count(self.handle))

    def kernel_timing(self, enable):
        """Bracket the fold + Cash kernel of every evaluation with CUDA events (measurement aid)."""
        check(self.lib.gpy_kernel_timing(self.handle, int(bool(enable))))

    def kernel_time(self):
        """(summed kernel duration [ms], launches) since ``kernel_timing(True)``."""
        ms, n = C.c_double(0.0), C.c_int64(0)
        check(self.lib.gpy_kernel_time(self.handle, C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)
