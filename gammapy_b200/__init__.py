"""gammapy_b200: gammapy's forward-fold likelihood hot path (MapEvaluator.compute_npred + Cash) on B200.

The public surface mirrors the gammapy classes on that path (``MapDataset``, ``Datasets``, ``SkyModel`` ...);
all compute goes through ``libgpy_b200.so`` (hand-written sm_100a CUDA behind a C ABI, include/gpy_b200.h).
"""
from .datasets import Datasets, MapDataset, MapDatasetOnOff  # noqa: F401
from .irf import EDispKernel, EDispKernelMap, PSFKernel, PSFMap  # noqa: F401
from .maps import Map, MapAxis, WcsGeom, WcsNDMap  # noqa: F401
from .modeling import Fit, Parameter, Parameters  # noqa: F401
from .modeling.models import (  # noqa: F401
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
)
from .utils import SkyCoord  # noqa: F401

__version__ = "0.1.0"
