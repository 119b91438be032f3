"""The C-ABI library loads and exports every symbol include/gpy_b200.h declares (no compute without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "gpy_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpy_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    from gammapy_b200 import _lib

    assert header_functions() == sorted(_lib.SYMBOLS)


def test_library_exports_every_symbol():
    from gammapy_b200 import _lib

    lib = _lib.load()
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.gpy_abi_version() == 2


def test_no_cpu_fallback():
    """Without a CUDA device the context refuses to start (the product path has no CPU route)."""
    import ctypes as C

    import torch

    from gammapy_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    lib = _lib.load()
    handle = C.c_void_p()
    rc = lib.gpy_init(0, None, C.byref(handle))
    assert rc == -2 and handle.value is None
    assert b"no CPU fallback" in lib.gpy_last_error()
    with pytest.raises(_lib.GpyError):
        _lib.Context.get()
    from gammapy_b200 import configs

    with pytest.raises(_lib.GpyError):
        configs.make_c1(width=(1.0, 1.0), mask_radius=None).npred()


def test_struct_layout_matches_header():
    """ctypes mirrors of the C structs: sizes as the compiler lays them out (x86-64 SysV)."""
    import ctypes as C

    from gammapy_b200 import _lib

    assert C.sizeof(_lib.GeomDesc) == 64
    assert C.sizeof(_lib.SourceDesc) == 4 * (2 + 6 + 5 + 2 + 1)
    assert C.sizeof(_lib.BackgroundDesc) == 16
    assert C.sizeof(_lib.SourceState) == 32
    assert C.sizeof(_lib.DatasetDesc) == 64 + 8 * 8 + 8 + 8 * 3 + 8 + 8 + 8 + 8  # ... weight, compact (padded), mask_row_bytes
    assert C.sizeof(_lib.IrfMapsDesc) == 8 + 5 * 8 + 8 + 3 * 8  # nx, ny; binsz .. crval_lat; proj, n_rad; three pointers


def test_every_python_module_imports_without_a_gpu():
    """The host layer must import on a CPU-only machine (the driver's build check and the CPU test tier run there); the
    GPU is only touched when something is evaluated."""
    import importlib
    import pkgutil

    import gammapy_b200

    names = [m.name for m in pkgutil.walk_packages(gammapy_b200.__path__, "gammapy_b200.")
             if not m.name.rsplit(".", 1)[-1].startswith("lib")]
    assert len(names) >= 18
    for name in names:
        importlib.import_module(name)
