"""Build the reference's own Cython Cash kernels into ``oracle/_ref`` (TEST INFRASTRUCTURE ONLY).

Compiles ``/root/reference/gammapy/stats/fit_statistics_cython.pyx`` *where it lies* (no source is
copied into the repo): ``cython`` writes the generated C into ``oracle/_ref/`` and gcc builds the
CPython extension there.  ``oracle/_ref/`` is git-ignored but travels to the GPU box with gpurun.

Run:  python -m oracle.build_ref         (no-op if /root/reference is absent or the .so is current)
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
PYX = "/root/reference/gammapy/stats/fit_statistics_cython.pyx"
MODNAME = "fit_statistics_cython"


def so_path():
    return os.path.join(REF_DIR, MODNAME + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False, verbose=False):
    """Return the path of the built extension, or None when the reference tree is not available."""
    out = so_path()
    if not os.path.exists(PYX):
        return out if os.path.exists(out) else None
    if os.path.exists(out) and not force and os.path.getmtime(out) >= os.path.getmtime(PYX):
        return out
    import numpy

    os.makedirs(REF_DIR, exist_ok=True)
    c_file = os.path.join(REF_DIR, MODNAME + ".c")
    cmds = [
        [sys.executable, "-m", "cython", "-3", PYX, "-o", c_file],
        ["gcc", "-O2", "-shared", "-fPIC", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
         "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(), c_file, "-o", out, "-lm"],
    ]
    for cmd in cmds:
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True, capture_output=not verbose)
    return out


def load():
    """Import the compiled reference module (building it first when possible); None if unavailable."""
    path = build()
    if path is None or not os.path.exists(path):
        return None
    import importlib.util

    spec = importlib.util.spec_from_file_location(MODNAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
