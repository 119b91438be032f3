"""Compile ``libgpy_b200.so`` (sm_100a) in-tree with nvcc, on first use or when a source is newer (no JIT cache).

The built library is git-ignored (history stays source-only) but is part of the working tree ``gpurun`` ships."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgpy_b200.so")
SOURCES = [os.path.join(CSRC, "gpy_b200.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "kernels.cuh"), os.path.join(CSRC, "tsmap.cuh"), os.path.join(HERE, "..", "include", "gpy_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--shared",
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off",
    "-Xptxas", "-v",
    "--cudart", "shared",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in DEPS if os.path.exists(p))


def build(force=False, verbose=False):
    """Build the shared library when sources are newer; returns its path."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("GPY_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + SOURCES + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(" ".join(cmd))
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libgpy_b200.so")
    with open(os.path.join(HERE, "csrc", "ptxas_info.txt"), "w") as fh:
        # registers / spills / shared memory per kernel (-Xptxas -v); compile times change from build to build
        fh.write("".join(l for l in res.stderr.splitlines(keepends=True) if "Compile time" not in l))
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose=True))
