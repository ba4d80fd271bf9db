"""In-tree build of libgritas_b200.so (nvcc cross-compiles sm_100a without a GPU).

The product has no CPU implementation: if this library is missing or cannot be loaded,
gritas_b200.api raises -- it never falls back to anything under oracle/.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.environ.get("GRITAS_B200_LIBDIR") or os.path.join(HERE, "lib")   # override: experiments with alternative builds
LIB = os.path.join(LIBDIR, "libgritas_b200.so")
SOURCES = ["gritas_b200.cu", "xcov.cu"]
HEADERS = ["kernels.cuh", os.path.join("..", "..", "include", "gritas_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: gritas_b200 needs the CUDA toolkit to build its sm_100a library")
    return p


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + os.environ.get("GRITAS_B200_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


DRIVER = os.path.join(HERE, "bin", "gritas_b200_driver")


def build_driver(force: bool = False) -> str:
    """C++ stand-in for the Fortran caller (csrc/gritas_driver.cpp + m_gritas_grids.hpp), linked against the .so."""
    src = [os.path.join(CSRC, "gritas_driver.cpp"), os.path.join(CSRC, "m_gritas_grids.hpp")]
    if not force and os.path.exists(DRIVER) and all(os.path.getmtime(DRIVER) >= os.path.getmtime(s) for s in src + [LIB]):
        return DRIVER
    os.makedirs(os.path.dirname(DRIVER), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", DRIVER, src[0], "-L" + LIBDIR, "-lgritas_b200",
                           "-Wl,-rpath," + LIBDIR, "-Wl,-rpath,$ORIGIN/../lib"])
    return DRIVER


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
    print(build_driver(force=True))
