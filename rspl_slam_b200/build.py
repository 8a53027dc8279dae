"""In-tree build of the CUDA library (nvcc, sm_100a only). The built .so is git-ignored but
travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "librspl_ba.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared"]


CHECKED_LIB_PATH = os.path.join(_HERE, "..", "build", "checked", "librspl_ba_checked.so")


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise FileNotFoundError("nvcc not found")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(_HERE, "..", "include", "rspl_ba.h")]


def build_library(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> rspl_slam_b200/librspl_ba.so"""
    if not force and os.path.exists(LIB_PATH):
        t = os.path.getmtime(LIB_PATH)
        if all(os.path.getmtime(s) <= t for s in sources()):
            return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(CSRC, "capi.cu")]
    env = dict(os.environ)
    env.pop("CXX", None)  # the image exports CXX=/opt/gcc/bin/g++, which lacks parts of its runtime
    env.pop("CC", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


def build_checked(force: bool = False) -> str:
    """The same library with -DRSPL_BA_CHECKED (bounds / invariant asserts in the kernels, ba_math.cuh: BA_CHECK) into
    build/checked/ (git-ignored, travels to the GPU box); loaded through RSPL_BA_LIB by tests/test_checked_build.py."""
    out = os.path.abspath(CHECKED_LIB_PATH)
    if not force and os.path.exists(out) and all(os.path.getmtime(s) <= os.path.getmtime(out) for s in sources()):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-DRSPL_BA_CHECKED", "-o", out, os.path.join(CSRC, "capi.cu")]
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("nvcc (checked build) failed:\n" + r.stdout + r.stderr)
    return out
