"""Builds voxelis_b200/libvoxelis_b200.so (the C-ABI library) with nvcc for sm_100a, in-tree."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libvoxelis_b200.so")
SOURCES = ["vx_capi.cu"]
HEADERS = ["vx_device.cuh", "vx_build.cuh", "vx_bulk.cuh", "vx_read.cuh", "vx_release.cuh", "vx_dedup.cuh", "vx_stage.cuh", "vx_vtm.cuh", "vx_occupancy.cuh", "vx_terrain.cuh", "vx_voxelize.cuh", "vx_world.cuh", "vx_capi_batch.inl", "vx_capi_voxelize.inl", "vx_capi_vtm.inl",
           os.path.join("..", "..", "include", "voxelis_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-ldl"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    extra = os.environ.get("VX_NVCC_EXTRA", "").split()     # e.g. -DVX_MIN_CTAS=4 for A/B builds
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, *(["-Xptxas", "-v"] if verbose else []), "-o", LIB,
           *[os.path.join(CSRC, s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
