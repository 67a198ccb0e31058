"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/voxelis_b200.h declares (no compute calls — there is no GPU in the CPU suite)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "voxelis_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vx_[a-z0-9_]+)\s*\(", src)))


def test_header_and_python_binding_agree():
    from voxelis_b200 import api
    assert declared_symbols() == sorted(api.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    from voxelis_b200 import build
    lib_path = build.build()          # nvcc cross-compiles for sm_100a without a GPU
    L = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(L, name), name
    L.vx_abi_version.restype = ctypes.c_int
    assert L.vx_abi_version() == 1


def test_no_cpu_fallback_without_device():
    """With no CUDA device the product must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import voxelis_b200 as vx
    with pytest.raises(vx.VoxelisError):
        vx.VoxInterner.with_memory_budget(1 << 20)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "voxelis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "workloads.py", f


def test_cpp_mirror_compiles_and_links():
    """include/voxelis_b200.hpp (C++ mirror of the Rust API) + the README quick start compile with g++
    and link against the C-ABI library; the binary is run on the GPU box by test_gpu_cpp_mirror.py."""
    import subprocess
    from voxelis_b200 import build
    build.build()
    out = os.path.join(ROOT, "tests", "cpp", "readme_quickstart")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "readme_quickstart.cpp"), "-o", out,
           "-L", os.path.join(ROOT, "voxelis_b200"), "-lvoxelis_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "voxelis_b200")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert os.path.exists(out)
