"""Runs the C++ README quick start (tests/cpp/readme_quickstart.cpp over include/voxelis_b200.hpp)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_readme_quickstart_cpp():
    exe = os.path.join(ROOT, "tests", "cpp", "readme_quickstart")
    if not os.path.exists(exe):
        from voxelis_b200 import build
        build.build()
        subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), exe + ".cpp", "-o", exe,
                        "-L", os.path.join(ROOT, "voxelis_b200"), "-lvoxelis_b200",
                        "-Wl,-rpath," + os.path.join(ROOT, "voxelis_b200")], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and res.stdout.strip() == "ok", (res.returncode, res.stdout, res.stderr)
