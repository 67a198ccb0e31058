"""CPU-suite check of the occupancy kernels' per-word source: tests/cpp/occ_host_check.cu instantiates
occ_walk_line / occ_word (voxelis_b200/csrc/vx_occupancy.cuh, __host__ __device__) for the host and steps them over
pools downloaded from the oracle; the planes must equal the oracle's restatement of generate_occupancy_masks
(reference voxelis/src/utils/mesh.rs:418-596).  This is test infrastructure: the product library exports no host
compute, and the kernels proper are checked on the B200 (tests/test_gpu_occupancy.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_oracle_canonical import oracle_build
from test_oracle_occupancy import cell_offsets, chunk_set
from voxelis_b200 import workloads as wl

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp")


def load_stepper():
    out = os.path.join(HERE, "libocc_host_check.so")
    src = os.path.join(HERE, "occ_host_check.cu")
    hdr = os.path.join(HERE, "..", "..", "voxelis_b200", "csrc", "vx_occupancy.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        res = subprocess.run(["nvcc", "-std=c++17", "-O1", "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "20013",
                              "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, src],
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
    L = C.CDLL(out)
    L.occ_host_check.argtypes = [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] + [C.c_int] * 2 + [C.c_void_p] * 4
    L.occ_host_check_planes.argtypes = L.occ_host_check.argtypes
    return L


@pytest.fixture(scope="module")
def stepper():
    return load_stepper()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("entry", ["occ_host_check", "occ_host_check_planes"], ids=["word_owner", "morton_planes"])
@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("depth", [3, 5, 6])
def test_per_word_source_matches_oracle(stepper, oracle_api, depth, dtype, entry):
    masks, values = chunk_set(depth, dtype)
    it, roots, _ = oracle_build(oracle_api, depth, masks, values, dtype, budget=256 << 20)
    dl = it.download()
    children = np.ascontiguousarray(dl["children"], np.uint64)
    vals = np.ascontiguousarray(dl["values"].astype(wl.NP_DTYPE[dtype]))
    n = len(roots)
    for lod in range(0, depth + 1, 2 if depth > 3 else 1):
        ld = depth - lod
        S, G = 1 << ld, 64 >> ld
        per_builder = min(G ** 3, 5)
        for b0 in range(0, n, per_builder):
            idx = list(range(b0, min(b0 + per_builder, n)))
            offs = cell_offsets(S, len(idx), seed=b0 + lod)
            want = it.occupancy_masks(roots[idx], depth, offs, lod=lod)
            cell = np.zeros(G ** 3, np.uint64)
            for i, (ox, oy, oz) in zip(idx, offs):
                cell[((oy >> ld) * G + (oz >> ld)) * G + (ox >> ld)] = roots[i]
            M = 256
            ids, counts = np.zeros(M, np.uint64), np.zeros(M, np.uint64)
            glob = np.zeros(3 * 4096, np.uint64)
            pm = np.full((M, 3 * 4096), 0xDEADBEEF, np.uint64)      # every word of a live plane must be WRITTEN
            nm = getattr(stepper, entry)(_p(children), _p(vals), dtype, _p(cell), ld, M, _p(ids), _p(counts), _p(glob),
                                        _p(pm))
            assert nm == len(want["material_ids"]), (depth, lod, b0)
            assert np.array_equal(ids[:nm], want["material_ids"]) and np.array_equal(counts[:nm], want["material_counts"])
            assert np.array_equal(glob, want["global"]), (depth, lod, b0)
            assert np.array_equal(pm[:nm], want["per_material"]), (depth, lod, b0)
