"""ctypes view of oracle/liboracle.so — the CPU ORACLE (test infrastructure).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl
reference`` legs may import this module.  The product package ``voxelis_b200`` never does.

The classes mirror the reference's API names (``VoxInterner.with_memory_budget``,
``VoxTree``, ``Batch.set/fill/clear``, ``apply_batch``, ``get``) — reference:
voxelis/src/spatial/voxops.rs:8-35 — so parity tests read the same for both sides.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

U8, I32 = 0, 1
_NP = {U8: np.uint8, I32: np.int32}

STATS_FIELDS = [
    "requested_budget", "actual_budget", "node_size", "nodes_capacity", "total_allocations",
    "total_deallocations", "allocated_nodes", "recycled_nodes", "alive_nodes", "patterns",
    "total_cache_hits", "total_cache_misses", "branch_cache_hits", "branch_cache_misses",
    "leaf_cache_hits", "leaf_cache_misses", "collapsed_branches", "leaf_nodes", "branch_nodes",
    "max_alive_nodes", "max_node_id", "max_branch_ref_count", "max_leaf_ref_count",
    "max_generation", "generations_overflows",
]


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (g++ only)."""
    src = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "voxelis_oracle.hpp")]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp, u8p, u64p, i64p, u32p, u16p = (C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_int64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint16))
    L.orc_last_error.restype = C.c_char_p
    L.orc_interner_create.restype = vp
    L.orc_interner_create.argtypes = [C.c_size_t, C.c_int]
    L.orc_interner_destroy.argtypes = [vp]
    L.orc_tree_create.restype = vp
    L.orc_tree_create.argtypes = [C.c_int]
    L.orc_tree_destroy.argtypes = [vp]
    L.orc_tree_root.restype = C.c_uint64
    L.orc_tree_root.argtypes = [vp]
    L.orc_tree_dirty.argtypes = [vp]
    L.orc_tree_adopt_root.argtypes = [vp, C.c_uint64]
    L.orc_tree_adopt_root.restype = None
    L.orc_batch_blocks.restype = C.c_size_t
    L.orc_batch_blocks.argtypes = [C.c_int]
    L.orc_batch_set.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]
    L.orc_encode_child_index_path.restype = C.c_uint32
    L.orc_encode_child_index_path.argtypes = [C.c_int] * 3
    L.orc_path_mask.restype = C.c_uint32
    L.orc_path_mask.argtypes = [C.c_int, C.c_int]
    L.orc_id_pack.restype = C.c_uint64
    L.orc_id_pack.argtypes = [C.c_uint32, C.c_uint16, C.c_uint8, C.c_uint8, C.c_int]
    L.orc_tree_apply_batch.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int64, C.c_int]
    L.orc_apply_batches_fresh.argtypes = [vp, C.c_int, C.c_size_t, vp, vp, vp, vp, vp, vp, vp]
    L.orc_tree_fill.argtypes = [vp, vp, C.c_int64]
    L.orc_tree_clear.argtypes = [vp, vp]
    L.orc_tree_get.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, i64p]
    L.orc_tree_to_vec.argtypes = [vp, vp, vp]
    L.orc_root_to_vec.argtypes = [vp, C.c_uint64, C.c_int, vp]
    L.orc_root_to_vec_lod.argtypes = [vp, C.c_uint64, C.c_int, C.c_int, vp]
    L.orc_occupancy_masks.restype = C.c_longlong
    L.orc_occupancy_masks.argtypes = [vp, C.c_size_t, vp, C.c_int, vp, vp, vp, C.c_size_t, vp, vp, vp]
    for f in ("orc_point_in_or_on_cube", "orc_point_in_or_on_triangle", "orc_edge_quad_intersection",
              "orc_triangle_cube_intersection", "orc_point_in_quad"):
        getattr(L, f).argtypes = [vp, vp]
    L.orc_voxelize_chunk.argtypes = [C.c_int, vp, C.c_int, C.c_double, vp, C.c_size_t, vp, vp, vp, vp]
    L.orc_interner_ref.restype = C.c_uint32
    L.orc_interner_ref.argtypes = [vp, C.c_uint64]
    L.orc_interner_next_index.restype = C.c_uint32
    L.orc_interner_next_index.argtypes = [vp]
    L.orc_interner_free_count.restype = C.c_size_t
    L.orc_interner_free_count.argtypes = [vp]
    L.orc_interner_capacity.restype = C.c_size_t
    L.orc_interner_capacity.argtypes = [vp]
    L.orc_interner_stats.argtypes = [vp, vp]
    L.orc_interner_download.argtypes = [vp, vp, vp, vp, vp]
    L.orc_dag_signature.restype = C.c_longlong
    L.orc_dag_signature.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_int, vp, vp, vp, vp,
                                    C.c_size_t, vp, vp]
    L.orc_model_serialize.restype = C.c_int64
    L.orc_model_serialize.argtypes = [vp, C.c_size_t, vp, vp, vp, C.c_size_t]
    L.orc_model_deserialize.restype = C.c_int64
    L.orc_model_deserialize.argtypes = [vp, vp, C.c_size_t, vp, vp, C.c_size_t]
    L.orc_time_apply_fresh.restype = C.c_double
    L.orc_time_apply_fresh.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_size_t, vp, vp, C.c_int, vp]
    _lib = L
    return L


def _f64(*vals):
    return np.ascontiguousarray(np.array(vals, np.float64).ravel())


def point_in_or_on_cube(p, cube):
    """voxelis-math/src/lib.rs:129-151; cube = (min, max)."""
    return bool(lib().orc_point_in_or_on_cube(_ptr(_f64(p)), _ptr(_f64(*cube))))


def point_in_or_on_triangle(p, tri):
    """voxelis-math/src/lib.rs:153-178."""
    return bool(lib().orc_point_in_or_on_triangle(_ptr(_f64(p)), _ptr(_f64(*tri))))


def edge_quad_intersection(edge, quad):
    """voxelis-math/src/lib.rs:180-204."""
    return bool(lib().orc_edge_quad_intersection(_ptr(_f64(*edge)), _ptr(_f64(*quad))))


def point_in_quad(p, quad):
    """voxelis-math/src/lib.rs:206-214."""
    return bool(lib().orc_point_in_quad(_ptr(_f64(p)), _ptr(_f64(*quad))))


def triangle_cube_intersection(tri, cube):
    """voxelis-math/src/lib.rs:3-127."""
    return bool(lib().orc_triangle_cube_intersection(_ptr(_f64(*tri)), _ptr(_f64(*cube))))


def face_chunk_map(depth, chunk_world_size, mesh_min, vertices, faces):
    """Voxelizer::build_face_to_chunk_map (voxelis-voxelize/src/lib.rs:113-156) in plain Python: chunk position ->
    list of face indices, chunks in first-seen order (the reference keeps a hash map)."""
    vpa = 1 << depth
    voxel_size = float(chunk_world_size) / vpa
    inv = 1.0 / voxel_size
    out = {}
    v = np.asarray(vertices, np.float64) - np.asarray(mesh_min, np.float64)
    for fi, f in enumerate(np.asarray(faces)):
        tri = v[np.asarray(f) - 1]
        lo = np.floor(tri.min(0) * inv).astype(np.int64)
        hi = np.ceil(tri.max(0) * inv).astype(np.int64)
        cl = np.sign(lo) * (np.abs(lo) // vpa)                      # IVec3 / i32 truncates toward zero
        ch = np.sign(hi) * (np.abs(hi) // vpa)
        for cy in range(cl[1], ch[1] + 1):
            for cz in range(cl[2], ch[2] + 1):
                for cx in range(cl[0], ch[0] + 1):
                    out.setdefault((int(cx), int(cy), int(cz)), []).append(fi)
    return out


def voxelize_chunk(dtype, chunk_position, depth, chunk_world_size, mesh_min, faces, vertices):
    """Voxelizer::voxelize_chunk (voxelis-voxelize/src/lib.rs:159-249) -> (has_patches, masks[B][2], values[B][8])."""
    B = lib().orc_batch_blocks(depth)
    masks = np.zeros((B, 2), np.uint8)
    values = np.zeros((B, 8), _NP[dtype])
    faces = np.ascontiguousarray(faces, np.int32).reshape(-1, 3)
    vertices = np.ascontiguousarray(vertices, np.float64).reshape(-1, 3)
    pos = np.ascontiguousarray(chunk_position, np.int32)
    mm = np.ascontiguousarray(mesh_min, np.float64)
    rc = _check(lib().orc_voxelize_chunk(dtype, _ptr(pos), depth, float(chunk_world_size), _ptr(mm), len(faces),
                                         _ptr(faces), _ptr(vertices), _ptr(masks), _ptr(values)))
    return bool(rc), masks, values


class OracleError(RuntimeError):
    pass


class ReferencePanic(OracleError):
    """The restated reference code hit one of its own assert!/panic! sites."""


def _check(rc: int) -> int:
    if rc == -2:
        raise ReferencePanic(lib().orc_last_error().decode())
    if rc < 0:
        raise OracleError(lib().orc_last_error().decode())
    return rc


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ BlockId helpers
def id_pack(index, gen, types, mask, leaf) -> int:
    return lib().orc_id_pack(index, gen, types, mask, int(leaf))


def id_index(i): return i & 0xFFFFFFFF
def id_gen(i): return (i >> 32) & 0x7FFF
def id_is_leaf(i): return (i >> 63) == 1
def id_is_branch(i): return (i >> 63) == 0
def id_types(i): return (i >> 55) & 0xFF
def id_mask(i): return (i >> 47) & 0xFF


class VoxInterner:
    """VoxInterner::<T>::with_memory_budget — voxelis/src/interner/mod.rs:45-155."""

    def __init__(self, budget: int, dtype: int = U8):
        self.dtype = dtype
        self.h = lib().orc_interner_create(budget, dtype)
        if not self.h:
            raise ReferencePanic(lib().orc_last_error().decode())

    @classmethod
    def with_memory_budget(cls, budget: int, dtype: int = U8):
        return cls(budget, dtype)

    def close(self):
        if self.h:
            lib().orc_interner_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_ref(self, block_id: int) -> int:
        return lib().orc_interner_ref(self.h, block_id)

    @property
    def next_index(self): return lib().orc_interner_next_index(self.h)
    @property
    def free_count(self): return lib().orc_interner_free_count(self.h)
    @property
    def capacity(self): return lib().orc_interner_capacity(self.h)

    def stats(self) -> dict:
        a = np.zeros(len(STATS_FIELDS), np.uint64)
        lib().orc_interner_stats(self.h, _ptr(a))
        return {k: int(v) for k, v in zip(STATS_FIELDS, a)}

    def download(self):
        n = self.next_index
        ch = np.zeros((n, 8), np.uint64)
        va = np.zeros(n, np.int64)
        rf = np.zeros(n, np.uint32)
        ge = np.zeros(n, np.uint16)
        lib().orc_interner_download(self.h, _ptr(ch), _ptr(va), _ptr(rf), _ptr(ge))
        return {"children": ch, "values": va, "refs": rf, "gens": ge, "n": n}

    def apply_batches_fresh(self, depth, masks, values, has_fill=None, fills=None, has_patches=None):
        """Serial loop of voxelis-voxelize/src/lib.rs:357-361 over fresh trees."""
        n = masks.shape[0]
        roots = np.zeros(n, np.uint64)
        changed = np.zeros(n, np.uint8)
        masks = np.ascontiguousarray(masks, np.uint8)
        values = np.ascontiguousarray(values, _NP[self.dtype])
        hf = None if has_fill is None else np.ascontiguousarray(has_fill, np.uint8)
        fv = None if fills is None else np.ascontiguousarray(fills, np.int64)
        hp = None if has_patches is None else np.ascontiguousarray(has_patches, np.uint8)
        _check(lib().orc_apply_batches_fresh(
            self.h, depth, n, _ptr(masks), _ptr(values),
            None if hf is None else _ptr(hf), None if fv is None else _ptr(fv),
            None if hp is None else _ptr(hp), _ptr(roots), _ptr(changed)))
        return roots, changed

    def model_serialize(self, positions, roots) -> bytes:
        """VoxModel::serialize (world/voxmodel.rs:177-294): the VTM payload for chunks (positions[n][3], roots[n])."""
        positions = np.ascontiguousarray(positions, np.int32)
        roots = np.ascontiguousarray(roots, np.uint64)
        n = len(roots)
        size = _check(lib().orc_model_serialize(self.h, n, _ptr(positions), _ptr(roots), None, 0))
        out = np.zeros(max(size, 1), np.uint8)
        _check(lib().orc_model_serialize(self.h, n, _ptr(positions), _ptr(roots), _ptr(out), size))
        return out[:size].tobytes()

    def model_deserialize(self, data: bytes):
        """VoxModel::deserialize (world/voxmodel.rs:296-408) into this (fresh) interner -> (positions, roots)."""
        buf = np.frombuffer(data, np.uint8)
        cap = len(buf) // 25 + 1                     # a chunk record is at least 12 + 12 + 1 bytes
        pos = np.zeros((cap, 3), np.int32)
        roots = np.zeros(cap, np.uint64)
        n = _check(lib().orc_model_deserialize(self.h, _ptr(buf), len(buf), _ptr(pos), _ptr(roots), cap))
        return pos[:n].copy(), roots[:n].copy()

    def root_to_vec(self, root: int, depth: int, lod: int = 0):
        n = 1 << max(depth - lod, 0)
        out = np.zeros((n, n, n), _NP[self.dtype])  # [y][z][x]
        _check(lib().orc_root_to_vec_lod(self.h, int(root), depth, lod, _ptr(out)))
        return out

    def occupancy_masks(self, roots, depth: int, offsets, lod: int = 0, max_materials: int = 256):
        """generate_occupancy_masks (utils/mesh.rs:515-596) of every root into one OccupancyDataBuilder + build()
        (:263-285) -> dict(global[3*4096], active[6], materials[(id, count)], per_material[m][3*4096])."""
        roots = np.ascontiguousarray(roots, np.uint64)
        offsets = np.ascontiguousarray(offsets, np.uint32).reshape(len(roots), 3)
        glob = np.zeros(3 * 4096, np.uint64)
        active = np.zeros(6, np.uint64)
        ids = np.zeros(max_materials, np.uint64)
        counts = np.zeros(max_materials, np.uint64)
        pm = np.zeros((max_materials, 3 * 4096), np.uint64)
        n = _check(lib().orc_occupancy_masks(self.h, len(roots), _ptr(roots), max(depth - lod, 0), _ptr(offsets),
                                             _ptr(glob), _ptr(active), max_materials, _ptr(ids), _ptr(counts),
                                             _ptr(pm)))
        return {"global": glob, "active": active, "material_ids": ids[:n].copy(), "material_counts": counts[:n].copy(),
                "per_material": pm[:n].copy()}


class Batch:
    """Batch<T> — voxelis/src/core/batch.rs:39-45,63-81."""

    def __init__(self, max_depth: int, dtype: int = U8):
        self.max_depth, self.dtype = max_depth, dtype
        B = lib().orc_batch_blocks(max_depth)
        self.masks = np.zeros((B, 2), np.uint8)
        self.values = np.zeros((B, 8), _NP[dtype])
        self.to_fill = None
        self.has_patches = False

    def set(self, interner, pos, v) -> bool:  # batch.rs:145-175,211-213
        lib().orc_batch_set(_ptr(self.masks), _ptr(self.values), self.dtype, pos[0], pos[1], pos[2], int(v))
        self.has_patches = True
        return True

    def clear(self, interner=None):  # batch.rs:187-195
        self.masks[:] = 0
        self.values[:] = 0
        self.to_fill = None
        self.has_patches = False

    def fill(self, interner, v):  # batch.rs:178-184
        self.clear()
        self.to_fill = int(v)

    def size(self) -> int:  # batch.rs:110-121
        return int(np.count_nonzero((self.masks[:, 0] != 0) | (self.masks[:, 1] != 0)))


class VoxTree:
    """VoxTree<T> — voxelis/src/spatial/voxtree.rs:108-142."""

    def __init__(self, max_depth: int, dtype: int = U8):
        self.max_depth, self.dtype = max_depth, dtype
        self.h = lib().orc_tree_create(max_depth)
        if not self.h:
            raise ReferencePanic(lib().orc_last_error().decode())

    def __del__(self):
        try:
            if self.h:
                lib().orc_tree_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def create_batch(self) -> Batch:
        return Batch(self.max_depth, self.dtype)

    def apply_batch(self, interner: VoxInterner, batch: Batch) -> bool:
        return bool(_check(lib().orc_tree_apply_batch(
            interner.h, self.h, _ptr(batch.masks), _ptr(batch.values),
            int(batch.to_fill is not None), int(batch.to_fill or 0), int(batch.has_patches))))

    def get(self, interner: VoxInterner, pos):
        out = C.c_int64(0)
        rc = _check(lib().orc_tree_get(interner.h, self.h, pos[0], pos[1], pos[2], C.byref(out)))
        return out.value if rc == 1 else None

    def to_vec(self, interner: VoxInterner):
        n = 1 << self.max_depth
        out = np.zeros((n, n, n), _NP[self.dtype])  # [y][z][x]
        _check(lib().orc_tree_to_vec(interner.h, self.h, _ptr(out)))
        return out

    def fill(self, interner, v): _check(lib().orc_tree_fill(interner.h, self.h, int(v)))
    def clear(self, interner): _check(lib().orc_tree_clear(interner.h, self.h))
    def get_root_id(self) -> int: return lib().orc_tree_root(self.h)
    def adopt_root(self, root: int): lib().orc_tree_adopt_root(self.h, int(root))
    def is_empty(self) -> bool: return self.get_root_id() == 0
    def is_leaf(self) -> bool: return id_is_leaf(self.get_root_id())
    def is_dirty(self) -> bool: return bool(lib().orc_tree_dirty(self.h))
    def voxels_per_axis(self) -> int: return 1 << self.max_depth


def dag_signature(children: np.ndarray, values: np.ndarray, roots, depth: int, want_stream=False,
                  want_indeg=False):
    """Canonical signature of the DAG reachable from ``roots`` (see orc_dag_signature)."""
    children = np.ascontiguousarray(children, np.uint64)
    values = np.ascontiguousarray(values, np.int64)
    roots = np.ascontiguousarray(roots, np.uint64)
    n = children.shape[0]
    sig = np.zeros(2, np.uint64)
    per_depth = np.zeros((depth + 1, 2), np.uint64)
    totals = np.zeros(2, np.uint64)
    indeg = np.zeros(n, np.uint32) if want_indeg else None
    numbers = np.zeros(n, np.uint32)
    words = lib().orc_dag_signature(_ptr(children), _ptr(values), n, _ptr(roots), len(roots), depth,
                                    _ptr(sig), _ptr(per_depth), _ptr(totals), None, 0,
                                    None if indeg is None else _ptr(indeg), _ptr(numbers))
    if words < 0:
        raise OracleError("malformed DAG (index out of range or cycle)")
    out = {"sig": (int(sig[0]), int(sig[1])), "per_depth": [(int(b), int(l)) for b, l in per_depth],
           "branches": int(totals[0]), "leaves": int(totals[1]), "words": int(words), "numbers": numbers}
    if want_stream:
        stream = np.zeros(words, np.uint64)
        lib().orc_dag_signature(_ptr(children), _ptr(values), n, _ptr(roots), len(roots), depth,
                                None, None, None, _ptr(stream), words, None, None)
        out["stream"] = stream
    if want_indeg:
        out["indeg"] = indeg
    return out


def time_apply_fresh(dtype, depth, budget, masks, values, threads=1):
    """Seconds for fresh-tree apply_batch over the slab with ``threads`` private interners."""
    masks = np.ascontiguousarray(masks, np.uint8)
    values = np.ascontiguousarray(values, _NP[dtype])
    roots = np.zeros(masks.shape[0], np.uint64)
    t = lib().orc_time_apply_fresh(dtype, depth, budget, masks.shape[0], _ptr(masks), _ptr(values),
                                   threads, _ptr(roots))
    if t < 0:
        raise OracleError(lib().orc_last_error().decode())
    return t, roots
