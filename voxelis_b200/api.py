"""Host-side mirror of the reference's VoxInterner / VoxTree / Batch API over the C ABI.

Names, argument meaning and error behaviour follow the Rust traits (reference
voxelis/src/spatial/voxops.rs:8-35; Batch core/batch.rs; VoxTree spatial/voxtree.rs):
reference panics become ``VoxelisError`` with the status code of include/voxelis_b200.h.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

U8, I32 = 0, 1
_NP = {U8: np.uint8, I32: np.int32}
FLAG_FILL, FLAG_PATCHES = 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvoxelis_b200.so")
_lib = None

STATUS = {0: "VX_OK", -1: "VX_E_INVALID", -2: "VX_E_OOM", -3: "VX_E_CUDA", -4: "VX_E_UNSUPPORTED",
          -5: "VX_E_BOUNDS", -6: "VX_E_BUDGET", -7: "VX_E_POISONED"}

STATS_FIELDS = [
    "requested_budget", "actual_budget", "node_size", "nodes_capacity", "total_allocations",
    "total_deallocations", "allocated_nodes", "recycled_nodes", "alive_nodes", "patterns",
    "total_cache_hits", "total_cache_misses", "branch_cache_hits", "branch_cache_misses",
    "leaf_cache_hits", "leaf_cache_misses", "collapsed_branches", "leaf_nodes", "branch_nodes",
    "max_alive_nodes", "max_node_id", "max_branch_ref_count", "max_leaf_ref_count",
    "max_generation", "generations_overflows",
]

# every symbol include/voxelis_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "vx_last_error", "vx_abi_version", "vx_device_count", "vx_interner_create", "vx_interner_destroy",
    "vx_interner_reset", "vx_interner_reset_async", "vx_interner_capacity", "vx_interner_dtype", "vx_interner_device",
    "vx_interner_next_index", "vx_interner_get_ref", "vx_interner_get_value", "vx_interner_get_children",
    "vx_interner_stats", "vx_interner_download", "vx_interner_sync", "vx_interner_stream",
    "vx_batch_create", "vx_batch_destroy", "vx_batch_set", "vx_batch_fill", "vx_batch_clear",
    "vx_batch_masks", "vx_batch_values", "vx_batch_blocks", "vx_batch_to_fill", "vx_batch_size",
    "vx_batch_has_patches", "vx_batch_mark_patched", "vx_batch_max_depth", "vx_batch_dtype",
    "vx_batch_set_many", "vx_batch_assign", "vx_batch_touched_units", "vx_trees_forget",
    "vx_model_serialize", "vx_export_vtm", "vx_model_deserialize", "vx_import_vtm", "vx_tree_adopt_root",
    "vx_tree_create", "vx_tree_destroy", "vx_tree_root_id", "vx_tree_set_root_id", "vx_tree_max_depth",
    "vx_tree_voxels_per_axis", "vx_tree_is_empty", "vx_tree_is_leaf", "vx_tree_is_dirty",
    "vx_tree_mark_dirty", "vx_tree_clear_dirty", "vx_tree_apply_batch", "vx_apply_batches",
    "vx_apply_batches_slab", "vx_apply_batches_device", "vx_tree_get", "vx_tree_get_many",
    "vx_tree_to_vec", "vx_roots_to_vec", "vx_roots_to_vec_lod", "vx_occupancy_masks", "vx_terrain_heights_device", "vx_terrain_batches_device", "vx_random_batches_device", "vx_voxelize_plan", "vx_voxelize_chunks_device", "vx_tree_fill", "vx_tree_clear",
    "vx_interner_memory", "vx_interner_debug_memo", "vx_world_unique_id", "vx_world_create", "vx_world_destroy", "vx_world_size", "vx_world_rank",
    "vx_world_barrier", "vx_world_global_dedup", "vx_dedup_heights", "vx_dedup_pack", "vx_dedup_scatter", "vx_dedup_map_roots", "vx_interner_intern_records",
]


class VoxelisError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


def lib():
    """Loads libvoxelis_b200.so; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("VX_LIB", _LIB_PATH)      # VX_LIB: try an experimental build of the same ABI
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -m voxelis_b200.build` "
            "(nvcc, sm_100a). voxelis_b200 has no CPU fallback.")
    L = C.CDLL(path)
    vp, sz, i64, u64 = C.c_void_p, C.c_size_t, C.c_int64, C.c_uint64
    L.vx_last_error.restype = C.c_char_p
    L.vx_interner_create.restype = vp
    L.vx_interner_create.argtypes = [sz, C.c_int, C.c_int]
    L.vx_interner_destroy.argtypes = [vp]
    L.vx_interner_destroy.restype = None
    L.vx_interner_reset.argtypes = [vp]
    L.vx_interner_reset_async.argtypes = [vp, vp]
    L.vx_interner_capacity.restype = sz
    L.vx_interner_capacity.argtypes = [vp]
    L.vx_interner_dtype.argtypes = [vp]
    L.vx_interner_device.argtypes = [vp]
    L.vx_interner_next_index.restype = i64
    L.vx_interner_next_index.argtypes = [vp]
    L.vx_interner_get_ref.argtypes = [vp, u64, vp]
    L.vx_interner_get_value.argtypes = [vp, u64, vp]
    L.vx_interner_get_children.argtypes = [vp, u64, vp]
    L.vx_interner_stats.argtypes = [vp, vp]
    L.vx_interner_memory.argtypes = [vp, vp]
    L.vx_interner_debug_memo.argtypes = [vp, vp]
    L.vx_world_unique_id.argtypes = [vp]
    L.vx_world_create.restype = vp
    L.vx_world_create.argtypes = [C.c_int, C.c_int, vp, C.c_int]
    L.vx_world_destroy.restype = None
    L.vx_world_destroy.argtypes = [vp]
    L.vx_world_size.argtypes = [vp]
    L.vx_world_rank.argtypes = [vp]
    L.vx_world_barrier.argtypes = [vp, vp]
    L.vx_world_global_dedup.argtypes = [vp, vp, vp, sz, vp, vp, vp]
    L.vx_interner_debug_counters.argtypes = [vp, vp]
    L.vx_interner_profile_stages.argtypes = [vp, C.c_int]
    L.vx_interner_stage_ms.argtypes = [vp, vp, vp]
    L.vx_interner_host_trace.argtypes = [vp, vp]
    L.vx_interner_download.restype = i64
    L.vx_interner_download.argtypes = [vp, sz, vp, vp, vp, vp, vp]
    L.vx_interner_sync.argtypes = [vp]
    L.vx_interner_stream.restype = vp
    L.vx_interner_stream.argtypes = [vp]
    L.vx_batch_create.restype = vp
    L.vx_batch_create.argtypes = [C.c_uint8, C.c_int]
    L.vx_batch_destroy.argtypes = [vp]
    L.vx_batch_destroy.restype = None
    L.vx_batch_set.argtypes = [vp, C.c_int, C.c_int, C.c_int, i64]
    L.vx_batch_fill.argtypes = [vp, i64]
    L.vx_batch_clear.argtypes = [vp]
    L.vx_batch_masks.restype = vp
    L.vx_batch_masks.argtypes = [vp]
    L.vx_batch_values.restype = vp
    L.vx_batch_values.argtypes = [vp]
    L.vx_batch_blocks.restype = sz
    L.vx_batch_blocks.argtypes = [vp]
    L.vx_batch_to_fill.argtypes = [vp, vp]
    L.vx_batch_size.restype = sz
    L.vx_batch_size.argtypes = [vp]
    L.vx_batch_has_patches.argtypes = [vp]
    L.vx_batch_mark_patched.argtypes = [vp]
    L.vx_batch_mark_patched.restype = None
    L.vx_batch_set_many.argtypes = [vp, sz, vp, vp]
    L.vx_batch_assign.argtypes = [vp, vp, vp]
    L.vx_batch_touched_units.argtypes = [vp]
    L.vx_trees_forget.argtypes = [vp, sz]
    L.vx_tree_adopt_root.argtypes = [vp, u64]
    L.vx_batch_max_depth.restype = C.c_uint8
    L.vx_batch_max_depth.argtypes = [vp]
    L.vx_batch_dtype.argtypes = [vp]
    L.vx_tree_create.restype = vp
    L.vx_tree_create.argtypes = [C.c_uint8]
    L.vx_tree_destroy.argtypes = [vp]
    L.vx_tree_destroy.restype = None
    L.vx_tree_root_id.restype = u64
    L.vx_tree_root_id.argtypes = [vp]
    L.vx_tree_set_root_id.argtypes = [vp, vp, u64]
    L.vx_tree_max_depth.restype = C.c_uint8
    L.vx_tree_max_depth.argtypes = [vp]
    L.vx_tree_voxels_per_axis.restype = C.c_uint32
    L.vx_tree_voxels_per_axis.argtypes = [vp]
    for f in ("vx_tree_is_empty", "vx_tree_is_leaf", "vx_tree_is_dirty"):
        getattr(L, f).argtypes = [vp]
    L.vx_tree_mark_dirty.argtypes = [vp]
    L.vx_tree_mark_dirty.restype = None
    L.vx_tree_clear_dirty.argtypes = [vp]
    L.vx_tree_clear_dirty.restype = None
    L.vx_tree_apply_batch.argtypes = [vp, vp, vp]
    L.vx_apply_batches.argtypes = [vp, vp, vp, sz, vp]
    L.vx_apply_batches_slab.argtypes = [vp, C.c_uint8, sz, vp, vp, vp, vp, vp, vp]
    L.vx_apply_batches_device.argtypes = [vp, C.c_uint8, sz, vp, vp, vp, vp, vp, vp, vp]
    L.vx_tree_get.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
    L.vx_tree_get_many.argtypes = [vp, vp, sz, vp, vp, vp]
    L.vx_tree_to_vec.argtypes = [vp, vp, vp]
    L.vx_roots_to_vec.argtypes = [vp, C.c_uint8, sz, vp, vp]
    L.vx_roots_to_vec_lod.argtypes = [vp, C.c_uint8, C.c_uint8, sz, vp, vp]
    L.vx_occupancy_masks.argtypes = [vp, C.c_uint8, C.c_uint8, sz, vp, vp, vp, sz, C.c_uint32, vp, vp, vp, vp, vp, vp]
    L.vx_terrain_heights_device.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, i64, i64, vp, vp]
    L.vx_terrain_batches_device.argtypes = [vp, C.c_uint8, vp, vp, C.c_int, C.c_int, vp, vp, vp]
    L.vx_random_batches_device.argtypes = [vp, C.c_uint8, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, vp]
    L.vx_voxelize_plan.restype = i64
    L.vx_voxelize_plan.argtypes = [C.c_uint8, C.c_double, vp, sz, vp, sz, vp, vp, sz, vp, vp, sz, vp]
    L.vx_voxelize_chunks_device.argtypes = [vp, C.c_uint8, C.c_double, vp, sz, vp, sz, vp, sz, vp, sz, vp, vp, vp, vp, vp]
    L.vx_tree_fill.argtypes = [vp, vp, i64]
    L.vx_tree_clear.argtypes = [vp, vp]
    L.vx_model_serialize.restype = i64
    L.vx_model_serialize.argtypes = [vp, sz, vp, vp, vp, sz]
    L.vx_export_vtm.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_uint8, C.c_float, vp, sz, vp, vp, C.c_int]
    L.vx_model_deserialize.restype = i64
    L.vx_model_deserialize.argtypes = [vp, vp, sz, vp, vp, sz]
    L.vx_import_vtm.restype = i64
    L.vx_import_vtm.argtypes = [vp, C.c_char_p, vp, vp, vp, sz]
    L.vx_dedup_heights.argtypes = [vp, vp]
    L.vx_dedup_pack.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp, vp]
    L.vx_dedup_scatter.argtypes = [vp, sz, vp, vp, vp]
    L.vx_dedup_map_roots.argtypes = [vp, sz, vp, vp, vp]
    L.vx_interner_intern_records.argtypes = [vp, sz, vp, C.c_int, C.c_int, vp, vp]
    _lib = L
    return L


def _err(code: int):
    raise VoxelisError(code, lib().vx_last_error().decode())


def _ck(code: int) -> int:
    if code < 0:
        _err(code)
    return code


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    return lib().vx_device_count()


def id_index(i): return int(i) & 0xFFFFFFFF
def id_gen(i): return (int(i) >> 32) & 0x7FFF
def id_is_leaf(i): return (int(i) >> 63) == 1
def id_is_branch(i): return (int(i) >> 63) == 0
def id_types(i): return (int(i) >> 55) & 0xFF
def id_mask(i): return (int(i) >> 47) & 0xFF


class VoxInterner:
    """VoxInterner<T> — reference voxelis/src/interner/mod.rs:25-40,45-155."""

    def __init__(self, budget: int, dtype: int = U8, device: int = 0):
        self.dtype = dtype
        self.h = lib().vx_interner_create(budget, dtype, device)
        if not self.h:
            raise VoxelisError(-6, lib().vx_last_error().decode())

    @classmethod
    def with_memory_budget(cls, budget: int, dtype: int = U8, device: int = 0):
        return cls(budget, dtype, device)

    def close(self):
        if getattr(self, "h", None):
            lib().vx_interner_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self): _ck(lib().vx_interner_reset(self.h))
    def reset_async(self, stream: int = 0): _ck(lib().vx_interner_reset_async(self.h, C.c_void_p(stream or None)))
    def sync(self): _ck(lib().vx_interner_sync(self.h))
    @property
    def capacity(self) -> int: return lib().vx_interner_capacity(self.h)
    @property
    def next_index(self) -> int: return _ck(lib().vx_interner_next_index(self.h))
    @property
    def stream(self) -> int: return lib().vx_interner_stream(self.h) or 0

    def get_ref(self, block_id: int) -> int:
        out = C.c_uint32(0)
        _ck(lib().vx_interner_get_ref(self.h, int(block_id), C.byref(out)))
        return out.value

    def get_value(self, block_id: int) -> int:
        out = C.c_int64(0)
        _ck(lib().vx_interner_get_value(self.h, int(block_id), C.byref(out)))
        return out.value

    def get_children(self, block_id: int):
        out = np.zeros(8, np.uint64)
        _ck(lib().vx_interner_get_children(self.h, int(block_id), _ptr(out)))
        return out

    def stats(self) -> dict:
        a = np.zeros(len(STATS_FIELDS), np.uint64)
        _ck(lib().vx_interner_stats(self.h, _ptr(a)))
        return {k: int(v) for k, v in zip(STATS_FIELDS, a)}

    def memory(self) -> dict:
        """vx_interner_memory: bytes of device memory behind this interner (pools, tables, builder scratch)."""
        m = (C.c_uint64 * 8)()
        _ck(lib().vx_interner_memory(self.h, m))
        keys = ("pools_bytes", "table_bytes", "leaf_table_bytes", "bulk_scratch_bytes", "stage_bytes", "other_scratch_bytes",
                "pinned_host_bytes", "total_device_bytes")
        return {k: int(v) for k, v in zip(keys, m)}

    def debug_memo(self):
        m = (C.c_uint64 * 4)()
        _ck(lib().vx_interner_debug_memo(self.h, m))
        return [int(v) for v in m]

    def debug_counters(self) -> dict:
        a = np.zeros(8, np.uint64)
        _ck(lib().vx_interner_debug_counters(self.h, _ptr(a)))
        keys = ["leaf_calls", "branch_calls", "leaf_misses", "branch_misses", "collapsed", "probe_steps",
                "cache_hits_local", "recycled"]
        return {k: int(v) for k, v in zip(keys, a)}

    def profile_stages(self, on: bool = True) -> None:
        """Diagnostics: record CUDA events between the launches of every following apply call."""
        _ck(lib().vx_interner_profile_stages(self.h, 1 if on else 0))

    def stage_ms(self) -> list:
        """[(kernel name, device ms)] of the last apply call (empty unless profile_stages is on)."""
        ms = (C.c_float * 9)()
        names = (C.c_char_p * 9)()
        k = _ck(lib().vx_interner_stage_ms(self.h, ms, names))
        return [(names[i].decode(), float(ms[i])) for i in range(k)]

    def host_trace(self) -> dict:
        """Phases of the last vx_apply_batches call made while profile_stages was on (microseconds)."""
        a = np.zeros(8, np.float64)
        _ck(lib().vx_interner_host_trace(self.h, _ptr(a)))
        keys = ["host_lists_us", "host_enqueue_us", "host_wait_us", "dev_descriptors_memset_us", "dev_stage_us",
                "dev_build_us", "touched_units"]
        return {k: float(v) for k, v in zip(keys, a)}

    def download(self) -> dict:
        n = self.next_index
        ch = np.zeros((n, 8), np.uint64)
        va = np.zeros(n, np.int64)
        rf = np.zeros(n, np.uint32)
        ge = np.zeros(n, np.uint16)
        hs = np.zeros(n, np.uint64)
        got = _ck(lib().vx_interner_download(self.h, n, _ptr(ch), _ptr(va), _ptr(rf), _ptr(ge), _ptr(hs)))
        assert got == n
        return {"children": ch, "values": va, "refs": rf, "gens": ge, "hashes": hs, "n": n}

    # ---- multi-chunk slab entry (new; replaces the serial loop voxelis-voxelize/src/lib.rs:357-361)
    def apply_batches_slab(self, depth: int, masks, values, flags=None, fills=None, n=None,
                           roots_out=None, changed_out=None):
        """masks/values: numpy arrays [n][B][2] / [n][B][8] (host) or integer device pointers
        (then ``n`` is required).  Returns (roots, changed) as numpy arrays unless device
        outputs were supplied."""
        if isinstance(masks, np.ndarray):
            n = masks.shape[0]
            masks = np.ascontiguousarray(masks, np.uint8)
            values = np.ascontiguousarray(values, _NP[self.dtype])
        assert n is not None
        if flags is not None and isinstance(flags, np.ndarray):
            flags = np.ascontiguousarray(flags, np.uint8)
        if fills is not None and isinstance(fills, np.ndarray):
            fills = np.ascontiguousarray(fills, np.int64)
        roots = np.zeros(n, np.uint64) if roots_out is None else roots_out
        changed = np.zeros(n, np.uint8) if changed_out is None else changed_out
        _ck(lib().vx_apply_batches_slab(self.h, depth, n, _ptr(masks), _ptr(values), _ptr(flags), _ptr(fills),
                                        _ptr(roots), _ptr(changed)))
        return roots, changed

    def apply_batches_device(self, depth: int, n: int, d_masks: int, d_values: int, d_roots: int,
                             d_changed: int = 0, d_flags: int = 0, d_fills: int = 0, stream: int = 0):
        """Asynchronous, device pointers only (see vx_apply_batches_device)."""
        _ck(lib().vx_apply_batches_device(self.h, depth, n, C.c_void_p(d_masks), C.c_void_p(d_values),
                                          C.c_void_p(d_flags or None), C.c_void_p(d_fills or None),
                                          C.c_void_p(d_roots), C.c_void_p(d_changed or None),
                                          C.c_void_p(stream or None)))

    def terrain_heights_device(self, nx: int, nz: int, d_heights: int, seed: int = 0x5EED0000, height: int = 256,
                               x0: int = 0, z0: int = 0, stream: int = 0):
        """vx_terrain_heights_device: int32 heights[nx][nz] in device memory (== workloads.height_field)."""
        _ck(lib().vx_terrain_heights_device(self.h, nx, nz, seed, height, x0, z0, C.c_void_p(d_heights),
                                            C.c_void_p(stream or None)))

    def terrain_batches_device(self, depth: int, grid, d_heights: int, d_masks: int, d_values: int,
                               surface_only: bool = True, materials: int = 1, stream: int = 0):
        """vx_terrain_batches_device: the Batch arrays of a whole grid of terrain chunks, written in device memory
        (generate_terrain_batch[_3_mats], reference voxelis/src/utils/shapes.rs:273-357; == workloads.terrain_world)."""
        g = (C.c_uint32 * 3)(*[int(v) for v in grid])
        _ck(lib().vx_terrain_batches_device(self.h, depth, g, C.c_void_p(d_heights), int(bool(surface_only)),
                                            materials, C.c_void_p(d_masks), C.c_void_p(d_values),
                                            C.c_void_p(stream or None)))

    def random_batches_device(self, depth: int, n: int, d_masks: int, d_values: int, k: int = 255, cell: int = 1,
                              chunk0: int = 0, seed_base: int = 0x5EED0000, stream: int = 0):
        """vx_random_batches_device: n high-entropy chunks written in device memory (== workloads.p_random(k, cell) for
        chunk indices chunk0 .. chunk0 + n)."""
        _ck(lib().vx_random_batches_device(self.h, depth, n, seed_base, chunk0, k, cell, C.c_void_p(d_masks),
                                           C.c_void_p(d_values), C.c_void_p(stream or None)))

    def voxelize_chunks_device(self, depth: int, chunk_world_size: float, mesh_min, vertices, faces, plan,
                               d_masks: int, d_values: int, d_has_patches: int = 0):
        """Voxelizer::voxelize_chunk (reference voxelis-voxelize/src/lib.rs:159-249) for every chunk of ``plan``
        (from voxelize_plan) into the device slab masks[n][B][2] / values[n][B][8]."""
        positions, pair_chunk, pair_face = plan
        vertices = np.ascontiguousarray(vertices, np.float64).reshape(-1, 3)
        faces = np.ascontiguousarray(faces, np.int32).reshape(-1, 3)
        mm = np.ascontiguousarray(mesh_min, np.float64)
        _ck(lib().vx_voxelize_chunks_device(self.h, depth, float(chunk_world_size), _ptr(mm), len(vertices),
                                            _ptr(vertices), len(faces), _ptr(faces), len(positions), _ptr(positions),
                                            len(pair_chunk), _ptr(pair_chunk), _ptr(pair_face), C.c_void_p(d_masks),
                                            C.c_void_p(d_values), C.c_void_p(d_has_patches or None)))

    def model_serialize(self, positions, roots) -> bytes:
        """VoxModel::serialize (world/voxmodel.rs:177-294): VTM payload of chunks (positions[n][3], roots[n])."""
        positions = np.ascontiguousarray(positions, np.int32)
        roots = np.ascontiguousarray(roots, np.uint64)
        n = len(roots)
        size = _ck(lib().vx_model_serialize(self.h, n, _ptr(positions), _ptr(roots), None, 0))
        out = np.zeros(max(size, 1), np.uint8)
        got = _ck(lib().vx_model_serialize(self.h, n, _ptr(positions), _ptr(roots), _ptr(out), size))
        assert got == size
        return out[:size].tobytes()

    def export_vtm(self, path: str, name: str, max_depth: int, chunk_world_size: float, world_bounds, positions, roots,
                   compress: bool = True):
        """export_model_to_vtm (io/export.rs:90-151)."""
        positions = np.ascontiguousarray(positions, np.int32)
        roots = np.ascontiguousarray(roots, np.uint64)
        wb = np.ascontiguousarray(world_bounds, np.int32)
        _ck(lib().vx_export_vtm(self.h, path.encode(), name.encode(), max_depth, float(chunk_world_size), _ptr(wb),
                                len(roots), _ptr(positions), _ptr(roots), 1 if compress else 0))

    def model_deserialize(self, data: bytes):
        """VoxModel::deserialize (world/voxmodel.rs:296-408) into this FRESH interner -> (positions, roots)."""
        buf = np.frombuffer(data, np.uint8)
        cap = len(buf) // 25 + 1                     # a chunk record is at least 12 + 12 + 1 bytes
        pos = np.zeros((cap, 3), np.int32)
        roots = np.zeros(cap, np.uint64)
        n = _ck(lib().vx_model_deserialize(self.h, _ptr(buf), len(buf), _ptr(pos), _ptr(roots), cap))
        return pos[:n].copy(), roots[:n].copy()

    def import_vtm(self, path: str, max_chunks: int = 1 << 20):
        """import_model_from_vtm (io/import.rs:14-98) -> (info dict, positions, roots)."""
        class Info(C.Structure):
            _fields_ = [("flags", C.c_uint16), ("max_depth", C.c_uint8), ("chunk_world_size", C.c_float),
                        ("world_bounds", C.c_int32 * 3), ("name", C.c_char * 256)]
        info = Info()
        pos = np.zeros((max_chunks, 3), np.int32)
        roots = np.zeros(max_chunks, np.uint64)
        n = _ck(lib().vx_import_vtm(self.h, path.encode(), C.byref(info), _ptr(pos), _ptr(roots), max_chunks))
        meta = {"flags": info.flags, "max_depth": info.max_depth, "chunk_world_size": info.chunk_world_size,
                "world_bounds": tuple(info.world_bounds), "name": info.name.decode()}
        return meta, pos[:n].copy(), roots[:n].copy()

    def roots_to_vec(self, roots, depth: int, lod: int = 0):
        """to_vec for bare roots; ``lod`` > 0 unfolds only depth - lod levels (world/voxchunk.rs:267)."""
        roots = np.ascontiguousarray(roots, np.uint64)
        n = 1 << max(depth - lod, 0)
        out = np.zeros((len(roots), n, n, n), _NP[self.dtype])  # [r][y][z][x]
        _ck(lib().vx_roots_to_vec_lod(self.h, depth, lod, len(roots), _ptr(roots), _ptr(out)))
        return out

    def occupancy_masks(self, roots, depth: int, offsets, builder_of=None, n_builders: int = 1, lod: int = 0,
                        max_materials: int = 8):
        """generate_occupancy_masks (reference voxelis/src/utils/mesh.rs:515-596) of every root, chunk i into
        builder ``builder_of[i]`` at ``offsets[i]`` -> dict of host arrays: global[nb][3*4096], active[nb][6],
        n_materials[nb], material_ids / material_counts[nb][max_materials], per_material[nb][max_materials][3*4096]
        (rows past n_materials[b] are left zero)."""
        roots = np.ascontiguousarray(roots, np.uint64)
        offsets = np.ascontiguousarray(offsets, np.uint32).reshape(len(roots), 3)
        bo = None if builder_of is None else np.ascontiguousarray(builder_of, np.uint32)
        nb, M = n_builders, max_materials
        out = {"global": np.zeros((nb, 3 * 4096), np.uint64), "active": np.zeros((nb, 6), np.uint64),
               "n_materials": np.zeros(nb, np.uint32), "material_ids": np.zeros((nb, M), np.uint64),
               "material_counts": np.zeros((nb, M), np.uint64), "per_material": np.zeros((nb, M, 3 * 4096), np.uint64)}
        _ck(lib().vx_occupancy_masks(self.h, depth, lod, len(roots), _ptr(roots), _ptr(offsets),
                                     None if bo is None else _ptr(bo), nb, M, _ptr(out["global"]), _ptr(out["active"]),
                                     _ptr(out["n_materials"]), _ptr(out["material_ids"]), _ptr(out["material_counts"]),
                                     _ptr(out["per_material"])))
        return out


class World:
    """vx_world: this process's membership in a group of one-process-per-GPU ranks (NCCL underneath, loaded by the
    library).  ``unique_id()`` on rank 0, ship the 128 bytes to the other ranks with whatever the host has, then
    ``World(n_ranks, rank, uid, device)`` everywhere (collective)."""

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        _ck(lib().vx_world_unique_id(buf))
        return bytes(buf)

    def __init__(self, n_ranks: int, rank: int, uid: bytes, device: int = 0):
        self.n_ranks, self.rank, self.device = n_ranks, rank, device
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self.h = lib().vx_world_create(n_ranks, rank, buf, device)
        if not self.h:
            raise VoxelisError(-1, lib().vx_last_error().decode())

    @classmethod
    def from_torch_distributed(cls, device: int):
        """Rendezvous through an initialised torch.distributed group (plumbing only: 128 bytes broadcast)."""
        import torch
        import torch.distributed as dist
        n, r = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
        uid = cls.unique_id() if r == 0 else bytes(128)
        if n > 1:
            t = torch.tensor(list(uid), dtype=torch.uint8, device=torch.device("cuda", device))
            dist.broadcast(t, 0)
            uid = bytes(t.cpu().tolist())
        return cls(n, r, uid, device)

    def close(self):
        if getattr(self, "h", None):
            lib().vx_world_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def barrier(self, stream: int = 0):
        _ck(lib().vx_world_barrier(self.h, C.c_void_p(stream or None)))

    def global_dedup(self, local: "VoxInterner", shard: "VoxInterner", roots):
        """vx_world_global_dedup (collective).  Returns (global_roots uint64[n], summary dict)."""
        roots = np.ascontiguousarray(roots, np.uint64)
        out = np.zeros_like(roots)
        summ = (C.c_uint64 * 7)()
        _ck(lib().vx_world_global_dedup(self.h, local.h, shard.h, roots.size, _ptr(roots), _ptr(out), summ))
        keys = ("rounds", "branches", "leaves", "bytes_sent", "local_nodes_all_ranks", "this_shard_branches", "this_shard_leaves")
        return out, {k: int(v) for k, v in zip(keys, summ)}


def voxelize_plan(depth: int, chunk_world_size: float, mesh_min, vertices, faces):
    """Voxelizer::build_face_to_chunk_map (reference voxelis-voxelize/src/lib.rs:113-156) ->
    (positions[n][3] int32, pair_chunk[np] uint32, pair_face[np] uint32); host work, no device needed."""
    vertices = np.ascontiguousarray(vertices, np.float64).reshape(-1, 3)
    faces = np.ascontiguousarray(faces, np.int32).reshape(-1, 3)
    mm = np.ascontiguousarray(mesh_min, np.float64)
    npairs = C.c_size_t(0)
    cap_pairs = 2 * len(faces) + 64                     # one call when the guess holds (small faces), else a second
    cap_chunks = cap_pairs
    while True:
        positions = np.zeros((cap_chunks, 3), np.int32)
        pc = np.zeros(cap_pairs, np.uint32)
        pf = np.zeros(cap_pairs, np.uint32)
        n = _ck(lib().vx_voxelize_plan(depth, float(chunk_world_size), _ptr(mm), len(vertices), _ptr(vertices),
                                       len(faces), _ptr(faces), _ptr(positions), cap_chunks, _ptr(pc), _ptr(pf),
                                       cap_pairs, C.byref(npairs)))
        if n <= cap_chunks and npairs.value <= cap_pairs:
            break
        cap_chunks, cap_pairs = max(n, 1), max(npairs.value, 1)
    positions, pc, pf = positions[:n].copy(), pc[:npairs.value].copy(), pf[:npairs.value].copy()
    return positions, pc, pf


class Batch:
    """Batch<T> — reference voxelis/src/core/batch.rs:39-45."""

    def __init__(self, max_depth: int, dtype: int = U8):
        self.max_depth, self.dtype = max_depth, dtype
        self.h = lib().vx_batch_create(max_depth, dtype)
        if not self.h:
            raise VoxelisError(-1, lib().vx_last_error().decode())
        self._masks = self._values = None

    # Raw views of the batch's arrays (Batch::masks / values, batch.rs:86-105).  Asking for them tells the
    # library that the caller may write the arrays directly: such a batch must be mark_patched() after
    # writing, and apply then moves its masks over the bus instead of the one-bit-per-block map.
    @property
    def masks(self):
        if self._masks is None:
            B = lib().vx_batch_blocks(self.h)
            self._masks = np.ctypeslib.as_array(C.cast(lib().vx_batch_masks(self.h), C.POINTER(C.c_uint8)), (B, 2))
        return self._masks

    @property
    def values(self):
        if self._values is None:
            B = lib().vx_batch_blocks(self.h)
            ct = C.c_uint8 if self.dtype == U8 else C.c_int32
            self._values = np.ctypeslib.as_array(C.cast(lib().vx_batch_values(self.h), C.POINTER(ct)), (B, 8))
        return self._values

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self._masks = self._values = None
                lib().vx_batch_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set(self, interner, pos, v) -> bool:          # batch.rs:211-213
        return _ck(lib().vx_batch_set(self.h, int(pos[0]), int(pos[1]), int(pos[2]), int(v))) == 1

    def set_many(self, xyz, values) -> bool:
        """Array form of set(): xyz [n][3], values [n]."""
        xyz = np.ascontiguousarray(xyz, np.int32)
        values = np.ascontiguousarray(values, np.int64)
        return _ck(lib().vx_batch_set_many(self.h, xyz.shape[0], _ptr(xyz), _ptr(values))) == 1

    def assign(self, masks, values):
        """Takes over dense arrays in Batch layout (masks [B][2], values [B][8])."""
        masks = np.ascontiguousarray(masks, np.uint8)
        values = np.ascontiguousarray(values, _NP[self.dtype])
        B = lib().vx_batch_blocks(self.h)
        assert masks.size == 2 * B and values.size == 8 * B
        _ck(lib().vx_batch_assign(self.h, _ptr(masks), _ptr(values)))

    @property
    def touched_units(self) -> int: return _ck(lib().vx_batch_touched_units(self.h))

    def fill(self, interner, v): _ck(lib().vx_batch_fill(self.h, int(v)))   # batch.rs:218-221
    def clear(self, interner=None): _ck(lib().vx_batch_clear(self.h))       # batch.rs:223-225
    def size(self) -> int: return lib().vx_batch_size(self.h)
    def mark_patched(self): lib().vx_batch_mark_patched(self.h)

    @property
    def has_patches(self) -> bool: return bool(lib().vx_batch_has_patches(self.h))

    @property
    def to_fill(self):
        out = C.c_int64(0)
        return out.value if _ck(lib().vx_batch_to_fill(self.h, C.byref(out))) == 1 else None


class VoxTree:
    """VoxTree<T> — reference voxelis/src/spatial/voxtree.rs:108-142."""

    def __init__(self, max_depth: int, dtype: int = U8):
        self.max_depth, self.dtype = max_depth, dtype
        self.h = lib().vx_tree_create(max_depth)
        if not self.h:
            raise VoxelisError(-1, lib().vx_last_error().decode())

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().vx_tree_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def create_batch(self) -> Batch: return Batch(self.max_depth, self.dtype)   # voxtree.rs:296-301

    def apply_batch(self, interner: VoxInterner, batch: Batch) -> bool:        # voxtree.rs:303-328
        return _ck(lib().vx_tree_apply_batch(interner.h, self.h, batch.h)) == 1

    def get(self, interner: VoxInterner, pos):                                  # voxtree.rs:144-160
        out = C.c_int64(0)
        rc = _ck(lib().vx_tree_get(interner.h, self.h, int(pos[0]), int(pos[1]), int(pos[2]), C.byref(out)))
        return out.value if rc == 1 else None

    def get_many(self, interner: VoxInterner, xyz: np.ndarray):
        xyz = np.ascontiguousarray(xyz, np.int32)
        n = xyz.shape[0]
        found = np.zeros(n, np.uint8)
        vals = np.zeros(n, np.int64)
        _ck(lib().vx_tree_get_many(interner.h, self.h, n, _ptr(xyz), _ptr(found), _ptr(vals)))
        return found, vals

    def to_vec(self, interner: VoxInterner):                                    # utils/common.rs:158-246
        n = 1 << self.max_depth
        out = np.zeros((n, n, n), _NP[interner.dtype])  # [y][z][x]
        _ck(lib().vx_tree_to_vec(interner.h, self.h, _ptr(out)))
        return out

    def fill(self, interner, v): _ck(lib().vx_tree_fill(interner.h, self.h, int(v)))
    def clear(self, interner): _ck(lib().vx_tree_clear(interner.h, self.h))
    def get_root_id(self) -> int: return lib().vx_tree_root_id(self.h)
    def adopt_root(self, root: int): _ck(lib().vx_tree_adopt_root(self.h, int(root)))
    def is_empty(self) -> bool: return bool(lib().vx_tree_is_empty(self.h))
    def is_leaf(self) -> bool: return bool(lib().vx_tree_is_leaf(self.h))
    def is_dirty(self) -> bool: return bool(lib().vx_tree_is_dirty(self.h))
    def mark_dirty(self): lib().vx_tree_mark_dirty(self.h)
    def clear_dirty(self): lib().vx_tree_clear_dirty(self.h)
    def voxels_per_axis(self) -> int: return lib().vx_tree_voxels_per_axis(self.h)


class ChunkSet:
    """n (tree, batch) pairs with their handle arrays built once — the grid driver's view of a world
    (world/voxmodel.rs:27-59 keeps coord -> chunk; voxelis-voxelize/src/lib.rs:357-361 loops over it)."""

    def __init__(self, trees, batches):
        assert len(trees) == len(batches)
        self.trees, self.batches, self.n = list(trees), list(batches), len(trees)
        self.th = (C.c_void_p * self.n)(*[t.h for t in self.trees])
        self.bh = (C.c_void_p * self.n)(*[b.h for b in self.batches])
        self.changed = np.zeros(self.n, np.uint8)

    def apply(self, interner: VoxInterner):
        _ck(lib().vx_apply_batches(interner.h, self.th, self.bh, self.n, _ptr(self.changed)))
        return self.changed

    def forget(self):
        """After interner.reset(): every tree is empty again."""
        _ck(lib().vx_trees_forget(self.th, self.n))

    def roots(self) -> np.ndarray:
        f = lib().vx_tree_root_id
        return np.array([f(h) for h in self.th], np.uint64)


def apply_batches(interner: VoxInterner, trees, batches):
    """New multi-chunk entry (vx_apply_batches): result == serial application in index order."""
    return ChunkSet(trees, batches).apply(interner).astype(bool)


def trees_forget(trees):
    """After interner.reset(): make the trees empty again (vx_trees_forget)."""
    n = len(trees)
    _ck(lib().vx_trees_forget((C.c_void_p * n)(*[t.h for t in trees]), n))
