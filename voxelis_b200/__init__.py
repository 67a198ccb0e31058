"""voxelis_b200 — B200-native batched SVO-DAG build path behind the reference's
VoxTree / VoxInterner / Batch API (voxelis/src/spatial/voxops.rs:8-35).

This package is a thin ctypes mirror of the C ABI in include/voxelis_b200.h; every compute
call runs hand-written sm_100a kernels in libvoxelis_b200.so.  There is no CPU fallback:
importing works anywhere, but ``lib()`` raises if the CUDA library is missing and interner
creation fails without a CUDA device.
"""
from .api import (I32, U8, Batch, ChunkSet, VoxelisError, VoxInterner, VoxTree, World, apply_batches, device_count, voxelize_plan,  # noqa: F401
                  trees_forget,
                  id_index, id_is_branch, id_is_leaf, id_mask, id_types, lib)

__all__ = ["U8", "I32", "Batch", "VoxInterner", "VoxTree", "VoxelisError", "apply_batches", "trees_forget", "ChunkSet", "device_count", "voxelize_plan", "lib", "World"]
