import numpy as np, sys
sys.path.insert(0, "/root/repo")
import voxelis_b200 as vx
from voxelis_b200 import workloads as wl
parts = [wl.terrain_world((3, 2, 3), 5, "surface_and_below", wl.U8, materials=3),
         wl.batch_from_function(5, wl.p_random(255), wl.U8, 3),
         wl.named_workload("checkerboard", 4, 5, wl.U8), wl.named_workload("sum", 2, 5, wl.U8)]
masks = np.concatenate([p[0] for p in parts]); values = np.concatenate([p[1] for p in parts])
it = vx.VoxInterner.with_memory_budget(64 << 20)
roots, ch = it.apply_batches_slab(5, masks, values)
d = it.roots_to_vec(roots[:4], 5)
# edit + release path
t = vx.VoxTree(5); b = t.create_batch()
b.masks[:] = masks[20]; b.values[:] = values[20]; b.mark_patched(); t.apply_batch(it, b)
b.masks[:] = masks[21]; b.values[:] = values[21]; t.apply_batch(it, b); t.clear(it)
# single batch through the in-place path: a solid chunk (solid_unit), the same with a fill, a nearly solid one
mu, vu = wl.named_workload("uniform", 1)
t2 = vx.VoxTree(5); b2 = t2.create_batch(); b2.assign(mu[0], vu[0]); t2.apply_batch(it, b2); assert t2.is_leaf()
b2.clear(); b2.fill(it, 3); t2.apply_batch(it, b2)
vu2 = vu.copy(); vu2[0, 700, 3] = 9
b2.clear(); b2.assign(mu[0], vu2[0]); t2.apply_batch(it, b2); t2.clear(it)
# i32 + D6
m2, v2 = wl.batch_from_function(6, wl.p_random(4), wl.I32, 1)
it2 = vx.VoxInterner.with_memory_budget(64 << 20, vx.I32); it2.apply_batches_slab(6, m2, v2)
print("sanitizer workload done", it.stats()["alive_nodes"], it2.stats()["alive_nodes"])
# host batch handles: staging kernels (occupancy-bitmap and mask variants), listed plan kernel, slices
trees = [vx.VoxTree(5) for _ in range(24)]
batches = [t.create_batch() for t in trees]
for i, b in enumerate(batches):
    if i % 3 == 0:
        b.masks[:] = masks[i % len(masks)]; b.values[:] = values[i % len(values)]; b.mark_patched()
    elif i % 3 == 1:
        b.assign(masks[i % len(masks)], values[i % len(values)])
    else:
        b.set_many(np.random.default_rng(i).integers(0, 32, (300, 3)), np.random.default_rng(i).integers(0, 4, 300))
it3 = vx.VoxInterner.with_memory_budget(64 << 20)
vx.apply_batches(it3, trees[1::3] + trees[2::3], batches[1::3] + batches[2::3])      # API-only batches: occ kernel
vx.apply_batches(it3, trees[0::3], batches[0::3])                                    # raw batches: masks kernel
# VTM export / import
r3 = np.array([t.get_root_id() for t in trees], np.uint64)
pos = np.zeros((len(trees), 3), np.int32)
payload = it3.model_serialize(pos, r3)
it4 = vx.VoxInterner.with_memory_budget(64 << 20)
_, r4 = it4.model_deserialize(payload)
assert np.array_equal(it4.roots_to_vec(r4[:2], 5), it3.roots_to_vec(r3[:2], 5))
print("sanitizer workload (handles + VTM) done", it3.stats()["alive_nodes"], len(payload))
