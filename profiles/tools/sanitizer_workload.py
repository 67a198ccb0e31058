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
# i32 + D6
m2, v2 = wl.batch_from_function(6, wl.p_random(4), wl.I32, 1)
it2 = vx.VoxInterner.with_memory_budget(64 << 20, vx.I32); it2.apply_batches_slab(6, m2, v2)
print("sanitizer workload done", it.stats()["alive_nodes"], it2.stats()["alive_nodes"])
