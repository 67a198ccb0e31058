"""compute-sanitizer workload for the kernels either side of the path: occupancy planes (shared-memory path, word-owner
path, levels of detail), device terrain generation, voxeliser.  Small sizes: the sanitizer is ~50x slower."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import voxelis_b200 as vx, meshes
from voxelis_b200 import workloads as wl
dev = torch.device("cuda", 0)
# terrain on the device -> build -> occupancy (shared-memory path; M = 2 and M = 5), several LODs, i32 too
for dtype, tdt in ((vx.U8, torch.uint8), (vx.I32, torch.int32)):
    it = vx.VoxInterner.with_memory_budget(64 << 20, dtype)
    grid, depth, N, B = (4, 4, 2), 5, 32, 4096
    n = grid[0] * grid[1] * grid[2]
    h = torch.empty((grid[0] * N, grid[2] * N), dtype=torch.int32, device=dev)
    m = torch.empty((n, B, 2), dtype=torch.uint8, device=dev); v = torch.empty((n, B, 8), dtype=tdt, device=dev)
    roots = torch.zeros(n, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    it.terrain_heights_device(grid[0] * N, grid[2] * N, h.data_ptr(), wl.SEED_BASE, grid[1] * N)
    it.terrain_batches_device(depth, grid, h.data_ptr(), m.data_ptr(), v.data_ptr(), False, 3)
    it.apply_batches_device(depth, n, m.data_ptr(), v.data_ptr(), roots.data_ptr()); it.sync()
    r = roots.cpu().numpy().astype(np.uint64)
    idx = np.arange(n); cx, cy, cz = idx // 8, (idx // 2) % 4, idx % 2
    bo = ((cx // 2) * 2 + cy // 2).astype(np.uint32); offs = np.stack([(cx % 2) * 32, (cy % 2) * 32, cz * 32], 1)
    for M in (2, 5):
        try:
            out = it.occupancy_masks(r, depth, offs, bo, 4, max_materials=M)
        except vx.VoxelisError as e:
            print("expected (3 materials, room for 2):", e)
    for lod in (1, 2, 3, 4):
        S = 32 >> lod
        it.occupancy_masks(r[:8], depth, [((i % 2) * S, (i // 2 % 2) * S, (i // 4) * S) for i in range(8)], lod=lod)
# many materials: word-owner kernels (spill path)
mm, vv = wl.batch_from_function(5, wl.p_random(255), wl.U8, 2)
it = vx.VoxInterner.with_memory_budget(64 << 20)
rr, _ = it.apply_batches_slab(5, mm, vv)
out = it.occupancy_masks(rr, 5, [(0, 0, 0), (32, 32, 32)], max_materials=255)
# voxeliser
verts, faces = meshes.uv_sphere((0.9, 0.8, 0.85), 0.7, 10, 6)
plan = vx.voxelize_plan(4, 0.5, verts.min(0), verts, faces)
nchunks = len(plan[0])
m = torch.empty((nchunks, 512, 2), dtype=torch.uint8, device=dev); v = torch.empty((nchunks, 512, 8), dtype=torch.uint8, device=dev)
hp = torch.empty(nchunks, dtype=torch.uint8, device=dev); torch.cuda.synchronize()
it.voxelize_chunks_device(4, 0.5, verts.min(0), verts, faces, plan, m.data_ptr(), v.data_ptr(), hp.data_ptr())
print("sanitizer workload 2 done", int(out["n_materials"][0]), nchunks, int(hp.sum()))
