"""Voxeliser on the device: a UV sphere of ~65 k triangles, radius 4 chunk-lengths, 32^3 voxels per
chunk -> plan on the host, vx_voxelize_chunks_device, vx_apply_batches_device; next to the CPU oracle's
voxelize_chunk on a bounded sample of the same chunks.  Usage: python profiles/tools/voxelize_bench.py [nu nv]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import meshes
import voxelis_b200 as vx
from voxelis_b200 import workloads as wl
from oracle import oracle as o

nu, nv = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 128)
depth, cws, dtype = 5, 1.0, vx.I32
verts, faces = meshes.uv_sphere((4.3, 4.3, 4.3), 4.0, nu, nv)
mesh_min = verts.min(0)
t0 = time.perf_counter()
plan = vx.voxelize_plan(depth, cws, mesh_min, verts, faces)
plan_ms = (time.perf_counter() - t0) * 1e3
positions, pc, pf = plan
n, B = len(positions), wl.blocks_per_chunk(depth)
dev = torch.device("cuda", 0)
it = vx.VoxInterner.with_memory_budget(256 << 20, dtype, 0)
m = torch.empty((n, B, 2), dtype=torch.uint8, device=dev)
v = torch.empty((n, B, 8), dtype=torch.int32, device=dev)
hp = torch.empty(n, dtype=torch.uint8, device=dev)
roots = torch.zeros(n, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
def vox():
    it.voxelize_chunks_device(depth, cws, mesh_min, verts, faces, plan, m.data_ptr(), v.data_ptr(), hp.data_ptr())
vox()
if os.environ.get("VX_ONE_SHOT"):
    sys.exit(0)
t0 = time.perf_counter()
for _ in range(5):
    vox()
vox_ms = (time.perf_counter() - t0) / 5 * 1e3
it.apply_batches_device(depth, n, m.data_ptr(), v.data_ptr(), roots.data_ptr())
it.sync()
nodes = it.stats()["alive_nodes"]
set_voxels = int(torch.count_nonzero(v).item())
with_patches = int(hp.sum().item())
# oracle on a bounded sample
sample = np.linspace(0, n - 1, 24).astype(int)
gm, gv = m.cpu().numpy(), v.cpu().numpy()
t0 = time.perf_counter()
same = True
for c in sample:
    has, om, ov = o.voxelize_chunk(dtype, positions[c], depth, cws, mesh_min, faces[pf[pc == c]], verts)
    same &= bool(np.array_equal(om, gm[c]) and np.array_equal(ov, gv[c]))
cpu_s = time.perf_counter() - t0
print(json.dumps({"mesh": f"uv sphere {len(faces)} faces, radius 4 chunks, depth 5, i32", "chunks_planned": n,
                  "chunks_with_patches": with_patches, "pairs": int(len(pc)), "set_voxels": set_voxels,
                  "plan_host_ms": plan_ms, "voxelize_call_ms": vox_ms, "voxelize_chunks_per_s": n / vox_ms * 1e3,
                  "voxelize_faces_per_s": len(faces) / vox_ms * 1e3, "dag_nodes_after_build": nodes,
                  "oracle_sample_chunks": len(sample), "oracle_chunks_per_s_1_thread": len(sample) / cpu_s,
                  "sample_batches_identical": bool(same)}))
