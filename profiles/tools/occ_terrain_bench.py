"""The steps either side of the build path, on the perlin world (64x8x64 chunks of 32^3 u8), all in device memory:
  1. vx_terrain_heights_device + vx_terrain_batches_device  (utils/shapes.rs:273-357)
  2. vx_apply_batches_device                                (the hot path)
  3. vx_occupancy_masks, 2x2x2 chunks per 64^3 builder      (utils/mesh.rs:418-596)
timed with CUDA events on the interner's stream, next to the CPU oracle's generate_occupancy_masks restatement on a
bounded sample.  Usage: python profiles/tools/occ_terrain_bench.py [surface_only|surface_and_below]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import voxelis_b200 as vx
from voxelis_b200 import workloads as wl
from oracle import oracle as o

variant = sys.argv[1] if len(sys.argv) > 1 else "surface_and_below"
materials = 1 if variant == "surface_only" else 3
depth, grid = 5, (64, 8, 64)
gx, gy, gz = grid
N = 1 << depth
B = wl.blocks_per_chunk(depth)
n = gx * gy * gz
dev = torch.device("cuda", 0)
it = vx.VoxInterner.with_memory_budget(256 << 20, vx.U8, 0)
stream = torch.cuda.ExternalStream(it.stream, device=dev)
h = torch.empty((gx * N, gz * N), dtype=torch.int32, device=dev)
m = torch.empty((n, B, 2), dtype=torch.uint8, device=dev)
v = torch.empty((n, B, 8), dtype=torch.uint8, device=dev)
roots = torch.zeros(n, dtype=torch.int64, device=dev)
changed = torch.zeros(n, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()

def ev():
    return torch.cuda.Event(enable_timing=True)

def gen():
    it.terrain_heights_device(gx * N, gz * N, h.data_ptr(), wl.SEED_BASE, gy * N)
    it.terrain_batches_device(depth, grid, h.data_ptr(), m.data_ptr(), v.data_ptr(), variant == "surface_only", materials)

def build():
    it.reset_async()
    it.apply_batches_device(depth, n, m.data_ptr(), v.data_ptr(), roots.data_ptr(), changed.data_ptr())

def timed(f, reps=20, warm=3):
    for _ in range(warm):
        f()
    it.sync()
    a, b = ev(), ev()
    a.record(stream)
    for _ in range(reps):
        f()
    b.record(stream)
    it.sync()
    return a.elapsed_time(b) / reps

ms_gen = timed(gen)
ms_build = timed(build)
ms_both = timed(lambda: (gen(), build()))
# occupancy: 2x2x2 chunks per builder (mesh.rs:598-606), device outputs
hroots = roots.cpu().numpy().astype(np.uint64)
idx = np.arange(n)
cx, cy, cz = idx // (gy * gz), (idx // gz) % gy, idx % gz
builder_of = ((cx // 2) * (gy // 2) + cy // 2) * (gz // 2) + cz // 2
offsets = np.stack([(cx % 2) * 32, (cy % 2) * 32, (cz % 2) * 32], 1).astype(np.uint32)
nb = n // 8
M = 3 if materials == 3 else 2
L = vx.lib()
import ctypes as C
d_global = torch.empty((nb, 3 * 4096), dtype=torch.int64, device=dev)
d_active = torch.empty((nb, 6), dtype=torch.int64, device=dev)
d_nmat = torch.empty(nb, dtype=torch.int32, device=dev)
d_ids = torch.empty((nb, M), dtype=torch.int64, device=dev)
d_counts = torch.empty((nb, M), dtype=torch.int64, device=dev)
d_pm = torch.empty((nb, M, 3 * 4096), dtype=torch.int64, device=dev)
torch.cuda.synchronize()
bo = np.ascontiguousarray(builder_of, np.uint32)
vp = lambda t: C.c_void_p(t.data_ptr())
npp = lambda a: a.ctypes.data_as(C.c_void_p)
def occ():
    rc = L.vx_occupancy_masks(it.h, depth, 0, n, npp(hroots), npp(offsets), npp(bo), nb, M, vp(d_global), vp(d_active),
                              vp(d_nmat), vp(d_ids), vp(d_counts), vp(d_pm))
    assert rc == 0, L.vx_last_error()
occ()
t0 = time.perf_counter()
a, b = ev(), ev()
a.record(stream)
for _ in range(5):
    occ()
b.record(stream)
it.sync()
ms_occ_wall = (time.perf_counter() - t0) / 5 * 1e3
ms_occ_dev = a.elapsed_time(b) / 5          # includes the host-side cell placement between launches
nmat = d_nmat.cpu().numpy()
out_bytes = int(nb * (3 * 4096 * 8 + 48) + int(nmat.sum()) * 3 * 4096 * 8)
# CPU oracle on a bounded sample of builders
sample = 64
em, evv = wl.terrain_world(grid, depth, variant, wl.U8, materials=materials)
c = o.VoxInterner(256 << 20, 0)
croots, _ = c.apply_batches_fresh(depth, em, evv)
same_voxels = bool(np.array_equal(m.cpu().numpy(), em) and np.array_equal(v.cpu().numpy(), evv))
pick = np.linspace(0, nb - 1, sample).astype(int)
t0 = time.perf_counter()
ok = True
for bsel in pick:
    sel = np.nonzero(builder_of == bsel)[0]
    want = c.occupancy_masks(croots[sel], depth, offsets[sel], max_materials=M)
    k = len(want["material_ids"])
    ok &= bool(np.array_equal(want["global"], d_global[bsel].cpu().numpy().view(np.uint64)))
    ok &= k == nmat[bsel] and bool(np.array_equal(want["per_material"], d_pm[bsel, :k].cpu().numpy().view(np.uint64)))
cpu_s = time.perf_counter() - t0
print(json.dumps({"world": f"perlin 64x8x64 d5 u8 {variant} materials={materials}", "chunks": n,
                  "generate_ms": ms_gen, "generate_chunks_per_s": n / ms_gen * 1e3,
                  "generate_bytes_written": int(m.numel() + v.numel()), "generate_GBps": (m.numel() + v.numel()) / ms_gen / 1e6,
                  "build_ms": ms_build, "generate_plus_build_ms": ms_both,
                  "generate_plus_build_chunks_per_s": n / ms_both * 1e3,
                  "generated_equals_numpy_world": same_voxels,
                  "occupancy_builders": nb, "occupancy_ms_device_span": ms_occ_dev, "occupancy_ms_wall": ms_occ_wall,
                  "occupancy_output_bytes": out_bytes, "occupancy_output_GBps_wall": out_bytes / ms_occ_wall / 1e6,
                  "occupancy_chunks_per_s_wall": n / ms_occ_wall * 1e3,
                  "occupancy_matches_oracle_on_sample": bool(ok), "oracle_sample_builders": sample,
                  "oracle_chunks_per_s_incl_compare": sample * 8 / cpu_s}))
