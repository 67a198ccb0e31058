"""Workload for ncu captures of the generation and occupancy kernels: the perlin world generated on the device,
built, and unfolded into 4096 occupancy builders once.  Usage: ncu -k regex:'occ_|terrain_' ... python this.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import voxelis_b200 as vx
from voxelis_b200 import workloads as wl

variant = sys.argv[1] if len(sys.argv) > 1 else "surface_and_below"
materials = 1 if variant == "surface_only" else 3
depth, grid = 5, (64, 8, 64)
gx, gy, gz = grid
N, B, n = 32, 4096, gx * gy * gz
dev = torch.device("cuda", 0)
it = vx.VoxInterner.with_memory_budget(256 << 20, vx.U8, 0)
h = torch.empty((gx * N, gz * N), dtype=torch.int32, device=dev)
m = torch.empty((n, B, 2), dtype=torch.uint8, device=dev)
v = torch.empty((n, B, 8), dtype=torch.uint8, device=dev)
roots = torch.zeros(n, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
it.terrain_heights_device(gx * N, gz * N, h.data_ptr(), wl.SEED_BASE, gy * N)
it.terrain_batches_device(depth, grid, h.data_ptr(), m.data_ptr(), v.data_ptr(), variant == "surface_only", materials)
it.apply_batches_device(depth, n, m.data_ptr(), v.data_ptr(), roots.data_ptr())
it.sync()
hroots = roots.cpu().numpy().astype(np.uint64)
idx = np.arange(n)
cx, cy, cz = idx // (gy * gz), (idx // gz) % gy, idx % gz
bo = np.ascontiguousarray(((cx // 2) * (gy // 2) + cy // 2) * (gz // 2) + cz // 2, np.uint32)
offsets = np.stack([(cx % 2) * 32, (cy % 2) * 32, (cz % 2) * 32], 1).astype(np.uint32)
nb, M = n // 8, (3 if materials == 3 else 2)
outs = [torch.empty(s, dtype=t, device=dev) for s, t in (((nb, 3 * 4096), torch.int64), ((nb, 6), torch.int64),
        ((nb,), torch.int32), ((nb, M), torch.int64), ((nb, M), torch.int64), ((nb, M, 3 * 4096), torch.int64))]
torch.cuda.synchronize()
npp = lambda a: a.ctypes.data_as(C.c_void_p)
rc = vx.lib().vx_occupancy_masks(it.h, depth, 0, n, npp(hroots), npp(offsets), npp(bo), nb, M,
                                 *[C.c_void_p(t.data_ptr()) for t in outs])
assert rc == 0
print("materials per builder:", np.bincount(outs[2].cpu().numpy()))
