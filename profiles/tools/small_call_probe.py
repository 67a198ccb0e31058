"""Per-launch CUDA-event times of small apply calls: the first 1/8 .. 1/512 of the perlin world (what one GPU of a
strong-scaling run gets), and an EMPTY slab of the same sizes (the floor of the five launches)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import voxelis_b200 as vx
from voxelis_b200 import workloads as wl

m, v = wl.terrain_world((64, 8, 64), 5, "surface_only", wl.U8)
for label, mm, vv in (("perlin", m, v), ("empty", np.zeros_like(m), np.zeros_like(v))):
    for n in (32768, 4096, 512, 64):
        it = vx.VoxInterner.with_memory_budget(256 << 20)
        dm, dv = torch.from_numpy(mm[:n]).cuda(), torch.from_numpy(vv[:n]).cuda()
        dr = torch.zeros(n, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for rep in range(23):   # whole call without the stage events
            it.reset_async()
            it.sync()
            s = torch.cuda.current_stream()
            t0 = __import__("time").perf_counter()
            it.apply_batches_device(5, n, dm.data_ptr(), dv.data_ptr(), dr.data_ptr())
            it.sync()
            if rep >= 3:
                tot += (__import__("time").perf_counter() - t0) / 20
        it.profile_stages(True)
        acc = {}
        for rep in range(13):
            it.reset_async()
            it.apply_batches_device(5, n, dm.data_ptr(), dv.data_ptr(), dr.data_ptr())
            it.sync()
            if rep >= 3:
                for k, ms in it.stage_ms():
                    acc[k] = acc.get(k, 0) + ms / 10
        print(label, n, "wall %.1f us" % (tot * 1e6), "stages sum %.1f us" % (sum(acc.values()) * 1e3),
              {k.replace("bulk_", "").replace("_kernel", ""): round(x * 1e3, 1) for k, x in acc.items()}, flush=True)
        del it, dm, dv
