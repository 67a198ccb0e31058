"""Where does the plan launch spend its time?  Per-launch CUDA-event times of one apply call on (a) an all-empty world,
(b) the perlin world, (c) the perlin world with every non-empty chunk moved to the front — same bytes of masks each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import voxelis_b200 as vx
from voxelis_b200 import workloads as wl

m, v = wl.terrain_world((64, 8, 64), 5, "surface_only", wl.U8)
n = m.shape[0]
worlds = {"empty": (np.zeros_like(m), np.zeros_like(v)), "perlin": (m, v)}
order = np.argsort(~m[:, :, 0].any(axis=1), kind="stable")
worlds["perlin_nonempty_first"] = (m[order], v[order])
for name, (mm, vv) in worlds.items():
    it = vx.VoxInterner.with_memory_budget(256 << 20)
    dm, dv = torch.from_numpy(mm).cuda(), torch.from_numpy(vv).cuda()
    dr = torch.zeros(n, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    it.profile_stages(True)
    acc = {}
    for rep in range(13):
        it.reset_async()
        it.apply_batches_device(5, n, dm.data_ptr(), dv.data_ptr(), dr.data_ptr())
        it.sync()
        if rep >= 3:
            for k, ms in it.stage_ms():
                acc[k] = acc.get(k, 0) + ms / 10
    print(name, {k: round(x * 1e3, 1) for k, x in acc.items()}, flush=True)
    del it, dm, dv
