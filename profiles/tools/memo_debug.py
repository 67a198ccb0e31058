"""Diagnostic: does the bulk builder's unit memo alias repeated units?  (vx_interner_debug_memo + the local-cache counter:
a unit that is aliased never runs block_node, so cache_hits_local stays at one unit's worth.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import voxelis_b200 as vx
from voxelis_b200 import workloads as wl

name, n = "checkerboard", 4096
m, v = wl.named_workload(name, n)
dm, dv = torch.from_numpy(m).cuda(), torch.from_numpy(v).cuda()
dr = torch.zeros(n, dtype=torch.int64, device="cuda")
stream = torch.cuda.Stream()
torch.cuda.synchronize()
for mode in ("own stream + sync", "torch stream + sync", "torch stream, 5 calls back to back", "own stream, 5 calls back to back"):
    it = vx.VoxInterner.with_memory_budget(256 << 20)
    s = stream.cuda_stream if "torch" in mode else 0
    reps = 5 if "back" in mode else 1
    for outer in range(2):
        for rep in range(reps):
            it.reset_async(s)
            it.apply_batches_device(5, n, dm.data_ptr(), dv.data_ptr(), dr.data_ptr(), 0, stream=s)
        torch.cuda.synchronize()
        it.sync()
        print(mode, outer, "memo", [hex(x) for x in it.debug_memo()],
              {k: v for k, v in it.debug_counters().items() if k in ("cache_hits_local", "probe_steps")}, flush=True)
    del it
