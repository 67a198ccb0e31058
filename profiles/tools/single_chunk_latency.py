"""BASELINE config 1 from the caller's side: wall clock of tree.apply_batch on a host Batch (one 32^3 chunk), and the kernel's share."""
import time, numpy as np, sys
sys.path.insert(0, '.')
import voxelis_b200 as vx
from voxelis_b200 import workloads as wl
BUDGET = 256 << 20
for name in ["uniform", "checkerboard", "random255"]:
    itl = vx.VoxInterner.with_memory_budget(BUDGET, vx.U8, 0)
    mu, vu = wl.named_workload(name, 1)
    tr = vx.VoxTree(5, vx.U8)
    b = tr.create_batch(); b.assign(mu[0], vu[0])
    ts = []
    for i in range(300):
        itl.reset(); vx.trees_forget([tr])
        t0 = time.perf_counter(); tr.apply_batch(itl, b); ts.append(time.perf_counter() - t0)
    ts = np.array(ts[20:]) * 1e6
    itl.profile_stages(True)
    ks = []
    for i in range(50):
        itl.reset(); vx.trees_forget([tr]); tr.apply_batch(itl, b)
        ks.append(sum(ms for _, ms in itl.stage_ms()) * 1e3)
    print(name, "wall median %.1f us  p10 %.1f;  kernel (CUDA events) median %.1f us" % (np.median(ts), np.percentile(ts, 10), np.median(ks)))
