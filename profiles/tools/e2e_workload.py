"""The end-to-end leg of bench.py on its own (for ncu / A-B runs): the perlin world as n Batch + n VoxTree
handles, a few vx_apply_batches calls.  Usage: python profiles/tools/e2e_workload.py [calls]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
import voxelis_b200 as vx

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 5
masks, values = bench.make_world(0)
n = masks.shape[0]
it = vx.VoxInterner.with_memory_budget(bench.BUDGET, vx.U8, 0)
trees = [vx.VoxTree(bench.DEPTH, vx.U8) for _ in range(n)]
batches = [t.create_batch() for t in trees]
for b, m, v in zip(batches, masks, values):
    b.assign(m, v)
cs = vx.ChunkSet(trees, batches)
ts = []
for _ in range(calls):
    it.reset_async()
    cs.forget()
    t0 = time.perf_counter()
    cs.apply(it)
    ts.append((time.perf_counter() - t0) * 1e3)
it.profile_stages(True)
it.reset_async(); cs.forget(); cs.apply(it)
print("apply ms:", " ".join(f"{t:.3f}" for t in ts), "| phases:", {k: round(v) for k, v in it.host_trace().items()})
