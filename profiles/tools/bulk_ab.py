import sys, os, json, numpy as np, torch
sys.path.insert(0, "/root/repo")
import voxelis_b200 as vx
from voxelis_b200 import workloads as wl
dev = torch.device("cuda:0")
def run(name, masks, values, depth, budget, dtype=vx.U8, steps=20):
    n = masks.shape[0]
    dm = torch.from_numpy(masks).to(dev); dv = torch.from_numpy(values).to(dev)
    roots = torch.zeros(n, dtype=torch.int64, device=dev); ch = torch.zeros(n, dtype=torch.uint8, device=dev)
    out = {}
    for builder in ("fused", "bulk"):
        os.environ["VX_BUILDER"] = builder
        it = vx.VoxInterner.with_memory_budget(budget, dtype)
        st = torch.cuda.Stream(dev)
        ts = []; stages = None
        for i in range(2 * (steps + 3)):
            if i == steps + 3:
                total_ms = float(np.median(ts)); ts = []      # second half: per-launch events (no launch chaining)
                it.profile_stages(True)
            it.reset_async(st.cuda_stream)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st)
            it.apply_batches_device(depth, n, dm.data_ptr(), dv.data_ptr(), roots.data_ptr(), ch.data_ptr(), stream=st.cuda_stream)
            e1.record(st); st.synchronize()
            if i % (steps + 3) >= 3: ts.append(e0.elapsed_time(e1))
            stages = it.stage_ms()
        it.sync()
        s = it.stats()
        out[builder] = dict(ms=total_ms, ms_profiled=float(np.median(ts)), nodes=s["alive_nodes"], stages=[(a, round(b, 4)) for a, b in stages],
                            roots_sum=int(roots.sum().item() & 0xFFFFFFFF))
        del it
    print(name, json.dumps(out), flush=True)
which = sys.argv[1:] or ["perlin", "checker", "random", "below", "d7", "i32"]
if "sizes" in which:      # where the two builders cross over (auto threshold in use_bulk_builder)
    m, v = wl.terrain_world((16, 4, 16), 5, "surface_only", wl.U8)
    keep = np.flatnonzero(m[:, :, 0].any(1))
    for n in (16, 64, 128, 256, 1024):
        idx = keep[:n] if len(keep) >= n else np.arange(n)
        run(f"perlin_nonempty_x{n}", m[idx].copy(), v[idx].copy(), 5, 256 << 20)
if "perlin" in which:
    m, v = wl.terrain_world((64, 8, 64), 5, "surface_only", wl.U8); run("perlin", m, v, 5, 256 << 20)
if "below" in which:
    m, v = wl.terrain_world((64, 8, 64), 5, "surface_and_below", wl.U8); run("below", m, v, 5, 256 << 20)
if "checker" in which:
    m, v = wl.named_workload("checkerboard", 4096, 5, wl.U8); run("checker4096", m, v, 5, 256 << 20)
if "random" in which:
    m, v = wl.batch_from_function(5, wl.p_random(255), wl.U8, 1024); run("random255x1024", m, v, 5, 2 << 30, steps=5)
if "d7" in which:
    m, v = wl.batch_from_function(7, wl.p_random(255), wl.U8, 8); run("d7x8", m, v, 7, 2 << 30, steps=5)
if "i32" in which:
    m, v = wl.named_workload("sum", 512, 5, wl.I32); run("sum_i32x512", m, v, 5, 1 << 30, vx.I32, steps=5)
