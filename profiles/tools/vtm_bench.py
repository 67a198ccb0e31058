"""VTM export / import of the perlin world from the device pools, next to the CPU oracle's serialise
(world/voxmodel.rs:177-294 restated).  Usage: python profiles/tools/vtm_bench.py"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
import voxelis_b200 as vx
from oracle import oracle as o

masks, values = bench.make_world(0)
n = masks.shape[0]
gx, gy, gz = bench.GRID
idx = np.arange(n)
positions = np.stack([idx // (gy * gz), (idx // gz) % gy, idx % gz], 1).astype(np.int32)
it = vx.VoxInterner.with_memory_budget(bench.BUDGET, vx.U8, 0)
roots, changed = it.apply_batches_slab(bench.DEPTH, masks, values)
def timed(f, reps=5):
    f()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = f()
    return (time.perf_counter() - t0) / reps * 1e3, r
ms_ser, payload = timed(lambda: it.model_serialize(positions, roots))
tmp = tempfile.mkdtemp()
ms_raw, _ = timed(lambda: it.export_vtm(os.path.join(tmp, "w.vtm"), "dunes", 5, 1.0, bench.GRID, positions, roots, compress=False))
ms_zst, _ = timed(lambda: it.export_vtm(os.path.join(tmp, "wz.vtm"), "dunes", 5, 1.0, bench.GRID, positions, roots, compress=True))
def imp(path):
    g2 = vx.VoxInterner.with_memory_budget(bench.BUDGET, vx.U8, 0)
    r = g2.import_vtm(path, max_chunks=n)
    g2.sync()
    return g2, r
ms_imp, (g2, (meta, pos2, roots2)) = timed(lambda: imp(os.path.join(tmp, "w.vtm")), reps=3)
same = bool(np.array_equal(g2.roots_to_vec(roots2[:256], 5), it.roots_to_vec(roots[:256], 5)))
c = o.VoxInterner(bench.BUDGET, 0)
croots, _ = c.apply_batches_fresh(bench.DEPTH, masks, values)
ms_cpu, cpayload = timed(lambda: c.model_serialize(positions, croots), reps=3)
print(json.dumps({"world": "perlin 64x8x64 d5 u8", "alive_nodes": it.stats()["alive_nodes"], "chunks": n,
                  "payload_bytes": len(payload), "oracle_payload_bytes": len(cpayload),
                  "gpu_model_serialize_ms": ms_ser, "gpu_export_vtm_raw_ms": ms_raw, "gpu_export_vtm_zstd7_ms": ms_zst,
                  "file_bytes_raw": os.path.getsize(os.path.join(tmp, "w.vtm")), "file_bytes_zstd": os.path.getsize(os.path.join(tmp, "wz.vtm")),
                  "gpu_import_vtm_raw_ms": ms_imp, "import_voxels_equal_first_256_chunks": same,
                  "cpu_oracle_model_serialize_ms": ms_cpu}))
