"""DRAM bytes of one apply call, measured live.  bench.py runs THIS file under

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --print-units base --csv ...

(one warm-up step + one measured step of the named workload, device-resident batches, the same generators as the
bench) and reads the launches after the last interner reset (`init_scalars_kernel`) = the measured apply call.
`parse()` is the reader; a number taken under the profiler is only ever used for BYTES, never for time.
"""
from __future__ import annotations

import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def inputs(workload: str):
    from voxelis_b200 import workloads as wl
    if workload == "perlin":
        return wl.terrain_world((64, 8, 64), 5, "surface_only", wl.U8), 5, 256 << 20
    if workload == "below":
        return wl.terrain_world((64, 8, 64), 5, "surface_and_below", wl.U8, materials=3), 5, 256 << 20
    name, _, n = workload.partition(":")
    return wl.named_workload(name, int(n or 4096)), 5, (2 << 30) if name.startswith("random") else (256 << 20)


def main():
    import torch
    import voxelis_b200 as vx
    (masks, values), depth, budget = inputs(sys.argv[1])
    dev = torch.device("cuda", 0)
    n = masks.shape[0]
    dm, dv = torch.from_numpy(masks).to(dev), torch.from_numpy(values).to(dev)
    dr = torch.zeros(n, dtype=torch.int64, device=dev)
    it = vx.VoxInterner.with_memory_budget(budget, vx.U8, 0)
    torch.cuda.synchronize()
    for _ in range(2):
        it.reset_async()
        it.apply_batches_device(depth, n, dm.data_ptr(), dv.data_ptr(), dr.data_ptr())
        it.sync()


def parse(path: str):
    """-> {"dram_bytes": read + write of the last apply call, "read", "write", "launches": [(kernel, bytes, ns)]}"""
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    per = {}
    order = []
    for r in rows:
        k = int(r["ID"])
        if k not in per:
            per[k] = {"name": r["Kernel Name"], "read": 0.0, "write": 0.0, "ns": 0.0}
            order.append(k)
        v = float(r["Metric Value"].replace(",", "") or 0)
        m = r["Metric Name"]
        if m == "dram__bytes_read.sum":
            per[k]["read"] = v
        elif m == "dram__bytes_write.sum":
            per[k]["write"] = v
        elif m == "gpu__time_duration.sum":
            per[k]["ns"] = v
    last_reset = max((i for i, k in enumerate(order) if "init_scalars_kernel" in per[k]["name"]), default=-1)
    call = [per[k] for k in order[last_reset + 1:]]
    rd, wr = sum(c["read"] for c in call), sum(c["write"] for c in call)
    return {"dram_bytes": int(rd + wr), "read": int(rd), "write": int(wr),
            "launches": [(c["name"].split("(")[0][-48:], int(c["read"] + c["write"]), int(c["ns"])) for c in call]}


if __name__ == "__main__":
    main()
