"""Run under torchrun (one process per GPU): every rank builds its own chunks into a private interner, the ranks
merge through vx_world_global_dedup (the C entry: NCCL send / recv issued by the library), rank 0 gathers the shards
and checks them against ONE interner built by the oracle over all ranks' chunks (the reference's model,
world/voxmodel.rs:31-32): same unique branch / leaf counts, isomorphic DAG.  Prints DEDUP_OK."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import voxelis_b200 as vx  # noqa: E402
from voxelis_b200 import dedup  # noqa: E402
from voxelis_b200 import workloads as wl  # noqa: E402


def chunks_of(rank, depth, dtype):
    m1, v1 = wl.terrain_world((4, 2, 4), depth, "surface_and_below", dtype, x_chunk_offset=2 * rank, materials=3)   # overlaps the neighbour
    m2, v2 = wl.batch_from_function(depth, wl.p_random(4), dtype, 3, chunk_arg=[50 + rank, 7, 8])
    m3, v3 = wl.named_workload("sum", 2, depth, dtype)
    return np.concatenate([m1, m2, m3]), np.concatenate([v1, v2, v3])


def main():
    rank, world_size, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world = vx.World.from_torch_distributed(local_rank)
    ok = True
    for dtype in (wl.U8, wl.I32):
        depth = 5
        masks, values = chunks_of(rank, depth, dtype)
        local = vx.VoxInterner.with_memory_budget(128 << 20, dtype, local_rank)
        roots, _ = local.apply_batches_slab(depth, masks, values)
        shard = vx.VoxInterner.with_memory_budget(128 << 20, dtype, local_rank)
        groots, summ = world.global_dedup(local, shard, roots)
        d = shard.download()
        payload = {"n": d["n"], "children": d["children"], "values": d["values"], "groots": groots}
        gathered = [None] * world_size if rank == 0 else None
        dist.gather_object(payload, gathered, 0)
        if rank == 0:
            from oracle import oracle
            ref = oracle.VoxInterner(512 << 20, dtype)
            allm, allv = zip(*[chunks_of(r, depth, dtype) for r in range(world_size)])
            rroots, _ = ref.apply_batches_fresh(depth, np.concatenate(allm), np.concatenate(allv))
            rd = ref.download()
            rs = oracle.dag_signature(rd["children"], rd["values"], rroots, depth, want_stream=True)
            ok &= (summ["branches"], summ["leaves"]) == (rs["branches"], rs["leaves"])
            ok &= sum(g["n"] - 1 for g in gathered) == rs["branches"] + rs["leaves"]

            class _S:                                  # merged_pools wants objects with download()
                def __init__(self, g): self.g = g
                def download(self): return self.g
            children, vals, remap = dedup.merged_pools([_S(g) for g in gathered])
            ms = oracle.dag_signature(children, vals, remap(np.concatenate([g["groots"] for g in gathered])), depth, want_stream=True)
            ok &= bool(np.array_equal(ms["stream"], rs["stream"]))
            print(f"dtype {dtype}: G={world_size} global {summ['branches']}+{summ['leaves']} vs one interner "
                  f"{rs['branches']}+{rs['leaves']}; local nodes {summ['local_nodes_all_ranks']}; ok={ok}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    world.close()
    dist.destroy_process_group()
    if rank == 0 and int(flag.item()) == 1:
        print("DEDUP_OK", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
