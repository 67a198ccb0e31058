"""Global dedup variant (BASELINE.json config 5): per-rank interners merged into hash-partitioned global
shards must hold exactly the nodes ONE shared interner would hold (the reference's model,
world/voxmodel.rs:31-32) — same unique branch / leaf counts, isomorphic DAG."""
import numpy as np
import pytest

from voxelis_b200 import workloads as wl

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_merge_matches_single_interner(gpu_api, oracle_api, G, dtype):
    from voxelis_b200 import dedup
    depth, per_rank = 4, 6
    parts_m, parts_v = [], []
    for r in range(G):      # overlapping content across ranks: the same terrain strip + rank-specific noise
        m1, v1 = wl.terrain_world((2, 1, 2), depth, "surface_and_below", dtype, x_chunk_offset=0, materials=3)
        m2, v2 = wl.batch_from_function(depth, wl.p_random(4), dtype, 2, chunk_arg=[100 + r, 7])
        parts_m.append(np.concatenate([m1, m2]))
        parts_v.append(np.concatenate([v1, v2]))
    assert parts_m[0].shape[0] == per_rank
    locals_, roots = [], []
    for r in range(G):
        it = gpu_api.VoxInterner.with_memory_budget(32 << 20, dtype)
        rt, _ = it.apply_batches_slab(depth, parts_m[r], parts_v[r])
        locals_.append(it)
        roots.append(rt)
    shards, groots, summary = dedup.global_dedup_local(locals_, roots, 32 << 20, dtype)
    # the reference's model: every chunk of every rank applied to ONE interner
    ref = oracle_api.VoxInterner(64 << 20, dtype)
    rroots, _ = ref.apply_batches_fresh(depth, np.concatenate(parts_m), np.concatenate(parts_v))
    rd = ref.download()
    rs = oracle_api.dag_signature(rd["children"], rd["values"], rroots, depth, want_stream=True)
    assert (summary["branches"], summary["leaves"]) == (rs["branches"], rs["leaves"])
    assert sum(s.next_index - 1 for s in shards) == rs["branches"] + rs["leaves"]
    # merged DAG isomorphic to the single-interner DAG
    children, values, remap = dedup.merged_pools(shards)
    mroots = remap(np.concatenate(groots))
    ms = oracle_api.dag_signature(children, values, mroots, depth, want_stream=True)
    assert np.array_equal(ms["stream"], rs["stream"])
    assert ms["per_depth"] == rs["per_depth"]
    # per-GPU interners together hold at least as many nodes as the merged shards
    assert sum(it.next_index - 1 for it in locals_) >= rs["branches"] + rs["leaves"]
    # every owner only holds keys that hash to it: owner bits of its ids
    for o, rt in enumerate(groots):
        assert all(dedup.owner_of(g) < G for g in rt if g)


@pytest.mark.parametrize("dtype", [wl.U8, wl.I32], ids=["u8", "i32"])
def test_world_entry_single_rank(gpu_api, oracle_api, dtype):
    """vx_world_global_dedup through the C entry with a one-rank world (NCCL send / recv to self): the shard ends up
    holding exactly the nodes of ONE interner over the same chunks, and the global roots describe the same DAG."""
    from voxelis_b200 import dedup
    vx = gpu_api
    depth = 5
    m1, v1 = wl.terrain_world((3, 2, 3), depth, "surface_and_below", dtype, materials=3)
    m2, v2 = wl.batch_from_function(depth, wl.p_random(4), dtype, 3)
    masks, values = np.concatenate([m1, m2]), np.concatenate([v1, v2])
    world = vx.World(1, 0, vx.World.unique_id(), 0)
    local = vx.VoxInterner.with_memory_budget(64 << 20, dtype)
    roots, _ = local.apply_batches_slab(depth, masks, values)
    for rep in range(2):                                   # the world's buffers are reused by a second merge
        shard = vx.VoxInterner.with_memory_budget(64 << 20, dtype)
        groots, summ = world.global_dedup(local, shard, roots)
        ref = oracle_api.VoxInterner(64 << 20, dtype)
        rroots, _ = ref.apply_batches_fresh(depth, masks, values)
        rd = ref.download()
        rs = oracle_api.dag_signature(rd["children"], rd["values"], rroots, depth, want_stream=True)
        assert (summ["branches"], summ["leaves"]) == (rs["branches"], rs["leaves"])
        assert shard.next_index - 1 == rs["branches"] + rs["leaves"] == summ["local_nodes_all_ranks"]
        children, vals, remap = dedup.merged_pools([shard])
        ms = oracle_api.dag_signature(children, vals, remap(groots), depth, want_stream=True)
        assert np.array_equal(ms["stream"], rs["stream"])
    world.close()


def test_world_entry_two_gpus(gpu_api):
    """The real exchange: torchrun with one process per GPU (tests/multigpu_dedup_check.py); needs >= 2 devices."""
    import os
    import subprocess
    import sys
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least two GPUs (gpurun --gpus 2)")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(here, "multigpu_dedup_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "DEDUP_OK" in res.stdout
