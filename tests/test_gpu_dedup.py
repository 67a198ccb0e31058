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
