"""Global dedup across per-GPU interners — BASELINE.json config 5, SURVEY §8e.

The reference keeps ONE interner for the whole world (voxelis/src/world/voxmodel.rs:31-32); its
design document proposes sharding the pattern map by ``hash % N`` once it becomes the bottleneck
(Voxelis Bible §3.9, §13).  Here every GPU first builds its chunks into a private interner (no
communication), then ``global_dedup`` merges them into hash-partitioned *global shards*:

    for height h = 0 (leaves) .. max:                      # parents need their children's global ids
        pack local nodes of height h by owner = hash(children as global ids) mod G   (vx_dedup_pack)
        all-to-all the 72-byte records to their owners                               (NCCL)
        owners intern them into their shard                                          (vx_interner_intern_records)
        all-to-all the 8-byte global ids back, gmap[local node] = global id          (vx_dedup_scatter)

The exchange goes through a *transport*: ``DistTransport`` = ``torch.distributed.all_to_all_single``
over NCCL (one process per GPU); ``LocalTransport`` runs G logical ranks inside one process on one
device (same kernels, the all-to-all degenerates to slicing) so the path is testable on a single GPU.
PyTorch is plumbing here (device buffers + the collective); every step that touches node data is a
kernel behind the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .api import VoxInterner, _ck, lib

REC_WORDS = 9


def owner_of(gid: int) -> int:
    return (int(gid) >> 44) & 7


def strip_owner(gid: int) -> int:
    return int(gid) & ~(7 << 44)


def _p(t):
    return C.c_void_p(t.data_ptr() if t is not None and t.numel() else 0)


def _sync(device):
    """The library's kernels run on the interner's own stream and synchronise on return; torch (and
    NCCL) work is queued on torch's streams.  Before handing a torch tensor to the library, wait for
    whatever torch still has in flight for it (an all-to-all, a cat, a slice copy)."""
    torch.cuda.current_stream(device).synchronize()


class _Rank:
    """State of one (logical) rank during the merge."""

    def __init__(self, rank: int, local: VoxInterner, shard: VoxInterner, device):
        self.rank, self.local, self.shard, self.device = rank, local, shard, device
        n = local.next_index
        self.heights = torch.empty(max(n, 1), dtype=torch.uint8, device=device)
        self.gmap = torch.zeros(max(n, 1), dtype=torch.int64, device=device)
        self.max_height = _ck(lib().vx_dedup_heights(local.h, _p(self.heights)))
        self.created = [0, 0]  # [branches, leaves] created in this rank's shard

    def pack(self, height: int, G: int):
        _sync(self.device)
        counts = np.zeros(G, np.uint64)
        _ck(lib().vx_dedup_pack(self.local.h, height, _p(self.heights), _p(self.gmap), G,
                                counts.ctypes.data_as(C.c_void_p), None, None))
        total = int(counts.sum())
        records = torch.empty(max(total, 1) * REC_WORDS, dtype=torch.int64, device=self.device)
        src = torch.empty(max(total, 1), dtype=torch.int32, device=self.device)
        if total:
            _ck(lib().vx_dedup_pack(self.local.h, height, _p(self.heights), _p(self.gmap), G,
                                    counts.ctypes.data_as(C.c_void_p), _p(records), _p(src)))
        return counts.astype(np.int64), records[: total * REC_WORDS], src[:total]

    def intern(self, records: torch.Tensor, leaf_round: bool) -> torch.Tensor:
        n = records.numel() // REC_WORDS
        ids = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        created = C.c_uint64(0)
        _sync(self.device)
        if n:
            _ck(lib().vx_interner_intern_records(self.shard.h, n, _p(records), self.rank, int(leaf_round), _p(ids),
                                                 C.byref(created)))
        self.created[1 if leaf_round else 0] += created.value
        return ids[:n]

    def scatter(self, src: torch.Tensor, ids: torch.Tensor):
        _sync(self.device)
        if src.numel():
            _ck(lib().vx_dedup_scatter(self.local.h, src.numel(), _p(src), _p(ids), _p(self.gmap)))

    def map_roots(self, roots: torch.Tensor) -> torch.Tensor:
        out = torch.zeros_like(roots)
        _sync(self.device)
        if roots.numel():
            _ck(lib().vx_dedup_map_roots(self.local.h, roots.numel(), _p(roots), _p(self.gmap), _p(out)))
        return out


def _split(t: torch.Tensor, counts, words: int):
    out, off = [], 0
    for c in counts:
        out.append(t[off * words:(off + int(c)) * words])
        off += int(c)
    return out


def global_dedup_local(locals_, roots_list, shard_budget: int, dtype: int, device=0):
    """G logical ranks in one process / on one device (tests, and the --gpus 1 form of config 5).
    locals_[r] = rank r's private interner, roots_list[r] = its chunk roots (numpy uint64).
    Returns (shards, global_roots_list, summary)."""
    G = len(locals_)
    dev = torch.device("cuda", device)
    ranks = [_Rank(r, locals_[r], VoxInterner.with_memory_budget(shard_budget, dtype, device), dev) for r in range(G)]
    max_h = max(r.max_height for r in ranks)
    sent = 0
    for h in range(max_h + 1):
        packed = [r.pack(h, G) for r in ranks]                       # (counts[G], records, src) per sender
        per_sender_parts = [_split(rec, cnt, REC_WORDS) for cnt, rec, _ in packed]
        answers = [[None] * G for _ in range(G)]                      # answers[sender][owner]
        for o in range(G):                                            # "all-to-all": owner o receives from every sender
            inbox = torch.cat([per_sender_parts[s][o] for s in range(G)]) if G > 1 else per_sender_parts[0][0]
            sent += inbox.numel() * 8
            ids = ranks[o].intern(inbox.contiguous(), leaf_round=(h == 0))
            off = 0
            for s in range(G):
                c = int(packed[s][0][o])
                answers[s][o] = ids[off:off + c]
                off += c
        for s in range(G):
            back = torch.cat(answers[s]) if G > 1 else answers[s][0]
            ranks[s].scatter(packed[s][2], back.contiguous())
    groots = []
    for r in range(G):
        rt = torch.from_numpy(np.ascontiguousarray(roots_list[r]).view(np.int64)).to(dev)
        groots.append(ranks[r].map_roots(rt).cpu().numpy().view(np.uint64))
    summary = {"G": G, "rounds": max_h + 1, "bytes_sent": sent,
               "branches": sum(r.created[0] for r in ranks), "leaves": sum(r.created[1] for r in ranks),
               "per_shard": [tuple(r.created) for r in ranks]}
    return [r.shard for r in ranks], groots, summary


_WORLDS: dict = {}


def world_for(device: int):
    """The process's vx_world (created once per device: ncclCommInitRank is a collective and not cheap)."""
    from .api import World
    if device not in _WORLDS:
        _WORLDS[device] = World.from_torch_distributed(device)
    return _WORLDS[device]


def global_dedup_dist(local: VoxInterner, roots: np.ndarray, shard_budget: int, dtype: int, device: int, world=None, shard=None):
    """One process per GPU: ONE C-ABI call, vx_world_global_dedup (pack kernel, NCCL send / recv of the records and of
    the ids, owner-side interning — voxelis_b200/csrc/vx_world.cuh).  torch.distributed is only used to hand the
    128-byte NCCL id to the other processes when no ``world`` is given.  Returns (shard, global_roots, summary)."""
    world = world or world_for(device)
    shard = shard or VoxInterner.with_memory_budget(shard_budget, dtype, device)   # a FRESH interner: this rank's global shard
    groots, summ = world.global_dedup(local, shard, roots)
    summary = {"G": world.n_ranks, "rounds": summ["rounds"], "bytes_sent": summ["bytes_sent"], "branches": summ["branches"],
               "leaves": summ["leaves"], "this_shard": (summ["this_shard_branches"], summ["this_shard_leaves"]),
               "local_nodes_all_ranks": summ["local_nodes_all_ranks"],
               "exchange": "vx_world_global_dedup: grouped ncclSend/ncclRecv issued by the library"}
    return shard, groots, summary


def global_dedup(local: VoxInterner, roots: np.ndarray, shard_budget: int, dtype: int, device: int, shard=None):
    """The merge for THIS process: the NCCL exchange when torch.distributed runs with more than one rank, else the
    single-rank form.  Returns (shard, global_roots, summary)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        shard, groots, summ = global_dedup_dist(local, roots, shard_budget, dtype, device, shard=shard)
        return shard, groots, summ
    shards, groots_l, summ = global_dedup_local([local], [roots], shard_budget, dtype, device)
    summ.setdefault("exchange", "single rank (no exchange)")
    return shards[0], groots_l[0], summ


def merged_pools(shards):
    """Downloads every shard and renumbers global ids into one index space (for the DAG checker):
    returns (children[n][8] uint64 with ids rewritten to plain BlockIds of the combined pool, values,
    remap(gid) -> combined id)."""
    dls = [s.download() for s in shards]
    offs = np.cumsum([0] + [d["n"] for d in dls])
    n = int(offs[-1])
    children = np.zeros((n, 8), np.uint64)
    values = np.zeros(n, np.int64)

    def remap_arr(g):
        g = g.astype(np.uint64)
        owner = ((g >> np.uint64(44)) & np.uint64(7)).astype(np.int64)
        idx = (g & np.uint64(0xFFFFFFFF)).astype(np.int64)
        hi = g & np.uint64(0xFFFF800000000000)                  # leaf | types | mask
        comb = (np.asarray(offs)[owner] + idx).astype(np.uint64)
        return np.where(g == 0, np.uint64(0), hi | comb)

    for o, d in enumerate(dls):
        children[offs[o]:offs[o + 1]] = remap_arr(d["children"].reshape(-1)).reshape(-1, 8)
        values[offs[o]:offs[o + 1]] = d["values"]
    return children, values, remap_arr
