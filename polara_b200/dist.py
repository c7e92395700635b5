"""Item-factor sharding across the GPUs of one box (one process per GPU).

Every rank owns a contiguous slice of the item factors, scores ALL users against its
slice with the fused kernel and emits one sorted candidate list per user; the only
data-path collective is one all-to-all by user range (each rank receives the lists of
"its" users from every peer) followed by a local k-way merge.
``torch.distributed`` is plumbing only.
"""
from __future__ import annotations

import torch


class ItemShard:
    """Contiguous item range of this rank plus the user range it owns after the exchange."""

    def __init__(self, rank, world, n_items):
        self.rank, self.world, self.n_items = int(rank), int(world), int(n_items)
        per = (self.n_items + self.world - 1) // self.world
        self.item_lo = min(self.n_items, self.rank * per)
        self.item_hi = min(self.n_items, (self.rank + 1) * per)

    def user_chunk(self, n_users):
        """users are split into `world` equal chunks (the last ones may be short/padded)."""
        return (n_users + self.world - 1) // self.world

    def user_range(self, n_users):
        c = self.user_chunk(n_users)
        return min(n_users, self.rank * c), min(n_users, (self.rank + 1) * c)


def exchange_candidates(cands, world, group=None):
    """``cands`` [m_pad, k, 2] int32 view of {score,id} lists for ALL users (m_pad divisible by
    world) -> [world, m_pad/world, k, 2]: for this rank's user chunk, the lists of every peer."""
    import torch.distributed as dist
    m_pad = cands.shape[0]
    assert m_pad % world == 0
    out = torch.empty_like(cands)
    try:
        dist.all_to_all_single(out, cands, group=group)
    except RuntimeError:
        # backends without all-to-all (CPU tests): same result via all-gather
        gathered = [torch.empty_like(cands) for _ in range(world)]
        dist.all_gather(gathered, cands, group=group)
        rank = dist.get_rank(group)
        chunk = m_pad // world
        out = torch.stack([g[rank * chunk:(rank + 1) * chunk] for g in gathered]).reshape(cands.shape)
    return out.view(world, m_pad // world, *cands.shape[1:])


def sharded_topk(eng, e, v_dev, rank_r, topk, seen, shard: ItemShard, n_users):
    """Fused scoring on this rank's item slice + exchange + merge.  Returns int64 ids
    [user_chunk x topk] of the users this rank owns (global item ids)."""
    v_slice = v_dev[shard.item_lo:shard.item_hi]
    chunk = shard.user_chunk(n_users)
    m_pad = chunk * shard.world
    cands = eng.score_topk_cands(e, v_slice, rank_r, topk, seen=seen, item_offset=shard.item_lo, m=n_users,
                                 m_alloc=m_pad)
    recv = exchange_candidates(cands, shard.world)
    return eng.merge_cands(recv, shard.world, chunk, topk)


def make_step(eng, p_dev, v_dev, rank_r, topk, shard=None, filter_seen=True):
    """One device-resident pass of the hot path: SpMM + fused scoring (+ exchange/merge)."""
    seen = (p_dev.indptr, p_dev.indices) if filter_seen else None
    n_users = p_dev.shape[0]

    def step():
        e = eng.spmm(p_dev, v_dev, ell=v_dev.shape[1])
        if shard is None:
            return eng.score_topk(e, v_dev, rank_r, topk, seen=seen)
        return sharded_topk(eng, e, v_dev, rank_r, topk, seen, shard, n_users)
    return step


def time_score_kernel(eng, p_dev, v_dev, rank_r, topk, shard=None, reps=3):
    """Average duration (ms) of the fused scoring kernel alone, read from the CUDA events the
    library records around that launch on the context stream."""
    e = eng.spmm(p_dev, v_dev, ell=v_dev.shape[1])
    seen = (p_dev.indptr, p_dev.indices)
    v_use = v_dev if shard is None else v_dev[shard.item_lo:shard.item_hi]
    off = 0 if shard is None else shard.item_lo
    total = 0.0
    for i in range(reps + 1):
        eng.score_topk(e, v_use, rank_r, topk, seen=seen, item_offset=off)
        ms = eng.last_score_kernel_ms()
        if i > 0:
            total += ms
    return total / reps


def make_e2e(model, shard=None):
    """The user-facing call: host CSR in, host recommendations out."""
    if shard is None:
        return model.get_recommendations
    model.shard = shard
    return model.get_recommendations
