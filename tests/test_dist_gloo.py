"""Host-side logic of the item-sharded path on CPU: shard bounds, user chunks and the candidate exchange
(gloo, world size 2).  The merge itself is a CUDA kernel and is covered by the GPU tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from polara_b200.dist import ItemShard, exchange_candidates


def test_item_shard_bounds_cover_everything():
    for world, n in ((2, 100_001), (4, 10), (8, 1_000_000), (3, 7)):
        shards = [ItemShard(r, world, n) for r in range(world)]
        assert shards[0].item_lo == 0 and shards[-1].item_hi == n
        for a, b in zip(shards[:-1], shards[1:]):
            assert a.item_hi == b.item_lo
        n_users = 1003
        chunk = shards[0].user_chunk(n_users)
        assert chunk * world >= n_users
        covered = sum(s.user_range(n_users)[1] - s.user_range(n_users)[0] for s in shards)
        assert covered == n_users


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, m_pad, k, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank r's list entry for (user u, slot j) encodes (r, u, j) so the exchanged layout can be checked
        u = torch.arange(m_pad).view(m_pad, 1, 1)
        j = torch.arange(k).view(1, k, 1)
        cands = torch.cat([(rank * 1_000_000 + u * 100 + j).expand(m_pad, k, 1),
                           (-(rank * 1_000_000 + u * 100 + j)).expand(m_pad, k, 1)], dim=2).to(torch.int32).contiguous()
        recv = exchange_candidates(cands, world)
        np.save(os.path.join(out_dir, "recv%d.npy" % rank), recv.numpy())
    finally:
        dist.destroy_process_group()


def test_exchange_candidates_gloo_world2(tmp_path):
    world, m_pad, k = 2, 8, 3
    port = _free_port()
    mp.spawn(_worker, args=(world, port, m_pad, k, str(tmp_path)), nprocs=world, join=True)
    chunk = m_pad // world
    for rank in range(world):
        recv = np.load(tmp_path / ("recv%d.npy" % rank))
        assert recv.shape == (world, chunk, k, 2)
        for src in range(world):
            for lu in range(chunk):
                for j in range(k):
                    u = rank * chunk + lu          # rank owns users [rank*chunk, (rank+1)*chunk)
                    assert recv[src, lu, j, 0] == src * 1_000_000 + u * 100 + j
                    assert recv[src, lu, j, 1] == -(src * 1_000_000 + u * 100 + j)
