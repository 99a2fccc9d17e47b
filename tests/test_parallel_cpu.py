"""World-size-2 gloo test (CPU) of the multi-GPU host logic: image sharding, rank seeds and the single all-gather of the
finished token grids (controlar_b200/parallel.py).  The decode loop itself has no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from controlar_b200.parallel import gather_token_grids, rank_seed, shard_bounds


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(8, world, rank)
        g = torch.Generator().manual_seed(rank_seed(3, world, rank))
        local = torch.randint(0, 16384, (hi - lo, 64), generator=g, dtype=torch.int32)
        allt = gather_token_grids(local)
        ret[rank] = (lo, hi, local, allt)
    finally:
        dist.destroy_process_group()


def _worker_uneven(rank: int, world: int, port: int, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(7, world, rank)                      # 7 images on 2 ranks: 4 + 3
        local = (torch.arange(lo, hi, dtype=torch.int32)[:, None] * 100 + torch.arange(5, dtype=torch.int32)[None, :])
        ret[rank] = gather_token_grids(local, n_images=7)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_uneven_shards():
    """n_images % world_size != 0 must neither hang nor mis-size the gather."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_uneven, args=(world, _free_port(), ret), nprocs=world, join=True)
    want = torch.arange(7, dtype=torch.int32)[:, None] * 100 + torch.arange(5, dtype=torch.int32)[None, :]
    assert torch.equal(ret[0], want) and torch.equal(ret[1], want)


def test_shard_bounds_cover_everything():
    for n in (1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_rank_seed_matches_reference_rule():
    assert [rank_seed(0, 8, r) for r in range(8)] == list(range(8))
    assert rank_seed(2, 4, 3) == 11          # sample_c2i_ddp.py:47: seed = global_seed * world_size + rank


def test_single_process_gather_is_identity():
    t = torch.arange(12, dtype=torch.int32).reshape(3, 4)
    assert gather_token_grids(t) is t


def test_gloo_world2_gather_token_grids():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    lo0, hi0, loc0, all0 = ret[0]
    lo1, hi1, loc1, all1 = ret[1]
    assert (lo0, hi0, lo1, hi1) == (0, 4, 4, 8)
    assert all0.shape == (8, 64) and all0.dtype == torch.int32
    assert torch.equal(all0, all1)                                   # every rank ends with the same global grid
    assert torch.equal(all0, torch.cat([loc0, loc1]))                # rank-major order
    assert not torch.equal(loc0, loc1)                               # different rank seeds
