"""World-size-2 gloo test (CPU) of the multi-GPU host logic: image sharding, rank seeds and the single all-gather of the
finished token grids (controlar_b200/parallel.py).  The decode loop itself has no collective."""
import os
import socket

import pytest
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


# ---- training step under torch DDP (the reference wraps the model in DistributedDataParallel, train_c2i_canny.py:166) ----
# The library's backward reaches the parameters through a torch.autograd.Function, so DDP's gradient hooks fire as usual and the
# bucketed all-reduce is torch's own: nothing to build, but it has to be shown to work.  The library calls are stubbed (no GPU
# here): each rank's "backward" returns gradients equal to rank + 1, so every .grad must end at the mean 1.5 on both ranks.
def _worker_ddp(rank: int, world: int, port: int, ret, find_unused: bool):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from torch.nn.parallel import DistributedDataParallel as DDP
        from controlar_b200 import engine
        from controlar_b200.autoregressive.models import gpt_t2i

        class Stub:
            grad_params = staticmethod(engine.ARTrainHandle.grad_params)

            def __init__(self, module, B, n):
                self.key = tuple(p.data_ptr() for p in module.parameters())
                self.max_batch, self.max_img_tokens, self.generation = B, n, 0

            def forward(self, idx, cond, feat, drop, mask, targets, valid):
                self.generation += 1
                return torch.zeros(idx.shape[0], idx.shape[1] + 1, 64), torch.tensor(float(rank))

            def backward(self, module, loss_grad=None, want_feat_grad=True):
                return {k: torch.full_like(p, float(rank + 1)) for k, p in self.grad_params(module)}, None

            def close(self):
                pass
        engine.ARTrainHandle = Stub
        torch.manual_seed(0)
        m = gpt_t2i.Transformer(gpt_t2i.ModelArgs(dim=128, n_layer=3, n_head=2, vocab_size=64, cls_token_num=1, block_size=16, num_classes=10,
                                                  model_type="c2i", adapter_size="small", condition_type="canny", token_dropout_p=0.0,
                                                  resid_dropout_p=0.0, ffn_dropout_p=0.0, class_dropout_prob=0.1)).train()
        # this library differentiates down to the control tokens: the control encoder (and the unused condition_embeddings table)
        # stay frozen, otherwise DDP waits for gradients that never come
        if not find_unused:
            trained = {id(p) for _, p in engine.ARTrainHandle.grad_params(m)}
            for p in m.parameters():
                p.requires_grad_(id(p) in trained)
        m.adapter.forward = lambda x: torch.zeros(x.shape[0], 16, 384)
        # find_unused: the reference's own call, DDP(model, find_unused_parameters=True) (train_c2i_canny.py:173), nothing frozen
        ddp = DDP(m, find_unused_parameters=find_unused)
        with torch.enable_grad():
            for _ in range(2):                                   # two iterations: the reducer must be re-armed after the first
                for p in m.parameters():
                    p.grad = None
                z = torch.randint(0, 64, (2, 16))
                _, loss = ddp(idx=z[:, :-1], cond_idx=torch.tensor([1, 2]), targets=z, condition=torch.zeros(2, 3, 64, 64))
                loss.backward()
        ret[rank] = {k: (float(p.grad.min()), float(p.grad.max())) for k, p in engine.ARTrainHandle.grad_params(m)}
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("find_unused", [False, True])
def test_ddp_over_the_library_backward(find_unused):
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker_ddp, args=(world, _free_port(), ret, find_unused), nprocs=world, join=True)
    for rank in range(world):
        assert len(ret[rank]) > 20
        for k, (lo, hi) in ret[rank].items():
            assert lo == hi == 1.5, (rank, k, lo, hi)
