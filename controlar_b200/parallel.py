"""Multi-GPU host logic of the conditional-decoding path: images are sharded over ranks, weights are
replicated, nothing is exchanged inside the decode loop, and the finished int32 token grids are gathered ONCE.

The reference has no collective here — its DDP sampler writes PNGs per rank and barriers
(autoregressive/sample/sample_c2i_ddp.py:146-156); the rank seed mirrors sample_c2i_ddp.py:47.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def rank_seed(global_seed: int, world_size: int, rank: int) -> int:
    """seed = global_seed * world_size + rank (sample_c2i_ddp.py:47)."""
    return int(global_seed) * int(world_size) + int(rank)


def shard_bounds(n_images: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous image range [lo, hi) of `rank`; the first n_images % world_size ranks take one extra image."""
    base, extra = divmod(int(n_images), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_token_grids(tokens: torch.Tensor, n_images: Optional[int] = None) -> torch.Tensor:
    """All-gather of this rank's int32 [B_loc, N] token grid into [n_images, N] (rank-major) — one collective per batch,
    enqueued on the current stream (NCCL on GPUs; gloo in the CPU tests).  Single-process: returns `tokens`.

    `n_images` = total number of images sharded with `shard_bounds` (default world * B_loc, i.e. equal shards).  When it is not
    a multiple of the world size the shards differ by one row: every rank pads its grid to ceil(n_images / world) rows (the
    collective needs identical shapes on all ranks) and the pad rows are cut out again after the gather."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tokens
    if tokens.dtype != torch.int32 or tokens.dim() != 2:
        raise ValueError("gather_token_grids expects an int32 [B_loc, N] tensor")
    world, rank = dist.get_world_size(), dist.get_rank()
    if n_images is None:
        n_images = world * tokens.shape[0]
    lo, hi = shard_bounds(n_images, world, rank)
    if hi - lo != tokens.shape[0]:
        raise ValueError(f"rank {rank} holds {tokens.shape[0]} grids but shard_bounds({n_images}, {world}, {rank}) is [{lo}, {hi})")
    rows = -(-int(n_images) // world)
    send = tokens.contiguous()
    if send.shape[0] < rows:
        send = torch.cat([send, send.new_zeros((rows - send.shape[0], send.shape[1]))])
    out = torch.empty((world * rows, tokens.shape[1]), dtype=tokens.dtype, device=tokens.device)
    dist.all_gather_into_tensor(out, send)
    if world * rows == n_images:
        return out
    keep = [out[r * rows: r * rows + (shard_bounds(n_images, world, r)[1] - shard_bounds(n_images, world, r)[0])] for r in range(world)]
    return torch.cat(keep)
