"""Multi-GPU data parallelism of the inference path: one process per GPU, contiguous image shards,
no collective on the data path, ONE gather of fixed-shape results at the end.

Mirrors `valid_fn` + `dist.all_gather_object` in the reference (main.py:260-302, :295-301), with
tensors of fixed shape instead of pickled Python dicts.  Results for a sharded batch equal the
reference run on each shard separately (the row-rank positional-encoding rule makes outputs depend
on the row's position inside its local batch)."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n items for `rank`; the first n % world ranks get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_predictions(local: Dict[str, torch.Tensor], n_total: int, group=None) -> Dict[str, torch.Tensor]:
    """All-gather per-rank result tensors (first dim = local rows) into global order.
    Shards may be ragged: every rank pads to the largest shard, then the padding is dropped."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = [shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world)]
    mx = max(max(sizes), 1)      # all_gather of zero-element tensors is backend-dependent: keep one padding row
    out = {}
    for k, t in local.items():
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        buf = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(buf, pad, group=group)
        out[k] = torch.cat([b[:s] for b, s in zip(buf, sizes)], dim=0)
    return out


def predict_sharded(engine, images: torch.Tensor, group=None) -> Dict[str, torch.Tensor]:
    """Every rank passes the same global batch (host tensor); each runs its contiguous shard on its own
    GPU and all ranks receive the full result.

    Every rank enters every collective: a rank whose shard is empty (fewer images than ranks) skips the engine
    and contributes zero-row tensors (`engine.empty_result()`), and a shard that would exceed the engine's
    capacity is rejected on ALL ranks before any collective is issued (shard sizes are a function of
    (n, world) only), so the job fails cleanly instead of hanging in all_gather until the NCCL timeout."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = int(images.shape[0])
    largest = shard_bounds(n, world, 0)[1] - shard_bounds(n, world, 0)[0]
    if largest > engine.max_batch:
        raise ValueError(f"{n} images over {world} ranks gives shards of {largest} rows, above the engine's max_batch "
                         f"{engine.max_batch}; the caller must chunk (chunk size is part of the semantics: SURVEY.md F3)")
    lo, hi = shard_bounds(n, world, rank)
    local = engine.predict(images[lo:hi].to(engine.device)) if hi > lo else engine.empty_result()
    return gather_predictions(local, n, group)
