"""Multi-GPU plumbing of the hot path: videos are independent, so they are sharded across ranks with
NO data-path collective; the only exchange is one all_gather of fixed-width per-video evaluation records at the end.

The reference has no counterpart (inference is 1 process x 1 GPU x batch 1, tools/eval_vidor.py:74-117; its only
multi-GPU code is the training-time nn.DataParallel of utils/DataParallel.py).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def video_cost(proposal, dim_feat: int) -> float:
    """Work estimate of one video: n * Tmax * D (what the per-frame stage scales with)."""
    if proposal.num_proposals == 0:
        return 0.0
    return float(proposal.num_proposals) * float(int(proposal.lengths.max())) * float(dim_feat)


def assign_lpt(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first greedy: VidOR video costs spread ~100x, so round-robin is not enough."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += costs[i]
    for s in shards:
        s.sort()
    return shards


def gather_records(records: torch.Tensor, group=None) -> torch.Tensor:
    """all_gather of a ragged [n_local, W] float64 record table: sizes first, then one padded payload."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return records
    world = dist.get_world_size(group)
    dev = records.device
    n = torch.tensor([records.shape[0]], dtype=torch.long, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    width = records.shape[1]
    pad = torch.zeros(max(max(sizes), 1), width, dtype=records.dtype, device=dev)
    pad[:records.shape[0]] = records
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)


def row_block(n_rows: int, rank: int, world: int):
    """[r0, r1) of the contiguous row block of ``rank`` when ``n_rows`` rows are split over ``world`` ranks (sizes differ by <= 1)."""
    base, rem = divmod(n_rows, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def gather_row_blocks(block: torch.Tensor, n_rows: int, group=None) -> torch.Tensor:
    """all_gather of the row blocks of a [n_rows, ...] matrix (each rank holds rows ``row_block(n_rows, rank, world)``): the one
    exchange of the single-video stress configuration (the pair matrix of one huge video is split by row blocks,
    the ~25 MB track table is replicated).  Blocks are padded to the largest block for the collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return block
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [row_block(n_rows, r, world) for r in range(world)]
    assert block.shape[0] == sizes[rank][1] - sizes[rank][0]
    cap = max(b - a for a, b in sizes)
    pad = torch.zeros((max(cap, 1),) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
    pad[:block.shape[0]] = block
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)


def traj_viou_row_sharded(boxes_list, dura: torch.Tensor, group=None):
    """Trajectory vIoU matrix + spans + mask of ONE video with its n x n pair matrix split by row blocks over the ranks:
    every rank computes rows [r0, r1) against all tracks (``geometry.traj_viou_batched`` on a sub-table), then the blocks are
    all-gathered so that every rank ends with the full matrices.  Same results as the single-GPU call."""
    from . import geometry
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    n = len(boxes_list)
    r0, r1 = row_block(n, rank, world)
    B = geometry.TrackTable.from_lists(boxes_list, dura)
    A = geometry.TrackTable.from_lists(boxes_list[r0:r1], dura[r0:r1])
    viou, spans, mask, _, _ = geometry.traj_viou_batched(A, B)
    viou = viou.reshape(r1 - r0, n)
    spans = spans.reshape(r1 - r0, n, 2)
    mask = mask.reshape(r1 - r0, n)
    return (gather_row_blocks(viou, n, group), gather_row_blocks(spans, n, group),
            gather_row_blocks(mask.to(torch.uint8), n, group).bool())
