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
