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


def flops_bigc(sum_len: float, n: int, t_max: int, dim_feat: int, E: int = 512, Q: int = 192, n_enc: int = 6, n_dec: int = 4,
               head: float = 2.0 * 192 * (2136 * 512 + 512 * 51)) -> float:
    """Useful flops of the BIG-C forward of one video (SURVEY.md 8d, K4/K5 formula with R = sum of track lengths)."""
    tc = (t_max - 1) // 2 + 1
    f = 2.0 * sum_len * (8 * E + E * E) + 2.0 * sum_len * (dim_feat * E + E * E)
    f += 2.0 * n * tc * 6 * E * E + 2.0 * n * 5 * E * E
    f += n_enc * (2.0 * n * 6 * E * E + 4.0 * n * n * E)
    f += n_dec * (2.0 * Q * 11 * E * E + 2.0 * n * E * E + 4.0 * Q * Q * E + 6.0 * Q * n * E)
    return f + head


def flops_grd(T: int, nq: int, H: int = 128, B: int = 10) -> float:
    """Useful flops of the grounding forward of one video (SURVEY.md 8d, K6 formula, re-associated context-query product)."""
    enc = lambda rows, seqs, L, k: rows * (4 * (2 * k * H + 2 * H * H) + 10 * H * H) + seqs * 4.0 * L * L * H
    f = 2.0 * T * 1024 * H + 2.0 * nq * 3 * 300 * H + 2.0 * nq * 2 * H
    f += enc(T, 1, T, 7) + enc(3 * nq, nq, 3, 3) + enc(nq * T, nq, T, 7)
    f += 2.0 * T * H * H + 2.0 * nq * T * H * 3 + 2.0 * nq * 3 * T * H + 2.0 * nq * T * 3 * H + 2.0 * nq * T * 3 * H + 2.0 * nq * T * 4 * H * H
    for out in (B, B, 2 * B):
        f += nq * T * (4 * (2 * 3 * H + 2 * H * H) + 2 * 3 * H + 2 * H * out)
    return f


def video_cost_flops(lengths, video_len: int, dim_feat: int, grounding: bool = True, nq_estimate: int = 450) -> float:
    """Scheduling cost of one video in useful flops: BIG-C over the RAGGED rows (the per-frame stage never stretches tracks here,
    so it scales with sum L, not n * Tmax) plus the grounding stage on ~``nq_estimate`` queries x ceil(video_len / 8) clips."""
    n = int(len(lengths))
    if n == 0:
        return 0.0
    c = flops_bigc(float(sum(int(x) for x in lengths)), n, int(max(lengths)), dim_feat)
    if grounding:
        c += flops_grd((int(video_len) + 7) // 8, nq_estimate)
    return c


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


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """One process per GPU: pin this process to the CPUs that are local to its GPU (``/sys/bus/pci/devices/<bus id>/local_cpulist``), so
    that pinned host buffers allocated afterwards are first-touched on the GPU's own NUMA node and the H2D copies do not cross the socket
    interconnect.  No reference counterpart (the reference copies pageable tensors with ``.to(device)``, tools/eval_vidor.py:96).  Best
    effort: returns what it did; leaves the affinity alone when sysfs has no topology for the device (containers) or there is one node."""
    import os
    info = {"device": int(device_index), "bound": False}
    try:
        p = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        info["pci"] = bus
        base = "/sys/bus/pci/devices/" + bus
        node = int(open(base + "/numa_node").read().strip())
        cpus = _parse_cpulist(open(base + "/local_cpulist").read())
        info["numa_node"], info["local_cpus"] = node, len(cpus)
        allowed = sorted(os.sched_getaffinity(0))
        info["allowed_cpus"] = len(allowed)
        target = sorted(set(cpus) & set(allowed))
        if node < 0 or not target or len(target) == len(allowed):
            return info
        os.sched_setaffinity(0, target)
        info["bound"] = True
    except Exception as e:          # no sysfs entry, no permission, ...: keep running unbound
        info["error"] = "%s: %s" % (type(e).__name__, e)
    return info
