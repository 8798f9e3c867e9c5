#!/usr/bin/env python
"""Bare pinned host -> device copy rate per GPU with every rank copying at once (torchrun, one rank per GPU), (a) as the process starts and
(b) after binding the process to the CPUs of its GPU's NUMA node and allocating the pinned buffer from there.
   python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/h2d_numa_probe.py [GB]"""
import json
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from vidsgg_big_b200 import shard  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gb * 1e9 / 4)
dst = torch.empty(n, dtype=torch.float32, device=dev)


def rate(host):
    for _ in range(2):
        dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        dst.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return 4 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9


def gather(x):
    t = torch.tensor([x], device=dev)
    out = [t.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(out, t)
    return [round(float(o.item()), 1) for o in out]


host = torch.empty(n, dtype=torch.float32, pin_memory=True)
host.fill_(1.0)
before = gather(rate(host))
del host
info = shard.bind_to_gpu_numa_node(local)
host = torch.empty(n, dtype=torch.float32, pin_memory=True)
host.fill_(1.0)                                     # first touch from the bound CPUs
after = gather(rate(host))
infos = [None] * world
if world > 1:
    dist.all_gather_object(infos, info)
else:
    infos = [info]
if rank == 0:
    print(json.dumps({"world": world, "gb_per_copy": gb, "cpu_count": os.cpu_count(), "h2d_gbs_before": before, "total_before": round(sum(before), 1),
                      "h2d_gbs_after_numa_binding": after, "total_after": round(sum(after), 1), "binding": infos}))
if world > 1:
    dist.destroy_process_group()
