"""BASELINE.json configs[4] on N GPUs: one video, 256 tracklets, 10k frames; the 256 x 256 pair matrix is split by row blocks over
the ranks (track table replicated), blocks all-gathered over NCCL.  Launch with torchrun (or plain python for N = 1):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/bench_geometry_stress_multi.py

Prints one JSON line on rank 0: per-rank kernel time (max over ranks), time incl. the all_gather, algorithmic GB/s, and the check
that the gathered matrix equals the single-GPU result bit for bit."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from vidsgg_big_b200 import synth, geometry, shard

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n, vlen = int(os.environ.get("N", 256)), 10000
P = synth.make_proposal(4242, n, vlen, 8, 36, min_len=2000, max_len=10000, with_features=False).to(dev)
boxes, dura = P.bboxes_list, P.traj_durations
d = dura.cpu().numpy()
s = np.maximum(d[:, None, 0], d[None, :, 0]); e = np.minimum(d[:, None, 1], d[None, :, 1])
ov = np.clip(e - s + 1, 0, None)
sumL = int(P.lengths.sum())
alg = 32 * int(ov.sum()) + 16 * 2 * sumL + n * n * 21
T = geometry.TrackTable.from_lists(boxes, dura)
full_v, full_sp, full_m, _, _ = geometry.traj_viou_batched(T, T)
for _ in range(3):
    shard.traj_viou_row_sharded(boxes, dura)
r0, r1 = shard.row_block(n, rank, world)
A = geometry.TrackTable.from_lists(boxes[r0:r1], dura[r0:r1])
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
k_ms, t_ms = [], []
for _ in range(10):
    flush.zero_()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    viou, spans, mask, _, _ = geometry.traj_viou_batched(A, T)
    e1.record()
    out = (shard.gather_row_blocks(viou.reshape(r1 - r0, n), n), shard.gather_row_blocks(spans.reshape(r1 - r0, n, 2), n),
           shard.gather_row_blocks(mask.reshape(r1 - r0, n), n))
    e2.record()
    torch.cuda.synchronize()
    k_ms.append(e0.elapsed_time(e1)); t_ms.append(e0.elapsed_time(e2))
t = torch.tensor([float(np.median(k_ms)), float(np.median(t_ms))], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
same = bool(torch.equal(out[0].reshape(-1), full_v) and torch.equal(out[1].reshape(-1, 2), full_sp) and torch.equal(out[2].reshape(-1), full_m))
if rank == 0:
    peak = 6555.5
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    k, tot = float(t[0]), float(t[1])
    print(json.dumps({"config": "stress: 1 video, %d tracklets, 10k frames, pair matrix row-sharded" % n, "n_gpus": world,
                      "frame_pairs": int(ov.sum()), "algorithmic_bytes": alg, "kernel_ms_max_over_ranks": k, "with_allgather_ms": tot,
                      "algorithmic_GBps_aggregate": alg / k / 1e6, "frac_of_hbm_peak_per_gpu": alg / k / 1e6 / peak / world,
                      "gathered_equals_single_gpu": same}))
if world > 1:
    dist.destroy_process_group()
