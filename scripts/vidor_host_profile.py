#!/usr/bin/env python
"""Where the wall time of one VidOR-shaped step goes on the HOST side: phases of Pipeline.launch / finish bracketed by device synchronisations
(so GPU time and host time separate), plus a cProfile of the un-synchronised step.   python scripts/vidor_host_profile.py [videos]"""
import cProfile
import os
import pstats
import sys
import time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from vidsgg_big_b200 import evalapi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda", 0)
seeds = [1000 + i for i in range(n)]
cfg, wl, props, graphs, feats = bench.make_videos("vidor", seeds, dev, seeds[0])
for p in props:
    f = p.features; p.to(dev); p.features = f
pipe = bench.Pipeline("vidor", "tf32+bf16x2", dev)
g2, _ = bench.gt_from_predictions(pipe, props, cfg, seeds, dev)
for _ in range(3):
    pipe.step(props, g2, gather=False)
torch.cuda.synchronize()
sync = torch.cuda.synchronize
T = {}


def mark(name, t0):
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    return time.perf_counter()


reps = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    pipe.step(props, g2, gather=False)
e1.record(); sync()
wall = e0.elapsed_time(e1) / reps
for _ in range(reps):
    t = time.perf_counter()
    h = pipe.launch(props); t = mark("launch: host (geometry + vsg_bigc_forward enqueue)", t)
    sync(); t = mark("launch: device wait (classification kernels)", t)
    tt, packed = h["tt"], h["packed"]
    q, s3, sp, _, off = packed.compact(); sync(); t = mark("finish: packed.compact (counts D2H + index build + 4 gathers)", t)
    rows = [i for i in range(len(props)) if off[i + 1] > off[i]]
    datas = [(q[off[i]:off[i + 1]], sp[off[i]:off[i + 1]], props[i].video_len) for i in rows]
    t = mark("finish: per-video slices", t)
    pooled, probs, mask = pipe.grd.forward_packed([props[i].i3d for i in rows], datas, **synth.GROUNDING_INFERENCE); t = mark("finish: grounding host (tables + vsg_grd_forward enqueue)", t)
    sync(); t = mark("finish: grounding device wait", t)
    PR = evalapi.PackedRelations.from_grounded(tt, packed, pooled, probs, mask, [p.video_len for p in props]); sync(); t = mark("finish: from_grounded", t)
    GT = pipe.pack_gt(g2)
    rec = evalapi.evaluate_packed(PR, GT, want_records=True); sync(); t = mark("finish: evaluate_packed (kernels + D2H + host records)", t)
print("one un-instrumented step: %.2f ms wall (CUDA events)" % wall)
tot = 0.0
for k, v in T.items():
    print("  %-70s %8.2f ms" % (k, v / reps)); tot += v / reps
print("  %-70s %8.2f ms" % ("sum of the synchronised phases", tot))
pr = cProfile.Profile()
pr.enable()
for _ in range(reps):
    pipe.step(props, g2, gather=False)
sync()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative")
import io
buf = io.StringIO(); st.stream = buf; st.print_stats(28)
print("\n".join(l[:170] for l in buf.getvalue().splitlines()[:48]))

# ---- host time of the pieces of Pipeline.launch (device idle at the start of each repetition) ----
from vidsgg_big_b200 import geometry  # noqa: E402
tt, pk = pipe._packs[id(props)]
H = {}
for _ in range(reps):
    sync()
    t = time.perf_counter()
    geometry.traj_viou_batched(tt, tt); dt = (time.perf_counter() - t) * 1e3; H["traj_viou_batched (host)"] = H.get("traj_viou_batched (host)", 0) + dt
    sync()
    t = time.perf_counter()
    packed = pipe.model.forward_packed(props, topk=wl["topk"], packed_videos=pk, sync=False)
    dt = (time.perf_counter() - t) * 1e3; H["forward_packed (host, enqueue only)"] = H.get("forward_packed (host, enqueue only)", 0) + dt
    sync()
    t = time.perf_counter()
    ws = torch.empty(20_000_000_000, dtype=torch.uint8, device=dev); dt = (time.perf_counter() - t) * 1e3
    H["torch.empty(20 GB) (host)"] = H.get("torch.empty(20 GB) (host)", 0) + dt
    del ws
print("host time of the launch pieces (ms):", {k: round(v / reps, 3) for k, v in H.items()})
print("allocator:", {k: torch.cuda.memory_stats()[k] for k in ("num_alloc_retries", "num_device_alloc", "num_device_free", "reserved_bytes.all.peak")})
