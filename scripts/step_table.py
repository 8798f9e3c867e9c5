#!/usr/bin/env python
"""Per-launch table of one instrumented bench step (every GEMM / cast / attention span with its shape, CUDA-event time and useful
TFLOP/s), for one or more precisions:   python scripts/step_table.py [vidvrd|vidor] [videos] [precision ...]"""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from vidsgg_big_b200 import linalg  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "vidvrd"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
precs = sys.argv[3:] or ["tf32+bf16x2", "bf16"]
dev = torch.device("cuda", 0)
if os.environ.get("VSG_KC"):                          # key-block size of the tcgen05 attention kernel (32 / 64)
    from vidsgg_big_b200._cabi import lib
    lib().vsg_mha_tc16_set_kc(int(os.environ["VSG_KC"]))
seeds = [1000 + i for i in range(n)]
cfg, wl, props, graphs, feats = bench.make_videos(kind, seeds, dev, seeds[0])
for p in props:
    f = p.features; p.to(dev); p.features = f
for prec in precs:
    pipe = bench.Pipeline(kind, prec, dev)
    pipe.model.backend = "py"                      # op-by-op launches so that every GEMM can be bracketed
    if kind == "vidor":
        pipe.grd.backend = "py"
        if os.environ.get("VSG_NOFUSE"):               # depthwise conv as a separate launch instead of inside the point-wise GEMM
            pipe.grd.fuse_dwconv = False
    g2, _ = bench.gt_from_predictions(pipe, props, cfg, seeds, dev)
    for _ in range(3):
        pipe.step(props, g2, gather=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pipe.step(props, g2, gather=False); e1.record(); torch.cuda.synchronize()
    linalg._Profile.begin()
    torch.cuda._sleep(int(60e-3 * 1.9e9))          # head start for the host: the event brackets then hold kernel time only
    h = pipe.launch(props)
    if kind == "vidor":
        h["packed"].counts
        torch.cuda._sleep(int(30e-3 * 1.9e9))
    pipe.finish(h, g2, gather=False)
    linalg._Profile.end()
    print("==== %s %s %d videos: step %.2f ms" % (kind, prec, n, e0.elapsed_time(e1)))
    agg = {}
    for k, stage, meta, ms, tf in linalg._Profile.table():
        key = (k, stage, meta)
        a = agg.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] = tf
    tot = 0.0
    for (k, stage, meta), (cnt, ms, tf) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-9s %-9s %-28s x%-3d %8.3f ms  %7.1f TFLOP/s" % (k, stage, meta, cnt, ms, tf))
        tot += ms
    print("total in spans %.2f ms" % tot)
