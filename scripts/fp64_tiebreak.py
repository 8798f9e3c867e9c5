#!/usr/bin/env python
"""Who is right on the near-ties?  The 200-video benchmark batch through the two fp32-class GPU modes and the fp32 CPU oracle; every video on
which a GPU mode and the fp32 oracle disagree is re-run through the SAME oracle code in float64 (weights, features and boxes cast to
double), and the three fp32-class results are compared with that float64 answer.
   python scripts/fp64_tiebreak.py [videos]        -> one JSON line"""
import copy
import json
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from oracle import bigc as ob  # noqa: E402  (the checker)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda", 0)
seeds = [1000 + i for i in range(n)]
cfg, wl, props, _, feats = bench.make_videos("vidvrd", seeds, dev, seeds[0], with_gt=False)
prep = bench.cpu_prepare("vidvrd", seeds, seeds[0], feats_from=(feats.cpu(), None))
for p in props:
    f = p.features; p.to(dev); p.features = f
key = lambda r: None if r is None else frozenset(map(tuple, torch.cat([r[0].cpu(), r[2].cpu()], 1).tolist()))
out = {}
with torch.no_grad():
    for prec in ("tf32+bf16x2", "fp16x3"):
        pipe = bench.Pipeline("vidvrd", prec, dev)
        out[prec] = [key(r) for r in pipe.model(props, topk=wl["topk"])]
    out["oracle_fp32"] = [key(r) for r in prep["trips"]]
    differ = [i for i in range(n) if len({out[k][i] for k in out}) > 1]
    st64 = {k: (v.double() if v.is_floating_point() else v) for k, v in prep["st"].items()}
    rep = []
    for i in differ:
        q = copy.copy(prep["props"][i])
        q.features, q.bboxes = q.features.double(), q.bboxes.double()
        k64 = key(ob.forward(st64, prep["cfg"], [q], wl["topk"])[0])
        rep.append({"video": i, "agree_with_fp64_oracle": {k: out[k][i] == k64 for k in out},
                    "triplets": {k: (None if out[k][i] is None else len(out[k][i])) for k in out}})
print(json.dumps({"videos": n, "videos_where_the_three_fp32_class_results_differ": differ, "fp64_verdict": rep}))
