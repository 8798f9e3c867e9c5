"""Wall-clock breakdown of one bench step with a synchronize after every stage (debug aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from vidsgg_big_b200 import evalapi, geometry, shard

dev = torch.device("cuda", 0)
pipe = bench.Pipeline("vidvrd", "3xtf32", dev)
cfg, wl, props, graphs, feats = bench.make_videos("vidvrd", 200, 1000, dev)
for g in graphs: g.to(dev)
for p in props:
    f = p.features; p.to(dev); p.features = f
for _ in range(3): pipe.step(props, graphs)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
acc = {}
for it in range(5):
    t = T(); tt = pipe._tt
    viou = geometry.traj_viou_batched(tt, tt); t1 = T(); acc["geometry"] = acc.get("geometry", 0) + t1 - t
    pk = pipe._pk
    logits, so, _ = pipe.model._encode2decode(pk); t2 = T(); acc["encode2decode"] = acc.get("encode2decode", 0) + t2 - t1
    packed = pipe.model._construct_triplets(pk, logits, so, 10, packed=True); t3 = T(); acc["triplets+counts D2H"] = acc.get("triplets+counts D2H", 0) + t3 - t2
    PR = evalapi.PackedRelations.from_packed_triplets(tt, packed); t4 = T(); acc["pack relations"] = acc.get("pack relations", 0) + t4 - t3
    GT = pipe.pack_gt(graphs)
    m = evalapi.match_relations(PR, GT, 0.5); t5 = T(); acc["match kernels"] = acc.get("match kernels", 0) + t5 - t4
    rec = evalapi.evaluate_packed(PR, GT, want_records=True); t6 = T(); acc["evaluate_packed (match+host)"] = acc.get("evaluate_packed (match+host)", 0) + t6 - t5
    allrec = shard.gather_records(torch.from_numpy(rec).to(dev)).cpu().numpy(); evalapi.metrics_from_records(allrec); t7 = T()
    acc["records->metrics"] = acc.get("records->metrics", 0) + t7 - t6
for k, v in acc.items(): print("%-32s %7.2f ms" % (k, 1e3 * v / 5))
