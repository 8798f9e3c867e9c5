"""Base-C training label assignment (tools/train_vidor.py:80-170, README.md:169 "around 1.5 hours" for 7000 VidOR-train videos)
on VidOR-train-shaped synthetic tracklets: batched GPU op vs the oracle's restatement of the reference loops on a sample."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vidsgg_big_b200 import synth, geometry
from oracle import geometry as og

V = int(os.environ.get("V", 1000))
dev = "cuda:0"
props, graphs = [], []
t0 = time.perf_counter()
for i in range(V):
    rng = np.random.default_rng(5000 + i)
    vlen, n = synth.vidor_video_shape(rng)
    P = synth.make_proposal(5000 + i, n, vlen, 8, 81, min_len=15, max_len=600, with_features=False)
    props.append(P); graphs.append(synth.make_gt_graph(5000 + i, P, 51, n_traj=(3, 10), n_rel=(5, 60), jitter_px=1.5))
gen_s = time.perf_counter() - t0
# CPU sample: the reference's loops (per-pair vIoU with .tolist(), then GT-relation x pair double loop)
ns = int(os.environ.get("CPU_SAMPLE", 3))
t0 = time.perf_counter()
for P, G in zip(props[:ns], graphs[:ns]):
    viou, _, _ = og.traj_viou_matrix(P.bboxes_list, P.traj_durations, G.traj_bboxes, G.traj_durations)
    so = torch.argmax(G.adj_matrix, dim=-1).t()
    gt5 = torch.cat([G.pred_cat_ids[:, None], G.traj_cat_ids[so], so], -1)
    og.label_maps(viou, gt5, 0.5, 51)
cpu_per_video = (time.perf_counter() - t0) / ns
for P, G in zip(props, graphs):
    P.to(dev); G.to(dev)
torch.cuda.synchronize()
geometry.prop_pair_to_gt_pred(props[:20], graphs[:20], 0.5, 51)
torch.cuda.synchronize()
t0 = time.perf_counter()
out, stats = geometry.prop_pair_to_gt_pred(props, graphs, 0.5, 51)
torch.cuda.synchronize()
gpu_s = time.perf_counter() - t0
print(json.dumps({"videos": V, "gpu_seconds": gpu_s, "gpu_videos_per_s": V / gpu_s, "extrapolated_7000_videos_s": 7000 * gpu_s / V,
                  "cpu_port_seconds_per_video": cpu_per_video, "cpu_port_extrapolated_7000_videos_h": 7000 * cpu_per_video / 3600,
                  "reference_published": "around 1.5 hours (README.md:169)", "stats": stats,
                  "videos_with_labels": sum(v is not None for v in out.values()), "synthetic_generation_s": gen_s}))
