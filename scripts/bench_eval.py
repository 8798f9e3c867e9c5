"""BASELINE.json configs[2]: eval_visual_relation vIoU matching of synthetic top-k predictions vs GT, VidVRD-test sized
(200 videos, ~1000 predictions and ~25 GT relations per video).  Times the packed path (kernels + native host pass),
the dict-compatible path (incl. packing the dicts) and the CPU oracle on a sample; reports the K3 algorithmic bytes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vidsgg_big_b200 import synth, geometry, evalapi, convert
from oracle import evalapi as oe

V, M = int(os.environ.get("V", 200)), int(os.environ.get("M", 1000))
dev = "cuda:0"
props, graphs, trips = [], [], []
for i in range(V):
    rng = np.random.default_rng(7000 + i)
    vlen, n = synth.vidvrd_video_shape(rng)
    P = synth.make_proposal(7000 + i, n, vlen, 8, 36, min_len=20, max_len=150, with_features=False)
    G = synth.make_gt_graph(7000 + i, P, 133, n_traj=(3, 8), n_rel=(5, 60), jitter_px=2.0)
    props.append(P); graphs.append(G); trips.append(synth.make_predictions(7000 + i, P, G, 133, m=M, p_from_gt=0.5))
cv = convert.EvalFmtCvtor("vidvrd")
ns = int(os.environ.get("CPU_SAMPLE", 10))
gts, prs = {}, {}
for P, G, T in zip(props[:max(ns, 40)], graphs, trips):
    gts.update(cv.to_eval_format_gt(G)); prs.update(cv.to_eval_format_pr(P, T))
sub = {k: gts[k] for k in list(gts)[:ns]}
t0 = time.perf_counter(); ref = oe.evaluate(sub, prs); cpu_per_video = (time.perf_counter() - t0) / ns
# dict path on 40 videos
torch.cuda.synchronize(); t0 = time.perf_counter(); d = evalapi.eval_visual_relation(gts, prs); torch.cuda.synchronize()
dict_per_video = (time.perf_counter() - t0) / len(gts)
# packed path on all videos
for P, G in zip(props, graphs): P.to(dev); G.to(dev)
tt, gt_t = geometry.TrackTable.from_containers(props), geometry.TrackTable.from_containers(graphs)
PR = evalapi.PackedRelations.from_triplets(tt, trips)
GT = evalapi.PackedRelations.from_gt_graphs(gt_t, graphs)
for _ in range(3): evalapi.evaluate_packed(PR, GT, want_records=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): rec = evalapi.evaluate_packed(PR, GT, want_records=True)
torch.cuda.synchronize(); packed_s = (time.perf_counter() - t0) / 10
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); m = evalapi.match_relations(PR, GT, 0.5, keep_ov=True); ev1.record(); torch.cuda.synchronize()
kern_ms = ev0.elapsed_time(ev1)
# algorithmic bytes: 64*ov per same-triplet overlapping (pred, gt) + 32*(sumL_pred + sumL_gt) + 8*pairs
pr, gr = PR.rel.cpu().numpy(), GT.rel.cpu().numpy(); po, go = PR.vid_off_host, GT.vid_off_host
ovsum, pairs = 0, 0
for v in range(V):
    a, b = pr[po[v]:po[v + 1]], gr[go[v]:go[v + 1]]
    pairs += a.shape[0] * b.shape[0]
    same = (a[:, None, :3] == b[None, :, :3]).all(-1)
    ov = np.clip(np.minimum(a[:, None, 6], b[None, :, 6]) - np.maximum(a[:, None, 5], b[None, :, 5]), 0, None)
    ovsum += int((ov * same).sum())
alg = 64 * ovsum + 32 * (int((pr[:, 6] - pr[:, 5]).sum()) + int((gr[:, 6] - gr[:, 5]).sum())) + 8 * pairs
m_ap, rec_at, _ = evalapi.metrics_from_records(rec)
print(json.dumps({"videos": V, "preds_per_video": M, "gt_relations": int(GT.n_rel), "candidate_pairs": pairs, "matched_frame_pairs": ovsum,
                  "cpu_oracle_s_per_video": cpu_per_video, "dict_path_s_per_video": dict_per_video, "packed_path_s_per_video": packed_s / V,
                  "packed_path_videos_per_s": V / packed_s, "match_kernels_ms": kern_ms, "algorithmic_bytes": alg,
                  "algorithmic_GBps": alg / kern_ms / 1e6, "mAP": float(m_ap), "R@50": float(rec_at[50]), "tp_total": int(np.isfinite(m.hit.cpu().numpy()).sum())}))
