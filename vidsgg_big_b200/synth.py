"""Seeded synthetic inputs and weights.

The datasets (VidVRD / VidOR tracklets, RoI / I3D features) and the released checkpoints
are not available offline, so every test, golden fixture and benchmark uses inputs drawn
here.  Host-side generators use ``numpy.random.default_rng(seed)`` only (bit-stable across
machines), so a golden fixture needs to store the *outputs* of the reference, never the
inputs or weights.  ``device_*`` generators draw directly in HBM for benchmark-sized work.

State-dict key names / shapes follow the reference modules exactly so the same dict loads
into ``models/model_0v10.py:BIG_C``, ``models/model_0v7.py:BIG_C`` and
``models/grd_model_v5.py:DEBUG`` with ``load_state_dict(strict=True)``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .containers import TrajProposal, VideoGraph

# ----------------------------------------------------------------------------------------
# configs (dimensions from experiments/exp2/config_.py:3-30, exp5/config_.py:3-31,
# grounding_weights/config_.py:2-14,74-80)
# ----------------------------------------------------------------------------------------

def vidvrd_config(**over) -> dict:
    cfg = dict(
        num_enti_cats=36, num_pred_cats=133, dim_ffn=512, dim_enti=512, dim_pred=512, dim_att=512,
        dim_feat=2048, dim_clsme=300, dim_i3d=832, enco_pool_len=4, n_enco_layers=2, n_deco_layers=6,
        n_att_head=8, num_querys=192, neg_weight=0.1, positive_vIoU_th=0.5,
        EntiNameEmb_path=None, bias_matrix_path=None,
        cost_coeff_dict=dict(classification=1.0, adj_matrix=30.0),
        loss_coeff_dict=dict(classification=1.0, adj_matrix=30.0),
        variant="vidvrd",
    )
    cfg.update(over)
    return cfg


def vidor_config(**over) -> dict:
    cfg = dict(
        dataset_type="VidOR", num_enti_cats=81, num_pred_cats=51, dim_ffn=512, dim_enti=512, dim_pred=512,
        dim_att=512, dim_feat=1024, dim_clsme=300, enco_pool_len=4, n_enco_layers=6, n_deco_layers=4,
        n_att_head=8, num_querys=192, neg_weight=0.1, positive_vIoU_th=0.5,
        EntiNameEmb_path=None, use_clsme=True, bias_matrix_path=None,
        cost_coeff_dict=dict(classification=1.0, adj_matrix=30.0),
        loss_coeff_dict=dict(classification=1.0, adj_matrix=30.0),
        variant="vidor",
    )
    cfg.update(over)
    return cfg


def basec_config(**over) -> dict:
    """Base-C pairwise baseline, experiments/exp6/config_.py:1-17 (``rt_triplets_topk`` -1 = return all, 200 in config_rt200.py)."""
    cfg = dict(dataset_type="VidOR", num_enti_cats=81, num_pred_cats=51, dim_ffn=512, dim_enti=512, dim_pred=512, dim_att=512,
               dim_feat=1024, dim_clsme=300, enco_pool_len=4, positive_vIoU_th=0.5, rt_triplets_topk=-1,
               EntiNameEmb_path=None, use_clsme=True, bias_matrix_path=None)
    cfg.update(over)
    return cfg


def tiny_basec_config(**over) -> dict:
    cfg = basec_config(num_enti_cats=9, num_pred_cats=17, dim_ffn=64, dim_enti=64, dim_feat=96, dim_clsme=20)
    cfg.update(over)
    return cfg


def make_basec_state(seed: int, cfg: dict) -> "OrderedDict[str, torch.Tensor]":
    """Random-but-healthy Base-C weights with the reference's state_dict keys (models/model_pairwise_baseline.py:27-76)."""
    w = _W(seed)
    E, C, NP = cfg["dim_enti"], cfg["num_enti_cats"], cfg["num_pred_cats"]
    if cfg.get("EntiNameEmb_path") is not None or cfg.get("_force_entiemb", False):
        w.sd["EntiNameEmb"] = torch.from_numpy(w.rng.standard_normal((C, cfg["dim_clsme"]), dtype=np.float32) * np.float32(0.4))
    dirich = w.rng.dirichlet(np.ones(NP), size=(C, C)).astype(np.float32)
    w.sd["bias_matrix"] = torch.from_numpy(np.log(dirich + np.float32(1e-3)).astype(np.float32))
    w.linear("fc_feat2enti.0", E, cfg["dim_feat"]); w.linear("fc_feat2enti.2", E, E)
    w.linear("fc_bbox2enti.0", E, 8); w.linear("fc_bbox2enti.2", E, E)
    w.mat("conv_feat2enti.weight", E, 2 * E, 3); w.vec("conv_feat2enti.bias", E)
    w.linear("fc_enti2enco.0", E, E * cfg["enco_pool_len"]); w.linear("fc_enti2enco.2", E, E)
    dz = 2 * E + 2 * cfg["dim_clsme"]        # the reference sizes this layer for the classeme input regardless of use_clsme (:63-67)
    w.linear("fc_pred2logits.0", cfg["dim_ffn"], dz, gain=2.0); w.linear("fc_pred2logits.2", NP, cfg["dim_ffn"], gain=2.0)
    return w.sd


def grounding_config(**over) -> dict:
    cfg = dict(dim_feat=1024, dim_clsme=300, dim_hidden=128, num_bins=10,
               EntiNameEmb_path=None, PredNameEmb_path=None,
               loss_factor=dict(classification=1.0, centerness=1.0, regression=1.0))
    cfg.update(over)
    return cfg


GROUNDING_INFERENCE = dict(score_th=0.9, tiou_th=0.5, bins_th=0.2, nms_th=0.8)


def tiny_vidvrd_config(**over) -> dict:
    """Reduced dims for fast CPU fixtures (same structure as exp2)."""
    cfg = vidvrd_config(num_enti_cats=9, num_pred_cats=17, dim_ffn=64, dim_enti=64, dim_pred=64, dim_att=64,
                        dim_feat=96, dim_clsme=20, dim_i3d=40, n_enco_layers=2, n_deco_layers=3,
                        n_att_head=4, num_querys=24)
    cfg.update(over)
    return cfg


def tiny_vidor_config(**over) -> dict:
    cfg = vidor_config(num_enti_cats=11, num_pred_cats=13, dim_ffn=64, dim_enti=64, dim_pred=64, dim_att=64,
                       dim_feat=80, dim_clsme=20, n_enco_layers=3, n_deco_layers=2, n_att_head=4, num_querys=24)
    cfg.update(over)
    return cfg


# ----------------------------------------------------------------------------------------
# tracklets
# ----------------------------------------------------------------------------------------

def _random_walk_boxes(rng, L, w, h):
    """Smooth random-walk boxes (center N(0,4 px)/frame, 2 % size jitter), clipped, f32 xyxy."""
    bw0 = rng.uniform(0.08, 0.45) * w
    bh0 = rng.uniform(0.08, 0.45) * h
    cx0 = rng.uniform(bw0 / 2, w - bw0 / 2)
    cy0 = rng.uniform(bh0 / 2, h - bh0 / 2)
    cx = cx0 + np.cumsum(rng.normal(0, 4.0, L))
    cy = cy0 + np.cumsum(rng.normal(0, 4.0, L))
    bw = bw0 * (1.0 + rng.normal(0, 0.02, L))
    bh = bh0 * (1.0 + rng.normal(0, 0.02, L))
    x1 = np.clip(cx - bw / 2, 0, w - 2)
    y1 = np.clip(cy - bh / 2, 0, h - 2)
    x2 = np.clip(cx + bw / 2, x1 + 1, w - 1)
    y2 = np.clip(cy + bh / 2, y1 + 1, h - 1)
    return np.stack([x1, y1, x2, y2], 1).astype(np.float32)


def proposal_lengths(seed: int, n: int, video_len: int, min_len: int = 5, max_len: Optional[int] = None) -> np.ndarray:
    """Track lengths ``make_proposal`` will draw for this seed (its first draw), without building the boxes: lets a scheduler
    cost a video (``shard.video_cost``) before anything is generated."""
    rng = np.random.default_rng(seed)
    max_len = min(max_len or video_len, video_len)
    min_len = min(min_len, max_len)
    return rng.integers(min_len, max_len + 1, size=n)


def make_proposal(seed: int, n: int, video_len: int, dim_feat_total: int, num_enti_cats: int,
                  min_len: int = 5, max_len: Optional[int] = None, wh=(1280, 720),
                  with_features: bool = True, feat_scale: float = 0.5, name: Optional[str] = None) -> TrajProposal:
    """One video's proposals.  Spans closed, scores sorted descending (dataloader_vidvrd.py:40-46)."""
    rng = np.random.default_rng(seed)
    w, h = wh
    max_len = min(max_len or video_len, video_len)
    min_len = min(min_len, max_len)
    lens = rng.integers(min_len, max_len + 1, size=n)
    starts = np.array([rng.integers(0, video_len - L + 1) for L in lens], dtype=np.int64)
    duras = np.stack([starts, starts + lens - 1], 1).astype(np.int64)
    scores = np.sort(rng.uniform(0.3, 1.0, n).astype(np.float32))[::-1].copy()
    cats = rng.integers(1, num_enti_cats, size=n).astype(np.int64)
    boxes = np.concatenate([_random_walk_boxes(rng, int(L), w, h) for L in lens], 0) if n else np.zeros((0, 4), np.float32)
    feats = None
    if with_features and n:
        feats = (rng.standard_normal((int(lens.sum()), dim_feat_total), dtype=np.float32) * feat_scale)
    return TrajProposal(name or "synth_%06d" % seed, video_len, wh, torch.from_numpy(cats),
                        torch.from_numpy(scores), torch.from_numpy(duras), torch.from_numpy(boxes),
                        None if feats is None else torch.from_numpy(feats), torch.from_numpy(lens.astype(np.int64)))


def make_gt_graph(seed: int, proposal: TrajProposal, num_pred_cats: int, n_traj=(3, 8), n_rel=(5, 60),
                  jitter_px: float = 3.0) -> VideoGraph:
    """GT derived from proposals (SURVEY §8d cfg 3): pick tracks, jitter + round boxes, random sub-spans.

    ``traj_durations`` closed; ``pred_durations`` closed floats inside the s-o overlap.
    """
    rng = np.random.default_rng(seed + 77_000_000)
    n = proposal.num_proposals
    k = int(min(n, rng.integers(n_traj[0], n_traj[1] + 1)))
    pick = np.sort(rng.choice(n, size=k, replace=False))
    lens = proposal.lengths.numpy()
    offs = np.concatenate([[0], np.cumsum(lens)])
    duras = proposal.traj_durations.cpu().numpy()
    pb = proposal.bboxes.cpu().numpy()
    g_boxes, g_duras, g_cats = [], [], []
    for t in pick:
        s, e = duras[t]
        L = e - s + 1
        # GT tracks cover a random sub-span of the proposal so vIoU is < 1
        cut_s = int(rng.integers(0, max(1, L // 6)))
        cut_e = int(rng.integers(0, max(1, L // 6)))
        if L - cut_s - cut_e < 2:
            cut_s = cut_e = 0
        b = pb[offs[t] + cut_s: offs[t + 1] - cut_e]
        b = np.rint(b + rng.normal(0, jitter_px, b.shape)).astype(np.float32)
        b[:, 2] = np.maximum(b[:, 2], b[:, 0] + 1)
        b[:, 3] = np.maximum(b[:, 3], b[:, 1] + 1)
        g_boxes.append(b)
        g_duras.append([s + cut_s, e - cut_e])
        g_cats.append(int(proposal.cat_ids[t]))
    g_duras = np.asarray(g_duras, dtype=np.int64).reshape(-1, 2)
    # relations between temporally overlapping GT tracks
    pairs = [(a, b) for a in range(k) for b in range(k)
             if a != b and max(g_duras[a, 0], g_duras[b, 0]) <= min(g_duras[a, 1], g_duras[b, 1])]
    rel_s, rel_o, rel_c, rel_d = [], [], [], []
    if pairs:
        m = int(rng.integers(n_rel[0], n_rel[1] + 1))
        for _ in range(m):
            a, b = pairs[int(rng.integers(0, len(pairs)))]
            s = max(g_duras[a, 0], g_duras[b, 0]); e = min(g_duras[a, 1], g_duras[b, 1])
            ov = e - s + 1
            ln = int(rng.integers(min(30, ov), ov + 1))
            st = int(s + rng.integers(0, ov - ln + 1))
            rel_s.append(a); rel_o.append(b); rel_c.append(int(rng.integers(1, num_pred_cats)))
            rel_d.append([st, st + ln - 1])
    npred = len(rel_c)
    adj = np.zeros((2, npred, k), np.float32)
    for i in range(npred):
        adj[0, i, rel_s[i]] = 1.0
        adj[1, i, rel_o[i]] = 1.0
    g = VideoGraph(proposal.video_name, proposal.video_len, proposal.video_wh, g_cats, g_duras,
                   np.concatenate(g_boxes, 0) if g_boxes else np.zeros((0, 4), np.float32),
                   rel_c, np.asarray(rel_d, np.float32).reshape(-1, 2), adj,
                   lengths=[b.shape[0] for b in g_boxes])
    g.src_prop_ids = pick.tolist()      # which proposal each GT track was derived from (synthetic only)
    return g


def make_predictions(seed: int, proposal: TrajProposal, gt: VideoGraph, num_pred_cats: int, m: int = 200,
                     p_from_gt: float = 0.5, score_decimals: int = 2):
    """Synthetic classification-stage output ``(quintuples i64[m,5], score f32[m], span i64[m,2])``.

    About ``p_from_gt`` of the rows reuse a GT relation's categories on the proposal tracks the
    GT tracks were derived from (so the evaluation has hits); the rest are random overlapping
    pairs.  Scores are rounded to ``score_decimals`` so ties occur (stable-sort semantics matter).
    """
    rng = np.random.default_rng(seed + 33_000_000)
    n = proposal.num_proposals
    d = proposal.traj_durations.cpu().numpy()
    cats = proposal.cat_ids.cpu().numpy()
    ov_pairs = [(a, b) for a in range(n) for b in range(n)
                if a != b and max(d[a, 0], d[b, 0]) <= min(d[a, 1], d[b, 1])]
    if not ov_pairs:
        return None
    so_gt = torch.argmax(gt.adj_matrix, dim=-1).t().numpy() if gt.num_preds else np.zeros((0, 2), np.int64)
    rows, spans = [], []
    for _ in range(m):
        if gt.num_preds and rng.uniform() < p_from_gt:
            g = int(rng.integers(0, gt.num_preds))
            a, b = gt.src_prop_ids[so_gt[g, 0]], gt.src_prop_ids[so_gt[g, 1]]
            pc = int(gt.pred_cat_ids[g])
            if max(d[a, 0], d[b, 0]) > min(d[a, 1], d[b, 1]):
                a, b = ov_pairs[int(rng.integers(0, len(ov_pairs)))]
        else:
            a, b = ov_pairs[int(rng.integers(0, len(ov_pairs)))]
            pc = int(rng.integers(1, num_pred_cats))
        s, e = max(d[a, 0], d[b, 0]), min(d[a, 1], d[b, 1])
        if rng.uniform() < 0.5 and e - s >= 4:          # grounded sub-span
            ln = int(rng.integers(2, e - s + 2)); st = int(s + rng.integers(0, e - s + 2 - ln))
            s, e = st, st + ln - 1
        rows.append([pc, int(cats[a]), int(cats[b]), a, b]); spans.append([s, e])
    score = np.round(rng.uniform(0.05, 1.0, len(rows)), score_decimals).astype(np.float32)
    return (torch.tensor(rows, dtype=torch.long), torch.from_numpy(score), torch.tensor(spans, dtype=torch.long))


def make_gt_from_predictions(seed: int, proposal: TrajProposal, triplets, n_rel=(5, 40), jitter_px: float = 3.0,
                             p_wrong_pred: float = 0.3, num_pred_cats: int = 133) -> VideoGraph:
    """GT graph whose relations are drawn from a model's own predictions (so the evaluation has real matches to find):
    up to ``n_rel`` predicted (pred, subject track, object track) rows become GT relations on jittered + rounded copies of those
    proposal tracks, with a random sub-interval of the s-o overlap as duration; a fraction gets a different predicate (misses).
    ``triplets`` = (quintuples i64[m,5], scores, spans i64[m,2]) on any device, or None (-> falls back to ``make_gt_graph``)."""
    if triplets is None or triplets[0].shape[0] == 0:
        return make_gt_graph(seed, proposal, num_pred_cats)
    rng = np.random.default_rng(seed + 99_000_000)
    quint = triplets[0].cpu().numpy()
    spans = triplets[2].cpu().numpy()
    m = int(min(quint.shape[0], rng.integers(n_rel[0], n_rel[1] + 1)))
    rows = np.sort(rng.choice(quint.shape[0], size=m, replace=False))
    used = sorted(set(quint[rows, 3].tolist()) | set(quint[rows, 4].tolist()))
    remap = {t: i for i, t in enumerate(used)}
    lens = proposal.lengths.numpy()
    offs = np.concatenate([[0], np.cumsum(lens)])
    duras = proposal.traj_durations.cpu().numpy()
    pb = proposal.bboxes.cpu().numpy()
    cats = proposal.cat_ids.cpu().numpy()
    g_boxes, g_duras, g_cats = [], [], []
    for t in used:
        b = np.rint(pb[offs[t]:offs[t + 1]] + rng.normal(0, jitter_px, (int(lens[t]), 4))).astype(np.float32)
        b[:, 2] = np.maximum(b[:, 2], b[:, 0] + 1)
        b[:, 3] = np.maximum(b[:, 3], b[:, 1] + 1)
        g_boxes.append(b); g_duras.append(duras[t].tolist()); g_cats.append(int(cats[t]))
    rel_c, rel_d = [], []
    adj = np.zeros((2, m, len(used)), np.float32)
    for i, r in enumerate(rows):
        pc, _, _, st, ot = quint[r]
        s, e = int(spans[r, 0]), int(spans[r, 1])
        ln = int(rng.integers(max(1, (e - s + 1) // 2), e - s + 2))
        b0 = int(s + rng.integers(0, e - s + 2 - ln))
        if rng.uniform() < p_wrong_pred:
            pc = int(1 + (pc + rng.integers(0, num_pred_cats - 2)) % (num_pred_cats - 1))
        rel_c.append(int(pc)); rel_d.append([b0, b0 + ln - 1])
        adj[0, i, remap[int(st)]] = 1.0
        adj[1, i, remap[int(ot)]] = 1.0
    g = VideoGraph(proposal.video_name, proposal.video_len, proposal.video_wh, g_cats, np.asarray(g_duras, np.int64).reshape(-1, 2),
                   np.concatenate(g_boxes, 0), rel_c, np.asarray(rel_d, np.float32).reshape(-1, 2), adj, lengths=[b.shape[0] for b in g_boxes])
    g.src_prop_ids = used
    return g


def make_bipartite_case(seed: int, n_querys: int, num_pred_cats: int, n_gt: int, n_enti: int):
    """Inputs of ``BIG_C.bipartite_match`` (model_0v10.py:606-639): logits f32[Q,P], GT predicate ids i64[G], attention f32[2,Q,n] in (0,1)
    (softmax over tracklets x softmax over roles, like the decoder emits), aligned GT adjacency f32[2,G,n] (0/1)."""
    rng = np.random.default_rng(seed + 11_000_000)
    logit = torch.from_numpy(rng.standard_normal((n_querys, num_pred_cats), dtype=np.float32) * np.float32(2.0))
    gt_pred = torch.from_numpy(rng.integers(1, num_pred_cats, size=n_gt).astype(np.int64))
    raw = torch.from_numpy(rng.standard_normal((2, n_querys, n_enti), dtype=np.float32) * np.float32(3.0))
    att = torch.softmax(raw, -1) * torch.softmax(raw, 0)
    att[0, 0, 0] = 0.0                                   # exact zeros / ones hit the BCE log clamp (-100)
    att[1, 1, min(1, n_enti - 1)] = 1.0
    adj = torch.zeros(2, n_gt, n_enti)
    for g in range(n_gt):
        adj[0, g, int(rng.integers(0, n_enti))] = 1.0
        adj[1, g, int(rng.integers(0, n_enti))] = 1.0
    return logit, gt_pred, att, adj


def vidvrd_video_shape(rng) -> Tuple[int, int]:
    """(video_len, n) of SURVEY §8d config 2."""
    return int(rng.integers(90, 1201)), int(rng.integers(5, 51))


def vidor_video_shape(rng) -> Tuple[int, int]:
    """(video_len, n) of SURVEY §8d config 4: lognormal length (mean≈1040) clipped 120..5400."""
    sigma = 0.8
    mu = math.log(1040.0) - 0.5 * sigma * sigma
    vl = int(np.clip(rng.lognormal(mu, sigma), 120, 5400))
    return vl, int(rng.integers(10, 181))


# ----------------------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------------------

class _W:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def mat(self, name, *shape, fan=None, gain=1.0):
        if fan is None:
            recept = int(np.prod(shape[2:])) if len(shape) > 2 else 1
            fan = (shape[1] * recept, shape[0] * recept)
        std = gain * math.sqrt(2.0 / (fan[0] + fan[1]))
        self.sd[name] = torch.from_numpy(self.rng.standard_normal(shape, dtype=np.float32) * np.float32(std))

    def vec(self, name, n, mean=0.0, std=0.02):
        self.sd[name] = torch.from_numpy((mean + self.rng.standard_normal(n, dtype=np.float32) * std).astype(np.float32))

    def linear(self, prefix, out_f, in_f, gain=1.0):
        self.mat(prefix + ".weight", out_f, in_f, gain=gain)
        self.vec(prefix + ".bias", out_f)

    def norm(self, prefix, d):
        self.vec(prefix + ".weight", d, mean=1.0, std=0.05)
        self.vec(prefix + ".bias", d, std=0.02)

    def mha(self, prefix, d):
        self.mat(prefix + ".in_proj_weight", 3 * d, d)
        self.vec(prefix + ".in_proj_bias", 3 * d)
        self.linear(prefix + ".out_proj", d, d)


def make_bigc_state(seed: int, cfg: dict) -> "OrderedDict[str, torch.Tensor]":
    """Random-but-healthy BIG-C weights with the reference's state_dict keys.

    Keys: models/model_0v10.py:266-334 (vidvrd) / models/model_0v7.py:265-341 (vidor).
    Unlike ``_reset_parameters`` biases and LayerNorm affine terms are non-trivial so that
    parity tests exercise them.
    """
    w = _W(seed)
    E, P, A, F_ = cfg["dim_enti"], cfg["dim_pred"], cfg["dim_att"], cfg["dim_ffn"]
    Q, C, NP = cfg["num_querys"], cfg["num_enti_cats"], cfg["num_pred_cats"]
    vidor = cfg.get("variant") == "vidor"
    has_emb = (not vidor) or (cfg.get("EntiNameEmb_path") is not None) or cfg.get("_force_entiemb", False)
    if has_emb:
        w.sd["EntiNameEmb"] = torch.from_numpy(w.rng.standard_normal((C, cfg["dim_clsme"]), dtype=np.float32) * np.float32(0.4))
    if vidor:
        w.sd["pos_embedding"] = sine_pos_emb(Q, P)
    else:
        w.sd["pos_embedding"] = torch.from_numpy(w.rng.standard_normal((Q, P), dtype=np.float32) * np.float32(0.1))
    w.sd["pred_query_init"] = torch.from_numpy(w.rng.standard_normal((Q, P), dtype=np.float32) * np.float32(0.1))
    dirich = w.rng.dirichlet(np.ones(NP), size=(C, C)).astype(np.float32)
    w.sd["bias_matrix"] = torch.from_numpy(np.log(dirich + np.float32(1e-3)).astype(np.float32))
    w.linear("fc_feat2enti.0", E, cfg["dim_feat"]); w.linear("fc_feat2enti.2", E, E)
    w.linear("fc_bbox2enti.0", E, 8); w.linear("fc_bbox2enti.2", E, E)
    w.mat("conv_feat2enti.weight", E, 2 * E, 3); w.vec("conv_feat2enti.bias", E)
    w.linear("fc_enti2enco.0", E, E * cfg["enco_pool_len"]); w.linear("fc_enti2enco.2", E, E)
    if (not vidor) and cfg.get("dim_i3d"):
        w.linear("fc_i3d.0", E, cfg["dim_i3d"])
    for i in range(cfg["n_enco_layers"]):
        p = "encoder_layers.%d" % i
        w.mha(p + ".self_attn", E)
        w.linear(p + ".linear1", F_, E); w.linear(p + ".linear2", E, F_)
        w.norm(p + ".norm1", E); w.norm(p + ".norm2", E)
    for i in range(cfg["n_deco_layers"]):
        p = "decoder_layers.%d" % i
        w.mha(p + ".self_attn", P)
        for r in range(2):
            w.linear(p + ".fc_rolewise.%d.0" % r, P, E); w.linear(p + ".fc_rolewise.%d.2" % r, P, P)
        # gain > 1 keeps the role attention away from a uniform softmax
        w.linear(p + ".fc_enti2att", A, E, gain=3.0); w.linear(p + ".fc_pred2att", A, P, gain=3.0)
        w.linear(p + ".fc2.0", F_, P); w.linear(p + ".fc2.3", P, F_)
        for k in (1, 2, 3):
            w.norm(p + ".norm%d" % k, P)
    if vidor:
        dq = P + 2 * E + (2 * cfg["dim_clsme"] if cfg["use_clsme"] else 0)
        w.linear("fc_pred2logits.0", F_, dq); w.linear("fc_pred2logits.2", NP, F_)
    else:
        dq = P + 2 * cfg["dim_clsme"] + (4 * E if cfg.get("dim_i3d") else 2 * E)
        w.linear("fc_pred2logits", NP, dq)
    return w.sd


def sine_pos_emb(length: int, d_model: int) -> torch.Tensor:
    """models/model_0v7.py:228-237 ``SinePosEmb`` (also grd_model_v5.py:58-78 ``PosEncoder``)."""
    i = np.arange(d_model)
    freqs = np.where(i % 2 == 0, 10000.0 ** (-i / d_model), -(10000.0 ** ((1 - i) / d_model)))
    phases = np.where(i % 2 == 0, 0.0, np.pi / 2)
    freqs = torch.tensor(freqs.tolist(), dtype=torch.float32)[None, :]
    phases = torch.tensor(phases.tolist(), dtype=torch.float32)[None, :]
    pos = torch.arange(length)[:, None].repeat(1, d_model).float()
    return torch.sin(pos * freqs + phases)


def make_grounding_state(seed: int, cfg: dict, conv_scale: float = 0.35, head_gain: float = 30.0) -> "OrderedDict[str, torch.Tensor]":
    """Grounding (``DEBUG``) weights, keys per models/grd_model_v5.py:148-192.

    Depthwise / pointwise conv weights are scaled by ``conv_scale`` (SURVEY §8c caveat 1):
    un-scaled random conv stacks saturate the sigmoid heads and crash ``temporal_pooling``.
    """
    w = _W(seed)
    H, B = cfg["dim_hidden"], cfg["num_bins"]
    w.sd["EntiNameEmb"] = torch.from_numpy(w.rng.standard_normal((81, cfg["dim_clsme"]), dtype=np.float32) * np.float32(0.4))
    w.sd["PredNameEmb"] = torch.from_numpy(w.rng.standard_normal((51, cfg["dim_clsme"]), dtype=np.float32) * np.float32(0.4))
    w.linear("video_fc", H, cfg["dim_feat"]); w.linear("query_fc", H, cfg["dim_clsme"])
    w.linear("temp_fc", H, 2); w.linear("vq_fc", H, 4 * H)

    def dws(prefix, cin, cout, k, pw_gain=1.0):
        std_dw = conv_scale * math.sqrt(2.0 / k)
        std_pw = pw_gain * conv_scale * math.sqrt(2.0 / cin)
        w.sd[prefix + ".depth_wise.weight"] = torch.from_numpy(w.rng.standard_normal((cin, 1, k), dtype=np.float32) * np.float32(std_dw))
        w.vec(prefix + ".depth_wise.bias", cin)
        w.sd[prefix + ".point_wise.weight"] = torch.from_numpy(w.rng.standard_normal((cout, cin, 1), dtype=np.float32) * np.float32(std_pw))
        w.vec(prefix + ".point_wise.bias", cout)

    for name, k in (("video_encoder", 7), ("query_encoder", 3), ("combined_encoder", 7)):
        for c in range(4):
            dws("%s.convs.%d" % (name, c), H, H, k)
        w.mha(name + ".mh_attn", H)
        w.linear(name + ".fc", H, H)
        w.norm(name + ".normb", H)
        for c in range(4):
            w.norm("%s.norm_seq.%d" % (name, c), H)
        w.norm(name + ".norme", H)
    w.mat("proj2sim.weight", H, H)
    for head, out in (("cls_head", B), ("conf_head", B), ("regr_head", 2 * B)):
        for c in range(4):
            dws("%s.%d.0" % (head, c), H, H, 3)
        # final projection amplified so that logits span a few units (non-degenerate scores / masks)
        dws("%s.4" % head, H, out, 3, pw_gain=head_gain)
    return w.sd


def make_video_feature(seed: int, video_len: int, dim: int = 1024, scale: float = 0.05) -> torch.Tensor:
    """I3D clip features f32[T, dim], clips of 16 frames stride 8 => T = ceil(video_len / 8)."""
    rng = np.random.default_rng(seed + 55_000_000)
    T = (video_len + 7) // 8
    return torch.from_numpy(rng.standard_normal((T, dim), dtype=np.float32) * np.float32(scale))
