"""Oracle: BIG-C classification forward, torch-CPU fp32.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

A functional restatement over a plain ``state`` dict (reference state_dict keys) of
models/model_0v10.py (VidVRD) and models/model_0v7.py (VidOR) inference:
``_preprocess_proposal`` :391-430, ``encode2decode`` :434-475, ``prediction_head`` :478-507
(0v7 :483-513), ``construct_triplet`` :707-785.  It executes what the reference executes --
including the stretched ``[n, Tmax, D]`` tensors -- so that it can stand in for the
reference's CPU cost in ``bench.py``'s cpu_baseline.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .geometry import dura_intersection, stretch_index_map, unique_rows_with_groups


def _lin(x, st, name):
    return F.linear(x, st[name + ".weight"], st[name + ".bias"])


def _ln(x, st, name):
    return F.layer_norm(x, (x.shape[-1],), st[name + ".weight"], st[name + ".bias"], 1e-5)


def _mha(st, prefix, q_in, k_in, v_in, n_head):
    """nn.MultiheadAttention (batch of one sequence), eval mode: packed in-projection, heads of
    d/n_head scaled by 1/sqrt(d_head), softmax over keys, out-projection."""
    d = q_in.shape[-1]
    W, b = st[prefix + ".in_proj_weight"], st[prefix + ".in_proj_bias"]
    q = F.linear(q_in, W[:d], b[:d])
    k = F.linear(k_in, W[d:2 * d], b[d:2 * d])
    v = F.linear(v_in, W[2 * d:], b[2 * d:])
    dh = d // n_head
    q = q.reshape(-1, n_head, dh).transpose(0, 1)
    k = k.reshape(-1, n_head, dh).transpose(0, 1)
    v = v.reshape(-1, n_head, dh).transpose(0, 1)
    att = torch.softmax((q / math.sqrt(dh)) @ k.transpose(1, 2), dim=-1)
    o = (att @ v).transpose(0, 1).reshape(-1, d)
    return _lin(o, st, prefix + ".out_proj")


def box_motion_features(boxes: torch.Tensor, w, h) -> torch.Tensor:
    """8-d [ctx,dctx,cty,dcty,w,dw,h,dh] with a trailing zero on the deltas.  model_0v10.py:401-420."""
    b = boxes.clone()
    b[:, 0:4:2] /= w
    b[:, 1:4:2] /= h
    cx = (b[:, 2] + b[:, 0]) / 2
    cy = (b[:, 3] + b[:, 1]) / 2
    bw = b[:, 2] - b[:, 0]
    bh = b[:, 3] - b[:, 1]
    def delta(v):
        return torch.cat([v[1:] - v[:-1], v.new_zeros(1)])
    return torch.stack([cx, delta(cx), cy, delta(cy), bw, delta(bw), bh, delta(bh)], dim=1)


def preprocess(proposal):
    """Stretched ``[n,Tmax,8]`` box features and ``[n,Tmax,D]`` features.  model_0v10.py:391-430."""
    w, h = proposal.video_wh
    blist, flist = proposal.bboxes_list, proposal.features_list
    Tmax = max(int(b.shape[0]) for b in blist)
    tb, tf = [], []
    for b, f in zip(blist, flist):
        idx = torch.from_numpy(stretch_index_map(int(b.shape[0]), Tmax))
        tb.append(box_motion_features(b, w, h)[idx])
        tf.append(f[idx])
    return torch.stack(tb, 0), torch.stack(tf, 0)


def encoder_layer(st, p, x, n_head):
    """Post-norm encoder layer, no positional embedding.  model_0v10.py:103-117."""
    x = _ln(x + _mha(st, p + ".self_attn", x, x, x, n_head), st, p + ".norm1")
    y = _lin(F.relu(_lin(x, st, p + ".linear1")), st, p + ".linear2")
    return _ln(x + y, st, p + ".norm2")


def decoder_layer(st, p, query, pos, enco, n_head, dim_att, dim_enti):
    """Role-attention decoder layer.  model_0v10.py:178-225."""
    qk = query + pos
    query = _ln(query + _mha(st, p + ".self_attn", qk, qk, query, n_head), st, p + ".norm1")
    query = query + pos                                   # pos added again and kept in the residual
    e2a = _lin(enco, st, p + ".fc_enti2att")
    p2a = _lin(query, st, p + ".fc_pred2att")
    half = dim_att // 2
    logits = torch.stack([p2a[:, :half] @ e2a[:, :half].t(), p2a[:, half:] @ e2a[:, half:].t()], 0) / np.sqrt(dim_enti)
    att = torch.softmax(logits, dim=2) * torch.softmax(logits, dim=0)
    role = 0
    for r in range(2):
        vals = att[r] @ enco
        role = role + _lin(F.relu(_lin(vals, st, p + ".fc_rolewise.%d.0" % r)), st, p + ".fc_rolewise.%d.2" % r)
    query = _ln(query + role, st, p + ".norm2")
    ff = _lin(F.relu(_lin(query, st, p + ".fc2.0")), st, p + ".fc2.3")
    return _ln(query + ff, st, p + ".norm3"), att


def encode2decode(st: Dict[str, torch.Tensor], cfg: dict, proposal, return_intermediates=False):
    """model_0v10.py:434-475 / model_0v7.py:437-480."""
    vidor = cfg.get("variant") == "vidor"
    n = proposal.num_proposals
    E, F_in = cfg["dim_enti"], cfg["dim_feat"]
    tb, tf = preprocess(proposal)
    vis, extra = tf[:, :, :F_in], tf[:, :, F_in:]
    xb = F.relu(_lin(F.relu(_lin(tb, st, "fc_bbox2enti.0")), st, "fc_bbox2enti.2"))
    xv = F.relu(_lin(F.relu(_lin(vis, st, "fc_feat2enti.0")), st, "fc_feat2enti.2"))
    x = torch.cat([xb, xv], -1).permute(0, 2, 1)
    nodes = F.conv1d(x, st["conv_feat2enti.weight"], st["conv_feat2enti.bias"], stride=2, padding=1)
    pooled = F.adaptive_max_pool1d(nodes, cfg["enco_pool_len"]).reshape(n, -1)
    enti2enco = F.relu(_lin(F.relu(_lin(pooled, st, "fc_enti2enco.0")), st, "fc_enti2enco.2"))
    enco = enti2enco
    for i in range(cfg["n_enco_layers"]):
        enco = encoder_layer(st, "encoder_layers.%d" % i, enco, cfg["n_att_head"])
    query, pos = st["pred_query_init"], st["pos_embedding"]
    att = None
    for i in range(cfg["n_deco_layers"]):
        query, att = decoder_layer(st, "decoder_layers.%d" % i, query, pos, enco, cfg["n_att_head"],
                                   cfg["dim_att"], E)
    extra_avg = None
    if vidor:
        if cfg["use_clsme"] and ("EntiNameEmb" not in st):
            extra_avg = extra.mean(dim=1)
    elif cfg.get("dim_i3d"):
        extra_avg = extra.mean(dim=1)
    logits = prediction_head(st, cfg, query, att, proposal.cat_ids, extra_avg, enti2enco)
    if return_intermediates:
        return query, logits, att, dict(enti2enco=enti2enco, enco=enco, pooled=pooled, extra_avg=extra_avg)
    return query, logits, att


def prediction_head(st, cfg, query, att, cat_ids, extra_avg, enti_feat):
    """model_0v10.py:478-507 / model_0v7.py:483-513."""
    vidor = cfg.get("variant") == "vidor"
    so = torch.argmax(att, dim=-1)
    socat = cat_ids[so]
    bias = st["bias_matrix"][socat[0], socat[1], :]
    sf, of = enti_feat[so[0]], enti_feat[so[1]]
    if vidor:
        if cfg["use_clsme"]:
            if "EntiNameEmb" in st:
                sc, oc = st["EntiNameEmb"][socat[0]], st["EntiNameEmb"][socat[1]]
            else:
                sc, oc = extra_avg[so[0]], extra_avg[so[1]]
            z = torch.cat([query, sc, oc, sf, of], -1)
        else:
            z = torch.cat([query, sf, of], -1)
        logits = _lin(F.relu(_lin(z, st, "fc_pred2logits.0")), st, "fc_pred2logits.2")
    else:
        sc, oc = st["EntiNameEmb"][socat[0]], st["EntiNameEmb"][socat[1]]
        if cfg.get("dim_i3d"):
            si = F.relu(_lin(extra_avg[so[0]], st, "fc_i3d.0"))
            oi = F.relu(_lin(extra_avg[so[1]], st, "fc_i3d.0"))
            z = torch.cat([query, si, oi, sf, of, sc, oc], -1)
        else:
            z = torch.cat([query, sc, oc, sf, of], -1)
        logits = _lin(z, st, "fc_pred2logits")
    return logits + bias


def construct_triplet(proposal, logits, att, topk: int):
    """model_0v10.py:707-785: softmax/top-k, temporal-overlap filter, quintuple dedup keeping the
    max predicate score (first on ties), lexicographic order, background removed last."""
    Q = logits.shape[0]
    sc, cat = torch.topk(torch.softmax(logits, dim=-1), topk, dim=-1)
    sc, cat = sc.reshape(-1), cat.reshape(-1)
    qid = torch.arange(Q).repeat_interleave(topk)
    so = torch.argmax(att, dim=-1).t().repeat_interleave(topk, dim=0)
    duras = proposal.traj_durations
    n = duras.shape[0]
    inter, ok = dura_intersection(duras, duras)
    ok[range(n), range(n)] = False
    keep = ok[so[:, 0], so[:, 1]].nonzero(as_tuple=True)[0]
    if keep.numel() == 0:
        return None
    so, sc, cat, qid = so[keep], sc[keep], cat[keep], qid[keep]
    quint = torch.cat([cat[:, None], proposal.cat_ids[so], so], -1)
    trip_sc = torch.cat([sc[:, None], proposal.scores[so]], -1)
    uniq, groups = unique_rows_with_groups(quint)
    pick = torch.stack([g[trip_sc[g, 0].argmax()] for g in groups])
    u_sc, u_q = trip_sc[pick], qid[pick]
    u_span = inter[uniq[:, 3], uniq[:, 4], :]
    fg = uniq[:, 0] != 0
    return uniq[fg], u_sc[fg], u_span[fg], u_q[fg]


def forward(st, cfg, proposal_list, topk: int):
    """``BIG_C.forward`` in test mode.  model_0v10.py:359-388."""
    out = []
    for p in proposal_list:
        if p.num_proposals == 0:
            out.append(None)
            continue
        _, logits, att = encode2decode(st, cfg, p)
        out.append(construct_triplet(p, logits, att, topk))
    return out


def bipartite_cost(pred_logit, gt_pred, att_matrx, gt_adj_enti_align, c_cls: float, c_adj: float):
    """Cost matrix of the Hungarian matching, models/model_0v10.py:606-636: cross-entropy of every query against every GT predicate plus
    the binary cross-entropy between the query's attention rows and the GT's aligned adjacency rows, averaged over roles and tracklets
    (torch's BCE clamps its logs at -100).  -> f32[Q, G]."""
    logp = torch.log_softmax(pred_logit, dim=-1)                       # [Q, P]
    cost_cls = -logp[:, gt_pred]                                        # [Q, G]
    a = att_matrx[:, :, None, :]                                        # [2, Q, 1, n]
    t = gt_adj_enti_align[:, None, :, :]                                # [2, 1, G, n]
    bce = -(t * torch.clamp(torch.log(a), min=-100.0) + (1.0 - t) * torch.clamp(torch.log(1.0 - a), min=-100.0))
    cost_adj = bce.mean(dim=(0, 3))
    return cost_cls * c_cls + cost_adj * c_adj
