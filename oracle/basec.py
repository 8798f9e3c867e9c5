"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the Base-C pairwise baseline in inference mode
(reference models/model_pairwise_baseline.py).  Pinned against the unmodified reference by tests/golden/make_golden.py (basec).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .bigc import _lin, preprocess
from .geometry import dura_intersection, pair_ids, unique_rows_with_groups


def track_encoding(st: Dict[str, torch.Tensor], cfg: dict, proposal):
    """model_pairwise_baseline.py:170-191: per-frame MLPs on the stretched tensors, conv k3 s2, adaptive max-pool, fc_enti2enco;
    classeme = time-mean over the STRETCHED sequence when it comes from the features."""
    n = proposal.num_proposals
    F_in = cfg["dim_feat"]
    tb, tf = preprocess(proposal)
    vis, extra = tf[:, :, :F_in], tf[:, :, F_in:]
    xb = F.relu(_lin(F.relu(_lin(tb, st, "fc_bbox2enti.0")), st, "fc_bbox2enti.2"))
    xv = F.relu(_lin(F.relu(_lin(vis, st, "fc_feat2enti.0")), st, "fc_feat2enti.2"))
    x = torch.cat([xb, xv], -1).permute(0, 2, 1)
    nodes = F.conv1d(x, st["conv_feat2enti.weight"], st["conv_feat2enti.bias"], stride=2, padding=1)
    pooled = F.adaptive_max_pool1d(nodes, cfg["enco_pool_len"]).reshape(n, -1)
    enti2enco = F.relu(_lin(F.relu(_lin(pooled, st, "fc_enti2enco.0")), st, "fc_enti2enco.2"))
    clsme = extra.mean(dim=1) if (cfg["use_clsme"] and "EntiNameEmb" not in st) else None
    return enti2enco, clsme


def prediction_head(st, cfg, pairs, cat_ids, clsme, enti_feat):
    """model_pairwise_baseline.py:243-273."""
    socat = cat_ids[pairs]
    bias = st["bias_matrix"][socat[:, 0], socat[:, 1], :]
    sf, of = enti_feat[pairs[:, 0]], enti_feat[pairs[:, 1]]
    if cfg["use_clsme"]:
        if "EntiNameEmb" in st:
            sc, oc = st["EntiNameEmb"][socat[:, 0]], st["EntiNameEmb"][socat[:, 1]]
        else:
            sc, oc = clsme[pairs[:, 0]], clsme[pairs[:, 1]]
        z = torch.cat([sc, oc, sf, of], -1)
    else:
        z = torch.cat([sf, of], -1)
    return _lin(F.relu(_lin(z, st, "fc_pred2logits.0")), st, "fc_pred2logits.2") + bias


def forward_propagation(st, cfg, proposal, pairs):
    enti2enco, clsme = track_encoding(st, cfg, proposal)
    return prediction_head(st, cfg, pairs, proposal.cat_ids, clsme, enti2enco)


def construct_triplet(proposal, logits, pairs, topk: int, rt_topk: int):
    """model_pairwise_baseline.py:314-395: softmax / top-k per pair, temporal-overlap filter, lexicographic order of the (unique)
    quintuples, background removed, optionally the ``rt_topk`` best by mean score (descending)."""
    sc, cat = torch.topk(torch.softmax(logits, dim=-1), topk, dim=-1)
    sc, cat = sc.reshape(-1), cat.reshape(-1)
    so = pairs.repeat_interleave(topk, dim=0)
    duras = proposal.traj_durations
    n = duras.shape[0]
    inter, ok = dura_intersection(duras, duras)
    ok[range(n), range(n)] = False
    keep = ok[so[:, 0], so[:, 1]].nonzero(as_tuple=True)[0]
    if keep.numel() == 0:
        return None
    so, sc, cat = so[keep], sc[keep], cat[keep]
    quint = torch.cat([cat[:, None], proposal.cat_ids[so], so], -1)
    trip_sc = torch.cat([sc[:, None], proposal.scores[so]], -1)
    uniq, groups = unique_rows_with_groups(quint)
    pick = torch.stack([g[trip_sc[g, 0].argmax()] for g in groups])
    u_sc = trip_sc[pick]
    u_span = inter[uniq[:, 3], uniq[:, 4], :]
    fg = uniq[:, 0] != 0
    uniq, u_sc, u_span = uniq[fg], u_sc[fg], u_span[fg]
    if rt_topk > 0:
        top = u_sc.mean(dim=-1).argsort(descending=True, stable=True)[:rt_topk]
        uniq, u_sc, u_span = uniq[top], u_sc[top], u_span[top]
    return uniq, u_sc, u_span, torch.empty(u_sc.shape[0])


def forward(st, cfg, proposal_list, topk: int):
    out = []
    for p in proposal_list:
        if p.num_proposals == 0:
            out.append(None)
            continue
        pairs = pair_ids(p.num_proposals)
        logits = forward_propagation(st, cfg, p, pairs)
        out.append(construct_triplet(p, logits, pairs, topk, cfg["rt_triplets_topk"]))
    return out
