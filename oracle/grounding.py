"""Oracle: grounding stage ``DEBUG`` inference, torch-CPU fp32.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Functional restatement over a ``state`` dict of models/grd_model_v5.py: ``prepare_data``
:310-328, ``forward_propagation`` :331-373 (``QANetEncoderLayer`` :110-137, ``PosEncoder``
:58-78, ``DepthWiseSeparableConv1d`` :36-56), ``_forward_test_single`` :530-576,
``temporal_pooling`` :697-737, ``temporal_nms``/``_nms`` :667-695.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from .bigc import _lin, _ln, _mha
from .geometry import dura_intersection, generalized_tiou, tiou


def _pos_encoding(d_model: int, length: int) -> torch.Tensor:
    """(d_model, length) sine table.  grd_model_v5.py:58-78."""
    i = np.arange(d_model)
    freqs = [10000 ** (-k / d_model) if k % 2 == 0 else -10000 ** ((1 - k) / d_model) for k in range(d_model)]
    phases = [0 if k % 2 == 0 else np.pi / 2 for k in range(d_model)]
    fr = torch.Tensor(freqs)[:, None]
    ph = torch.Tensor(phases)[:, None]
    pos = torch.arange(length)[None, :].repeat(d_model, 1).float()
    return torch.sin(pos * fr + ph)


def _dws_conv(st, prefix, x):
    """Depthwise (k, pad k//2, groups=C) then pointwise conv over (N, C, T).  :36-56."""
    wd = st[prefix + ".depth_wise.weight"]
    k = wd.shape[-1]
    x = F.conv1d(x, wd, st[prefix + ".depth_wise.bias"], padding=k // 2, groups=wd.shape[0])
    return F.conv1d(x, st[prefix + ".point_wise.weight"], st[prefix + ".point_wise.bias"])


def qanet_encoder(st, p, x, num_conv=4, n_head=8):
    """x: (N, C, T).  grd_model_v5.py:110-137 in eval mode (dropout inactive)."""
    N, C, T = x.shape
    out = x + _pos_encoding(C, T)[None]
    res = out
    out = _ln(out.transpose(1, 2), st, p + ".normb").transpose(1, 2)
    for i in range(num_conv):
        out = F.relu(_dws_conv(st, "%s.convs.%d" % (p, i), out)) + res
        res = out
        out = _ln(out.transpose(1, 2), st, "%s.norm_seq.%d" % (p, i)).transpose(1, 2)
    seq = out.permute(0, 2, 1)                                   # (N, T, C)
    att = torch.stack([_mha(st, p + ".mh_attn", s, s, s, n_head) for s in seq], 0)
    out = att.permute(0, 2, 1) + res
    res = out
    out = _lin(_ln(out.transpose(1, 2), st, p + ".norme"), st, p + ".fc").transpose(1, 2)
    return F.relu(out) + res


def prepare_data(st, quintuples, spans, video_len):
    """grd_model_v5.py:310-328."""
    words = torch.stack([st["EntiNameEmb"][quintuples[:, 1]], st["PredNameEmb"][quintuples[:, 0]],
                         st["EntiNameEmb"][quintuples[:, 2]]], dim=1)
    return words, spans.float() / video_len


def forward_propagation(st, video_feature, words, so_span):
    """grd_model_v5.py:331-373.  Returns regrs (nq,T,2B), conf_logits (nq,T,B), cls_logits (nq,T,B)."""
    v = _lin(video_feature, st, "video_fc").t()[None]
    q = _lin(words, st, "query_fc").permute(0, 2, 1) + _lin(so_span, st, "temp_fc")[:, :, None]
    v = qanet_encoder(st, "video_encoder", v)
    q = qanet_encoder(st, "query_encoder", q)
    nq = q.shape[0]
    sim = torch.matmul(F.linear(v.transpose(1, 2), st["proj2sim.weight"]).expand(nq, -1, -1), q)
    s_r = torch.softmax(sim, dim=2)
    s_c = torch.softmax(sim, dim=1)
    s_rc = torch.matmul(s_r, s_c.transpose(1, 2))
    vv = v.expand(nq, -1, -1).transpose(1, 2)
    A = torch.matmul(s_r, q.transpose(1, 2))
    B = torch.matmul(s_rc, vv)
    comb = _lin(torch.cat([vv, A, A * vv, B * vv], -1), st, "vq_fc").transpose(1, 2)
    comb = qanet_encoder(st, "combined_encoder", comb)

    def head(name, sigmoid):
        y = comb
        for c in range(4):
            y = F.relu(_dws_conv(st, "%s.%d.0" % (name, c), y))
        y = _dws_conv(st, "%s.4" % name, y)
        return (torch.sigmoid(y) if sigmoid else y).transpose(1, 2)

    return head("regr_head", True), head("conf_head", False), head("cls_head", False)


def temporal_pooling(regrs, scores, num_bins, score_th, tiou_th):
    """grd_model_v5.py:697-737.  Per (query, bin): clips with score > score_th*top and gIoU with the
    top clip > tiou_th are pooled to [min start, max end]."""
    nq, T, _ = scores.shape
    r = regrs.reshape(nq, T, 2, num_bins)
    clip = torch.linspace(0, 1, T)
    start = clip[None, :, None] - r[:, :, 0, :]
    end = clip[None, :, None] + r[:, :, 1, :]
    out = torch.zeros(nq, num_bins, 2)
    for qi in range(nq):
        for k in range(num_bins):
            s = scores[qi, :, k]
            top, top_id = torch.max(s, dim=0)
            d = torch.stack([start[qi, :, k], end[qi, :, k]], -1)
            g = generalized_tiou(d[top_id][None], d)[0]
            keep = (s > score_th * top) & (g > tiou_th)
            sel = d[keep]
            out[qi, k, 0] = torch.min(sel[:, 0], dim=0)[0]       # raises on empty like the reference (:726)
            out[qi, k, 1] = torch.max(sel[:, 1], dim=0)[0]
    return out


def nms_1d(spans, probs, nms_th):
    """grd_model_v5.py:667-681: ascending argsort, repeatedly keep the last, drop tIoU >= th."""
    order = probs.argsort()
    mat = tiou(spans, spans)
    kept = []
    while order.numel() > 0:
        top = order[-1]
        kept.append(top)
        order = order[(mat[top, order[:-1]] < nms_th).nonzero(as_tuple=True)[0]]
    return torch.stack(kept)


def forward_test_single(st, cfg, video_feature, words, so_span, score_th, tiou_th, bins_th, nms_th):
    """grd_model_v5.py:530-576."""
    regrs, conf, cls = forward_propagation(st, video_feature, words, so_span)
    return postprocess(regrs, conf, cls, so_span, cfg["num_bins"], score_th, tiou_th, bins_th, nms_th)


def postprocess(regrs, conf, cls, so_span, B, score_th, tiou_th, bins_th, nms_th):
    """Everything of grd_model_v5.py:533-576 after the network (discrete decisions live here)."""
    scores = conf.sigmoid() * cls.sigmoid()
    probs = F.pad(scores.max(dim=1)[0], (0, 1), value=1.0)
    mask = probs > bins_th
    pooled = temporal_pooling(regrs, scores, B, score_th, tiou_th)
    ovl = []
    for k in range(B):
        clipped, ok = dura_intersection(so_span, pooled[:, k, :], broadcast=False)
        pooled[:, k, :] = so_span.clone()
        pooled[ok, k, :] = clipped[ok, :]
        ovl.append(ok)
    ovl = F.pad(torch.stack(ovl, -1), (0, 1), value=True)
    pooled = torch.cat([pooled, so_span[:, None, :]], 1)
    nms = torch.zeros_like(mask)
    for i in range(pooled.shape[0]):
        nms[i, nms_1d(pooled[i], probs[i], nms_th)] = True
    mask = mask & ovl & nms
    empty = (mask.sum(-1) == 0).nonzero(as_tuple=True)[0]
    if empty.numel() > 0:
        mask[empty, probs[empty].max(-1)[1]] = True
    weak = probs[:, :-1].max(-1)[0] <= bins_th
    probs[weak, -1] = 0.0
    return pooled, probs, mask


def forward(st, cfg, video_feature_list, data_list, score_th=0.5, tiou_th=0.5, bins_th=0.1, nms_th=0.5):
    """``DEBUG.forward(..., with_gt_data=False)`` in test mode.  grd_model_v5.py:196-221."""
    assert len(video_feature_list) == 1
    quint, spans, video_len = data_list[0]
    if quint is None or quint.shape[0] == 0:
        return None, None
    words, so = prepare_data(st, quint, spans, video_len)
    return forward_test_single(st, cfg, video_feature_list[0], words, so, score_th, tiou_th, bins_th, nms_th)


def expand_after_grounding(quintuples, cls_scores3, pooled, probs, mask, video_len):
    """Driver-side expansion, tools/eval_vidor.py:245-253: score = mean(cls scores)*bin prob,
    span = round(pooled*video_len) (long), rows selected by mask (row-major)."""
    nb = probs.shape[1]
    q = quintuples[:, None, :].repeat(1, nb, 1)[mask, :]
    s = (cls_scores3.mean(-1)[:, None] * probs)[mask]
    sp = torch.round((pooled * video_len)[mask, :]).type(torch.long)
    return q, s, sp
