"""Batched drivers replacing the reference's ``batch_size=1`` evaluation loops:

  inference_then_eval   tools/eval_vidvrd.py:42-162   BIG-C -> (convert) -> eval_visual_relation
  evaluate_cls_stage    tools/eval_vidor.py:19-138    BIG-C (VidOR) -> results dict + metrics
  evaluate_combined     tools/eval_vidor.py:141-280   cls-stage results -> grounding -> expansion -> eval

They take in-memory datasets (lists of ``TrajProposal`` / ``VideoGraph`` / I3D tensors already on the device), run the
whole list through the batched kernels and evaluate on the packed path; ``want_dicts=True`` additionally returns the
reference's ``{video_name: [relation dict, ...]}`` structure (for JSON export).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import evalapi, geometry
from .convert import EvalFmtCvtor


def _tables(proposals, gt_graphs):
    return geometry.TrackTable.from_containers(proposals), geometry.TrackTable.from_containers(gt_graphs, device=proposals[0].device)


def inference_then_eval(model, proposals: Sequence, gt_graphs: Sequence, topk: int = 10, viou_threshold: float = 0.5,
                        want_dicts: bool = False, dataset_type: str = "vidvrd"):
    """-> (mean_ap, rec_at_n, mprec_at_n[, predict_relations dict]).  Videos with zero proposals yield no predictions."""
    live = [i for i, p in enumerate(proposals) if p.num_proposals > 0]
    lp, lg = [proposals[i] for i in live], [gt_graphs[i] for i in live]
    with torch.no_grad():
        packed = model.forward_packed(lp, topk=topk)
    pt, gt_t = _tables(lp, lg)
    PR = evalapi.PackedRelations.from_packed_triplets(pt, packed)
    GT = evalapi.PackedRelations.from_gt_graphs(gt_t, lg)
    res = evalapi.evaluate_packed(PR, GT, viou_threshold)
    if not want_dicts:
        return res
    cv = EvalFmtCvtor(dataset_type)
    out = {}
    for p, t in zip(lp, packed.per_video()):
        out.update(cv.to_eval_format_pr(p, None if t is None else (t[0], t[1].mean(-1), t[2])))
    return res + (out,)


def evaluate_cls_stage(model, proposals: Sequence, gt_graphs: Sequence, topk: int = 3, viou_threshold: float = 0.5):
    """-> (metrics, infer_result_for_save) with ``infer_result_for_save[video_name] = [quintuples, scores(n,3), spans, query_ids]``
    (tools/eval_vidor.py:114), tensors on the CPU like the reference's pickle."""
    with torch.no_grad():
        res = model(list(proposals), topk=topk)
    save = {p.video_name: (None if r is None else [x.cpu() for x in r]) for p, r in zip(proposals, res)}
    live = [i for i, r in enumerate(res) if r is not None]
    lp, lg = [proposals[i] for i in live], [gt_graphs[i] for i in live]
    pt, gt_t = _tables(lp, lg)
    PR = evalapi.PackedRelations.from_triplets(pt, [(res[i][0], res[i][1].mean(-1), res[i][2]) for i in live])
    GT = evalapi.PackedRelations.from_gt_graphs(gt_t, lg)
    return evalapi.evaluate_packed(PR, GT, viou_threshold), save


def evaluate_combined(grd_model, cls_model, proposals: Sequence, video_features: Sequence[torch.Tensor], gt_graphs: Sequence,
                      topk: int = 3, score_th=0.9, tiou_th=0.5, bins_th=0.2, nms_th=0.8, viou_threshold: float = 0.5):
    """Classification + grounding + evaluation, all videos batched.  -> (mean_ap, rec_at_n, mprec_at_n, hit_infos)."""
    live = [i for i, p in enumerate(proposals) if p.num_proposals > 0]
    lp = [proposals[i] for i in live]
    with torch.no_grad():
        packed = cls_model.forward_packed(lp, topk=topk)
        q, s3, sp, _, off = packed.compact()
        keep = [k for k in range(len(lp)) if off[k + 1] > off[k]]
        if len(keep) != len(lp):      # videos whose classification stage produced nothing are dropped from the batch
            lp = [lp[k] for k in keep]
            live = [live[k] for k in keep]
            packed = cls_model.forward_packed(lp, topk=topk)
            q, s3, sp, _, off = packed.compact()
        datas = [(q[off[k]:off[k + 1]], sp[off[k]:off[k + 1]], lp[k].video_len) for k in range(len(lp))]
        pooled, probs, mask = grd_model.forward_packed([video_features[i] for i in live], datas, score_th=score_th, tiou_th=tiou_th,
                                                       bins_th=bins_th, nms_th=nms_th)
    lg = [gt_graphs[i] for i in live]
    pt, gt_t = _tables(lp, lg)
    PR = evalapi.PackedRelations.from_grounded(pt, packed, pooled, probs, mask, [p.video_len for p in lp])
    GT = evalapi.PackedRelations.from_gt_graphs(gt_t, lg)
    m_ap, rec, mprec, infos = evalapi.evaluate_packed(PR, GT, viou_threshold, with_infos=True)
    return m_ap, rec, mprec, {lp[k].video_name: v for k, v in infos.items()}
