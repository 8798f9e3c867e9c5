"""Oracle: tensors -> VidVRD-helper relation dicts.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Follows utils/evaluate.py:72-151
(``EvalFmtCvtor.to_eval_format_pr``) and the GT side of utils/evaluate.py:234-300 /
VidVRD-helper/dataset/dataset.py:173-208 (relation dict format).
"""
from __future__ import annotations

import torch


def default_names(prefix, n):
    return {i: "%s%d" % (prefix, i) for i in range(n)}


def cut(track, track_span_ho, span_ho):
    """utils/utils_func.py:523-536 ``traj_cutoff`` on half-open spans."""
    s0, e0 = track_span_ho
    s, e = span_ho
    assert len(track) == e0 - s0 and s0 <= s and e <= e0
    return track[s - s0: len(track) - (e0 - e)]


def to_eval_format_pr(proposal, triplets, enti_names, pred_names):
    """Prediction dicts of one video.  ``triplets`` = (quintuples i64[m,5], score f32[m], span i64[m,2] closed)
    or None.  Durations become half-open [s, e+1); trajectories are cut to the span."""
    if triplets is None:
        return {proposal.video_name: []}
    quint, score, span = triplets
    boxes = proposal.bboxes_list
    duras = proposal.traj_durations.clone().tolist()
    quint, score, span = quint.tolist(), score.tolist(), span.tolist()
    out = []
    for i in range(len(quint)):
        pc, sc, oc, st, ot = quint[i]
        if pc == 0:
            continue
        d = (span[i][0], span[i][1] + 1)
        sub = cut(boxes[st], (duras[st][0], duras[st][1] + 1), d)
        obj = cut(boxes[ot], (duras[ot][0], duras[ot][1] + 1), d)
        assert len(sub) == len(obj) == d[1] - d[0]
        out.append({
            "triplet": [enti_names[sc], pred_names[pc], enti_names[oc]],
            "duration": d,
            "score": float(score[i]),
            "sub_traj": sub.cpu().numpy().tolist(),
            "obj_traj": obj.cpu().numpy().tolist(),
        })
    return {proposal.video_name: out}


def to_eval_format_gt(gt_graph, enti_names, pred_names):
    """GT relation dicts of one video from a ``VideoGraph`` (closed spans -> half-open; integer boxes)."""
    so = torch.argmax(gt_graph.adj_matrix, dim=-1).t()
    tb = gt_graph.traj_bboxes
    td = gt_graph.traj_durations.tolist()
    pd = gt_graph.pred_durations.long().tolist()
    out = []
    for i in range(gt_graph.num_preds):
        s, o = so[i].tolist()
        d = (pd[i][0], pd[i][1] + 1)
        sub = cut(tb[s], (td[s][0], td[s][1] + 1), d)
        obj = cut(tb[o], (td[o][0], td[o][1] + 1), d)
        out.append({
            "triplet": [enti_names[int(gt_graph.traj_cat_ids[s])], pred_names[int(gt_graph.pred_cat_ids[i])],
                        enti_names[int(gt_graph.traj_cat_ids[o])]],
            "subject_tid": s, "object_tid": o,
            "duration": d,
            "sub_traj": [[int(v) for v in b] for b in sub.tolist()],
            "obj_traj": [[int(v) for v in b] for b in obj.tolist()],
        })
    return {gt_graph.video_name: out}
