"""Oracle: VidVRD-helper relation evaluation.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Pure Python / numpy float64, sequential
sums, like VidVRDhelperEvalAPIs/common.py and visual_relation_detection.py.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np

F32_EPS = np.finfo(np.float32).eps


def viou(traj_1, duration_1, traj_2, duration_2) -> float:
    """Volume IoU of two box lists with half-open durations.  common.py:65-106.

    Overlap = [max start, min end); volumes are summed over the whole passed lists; 0.0 when
    the durations do not overlap (touching ends count as no overlap).
    """
    s1, e1 = duration_1
    s2, e2 = duration_2
    if s1 >= e2 or e1 <= s2:
        return 0.
    lo, hi = max(s1, s2), min(e1, e2)
    acc = 0
    for f in range(lo, hi):
        b1 = traj_1[f - s1]
        b2 = traj_2[f - s2]
        ww = min(b1[2], b2[2]) - max(b1[0], b2[0]) + 1
        hh = min(b1[3], b2[3]) - max(b1[1], b2[1]) + 1
        acc += max(0, ww) * max(0, hh)
    vol1 = 0
    for b in traj_1:
        vol1 += (b[2] - b[0] + 1) * (b[3] - b[1] + 1)
    vol2 = 0
    for b in traj_2:
        vol2 += (b[2] - b[0] + 1) * (b[3] - b[1] + 1)
    return float(acc) / (vol1 + vol2 - acc)


def voc_ap(rec, prec) -> float:
    """VOC (non-07) AP: precision envelope integrated where recall changes.  common.py:4-37."""
    r = np.concatenate(([0.], rec, [1.]))
    p = np.concatenate(([0.], prec, [0.]))
    for k in range(p.size - 1, 0, -1):
        p[k - 1] = np.maximum(p[k - 1], p[k])
    step = np.where(r[1:] != r[:-1])[0]
    return np.sum((r[step + 1] - r[step]) * p[step + 1])


def _pr_curves(hit_scores, n_gt):
    tp = np.isfinite(hit_scores)
    ctp = np.cumsum(tp).astype(np.float32)
    cfp = np.cumsum(~tp).astype(np.float32)
    rec = ctp / np.maximum(n_gt, F32_EPS)
    prec = ctp / np.maximum(ctp + cfp, F32_EPS)
    return prec, rec


def detection_scores(gt_relations, pred_relations, viou_threshold, with_ids=False):
    """Greedy matching.  visual_relation_detection.py:7-34 (``_v2`` :124-156 adds gt2det_ids).

    Stable sort by score descending; each prediction takes the not-yet-detected GT with the
    same triplet and the largest min(sub vIoU, obj vIoU), requiring ``>= thr`` and a strict
    improvement (first GT wins ties); a GT is consumed once.
    """
    order = sorted(pred_relations, key=lambda r: r['score'], reverse=True)
    taken = np.zeros((len(gt_relations),), dtype=bool)
    gt2det = np.ones((len(gt_relations),), dtype=int) * (-1)
    hit = np.ones((len(order))) * -np.inf
    for pi, pr in enumerate(order):
        best, best_k = -float('Inf'), -1
        for gi, gt in enumerate(gt_relations):
            if taken[gi] or tuple(pr['triplet']) != tuple(gt['triplet']):
                continue
            ov = min(viou(pr['sub_traj'], pr['duration'], gt['sub_traj'], gt['duration']),
                     viou(pr['obj_traj'], pr['duration'], gt['obj_traj'], gt['duration']))
            if ov >= viou_threshold and ov > best:
                best, best_k = ov, gi
        if best_k >= 0:
            hit[pi] = pr['score']
            taken[best_k] = True
            gt2det[best_k] = pi
    prec, rec = _pr_curves(hit, len(gt_relations))
    if with_ids:
        return prec, rec, hit, gt2det
    return prec, rec, hit


def tagging_scores(gt_relations, pred_relations):
    """visual_relation_detection.py:37-58: triplets deduplicated in score order."""
    order = sorted(pred_relations, key=lambda r: r['score'], reverse=True)
    gt_set = set(tuple(r['triplet']) for r in gt_relations)
    seen, sc = [], []
    for r in order:
        t = tuple(r['triplet'])
        if t not in seen:
            seen.append(t)
            sc.append(r['score'])
    sc = np.asarray(sc)
    for i, t in enumerate(seen):
        if t not in gt_set:
            sc[i] = -np.inf
    prec, rec = _pr_curves(sc, len(gt_set))
    return prec, rec, sc


def evaluate(groundtruth, prediction, viou_threshold=0.5, det_nreturns=(50, 100), tag_nreturns=(1, 5, 10),
             with_infos=False):
    """visual_relation_detection.py:61-117 (``evaluate_v2`` :160-223 when ``with_infos``)."""
    video_ap = dict()
    pool_sc, pool_tp, p_at = defaultdict(list), defaultdict(list), defaultdict(list)
    n_gt_total = 0
    infos = {}
    for vid, gts in groundtruth.items():
        if len(gts) == 0:
            continue
        n_gt_total += len(gts)
        preds = prediction.get(vid, [])
        prec, rec, sc, g2d = detection_scores(gts, preds, viou_threshold, with_ids=True)
        infos[vid] = (sc, g2d)
        video_ap[vid] = voc_ap(rec, prec)
        tp = np.isfinite(sc)
        for k in det_nreturns:
            c = min(k, sc.size)
            pool_sc[k].append(sc[:c])
            pool_tp[k].append(tp[:c])
        tprec, _, _ = tagging_scores(gts, preds)
        for k in tag_nreturns:
            c = min(k, tprec.size)
            p_at[k].append(tprec[c - 1] if c > 0 else 0.)
    mean_ap = np.mean(list(video_ap.values()))
    rec_at = dict()
    for k in det_nreturns:
        sc = np.concatenate(pool_sc[k])
        tp = np.concatenate(pool_tp[k])
        tp = tp[np.argsort(sc)[::-1]]
        ctp = np.cumsum(tp).astype(np.float32)
        rec_at[k] = (ctp / np.maximum(n_gt_total, F32_EPS))[-1]
    mprec = {k: np.mean(p_at[k]) for k in tag_nreturns}
    if with_infos:
        return mean_ap, rec_at, mprec, infos
    return mean_ap, rec_at, mprec
