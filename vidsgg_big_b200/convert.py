"""Tensors -> VidVRD-helper relation dicts: mirror of ``utils/evaluate.py`` ``EvalFmtCvtor``.

Only needed when the dict / JSON format itself is wanted (result files, third-party evaluators): the evaluation fast path
(``evalapi.PackedRelations``) never materialises these lists.  One D2H read per tensor and numpy slicing, no per-frame loops.

Category tables are plain ``{id: name}`` dicts or sequences supplied by the caller (the reference hard-wires
``utils/categories_v2.py``); without them ids are rendered as ``"e<id>"`` / ``"p<id>"`` -- the evaluation only compares names
for equality, so any injective naming yields identical metrics.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch


class _Names(object):
    def __init__(self, table, prefix):
        self.table, self.prefix = table, prefix

    def __getitem__(self, i):
        if self.table is None:
            return "%s%d" % (self.prefix, int(i))
        return self.table[int(i)]


class EvalFmtCvtor(object):
    def __init__(self, dataset_type: str, enti_id2name=None, pred_id2name=None):
        self.dataset_type = dataset_type.lower()
        assert self.dataset_type in ("vidvrd", "vidor")
        self.entiId2Name = _Names(enti_id2name, "e")
        self.predId2Name = _Names(pred_id2name, "p")

    def _reset_video_name(self, video_name):
        """utils/evaluate.py:26-38: VidOR names ``<group>_<id>`` are reduced to ``<id>``."""
        if self.dataset_type == "vidor":
            parts = video_name.split('_')
            if len(parts) == 2:
                return parts[1]
        return video_name

    def to_eval_format_pr(self, proposal, pr_triplet, preserve_debug_info=False, use_pku=False):
        """utils/evaluate.py:72-151.  ``pr_triplet`` = (quintuples i64[m,5], score f[m], span i64[m,2] closed) or None.
        Durations become half-open ``[s, e+1)``, trajectories are cut to the span (``traj_cutoff``), background rows dropped."""
        video_name = self._reset_video_name(proposal.video_name)
        if pr_triplet is None:
            return {video_name: []}
        quint = pr_triplet[0].cpu().numpy()
        score = pr_triplet[1].cpu().numpy()
        span = pr_triplet[2].cpu().numpy()
        boxes = proposal.bboxes.cpu().numpy()
        dura = proposal.traj_durations.cpu().numpy()
        off = np.concatenate([[0], np.cumsum(proposal.lengths.numpy())])
        out = []
        for i in range(quint.shape[0]):
            pc, sc, oc, st, ot = (int(x) for x in quint[i])
            if pc == 0:
                continue
            s, e = int(span[i, 0]), int(span[i, 1]) + 1
            inside = dura[st, 0] <= s and e <= dura[st, 1] + 1 and dura[ot, 0] <= s and e <= dura[ot, 1] + 1
            assert inside, "span outside the subject / object track (%s)" % video_name
            sub = boxes[off[st] + s - dura[st, 0]: off[st] + e - dura[st, 0]]
            obj = boxes[off[ot] + s - dura[ot, 0]: off[ot] + e - dura[ot, 0]]
            rel = {"triplet": [self.entiId2Name[sc], self.predId2Name[pc], self.entiId2Name[oc]], "duration": (s, e),
                   "score": float(score[i]), "sub_traj": sub.tolist(), "obj_traj": obj.tolist()}
            if preserve_debug_info:
                rel["triplet_tid"] = (st, pc, ot)
                rel["debug_dura"] = {"s": (int(dura[st, 0]), int(dura[st, 1]) + 1), "o": (int(dura[ot, 0]), int(dura[ot, 1]) + 1), "inter": (s, e)}
            out.append(rel)
        return {video_name: out}

    def to_eval_format_gt(self, gt_graph):
        """GT relations of one video in the dict format (utils/evaluate.py:234-300; VidVRD-helper/dataset/dataset.py:173-208)."""
        video_name = self._reset_video_name(gt_graph.video_name)
        so = torch.argmax(gt_graph.adj_matrix, dim=-1).t().cpu().numpy()
        boxes = gt_graph.bboxes.cpu().numpy()
        dura = gt_graph.traj_durations.cpu().numpy()
        cats = gt_graph.traj_cat_ids.cpu().numpy()
        pcat = gt_graph.pred_cat_ids.cpu().numpy()
        pdur = gt_graph.pred_durations.long().cpu().numpy()
        off = np.concatenate([[0], np.cumsum(gt_graph.lengths.numpy())])
        out = []
        for i in range(gt_graph.num_preds):
            st, ot = int(so[i, 0]), int(so[i, 1])
            s, e = int(pdur[i, 0]), int(pdur[i, 1]) + 1
            sub = boxes[off[st] + s - dura[st, 0]: off[st] + e - dura[st, 0]]
            obj = boxes[off[ot] + s - dura[ot, 0]: off[ot] + e - dura[ot, 0]]
            out.append({"triplet": [self.entiId2Name[cats[st]], self.predId2Name[pcat[i]], self.entiId2Name[cats[ot]]],
                        "subject_tid": st, "object_tid": ot, "duration": (s, e),
                        "sub_traj": sub.astype(np.int64).tolist(), "obj_traj": obj.astype(np.int64).tolist()})
        return {video_name: out}
