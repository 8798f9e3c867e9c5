"""Batched drivers replacing the reference's ``batch_size=1`` evaluation loops:

  inference_then_eval   tools/eval_vidvrd.py:42-162   BIG-C -> (convert) -> eval_visual_relation
  evaluate_cls_stage    tools/eval_vidor.py:19-138    BIG-C (VidOR) -> results dict + metrics
  evaluate_combined     tools/eval_vidor.py:141-280   cls-stage results -> grounding -> expansion -> eval

They take in-memory datasets (lists of ``TrajProposal`` / ``VideoGraph`` / I3D tensors), run the whole list through the
batched kernels and evaluate on the packed path; ``want_dicts=True`` additionally returns the reference's
``{video_name: [relation dict, ...]}`` structure, and the ``save_*`` helpers write the reference's result files.

Videos WITHOUT predictions keep their ground truth, exactly like the reference loops: those only ``continue`` past the
prediction of a video with no proposals / no overlapping pair (tools/eval_vidvrd.py:126-128, tools/eval_vidor.py:100-103,
:227-229) and then evaluate against the complete GT file, so such a video scores AP = 0 and its GT relations stay in the
recall@K denominator and in the tagging means (visual_relation_detection.py:71-80).  Here a video without predictions
owns an empty segment of the packed prediction table.
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import evalapi, geometry
from .containers import TrajProposal
from .convert import EvalFmtCvtor


def _gt_relations(gt_graphs, device):
    gt_t = geometry.TrackTable.from_containers(gt_graphs, device=device)
    return evalapi.PackedRelations.from_gt_graphs(gt_t, gt_graphs)


def _device_of(model, proposals):
    dev = getattr(model, "device", None)
    if dev is not None:
        return dev
    for p in proposals:
        if p.num_proposals > 0:
            return p.device
    raise ValueError("no device: the model is not on a CUDA device and every video is empty")


def _empty_predictions(n_videos, device):
    z = lambda *s, dt=torch.long: torch.zeros(*s, dtype=dt, device=device)
    return evalapi.PackedRelations(z(0, 4, dt=torch.float32), z(1), z(0), z(0, 7), z(n_videos + 1), z(0, dt=torch.float64))


# ------------------------------------------------------------------------------------------------------
# staging of real (host) inputs: pinned buffers, asynchronous H2D on a copy stream
# ------------------------------------------------------------------------------------------------------
class StagedBatch(object):
    """Host ``TrajProposal``s of a batch packed field by field into pinned buffers and copied to the device with asynchronous
    H2D transfers on ``stream`` (the reference moves 2n+3 pageable tensors per video synchronously,
    dataloader_vidvrd.py:66-77).  ``proposals`` are device-side views (rows of the batch buffers); ``ready`` is recorded on
    ``stream`` after the last copy -- ``wait()`` makes the current stream wait for it."""

    FIELDS = (("features", torch.float32), ("bboxes", torch.float32), ("traj_durations", torch.long), ("cat_ids", torch.long),
              ("scores", torch.float32))

    def __init__(self, proposals: Sequence[TrajProposal], device, stream: Optional[torch.cuda.Stream] = None,
                 extra: Optional[Sequence[torch.Tensor]] = None):
        import copy
        device = torch.device(device)
        self.stream = stream if stream is not None else torch.cuda.Stream(device=device)
        live = [p for p in proposals if p.num_proposals > 0]
        self.nbytes = 0
        self._keep = []
        dev_buf = {}
        with torch.cuda.stream(self.stream):
            for name, dt in self.FIELDS:
                parts = [getattr(p, name) for p in live]
                if not parts or any(t is None for t in parts):
                    continue
                host_ids = [k for k, t in enumerate(parts) if not t.is_cuda]
                if host_ids:                                   # the host-resident pieces of this field: one pinned buffer, one async copy
                    hp = [parts[k].to(dt) for k in host_ids]
                    rows = sum(int(t.shape[0]) for t in hp)
                    host = torch.empty((rows,) + tuple(hp[0].shape[1:]), dtype=dt, pin_memory=True)
                    torch.cat(hp, 0, out=host)
                    dev = torch.empty(host.shape, dtype=dt, device=device)
                    dev.copy_(host, non_blocking=True)
                    self.nbytes += host.numel() * host.element_size()
                    self._keep.append(host)
                    r = 0
                    for k, t in zip(host_ids, hp):
                        parts[k] = dev[r:r + t.shape[0]]
                        r += t.shape[0]
                if len(host_ids) == len(parts):
                    dev_buf[name] = dev                        # every piece came from the host: the staged buffer IS the batch buffer
                else:
                    dev_buf[name] = torch.cat([t.to(dt) for t in parts], 0) if len(parts) > 1 else parts[0].to(dt)
            self.extra = None
            if extra is not None:
                self.extra = []
                for t in extra:
                    h = t if (t.is_cuda or t.is_pinned()) else t.pin_memory()
                    self.extra.append(h.to(device, non_blocking=True))
                    self.nbytes += 0 if t.is_cuda else t.numel() * t.element_size()
                    self._keep.append(h)
            self.ready = torch.cuda.Event()
            self.ready.record(self.stream)
        self.proposals: List[TrajProposal] = []
        r = n0 = 0
        for p in proposals:
            q = copy.copy(p)
            if p.num_proposals > 0:
                L, n = int(p.lengths.sum()), p.num_proposals
                q.bboxes = dev_buf["bboxes"][r:r + L]
                q.features = dev_buf["features"][r:r + L] if "features" in dev_buf else None
                q.traj_durations, q.cat_ids, q.scores = dev_buf["traj_durations"][n0:n0 + n], dev_buf["cat_ids"][n0:n0 + n], dev_buf["scores"][n0:n0 + n]
                r += L
                n0 += n
            self.proposals.append(q)

    def wait(self):
        torch.cuda.current_stream().wait_event(self.ready)
        return self.proposals


def stage(proposals, device, video_features=None):
    """Proposals (and optional per-video clip features) on ``device``; host inputs go through ``StagedBatch``."""
    if all(p.num_proposals == 0 or p.bboxes.is_cuda for p in proposals) and (video_features is None or all(v.is_cuda for v in video_features)):
        return list(proposals), video_features
    sb = StagedBatch(proposals, device, extra=video_features)
    return sb.wait(), sb.extra


# ------------------------------------------------------------------------------------------------------
# result writers (file names and payloads of the reference tools)
# ------------------------------------------------------------------------------------------------------
def save_infer_results(infer_result_for_save: Dict[str, Optional[list]], experiment_dir: str, dataset_type: str, save_tag: str = "") -> str:
    """``VidVRDtest_infer_result_<tag>.pkl`` (tools/eval_vidvrd.py:143-147) / ``VidORval_infer_results_<tag>.pkl``
    (tools/eval_vidor.py:124-129): ``{video_name: [quintuples, scores(n,3), spans, query_ids] | None}`` with CPU tensors."""
    name = "VidVRDtest_infer_result_{}.pkl" if dataset_type.lower() == "vidvrd" else "VidORval_infer_results_{}.pkl"
    path = os.path.join(experiment_dir, name.format(save_tag))
    with open(path, "wb") as f:
        pickle.dump(infer_result_for_save, f)
    return path


def save_predict_relations(predict_relations: Dict[str, list], experiment_dir: str, dataset_type: str, save_tag: str = "",
                           after_grounding: bool = False) -> str:
    """``<set>_predict_relations[_aft_grd]_<tag>.json`` (tools/eval_vidvrd.py:155-160, tools/eval_vidor.py:130-135, :272-277)."""
    prefix = "VidVRDtest" if dataset_type.lower() == "vidvrd" else "VidORval"
    name = "%s_predict_relations_%s{}.json" % (prefix, "aft_grd_" if after_grounding else "")
    path = os.path.join(experiment_dir, name.format(save_tag))
    with open(path, "w") as f:
        json.dump(predict_relations, f)
    return path


def save_hit_infos(hit_infos: dict, experiment_dir: str, save_tag: str = "") -> str:
    """``VidORval_hit_infos_aft_grd_<tag>.pkl`` (tools/eval_vidor.py:266-271)."""
    path = os.path.join(experiment_dir, "VidORval_hit_infos_aft_grd_{}.pkl".format(save_tag))
    with open(path, "wb") as f:
        pickle.dump(hit_infos, f)
    return path


# ------------------------------------------------------------------------------------------------------
def _classify(model, proposals, topk):
    """BIG-C over the videos that have proposals.  -> (live indices, their proposals, PackedTriplets or None)."""
    live = [i for i, p in enumerate(proposals) if p.num_proposals > 0]      # num_proposals == 0 -> no prediction (model_0v10.py:377-380)
    lp = [proposals[i] for i in live]
    packed = None
    if lp:
        with torch.no_grad():
            packed = model.forward_packed(lp, topk=topk)
    return live, lp, packed


def inference_then_eval(model, proposals: Sequence, gt_graphs: Sequence, topk: int = 10, viou_threshold: float = 0.5,
                        want_dicts: bool = False, dataset_type: str = "vidvrd"):
    """-> (mean_ap, rec_at_n, mprec_at_n[, predict_relations dict]).  Every video's GT is evaluated; videos with zero proposals
    or no overlapping pair contribute no predictions (and, like tools/eval_vidvrd.py:126-128, no entry in the dict)."""
    dev = _device_of(model, proposals)
    proposals, _ = stage(proposals, dev)
    V = len(proposals)
    live, lp, packed = _classify(model, proposals, topk)
    GT = _gt_relations(gt_graphs, dev)
    if lp:
        pt = geometry.TrackTable.from_containers(lp)
        PR = evalapi.PackedRelations.from_packed_triplets(pt, packed).spread(live, V)
    else:
        PR = _empty_predictions(V, dev)
    res = evalapi.evaluate_packed(PR, GT, viou_threshold)
    if not want_dicts:
        return res
    cv = EvalFmtCvtor(dataset_type)
    out = {}
    if lp:
        for p, t in zip(lp, packed.per_video()):
            if t is not None:
                out.update(cv.to_eval_format_pr(p, (t[0], t[1].mean(-1), t[2])))
    return res + (out,)


def evaluate_cls_stage(model, proposals: Sequence, gt_graphs: Sequence, topk: int = 3, viou_threshold: float = 0.5):
    """-> (metrics, infer_result_for_save) with ``infer_result_for_save[video_name] = [quintuples, scores(n,3), spans, query_ids]``
    or None (tools/eval_vidor.py:100-114), tensors on the CPU like the reference's pickle."""
    dev = _device_of(model, proposals)
    proposals, _ = stage(proposals, dev)
    V = len(proposals)
    live, lp, packed = _classify(model, proposals, topk)
    save = {p.video_name: None for p in proposals}
    GT = _gt_relations(gt_graphs, dev)
    if lp:
        for p, r in zip(lp, packed.per_video()):
            save[p.video_name] = None if r is None else [x.cpu() for x in r]
        pt = geometry.TrackTable.from_containers(lp)
        PR = evalapi.PackedRelations.from_packed_triplets(pt, packed).spread(live, V)
    else:
        PR = _empty_predictions(V, dev)
    return evalapi.evaluate_packed(PR, GT, viou_threshold), save


def evaluate_combined(grd_model, cls_model, proposals: Sequence, video_features: Sequence[torch.Tensor], gt_graphs: Sequence,
                      topk: int = 3, score_th=0.9, tiou_th=0.5, bins_th=0.2, nms_th=0.8, viou_threshold: float = 0.5):
    """Classification + grounding + evaluation, all videos batched.  -> (mean_ap, rec_at_n, mprec_at_n, hit_infos) with
    ``hit_infos[video_name] = (hit_scores, gt2det_ids)`` for every video that has GT (evaluate_v2, tools/eval_vidor.py:259-264)."""
    dev = _device_of(cls_model, proposals)
    proposals, video_features = stage(proposals, dev, video_features)
    V = len(proposals)
    live, lp, packed = _classify(cls_model, proposals, topk)
    GT = _gt_relations(gt_graphs, dev)
    PR = _empty_predictions(V, dev)
    if lp:
        q, s3, sp, _, off = packed.compact()
        with_rows = [k for k in range(len(lp)) if off[k + 1] > off[k]]      # videos with >= 1 classified triplet go through grounding
        if with_rows:
            datas = [(q[off[k]:off[k + 1]], sp[off[k]:off[k + 1]], lp[k].video_len) for k in with_rows]
            with torch.no_grad():
                pooled, probs, mask = grd_model.forward_packed([video_features[live[k]] for k in with_rows], datas, score_th=score_th,
                                                               tiou_th=tiou_th, bins_th=bins_th, nms_th=nms_th)
            pt = geometry.TrackTable.from_containers(lp)
            PR = evalapi.PackedRelations.from_grounded(pt, packed, pooled, probs, mask, [p.video_len for p in lp]).spread(live, V)
    m_ap, rec, mprec, infos = evalapi.evaluate_packed(PR, GT, viou_threshold, with_infos=True)
    return m_ap, rec, mprec, {gt_graphs[k].video_name: v for k, v in infos.items()}
