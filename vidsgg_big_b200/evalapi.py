"""Host-side mirror of ``VidVRDhelperEvalAPIs`` for visual relation detection (SURVEY.md §8b "Eval API"),
with the per-video vIoU + greedy matching done for ALL videos by one batched launch chain of
csrc/relmatch.cu.  No CPU fallback: the matching needs the CUDA library.

Mirrored reference API (same names / arguments / return values):
  viou(traj_1, duration_1, traj_2, duration_2)                     common.py:65-106
  voc_ap(rec, prec)                                                common.py:4-37
  eval_detection_scores(gt, pred, thr) / eval_detection_scores_v2  visual_relation_detection.py:7-34 / :124-156
  eval_tagging_scores(gt, pred)                                    visual_relation_detection.py:37-58
  evaluate == eval_visual_relation(groundtruth, prediction, ...)   visual_relation_detection.py:61-117
  evaluate_v2(...) -> (+det_infos)                                 visual_relation_detection.py:160-223
  eval_relation_with_gt(dataset_type, logger, prediction_results, json_results_path, return_hit_infos)  :226-265

Two input paths feed the same kernels:
  * dict path  -- the reference's dict-of-lists format (README.md:7-47); every relation's two box lists become
                  private f64 tracks, so arbitrary Python floats are honoured;
  * packed path -- ``PackedRelations.from_triplets`` / ``.from_gt_graphs``: relations index the proposal / GT track
                  tables already resident in HBM (f32), nothing is expanded.
AP / recall / tagging aggregation (row A14) stays on the host in numpy, verbatim in semantics.
"""
from __future__ import annotations

import ctypes as C
import json
from collections import defaultdict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _cabi
from ._cabi import VsgRelTable, check, lib, ptr, stream_ptr

F32_EPS = np.finfo(np.float32).eps


# ----------------------------------------------------------------------------------------------
class PackedRelations(object):
    """Relations of a batch of videos over a CSR track table (see include/vsg_b200.h ``VsgRelTable``)."""

    def __init__(self, boxes, off, tstart, rel, vid_off, scores=None, vol_full_track=False, triplets_host=None):
        self.boxes, self.off, self.tstart, self.rel, self.vid_off = boxes, off, tstart, rel, vid_off
        self.scores = scores
        self.vol_full_track = bool(vol_full_track)
        self.n_rel = int(rel.shape[0])
        self.n_vid = int(vid_off.shape[0]) - 1
        self._vid_off_h = None
        self._keep = (boxes, off, tstart, rel, vid_off, scores)
        self.triplets_host = triplets_host

    @property
    def vid_off_host(self) -> np.ndarray:
        if self._vid_off_h is None:
            self._vid_off_h = self.vid_off.cpu().numpy()
        return self._vid_off_h

    def spread(self, video_ids: Sequence[int], n_videos: int) -> "PackedRelations":
        """The same relations as segments of a LARGER video list: segment k of this table becomes video ``video_ids[k]``
        (ascending) of ``n_videos``; every other video gets an empty segment.  This is how videos without predictions stay in
        the evaluation with their ground truth (the reference only skips their predictions, tools/eval_vidor.py:100-103)."""
        ids = np.asarray(list(video_ids), dtype=np.int64)
        assert ids.size == self.n_vid and (ids.size == 0 or (np.all(np.diff(ids) > 0) and ids[0] >= 0 and ids[-1] < n_videos))
        if ids.size == n_videos:
            return self
        per = np.zeros(n_videos, np.int64)
        per[ids] = np.diff(self.vid_off_host)
        off = np.concatenate([[0], np.cumsum(per)]).astype(np.int64)
        out = PackedRelations(self.boxes, self.off, self.tstart, self.rel, torch.from_numpy(off).to(self.rel.device), self.scores,
                              self.vol_full_track, self.triplets_host)
        out._vid_off_h = off
        return out

    def table(self) -> VsgRelTable:
        t = VsgRelTable()
        t.boxes, t.off, t.tstart = self.boxes.data_ptr(), self.off.data_ptr(), self.tstart.data_ptr()
        t.rel, t.vid_off = self.rel.data_ptr(), self.vid_off.data_ptr()
        t.n_rel = self.n_rel
        t.box_f64 = 1 if self.boxes.dtype == torch.float64 else 0
        t.vol_full_track = 1 if self.vol_full_track else 0
        return t

    # ---- dict path ---------------------------------------------------------------------------
    @classmethod
    def from_dicts(cls, per_video: Sequence[List[dict]], vocab: Dict[tuple, int], device, with_scores: bool,
                   candidates: Optional[Sequence[set]] = None):
        """``candidates[v]`` (optional): the triplets that occur in video v's ground truth.  A prediction with any other
        triplet can never be matched (its vIoU is never evaluated by the reference either, visual_relation_detection.py:16-17),
        so its box lists are not converted -- it keeps a zero-length track and only takes part in the ranking."""
        rows, boxes, lens, tstart, scores, vid_off = [], [], [], [], [], [0]
        trk = 0
        empty = np.zeros((0, 4), np.float64)
        for v, rels in enumerate(per_video):
            cand = None if candidates is None else candidates[v]
            for r in rels:
                t = tuple(r["triplet"])
                tid = vocab.setdefault(t, len(vocab))
                s, e = int(r["duration"][0]), int(r["duration"][1])
                for key in ("sub_traj", "obj_traj"):
                    b = np.array(r[key], dtype=np.float64).reshape(-1, 4) if (cand is None or t in cand) else empty
                    if b is not empty and b.shape[0] < e - s:
                        # the kernels address rows by duration alone; the reference fails here too (common.py:88-90)
                        raise IndexError("relation %s of video %d: %s has %d boxes for a duration of %d frames"
                                         % (t, v, key, b.shape[0], e - s))
                    boxes.append(b)
                    lens.append(b.shape[0])
                    tstart.append(s)
                # the whole triplet is one vocabulary id (s_cat); p_cat / o_cat constant
                rows.append((tid, 0, 0, trk, trk + 1, s, e))
                trk += 2
                if with_scores:
                    scores.append(float(r["score"]))
            vid_off.append(len(rows))
        off = np.zeros(len(lens) + 1, np.int64)
        off[1:] = np.cumsum(np.asarray(lens, np.int64)) if lens else 0
        bx = np.concatenate(boxes, 0) if boxes else np.zeros((0, 4), np.float64)
        dev = torch.device(device)
        to = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=dt))).to(dev)
        return cls(to(bx, np.float64), to(off, np.int64), to(np.asarray(tstart, np.int64).reshape(-1), np.int64),
                   to(np.asarray(rows, np.int64).reshape(-1, 7), np.int64), to(vid_off, np.int64),
                   to(scores, np.float64) if with_scores else None, vol_full_track=True)

    # ---- packed path -------------------------------------------------------------------------
    @classmethod
    def from_triplets(cls, table, triplets: Sequence[Optional[tuple]]):
        """``table``: geometry.TrackTable of the batch's proposals; ``triplets[v]`` = (quintuples i64[m,5] =
        [pred,scat,ocat,sid,oid], score f[m], span i64[m,2] closed) or None, as returned by ``BIG_C.forward``
        (scores already reduced to one value per row, tools/eval_vidvrd.py:136)."""
        dev = table.boxes.device
        rows, scs, vid_off = [], [], [0]
        for v, t in enumerate(triplets):
            if t is not None and t[0].shape[0] > 0:
                q, s, sp = t[0].to(dev), t[1].to(dev), t[2].to(dev)
                base = int(table.seg_host[v]) if hasattr(table, "seg_host") else int(sum(table.counts[:v]))
                rows.append(torch.stack([q[:, 1], q[:, 0], q[:, 2], q[:, 3] + base, q[:, 4] + base, sp[:, 0], sp[:, 1] + 1], 1))
                scs.append(s.double())
                vid_off.append(vid_off[-1] + int(q.shape[0]))
            else:
                vid_off.append(vid_off[-1])
        rel = torch.cat(rows, 0).contiguous() if rows else torch.zeros(0, 7, dtype=torch.long, device=dev)
        scores = torch.cat(scs, 0).contiguous() if scs else torch.zeros(0, dtype=torch.float64, device=dev)
        return cls(table.boxes, table.off, table.dura[:, 0].contiguous(), rel,
                   torch.tensor(vid_off, dtype=torch.long).to(dev), scores, vol_full_track=False)

    @classmethod
    def from_packed_triplets(cls, table, packed, seg_host=None):
        """Batched twin of ``from_triplets`` for ``BIG_C.forward_packed`` output (bigc.PackedTriplets): a handful of
        device ops for the whole batch.  Score = mean of the three scores (tools/eval_vidvrd.py:136)."""
        dev = table.boxes.device
        q, s3, sp, _, off = packed.compact()
        counts = np.diff(off)
        base = np.concatenate([[0], np.cumsum(table.counts)])[:-1].astype(np.int64)
        base_rows = torch.from_numpy(np.repeat(base, counts)).to(dev)
        rel = torch.stack([q[:, 1], q[:, 0], q[:, 2], q[:, 3] + base_rows, q[:, 4] + base_rows, sp[:, 0], sp[:, 1] + 1], 1).contiguous()
        return cls(table.boxes, table.off, table.dura[:, 0].contiguous(), rel, torch.from_numpy(off).to(dev),
                   s3.mean(-1).double().contiguous(), vol_full_track=False)

    @classmethod
    def from_grounded(cls, table, packed, pooled_se, bins_probs, bins_mask, video_lens):
        """Classification output (bigc.PackedTriplets) + grounding output of the same queries (concatenated in video
        order) -> relations, i.e. the expansion of tools/eval_vidor.py:245-253 for the whole batch:
        score = mean(cls scores) * bin prob, span = round(pooled * video_len), one relation per kept bin."""
        dev = table.boxes.device
        q, s3, sp, _, off = packed.compact()
        counts = np.diff(off)
        V = counts.size
        base = np.concatenate([[0], np.cumsum(table.counts)])[:-1].astype(np.int64)
        base_rows = torch.from_numpy(np.repeat(base, counts)).to(dev)
        vid_rows = torch.from_numpy(np.repeat(np.arange(V, dtype=np.int64), counts)).to(dev)
        vlen = torch.as_tensor(np.asarray(video_lens, dtype=np.float32)).to(dev)[vid_rows]            # per query
        qi, bi = bins_mask.nonzero(as_tuple=True)                                                       # row-major, like [mask]
        score = (s3.mean(-1)[:, None] * bins_probs)[qi, bi]
        span = torch.round(pooled_se[qi, bi, :] * vlen[qi, None]).long()
        rel = torch.stack([q[qi, 1], q[qi, 0], q[qi, 2], q[qi, 3] + base_rows[qi], q[qi, 4] + base_rows[qi], span[:, 0], span[:, 1] + 1], 1)
        per_vid = torch.bincount(vid_rows[qi], minlength=V).cpu().numpy()
        vid_off = np.concatenate([[0], np.cumsum(per_vid)]).astype(np.int64)
        return cls(table.boxes, table.off, table.dura[:, 0].contiguous(), rel.contiguous(), torch.from_numpy(vid_off).to(dev),
                   score.double().contiguous(), vol_full_track=False)

    @classmethod
    def from_gt_graphs(cls, table, graphs: Sequence):
        """``table``: TrackTable of the batch's GT tracks; relation = (traj cats, pred cat, closed pred span)."""
        dev = table.boxes.device
        rows, vid_off = [], [0]
        base = 0
        for v, g in enumerate(graphs):
            if g.num_preds > 0:
                so = torch.argmax(g.adj_matrix, dim=-1).t().to(dev)
                cats = g.traj_cat_ids.to(dev)[so]
                pd = g.pred_durations.to(dev).long()
                rows.append(torch.stack([cats[:, 0], g.pred_cat_ids.to(dev), cats[:, 1], so[:, 0] + base, so[:, 1] + base,
                                         pd[:, 0], pd[:, 1] + 1], 1))
            vid_off.append(vid_off[-1] + g.num_preds)
            base += table.counts[v]
        rel = torch.cat(rows, 0).contiguous() if rows else torch.zeros(0, 7, dtype=torch.long, device=dev)
        return cls(table.boxes, table.off, table.dura[:, 0].contiguous(), rel,
                   torch.tensor(vid_off, dtype=torch.long).to(dev), None, vol_full_track=False)


class MatchResult(object):
    def __init__(self, order, hit, gt2det, ov, ov_off):
        self.order, self.hit, self.gt2det, self.ov, self.ov_off = order, hit, gt2det, ov, ov_off


def match_relations(pred: PackedRelations, gt: PackedRelations, viou_threshold: float, keep_ov: bool = False) -> MatchResult:
    """Batched greedy matching of every video (device tensors out; nothing synchronises)."""
    assert pred.n_vid == gt.n_vid
    dev = gt.rel.device
    V = gt.n_vid
    po, go = pred.vid_off_host, gt.vid_off_host
    ov_off_h = np.zeros(V + 1, np.int64)
    ov_off_h[1:] = np.cumsum((po[1:] - po[:-1]) * (go[1:] - go[:-1]))
    ov_off = torch.from_numpy(ov_off_h).to(dev)
    npred, ngt = pred.n_rel, gt.n_rel
    order = torch.empty(max(npred, 1), dtype=torch.int32, device=dev)
    hit = torch.empty(max(npred, 1), dtype=torch.float64, device=dev)
    gt2det = torch.empty(max(ngt, 1), dtype=torch.int32, device=dev)
    ov = torch.empty(max(int(ov_off_h[-1]), 1), dtype=torch.float64, device=dev)
    vol_p = torch.empty(max(2 * npred, 1), dtype=torch.float64, device=dev)
    vol_g = torch.empty(max(2 * ngt, 1), dtype=torch.float64, device=dev)
    taken = torch.empty(max(ngt + npred, 1), dtype=torch.uint8, device=dev)     # GT-taken flags + per-prediction candidate flags
    scores = pred.scores if pred.scores is not None else torch.zeros(max(npred, 1), dtype=torch.float64, device=dev)
    tp, tg = pred.table(), gt.table()
    from .linalg import _Profile
    with _Profile.span("rel_match"):
        check(lib().vsg_rel_viou_match(C.byref(tp), ptr(scores), C.byref(tg), V, ptr(ov_off), float(viou_threshold),
                                       ptr(order), ptr(ov), ptr(hit), ptr(gt2det), ptr(vol_p), ptr(vol_g), ptr(taken),
                                       stream_ptr(dev)), "vsg_rel_viou_match")
    return MatchResult(order[:npred], hit[:npred], gt2det[:ngt], ov if keep_ov else None, ov_off_h)


# ----------------------------------------------------------------------------------------------
# reference-compatible functions
# ----------------------------------------------------------------------------------------------
def _device():
    if not torch.cuda.is_available():
        raise _cabi.VsgError("vidsgg_big_b200.evalapi needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def viou(traj_1, duration_1, traj_2, duration_2) -> float:
    """common.py:65-106 on the GPU (f64)."""
    dev = _device()
    b1 = torch.from_numpy(np.asarray(traj_1, np.float64).reshape(-1, 4)).to(dev)
    b2 = torch.from_numpy(np.asarray(traj_2, np.float64).reshape(-1, 4)).to(dev)
    off1 = torch.tensor([0, b1.shape[0]], dtype=torch.long, device=dev)
    off2 = torch.tensor([0, b2.shape[0]], dtype=torch.long, device=dev)
    d1 = torch.tensor([[int(duration_1[0]), int(duration_1[1])]], dtype=torch.long, device=dev)
    d2 = torch.tensor([[int(duration_2[0]), int(duration_2[1])]], dtype=torch.long, device=dev)
    out = torch.empty(1, dtype=torch.float64, device=dev)
    check(lib().vsg_viou_pairs_f64(ptr(b1), ptr(off1), ptr(d1), ptr(b2), ptr(off2), ptr(d2), 1, ptr(out), stream_ptr(dev)),
          "vsg_viou_pairs_f64")
    return float(out.item())


def voc_ap(rec, prec, use_07_metric=False):
    """common.py:4-37 (host numpy; aggregation is not on the device path)."""
    if use_07_metric:
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            p = 0 if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
            ap = ap + p / 11.
        return ap
    r = np.concatenate(([0.], rec, [1.]))
    p = np.concatenate(([0.], prec, [0.]))
    p = np.maximum.accumulate(p[::-1])[::-1]
    idx = np.where(r[1:] != r[:-1])[0]
    return np.sum((r[idx + 1] - r[idx]) * p[idx + 1])


def _pr_curves(hit_scores, n_gt):
    tp = np.isfinite(hit_scores)
    ctp = np.cumsum(tp).astype(np.float32)
    cfp = np.cumsum(~tp).astype(np.float32)
    return ctp / np.maximum(ctp + cfp, F32_EPS), ctp / np.maximum(n_gt, F32_EPS)


def _match_dicts(gt_lists, pred_lists, thr):
    dev = _device()
    vocab: Dict[tuple, int] = {}
    g = PackedRelations.from_dicts(gt_lists, vocab, dev, with_scores=False)
    cands = [set(tuple(r["triplet"]) for r in rels) for rels in gt_lists]
    p = PackedRelations.from_dicts(pred_lists, vocab, dev, with_scores=True, candidates=cands)
    m = match_relations(p, g, thr)
    return m.hit.cpu().numpy(), m.gt2det.cpu().numpy().astype(int), p.vid_off_host, g.vid_off_host


def eval_detection_scores_v2(gt_relations, pred_relations, viou_threshold):
    hit, g2d, _, _ = _match_dicts([gt_relations], [pred_relations], viou_threshold)
    prec, rec = _pr_curves(hit, len(gt_relations))
    return prec, rec, hit, g2d


def eval_detection_scores(gt_relations, pred_relations, viou_threshold):
    return eval_detection_scores_v2(gt_relations, pred_relations, viou_threshold)[:3]


def _tagging_from_ids(gt_trip, pred_trip_sorted, pred_scores_sorted):
    """visual_relation_detection.py:37-58 on hashable triplet keys."""
    gt_set = set(gt_trip)
    seen, sc = {}, []
    for t, s in zip(pred_trip_sorted, pred_scores_sorted):
        if t not in seen:
            seen[t] = len(sc)
            sc.append(s)
    sc = np.asarray(sc, dtype=np.float64) if sc else np.asarray([], dtype=np.float64)
    for t, i in seen.items():
        if t not in gt_set:
            sc[i] = -np.inf
    prec, rec = _pr_curves(sc, len(gt_set))
    return prec, rec, sc


def eval_tagging_scores(gt_relations, pred_relations):
    order = sorted(pred_relations, key=lambda x: x['score'], reverse=True)
    return _tagging_from_ids([tuple(r['triplet']) for r in gt_relations], [tuple(r['triplet']) for r in order],
                             [r['score'] for r in order])


def _aggregate(vids, n_gt_per_vid, hits, tag_precs, det_nreturns, tag_nreturns):
    """visual_relation_detection.py:94-109."""
    video_ap, pool_sc, pool_tp, p_at = dict(), defaultdict(list), defaultdict(list), defaultdict(list)
    total_gt = 0
    for vid, n_gt, sc, tprec in zip(vids, n_gt_per_vid, hits, tag_precs):
        total_gt += n_gt
        prec, rec = _pr_curves(sc, n_gt)
        video_ap[vid] = voc_ap(rec, prec)
        tp = np.isfinite(sc)
        for k in det_nreturns:
            c = min(k, sc.size)
            pool_sc[k].append(sc[:c])
            pool_tp[k].append(tp[:c])
        for k in tag_nreturns:
            c = min(k, tprec.size)
            p_at[k].append(tprec[c - 1] if c > 0 else 0.)
    mean_ap = np.mean(list(video_ap.values()))
    rec_at = dict()
    for k in det_nreturns:
        sc = np.concatenate(pool_sc[k])
        tp = np.concatenate(pool_tp[k])[np.argsort(sc)[::-1]]
        ctp = np.cumsum(tp).astype(np.float32)
        # no prediction in ANY video: the reference indexes an empty array here (IndexError, visual_relation_detection.py:105);
        # a batched driver that keeps prediction-less videos reports the recall that situation means
        rec_at[k] = (ctp / np.maximum(total_gt, F32_EPS))[-1] if ctp.size else np.float32(0.0)
    return mean_ap, rec_at, {k: np.mean(p_at[k]) for k in tag_nreturns}


def per_video_records(vids, n_gt_per_vid, hits, tag_precs, det_nreturns=(50, 100), tag_nreturns=(1, 5, 10)) -> np.ndarray:
    """Fixed-width per-video records ``[vid, AP, n_gt, tp@K..., P@K...]`` (f64): all that the final aggregation of
    visual_relation_detection.py:94-109 needs, so ranks only exchange 64 bytes per video."""
    rows = []
    for vid, n_gt, sc, tprec in zip(vids, n_gt_per_vid, hits, tag_precs):
        prec, rec = _pr_curves(sc, n_gt)
        tp = np.isfinite(sc)
        row = [float(vid), float(voc_ap(rec, prec)), float(n_gt)]
        row += [float(tp[:min(k, sc.size)].sum()) for k in det_nreturns]
        row += [float(tprec[min(k, tprec.size) - 1]) if min(k, tprec.size) > 0 else 0. for k in tag_nreturns]
        rows.append(row)
    return np.asarray(rows, dtype=np.float64).reshape(-1, 3 + len(det_nreturns) + len(tag_nreturns))


def metrics_from_records(records: np.ndarray, det_nreturns=(50, 100), tag_nreturns=(1, 5, 10)):
    """Finish mAP / recall@K / P@K from gathered records; equals ``_aggregate`` (recall@K = sum TP@K / sum GT in float32)."""
    mean_ap = np.mean(records[:, 1])
    total_gt = int(records[:, 2].sum())
    rec_at = {}
    for i, k in enumerate(det_nreturns):
        rec_at[k] = np.float32(records[:, 3 + i].sum()) / np.maximum(total_gt, F32_EPS)
    nd = len(det_nreturns)
    mprec = {k: np.mean(records[:, 3 + nd + i].astype(np.float32) if False else records[:, 3 + nd + i]) for i, k in enumerate(tag_nreturns)}
    return mean_ap, rec_at, mprec


def evaluate_v2(groundtruth, prediction, viou_threshold=0.5, det_nreturns=[50, 100], tag_nreturns=[1, 5, 10]):
    """visual_relation_detection.py:160-223; all videos matched by one batched device call."""
    vids = [v for v, g in groundtruth.items() if len(g) > 0]        # videos without GT are skipped (:73-74)
    gt_lists = [groundtruth[v] for v in vids]
    pr_lists = [prediction.get(v, []) for v in vids]
    hit, g2d, po, go = _match_dicts(gt_lists, pr_lists, viou_threshold)
    hits, tags, infos = [], [], {}
    for i, v in enumerate(vids):
        h = hit[po[i]:po[i + 1]]
        hits.append(h)
        infos[v] = (h, g2d[go[i]:go[i + 1]])
        tags.append(eval_tagging_scores(gt_lists[i], pr_lists[i])[0])
    mean_ap, rec_at, mprec = _aggregate(vids, [len(g) for g in gt_lists], hits, tags, det_nreturns, tag_nreturns)
    return mean_ap, rec_at, mprec, infos


def evaluate(groundtruth, prediction, viou_threshold=0.5, det_nreturns=[50, 100], tag_nreturns=[1, 5, 10]):
    """``eval_visual_relation`` (visual_relation_detection.py:61-117)."""
    return evaluate_v2(groundtruth, prediction, viou_threshold, det_nreturns, tag_nreturns)[:3]


eval_visual_relation = evaluate


def evaluate_packed(pred: PackedRelations, gt: PackedRelations, viou_threshold=0.5, det_nreturns=(50, 100),
                    tag_nreturns=(1, 5, 10), with_infos=False, want_records=False):
    """Packed fast path: same numbers as ``evaluate`` on the equivalent dicts, no dict materialisation."""
    m = match_relations(pred, gt, viou_threshold)
    if want_records and not with_infos:
        return _records_native(pred, gt, m, det_nreturns, tag_nreturns)
    hit = m.hit.cpu().numpy()
    order = m.order.cpu().numpy()
    g2d = m.gt2det.cpu().numpy().astype(int)
    prel = pred.rel[:, :3].cpu().numpy()
    grel = gt.rel[:, :3].cpu().numpy()
    psc = pred.scores.cpu().numpy()
    po, go = pred.vid_off_host, gt.vid_off_host
    vids, ngt, hits, tags, infos = [], [], [], [], {}
    for v in range(gt.n_vid):
        n_g = int(go[v + 1] - go[v])
        if n_g == 0:
            continue
        o = order[po[v]:po[v + 1]]
        trip = [tuple(t) for t in prel[po[v]:po[v + 1]][o].tolist()]
        sc = psc[po[v]:po[v + 1]][o].tolist()
        vids.append(v); ngt.append(n_g)
        hits.append(hit[po[v]:po[v + 1]])
        tags.append(_tagging_from_ids([tuple(t) for t in grel[go[v]:go[v + 1]].tolist()], trip, sc)[0])
        infos[v] = (hits[-1], g2d[go[v]:go[v + 1]])
    if want_records:
        return per_video_records(vids, ngt, hits, tags, det_nreturns, tag_nreturns)
    res = _aggregate(vids, ngt, hits, tags, list(det_nreturns), list(tag_nreturns))
    return res + (infos,) if with_infos else res


def _records_native(pred: PackedRelations, gt: PackedRelations, m: MatchResult, det_nreturns, tag_nreturns) -> np.ndarray:
    """Per-video records through the native host pass (csrc/evalhost.cu): three D2H reads, no Python loop over videos."""
    hit = np.ascontiguousarray(m.hit.cpu().numpy())
    order = np.ascontiguousarray(m.order.cpu().numpy())
    ptrip = np.ascontiguousarray(pred.rel[:, :3].cpu().numpy())
    if getattr(gt, "_trip_host", None) is None:
        gt._trip_host = np.ascontiguousarray(gt.rel[:, :3].cpu().numpy())
    gtrip = gt._trip_host
    po = np.ascontiguousarray(pred.vid_off_host.astype(np.int64))
    go = np.ascontiguousarray(gt.vid_off_host.astype(np.int64))
    det = np.asarray(list(det_nreturns), dtype=np.int32)
    tag = np.asarray(list(tag_nreturns), dtype=np.int32)
    rec = np.zeros((gt.n_vid, 3 + det.size + tag.size), dtype=np.float64)
    hp = lambda a: C.c_void_p(a.ctypes.data)
    n = lib().vsg_eval_records_host(hp(hit), hp(order), hp(ptrip), hp(po), hp(gtrip), hp(go), gt.n_vid, hp(det), int(det.size),
                                    hp(tag), int(tag.size), hp(rec))
    if n < 0:
        check(n, "vsg_eval_records_host")
    return rec[:n]


def eval_relation_with_gt(dataset_type, logger=None, prediction_results=None, json_results_path=None,
                          return_hit_infos=False, gt_relations_path=None):
    """visual_relation_detection.py:226-265.  ``gt_relations_path`` (extra, optional) overrides the reference's
    hard-coded relative paths (:247-253)."""
    print_func = print if logger is None else logger.info
    if prediction_results is None:
        print_func("loading json results from {}".format(json_results_path))
        with open(json_results_path, 'r') as f:
            prediction_results = json.load(f)
        print_func("Done.")
    else:
        assert json_results_path is None
    if gt_relations_path is None:
        gt_relations_path = ("datasets/GT_json_for_eval/VidVRDtest_gts.json" if dataset_type.lower() == "vidvrd"
                             else "datasets/GT_json_for_eval/VidORval_gts.json")
    with open(gt_relations_path, 'r') as f:
        gt_relations = json.load(f)
    print_func('Computing average precision AP over {} videos...'.format(len(gt_relations)))
    mean_ap, rec_at_n, mprec_at_n, hit_infos = evaluate_v2(gt_relations, prediction_results, viou_threshold=0.5)
    print_func('detection mean AP (used in challenge): {}'.format(mean_ap))
    print_func('detection recall: {}'.format(rec_at_n))
    print_func('tagging precision: {}'.format(mprec_at_n))
    if return_hit_infos:
        return hit_infos
