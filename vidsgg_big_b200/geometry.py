"""Host-side mirror of the reference geometry helpers (SURVEY.md §8b "Geometry helpers"), backed by
the sm_100a kernels of csrc/geometry.cu through the C ABI.  CUDA tensors only -- no CPU fallback.

Reference functions mirrored (same names, argument meaning and results):
  dura_intersection_ts   utils/utils_func.py:347-373
  vIoU_ts                utils/utils_func.py:437-471
  trajid2pairid          models/model_pairwise_baseline.py:104-111 / tools/train_vidor.py:73-78
  tIoU, generalized_tIoU utils/utils_func.py:375-410 (models/grd_model_v5.py:18-33)
  unique_with_idx_nd     utils/utils_func.py:330-345
  stack_with_repeat_2d   models/model_0v10.py:18-46
and the loops built on them:
  traj_viou_matrix       models/model_0v10.py:565-581, tools/train_vidor.py:107-122 (one launch for a batch of videos)
  enti_viou_align        models/model_0v10.py:559-604
  prop_pair_to_gt_pred   tools/train_vidor.py:143-159 (labels as a dense uint8 matrix, see ``pair_labels``)
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _cabi
from ._cabi import check, lib, ptr, require_cuda, stream_ptr

TILE, CHUNK = 32, 256  # csrc/geometry.cu T2_TA / T2_TB, T2_CH


def cat_rows(xs):
    """Concatenate along rows; zero-copy when the pieces are adjacent row views of one device buffer."""
    if len(xs) == 1:
        return xs[0].contiguous()
    x0 = xs[0]
    if x0.is_contiguous() and x0.dim() >= 1:
        row = x0.stride(0) * x0.element_size() if x0.dim() > 1 else x0.element_size()
        nxt, ok = x0.data_ptr(), True
        for x in xs:
            if (not x.is_contiguous()) or x.data_ptr() != nxt or x.dtype != x0.dtype or x.shape[1:] != x0.shape[1:] \
                    or x.untyped_storage().data_ptr() != x0.untyped_storage().data_ptr():
                ok = False
                break
            nxt += x.shape[0] * row
        if ok:
            total = sum(int(x.shape[0]) for x in xs)
            return x0.as_strided((total,) + tuple(x0.shape[1:]), x0.stride())
    return torch.cat(xs, 0)


# --------------------------------------------------------------------------------------------
class TrackTable(object):
    """Packed tracks of a batch of videos in HBM: boxes f32[sum L,4], off i64[n+1], dura i64[n,2] (closed),
    seg i32[V+1] (track range of each video)."""

    def __init__(self, boxes, off, dura, seg, counts: List[int], seg_len: Optional[List[int]] = None):
        self.boxes, self.off, self.dura, self.seg, self.counts = boxes, off, dura, seg, counts
        self.seg_len = seg_len      # frames on each video's absolute frame axis (video_len); needed by the tiled kernel

    @property
    def n_tracks(self):
        return int(self.dura.shape[0])

    @classmethod
    def from_containers(cls, items: Sequence, device=None) -> "TrackTable":
        """``items``: TrajProposal / VideoGraph objects (packed ``bboxes``/``lengths``) of the batch."""
        boxes_l, lens_l, dura_l, counts, seg_len = [], [], [], [], []
        for it in items:
            seg_len.append(int(getattr(it, "video_len", 0)))
            bx = it.bboxes
            du = it.traj_durations
            device = device or bx.device
            boxes_l.append(bx.to(device, torch.float32))
            dura_l.append(du.to(device, torch.long))
            lens_l.append(it.lengths)
            counts.append(int(du.shape[0]))
        require_cuda(*boxes_l)
        boxes, dura = cat_rows(boxes_l), cat_rows(dura_l)
        lens = torch.cat([torch.as_tensor(l, dtype=torch.long) for l in lens_l]) if lens_l else torch.zeros(0, dtype=torch.long)
        off = torch.zeros(lens.numel() + 1, dtype=torch.long)
        off[1:] = torch.cumsum(lens, 0)
        seg = torch.zeros(len(counts) + 1, dtype=torch.int32)
        seg[1:] = torch.cumsum(torch.tensor(counts, dtype=torch.int32), 0)
        return cls(boxes, off.to(device), dura, seg.to(device), counts, seg_len if all(x > 0 for x in seg_len) else None)

    @classmethod
    def from_lists(cls, boxes_list: Sequence[torch.Tensor], dura: torch.Tensor) -> "TrackTable":
        """Reference-style arguments of one video: list of [L_i,4] tensors + [n,2] closed spans."""
        require_cuda(dura, *boxes_list)
        n = len(boxes_list)
        boxes = torch.cat([b.float() for b in boxes_list], 0) if n else torch.zeros(0, 4, device=dura.device)
        lens = torch.tensor([int(b.shape[0]) for b in boxes_list], dtype=torch.long)
        off = torch.zeros(n + 1, dtype=torch.long)
        off[1:] = torch.cumsum(lens, 0)
        seg = torch.tensor([0, n], dtype=torch.int32)
        return cls(boxes.contiguous(), off.to(dura.device), dura.long().contiguous(), seg.to(dura.device), [n])


def dura_intersection_ts(dura1: torch.Tensor, dura2: torch.Tensor, broadcast: bool = True):
    """Closed-span intersection, bit-exact on int64 (utils/utils_func.py:347-373)."""
    assert isinstance(dura1, torch.Tensor) and isinstance(dura2, torch.Tensor)
    require_cuda(dura1, dura2)
    n1, n2 = dura1.shape[0], dura2.shape[0]
    if dura1.dtype == torch.long and dura2.dtype == torch.long:
        dt, a, b = 0, dura1.contiguous(), dura2.contiguous()
    else:
        dt, a, b = 1, dura1.float().contiguous(), dura2.float().contiguous()
    # the reference asserts every input span is valid (:351-353)
    assert bool((a[:, 0] <= a[:, 1]).all()) and bool((b[:, 0] <= b[:, 1]).all())
    if broadcast:
        inter = torch.empty(n1, n2, 2, dtype=a.dtype, device=a.device)
        mask = torch.empty(n1, n2, dtype=torch.uint8, device=a.device)
    else:
        assert n1 == n2
        inter = torch.empty(n1, 2, dtype=a.dtype, device=a.device)
        mask = torch.empty(n1, dtype=torch.uint8, device=a.device)
    check(lib().vsg_dura_intersection_ex(ptr(a), n1, ptr(b), n2, 1 if broadcast else 0, dt, ptr(inter), ptr(mask),
                                         stream_ptr(a.device)), "vsg_dura_intersection_ex")
    return inter, mask.bool()


def _tiou(duras1, duras2, broadcast, generalized):
    assert isinstance(duras1, torch.Tensor) and isinstance(duras2, torch.Tensor)
    require_cuda(duras1, duras2)
    n1, n2 = duras1.shape[0], duras2.shape[0]
    if duras1.dtype == torch.long and duras2.dtype == torch.long:
        dt, a, b = 0, duras1.contiguous(), duras2.contiguous()
    else:
        dt, a, b = 1, duras1.float().contiguous(), duras2.float().contiguous()
    if not broadcast:
        assert duras1.shape == duras2.shape
    out = torch.empty((n1, n2) if broadcast else (n1,), dtype=torch.float32, device=a.device)
    check(lib().vsg_tiou(ptr(a), n1, ptr(b), n2, 1 if broadcast else 0, 1 if generalized else 0, dt, ptr(out), stream_ptr(a.device)),
          "vsg_tiou")
    return out


def tIoU(duras1: torch.Tensor, duras2: torch.Tensor, broadcast: bool = True) -> torch.Tensor:
    """Temporal IoU of closed spans, zero where they do not touch (utils/utils_func.py:375-390)."""
    return _tiou(duras1, duras2, broadcast, False)


def generalized_tIoU(duras1: torch.Tensor, duras2: torch.Tensor, broadcast: bool = True) -> torch.Tensor:
    """tIoU without the zeroing, in [-1, 1] (utils/utils_func.py:393-410, models/grd_model_v5.py:18-33)."""
    return _tiou(duras1, duras2, broadcast, True)


def unique_with_idx_nd(tensor: torch.Tensor):
    """``(unique rows in lexicographic order, tuple of index tensors)`` -- utils/utils_func.py:330-345: ``torch.unique(dim=0)`` plus, per
    unique row, the ascending original indices of its occurrences.  int64 input of shape (N, d1, ..., dk), N <= 8192; one D2H read of
    the group count / sizes (the result is a Python tuple of ragged tensors)."""
    require_cuda(tensor)
    assert tensor.dtype == torch.long, "unique_with_idx_nd: int64 rows"
    n = tensor.shape[0]
    if n == 0:
        return tensor[:0], tuple()
    flat = tensor.reshape(n, -1).contiguous()
    d = max(int(flat.shape[1]), 1)
    dev = tensor.device
    order = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    group = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    ng = torch.empty(1, dtype=torch.int32, device=dev)
    check(lib().vsg_unique_rows(ptr(flat), n, d, ptr(order), ptr(group), ptr(ng), stream_ptr(dev)), "vsg_unique_rows")
    if n == 0:
        return tensor[:0], tuple()
    order, group = order[:n].long(), group[:n].long()
    counts = torch.bincount(group).tolist()                      # host sizes of the ragged result
    index_map = torch.split(order, counts)
    first = torch.cumsum(torch.tensor([0] + counts[:-1]), 0).to(dev)
    return tensor[order[first]], tuple(index_map)


def stack_with_repeat_2d(tensor_list: Sequence[torch.Tensor], dim: int) -> torch.Tensor:
    """models/model_0v10.py:18-46: stretch every 2-D tensor to the longest one by repeating element i ``ceil((max_L - i) / L)`` times
    along the ragged axis, then ``torch.stack(..., dim)``.  (The B200 hot path never builds this tensor -- its kernels apply the same
    index map on the fly; this is the helper for code that calls the reference function.)"""
    assert len(tensor_list) > 0 and tensor_list[0].dim() == 2
    require_cuda(*tensor_list)
    rows = [int(t.shape[0]) for t in tensor_list]
    cols = [int(t.shape[1]) for t in tensor_list]
    if all(r == rows[0] for r in rows):
        repeat_dim = 1
    elif all(c == cols[0] for c in cols):
        repeat_dim = 0
    else:
        assert False
    dev = tensor_list[0].device
    srcs = [t.float() if repeat_dim == 0 else t.float().t() for t in tensor_list]        # ragged axis first
    lens = [int(t.shape[0]) for t in srcs]
    width, tmax, n = int(srcs[0].shape[1]), max(lens), len(srcs)
    src = torch.cat([t.contiguous() for t in srcs], 0)
    off = torch.zeros(n + 1, dtype=torch.long)
    off[1:] = torch.cumsum(torch.tensor(lens), 0)
    out = torch.empty(n, tmax, width, dtype=torch.float32, device=dev)
    check(lib().vsg_stretch_rows(ptr(src), width, width, ptr(off.to(dev)), n, tmax, ptr(out), stream_ptr(dev)), "vsg_stretch_rows")
    parts = list(out.unbind(0)) if repeat_dim == 0 else [o.t() for o in out.unbind(0)]
    return torch.stack(parts, dim=dim).to(tensor_list[0].dtype)


def trajid2pairid(num_prop: int, device="cuda") -> torch.Tensor:
    """All ordered (s,o), s != o, row-major, int64[n(n-1),2] (model_pairwise_baseline.py:104-111)."""
    out = torch.empty(max(num_prop * (num_prop - 1), 0), 2, dtype=torch.long, device=device)
    check(lib().vsg_pair_ids(int(num_prop), ptr(out), stream_ptr(out.device)), "vsg_pair_ids")
    return out


def track_volumes(table: TrackTable) -> torch.Tensor:
    vol = torch.empty(table.n_tracks, dtype=torch.float32, device=table.boxes.device)
    check(lib().vsg_track_volumes(ptr(table.boxes), ptr(table.off), table.n_tracks, ptr(vol), stream_ptr(vol.device)),
          "vsg_track_volumes")
    return vol


def traj_viou_batched(A: TrackTable, B: TrackTable, want_spans=True, want_mask=True, want_viou=True,
                      want_inter=False, variant: int = 1):
    """vIoU matrices of every video of the batch in ONE launch.

    Returns ``(viou f32[P], spans i64[P,2], mask u8[P], seg_out i64[V+1] (host list), inter f32[P])``; video ``v``'s
    ``nA_v x nB_v`` block is rows ``seg_out[v]:seg_out[v+1]`` (row-major).  Entries not requested are None.
    """
    assert len(A.counts) == len(B.counts)
    same = A is B
    dev = A.boxes.device
    # per-video output offsets: cached on A for a given B (a resident batch is evaluated many times; the upload is a blocking H2D copy)
    cache = A.__dict__.setdefault("_seg_out_cache", {})
    hit = cache.get(id(B))
    if hit is not None and hit[0] is B:
        _, seg_out_h, seg_out = hit
    else:
        sizes = [a * b for a, b in zip(A.counts, B.counts)]
        seg_out_h = [0]
        for s in sizes:
            seg_out_h.append(seg_out_h[-1] + s)
        seg_out = torch.tensor(seg_out_h, dtype=torch.long).to(dev)
        cache.clear()
        cache[id(B)] = (B, seg_out_h, seg_out)
    P = seg_out_h[-1]
    spans = torch.empty(P, 2, dtype=torch.long, device=dev) if want_spans else None
    mask = torch.empty(P, dtype=torch.uint8, device=dev) if want_mask else None
    viou = torch.empty(P, dtype=torch.float32, device=dev) if want_viou else None
    inter = torch.empty(P, dtype=torch.float32, device=dev) if want_inter else None
    volA = torch.empty(max(A.n_tracks, 1), dtype=torch.float32, device=dev)
    volB = volA if same else torch.empty(max(B.n_tracks, 1), dtype=torch.float32, device=dev)
    if variant == 2:
        assert not want_inter
        seg_len = A.seg_len or B.seg_len
        if seg_len is None:      # no video_len known: one D2H read of the largest end frame
            hi = 0
            if A.n_tracks:
                hi = max(hi, int(A.dura[:, 1].max().item()) + 1)
            if B.n_tracks:
                hi = max(hi, int(B.dura[:, 1].max().item()) + 1)
            seg_len = [hi] * len(A.counts)
        n_jobs = sum(((a + TILE - 1) // TILE) * ((b + TILE - 1) // TILE) * max(1, (l + CHUNK - 1) // CHUNK)
                     for a, b, l in zip(A.counts, B.counts, seg_len))
        seg_len_d = torch.tensor(seg_len, dtype=torch.long).to(dev)
        jobs_ws = torch.empty((len(A.counts) + 1) * 3, dtype=torch.long, device=dev)
        njobs_ws = torch.empty(1, dtype=torch.long, device=dev)
        part_ws = torch.empty(max(n_jobs, 1) * TILE * TILE, dtype=torch.float32, device=dev)
        check(lib().vsg_traj_viou_matrix_tiled(
            ptr(A.boxes), ptr(A.off), ptr(A.dura), A.n_tracks, ptr(B.boxes), ptr(B.off), ptr(B.dura), B.n_tracks,
            ptr(A.seg), ptr(B.seg), ptr(seg_out), ptr(seg_len_d), len(A.counts), P, n_jobs, ptr(spans), ptr(mask), ptr(viou),
            ptr(volA), ptr(volB), ptr(jobs_ws), ptr(njobs_ws), ptr(part_ws), stream_ptr(dev)), "vsg_traj_viou_matrix_tiled")
    else:
        check(lib().vsg_traj_viou_matrix(
            ptr(A.boxes), ptr(A.off), ptr(A.dura), A.n_tracks, ptr(B.boxes), ptr(B.off), ptr(B.dura), B.n_tracks,
            ptr(A.seg), ptr(B.seg), ptr(seg_out), len(A.counts), P, ptr(spans), ptr(mask), ptr(viou), ptr(inter),
            ptr(volA), ptr(volB), variant, stream_ptr(dev)), "vsg_traj_viou_matrix")
    return viou, spans, mask, seg_out_h, inter


def traj_viou_matrix(boxes_a: Sequence[torch.Tensor], dura_a: torch.Tensor,
                     boxes_b: Sequence[torch.Tensor], dura_b: torch.Tensor, variant: int = 1):
    """One video, reference-style arguments.  Returns ``(viou f32[nA,nB], inter i64[nA,nB,2], mask bool[nA,nB])`` --
    what the loop at models/model_0v10.py:569-581 produces (zeros where the spans do not overlap)."""
    A = TrackTable.from_lists(boxes_a, dura_a)
    B = TrackTable.from_lists(boxes_b, dura_b)
    nA, nB = A.n_tracks, B.n_tracks
    viou, spans, mask, _, _ = traj_viou_batched(A, B, variant=variant)
    return viou.view(nA, nB), spans.view(nA, nB, 2), mask.view(nA, nB).bool()


def vIoU_ts(traj_1: torch.Tensor, traj_2: torch.Tensor, dura_1, dura_2) -> torch.Tensor:
    """Volume IoU of two trajectories given *relative* closed overlap spans (utils/utils_func.py:437-471):
    full-track volumes (+1 pixel convention), intersection over the given slices only."""
    assert isinstance(traj_1, torch.Tensor) and isinstance(traj_2, torch.Tensor)
    require_cuda(traj_1, traj_2)
    s1, e1 = int(dura_1[0]), int(dura_1[1])
    s2, e2 = int(dura_2[0]), int(dura_2[1])
    a = traj_1.float()[s1:e1 + 1].contiguous()
    b = traj_2.float()[s2:e2 + 1].contiguous()
    assert a.shape == b.shape
    dev = traj_1.device
    n = a.shape[0]
    span = torch.tensor([[0, n - 1]], dtype=torch.long, device=dev)
    A = TrackTable.from_lists([a], span)
    B = TrackTable.from_lists([b], span)
    _, _, _, _, inter = traj_viou_batched(A, B, want_spans=False, want_mask=False, want_viou=False, want_inter=True)
    full = TrackTable.from_lists([traj_1.float().contiguous(), traj_2.float().contiguous()],
                                 torch.tensor([[0, traj_1.shape[0] - 1], [0, traj_2.shape[0] - 1]], dtype=torch.long, device=dev))
    vol = track_volumes(full)
    it = inter[0]
    return it / (vol[0] + vol[1] - it)


def pair_labels(viou: torch.Tensor, gt_so: torch.Tensor, th: float) -> torch.Tensor:
    """Base-C label assignment (tools/train_vidor.py:143-159): ``bool[n_gt_pred, n(n-1)]`` in ``trajid2pairid`` order."""
    require_cuda(viou, gt_so)
    n, ng = viou.shape
    npred = gt_so.shape[0]
    out = torch.zeros(npred, max(n * (n - 1), 0), dtype=torch.uint8, device=viou.device)
    check(lib().vsg_pair_labels(ptr(viou.float().contiguous()), n, ng, ptr(gt_so.long().contiguous()), npred, float(th),
                                ptr(out), stream_ptr(viou.device)), "vsg_pair_labels")
    return out.bool()


def enti_viou_align(gt_adj: torch.Tensor, proposal, gt_graph, positive_vIoU_th: float, gt_closed: bool = True):
    """Training label assignment of models/model_0v10.py:559-604 on top of the vIoU-matrix kernel.

    ``gt_graph.traj_durations`` must be closed spans (``gt_closed``); unlike the reference (:567) the input
    is not mutated.  The force-assign / row-argmax steps (:583-602) are tiny index ops on the matrix.
    """
    assert gt_closed
    A = TrackTable.from_containers([proposal])
    B = TrackTable.from_containers([gt_graph], device=A.boxes.device)
    nP, nG = A.n_tracks, B.n_tracks
    viou, _, _, _, _ = traj_viou_batched(A, B, want_spans=False, want_mask=False)
    viou = viou.view(nP, nG)
    hit = viou > positive_vIoU_th
    best_prop = torch.argmax(viou, dim=0)
    orphan = hit.sum(dim=0) == 0
    hit[best_prop[orphan], orphan.nonzero(as_tuple=True)[0]] = True
    row_has = hit.sum(dim=1) > 0
    best_gt = torch.argmax(viou, dim=1)
    aligned = gt_adj.to(viou.device)[:, :, best_gt] * row_has[None, None, :].float()
    return aligned, viou


def prop_pair_to_gt_pred(proposals: Sequence, gt_graphs: Sequence, positive_vIoU_th: float, num_pred_cats: int):
    """Base-C training label assignment for a whole dataset (tools/train_vidor.py:80-170 "takes around 1.5 hours" + the
    multi-hot construction of :242-256): ONE launch computes every video's proposal x GT-trajectory vIoU matrix, then a
    label kernel per video marks the ordered proposal pairs (s, o) whose subject AND object exceed the threshold for a GT relation.

    Returns ``(label_maps, stats)``: ``label_maps[video_name] = None | (pairid2trajids i64[num_pairs,2], multihot f32[num_pairs,P])``
    with pairs in the reference's order (first hit: GT relation major, pair minor); ``stats`` = hit counters the reference prints.
    GT ``traj_durations`` are closed spans.
    """
    live = [i for i, g in enumerate(gt_graphs) if g.num_trajs > 0 and g.num_preds > 0 and proposals[i].num_proposals > 0]
    out = {g.video_name: None for g in gt_graphs}
    stats = dict(hit_gt_traj=0, gt_traj=0, hit_gt_pred=0, gt_pred=0)
    if not live:
        return out, stats
    lp, lg = [proposals[i] for i in live], [gt_graphs[i] for i in live]
    A = TrackTable.from_containers(lp)
    B = TrackTable.from_containers(lg, device=A.boxes.device)
    viou, _, _, seg, _ = traj_viou_batched(A, B, want_spans=False, want_mask=False)
    dev = viou.device
    for k, (p, g) in enumerate(zip(lp, lg)):
        n, ng = p.num_proposals, g.num_trajs
        v = viou[seg[k]:seg[k + 1]].view(n, ng)
        stats["gt_traj"] += ng
        stats["hit_gt_traj"] += int((v > positive_vIoU_th).any(dim=0).sum())
        so = torch.argmax(g.adj_matrix.to(dev), dim=-1).t().contiguous()
        stats["gt_pred"] += g.num_preds
        if n < 2:
            continue
        lab = pair_labels(v, so, positive_vIoU_th)                       # bool [n_gt_pred, n(n-1)]
        gi, pi = lab.nonzero(as_tuple=True)                              # sorted by GT relation, then pair: the reference's loop order
        stats["hit_gt_pred"] += int(torch.unique(gi).numel())
        if gi.numel() == 0:
            continue
        uniq, inv = torch.unique(pi, return_inverse=True)               # sorted by pair id; re-order by first appearance
        first = torch.full((uniq.numel(),), gi.numel(), dtype=torch.long, device=dev).scatter_reduce_(
            0, inv, torch.arange(gi.numel(), device=dev), reduce="amin")
        order = torch.argsort(first)
        rank = torch.empty_like(order)
        rank[order] = torch.arange(order.numel(), device=dev)
        pair_ids_sorted = uniq[order]
        s = pair_ids_sorted // (n - 1)
        r = pair_ids_sorted % (n - 1)
        o = r + (r >= s).long()
        multihot = torch.zeros(uniq.numel(), num_pred_cats, device=dev)
        multihot[rank[inv], g.pred_cat_ids.to(dev)[gi]] = 1
        out[g.video_name] = (torch.stack([s, o], 1), multihot)
    return out, stats
