"""Oracle: span / pair / trajectory-vIoU geometry.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch


def dura_intersection(d1: torch.Tensor, d2: torch.Tensor, broadcast: bool = True):
    """Closed-span intersection.  Follows utils/utils_func.py:347-373.

    Returns ``(inter[n1,n2,2] = (max start, min end), mask = start<=end)``; pairs that do
    not overlap keep their *inverted* span.  Works on int64 (bit-exact) and float spans.
    """
    assert bool((d1[:, 0] <= d1[:, 1]).all()) and bool((d2[:, 0] <= d2[:, 1]).all())
    if broadcast:
        lo = torch.maximum(d1[:, None, 0], d2[None, :, 0])
        hi = torch.minimum(d1[:, None, 1], d2[None, :, 1])
    else:
        assert d1.shape[0] == d2.shape[0]
        lo = torch.maximum(d1[:, 0], d2[:, 0])
        hi = torch.minimum(d1[:, 1], d2[:, 1])
    inter = torch.stack([lo, hi], dim=-1)
    return inter, lo <= hi


def pair_ids(n: int) -> torch.Tensor:
    """All ordered (s,o), s != o, row-major.  models/model_pairwise_baseline.py:104-111,
    tools/train_vidor.py:73-78."""
    rows = [(s, o) for s in range(n) for o in range(n) if s != o]
    return torch.tensor(rows, dtype=torch.long).reshape(-1, 2)


def viou_single(traj1: torch.Tensor, traj2: torch.Tensor, rel1, rel2) -> torch.Tensor:
    """One pair's volume IoU from *relative* closed overlap spans.  utils/utils_func.py:437-471.

    Areas use the +1 pixel convention and are summed over the FULL tracks; the intersection
    only over the overlap slice.
    """
    a = traj1.float()
    b = traj2.float()
    vol_a = ((a[:, 2] - a[:, 0] + 1) * (a[:, 3] - a[:, 1] + 1)).sum()
    vol_b = ((b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)).sum()
    sa = a[int(rel1[0]): int(rel1[1]) + 1]
    sb = b[int(rel2[0]): int(rel2[1]) + 1]
    assert sa.shape == sb.shape
    w = (torch.minimum(sa[:, 2], sb[:, 2]) - torch.maximum(sa[:, 0], sb[:, 0]) + 1).clamp(min=0.0)
    h = (torch.minimum(sa[:, 3], sb[:, 3]) - torch.maximum(sa[:, 1], sb[:, 1]) + 1).clamp(min=0.0)
    inter = (w * h).sum()
    return inter / (vol_a + vol_b - inter)


def traj_viou_matrix(boxes_a: Sequence[torch.Tensor], dura_a: torch.Tensor,
                     boxes_b: Sequence[torch.Tensor], dura_b: torch.Tensor):
    """vIoU of every temporally-overlapping (a, b) pair, zeros elsewhere.

    The per-pair Python loop of models/model_0v10.py:565-581 (twin tools/train_vidor.py:107-122).
    Returns ``(viou f32[nA,nB], inter i64[nA,nB,2], mask bool[nA,nB])``.
    """
    inter, mask = dura_intersection(dura_a, dura_b)
    rel_a = inter - dura_a[:, 0, None, None]
    rel_b = inter - dura_b[None, :, 0, None]
    out = torch.zeros(mask.shape, dtype=torch.float32)
    ia, ib = mask.nonzero(as_tuple=True)
    for p, g in zip(ia.tolist(), ib.tolist()):
        out[p, g] = viou_single(boxes_a[p], boxes_b[g], rel_a[p, g], rel_b[p, g])
    return out, inter, mask


def traj_viou_matrix_np(boxes_a: np.ndarray, off_a: np.ndarray, dura_a: np.ndarray,
                        boxes_b: np.ndarray, off_b: np.ndarray, dura_b: np.ndarray) -> np.ndarray:
    """Vectorised-per-pair numpy twin of :func:`traj_viou_matrix` for larger test sizes
    (same arithmetic in float32, same summation unit = one pair; numpy pairwise summation)."""
    nA, nB = len(dura_a), len(dura_b)
    def vols(bx, off):
        ar = (bx[:, 2] - bx[:, 0] + np.float32(1)) * (bx[:, 3] - bx[:, 1] + np.float32(1))
        return np.array([ar[off[i]:off[i + 1]].sum(dtype=np.float32) for i in range(len(off) - 1)], np.float32)
    va, vb = vols(boxes_a, off_a), vols(boxes_b, off_b)
    out = np.zeros((nA, nB), np.float32)
    for i in range(nA):
        for j in range(nB):
            s = max(dura_a[i, 0], dura_b[j, 0]); e = min(dura_a[i, 1], dura_b[j, 1])
            if s > e:
                continue
            A = boxes_a[off_a[i] + s - dura_a[i, 0]: off_a[i] + e - dura_a[i, 0] + 1]
            B = boxes_b[off_b[j] + s - dura_b[j, 0]: off_b[j] + e - dura_b[j, 0] + 1]
            w = np.maximum(np.minimum(A[:, 2], B[:, 2]) - np.maximum(A[:, 0], B[:, 0]) + np.float32(1), np.float32(0))
            h = np.maximum(np.minimum(A[:, 3], B[:, 3]) - np.maximum(A[:, 1], B[:, 1]) + np.float32(1), np.float32(0))
            it = (w * h).sum(dtype=np.float32)
            out[i, j] = it / (va[i] + vb[j] - it)
    return out


def tiou(d1, d2, broadcast=True):
    """utils/utils_func.py:375-390 (zero where the closed spans do not touch)."""
    if broadcast:
        a0, a1, b0, b1 = d1[:, None, 0], d1[:, None, 1], d2[None, :, 0], d2[None, :, 1]
    else:
        a0, a1, b0, b1 = d1[:, 0], d1[:, 1], d2[:, 0], d2[:, 1]
    keep = (a1 >= b0) * (b1 >= a0)
    val = (torch.minimum(a1, b1) - torch.maximum(a0, b0)) / (torch.maximum(a1, b1) - torch.minimum(a0, b0))
    val[torch.logical_not(keep)] = 0
    return val


def generalized_tiou(d1, d2, broadcast=True):
    """utils/utils_func.py:393-410 / models/grd_model_v5.py:18-33 (no zeroing; may be negative)."""
    if broadcast:
        a0, a1, b0, b1 = d1[:, None, 0], d1[:, None, 1], d2[None, :, 0], d2[None, :, 1]
    else:
        a0, a1, b0, b1 = d1[:, 0], d1[:, 1], d2[:, 0], d2[:, 1]
    return (torch.minimum(a1, b1) - torch.maximum(a0, b0)) / (torch.maximum(a1, b1) - torch.minimum(a0, b0))


def unique_rows_with_groups(t: torch.Tensor):
    """utils/utils_func.py:330-345: lexicographically sorted unique rows + per-group original
    indices in ascending order."""
    uniq, counts = torch.unique(t, return_counts=True, dim=0)
    groups = []
    for u in uniq:
        groups.append((t == u[None]).reshape(t.shape[0], -1).all(dim=-1).nonzero(as_tuple=True)[0])
    return uniq, tuple(groups)


def stretch_index_map(L: int, Tmax: int) -> np.ndarray:
    """Source frame of every stretched position: frame i of an L-frame track is repeated
    ceil((Tmax-i)/L) times.  models/model_0v10.py:18-46 (``stack_with_repeat_2d``)."""
    n_pad = L - (Tmax % L)
    tot = np.array([1] * Tmax + [0] * n_pad).reshape(-1, L)
    reps = tot.sum(0)
    return np.repeat(np.arange(L), reps)


def enti_viou_align(gt_adj: torch.Tensor, boxes_p, dura_p, boxes_g, dura_g_closed, th: float):
    """Training label assignment on top of the vIoU matrix.  models/model_0v10.py:559-604.

    ``dura_g_closed`` must already be closed (the reference converts in place at :567).
    Returns ``(gt_adj_enti_align f32[2,n_gt_pred,nP], viou f32[nP,nG])``.
    """
    viou, _, _ = traj_viou_matrix(boxes_p, dura_p, boxes_g, dura_g_closed)
    nP, nG = viou.shape
    hit = viou > th
    best_prop = torch.argmax(viou, dim=0)
    orphan = hit.sum(dim=0) == 0
    hit[best_prop[orphan], orphan] = True
    assert int(hit.sum()) >= nG
    out = torch.zeros(2, gt_adj.shape[1], nP)
    for p in range(nP):
        if int(hit[p].sum()) > 0:
            out[:, :, p] = gt_adj[:, :, int(torch.argmax(viou[p]))]
    return out, viou


def pair_labels(viou: torch.Tensor, gt_so: torch.Tensor, th: float) -> torch.Tensor:
    """Base-C label assignment, tools/train_vidor.py:143-159: ordered pair (s,o) is positive for
    GT relation g iff viou[s, g_s] > th and viou[o, g_o] > th.  Returns bool[n_gt_pred, n(n-1)]
    in ``pair_ids`` order."""
    n = viou.shape[0]
    pid = pair_ids(n)
    out = torch.zeros(gt_so.shape[0], pid.shape[0], dtype=torch.bool)
    for g in range(gt_so.shape[0]):
        gs, go = gt_so[g].tolist()
        for k in range(pid.shape[0]):
            s, o = pid[k].tolist()
            out[g, k] = bool(viou[s, gs] > th) and bool(viou[o, go] > th)
    return out


def label_maps(viou: torch.Tensor, gt_5tuples: torch.Tensor, th: float, num_pred_cats: int):
    """Base-C training labels of one video, as consumed by the trainer: the dict loop of tools/train_vidor.py:143-159
    (pairs keyed in order of first hit: outer loop GT relation, inner loop pair order) followed by :242-256
    (``pairid2trajids`` + multi-hot predicate matrix).  Returns None when nothing hits."""
    n = viou.shape[0]
    pid = pair_ids(n)
    table = {}
    for g in range(gt_5tuples.shape[0]):
        gs, go = gt_5tuples[g, 3:].tolist()
        for k in range(pid.shape[0]):
            s, o = pid[k].tolist()
            if bool(viou[s, gs] > th) and bool(viou[o, go] > th):
                table.setdefault((s, o), []).append(int(gt_5tuples[g, 0]))
    if not table:
        return None
    pairs = torch.tensor(list(table.keys()), dtype=torch.long)
    multihot = torch.zeros(len(table), num_pred_cats)
    for i, cats in enumerate(table.values()):
        multihot[i, cats] = 1
    return pairs, multihot
