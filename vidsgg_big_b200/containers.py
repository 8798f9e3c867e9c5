"""Input containers of the per-video relation hot path.

Mirrors the attribute surface of the reference containers

* ``TrajProposal``  -- dataloaders/dataloader_vidvrd.py:14-82 (VidOR twin
  dataloaders/dataloader_vidor_v3.py:18-99)
* ``VideoGraph``    -- dataloaders/dataloader_vidvrd.py:84-143

but stores every per-tracklet tensor *packed* (CSR): one ``[sum(L_i), 4]`` box
buffer and one ``[sum(L_i), D]`` feature buffer per video plus ``lengths``.  The
reference's ``bboxes_list`` / ``features_list`` attributes are exposed as zero-copy
row views into those buffers, so code written against the reference containers keeps
working while the CUDA path consumes the packed buffers without a gather.

Spans are **closed** ``[s, e]`` with ``L_i == e - s + 1`` exactly as in the reference
(dataloader_vidvrd.py:33-34).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def _as_long(x, device=None):
    return torch.as_tensor(x, dtype=torch.long, device=device)


class TrajProposal(object):
    """Tracklet proposals of one video (packed storage, reference attribute names)."""

    def __init__(self, video_name: str, video_len: int, video_wh: Tuple[int, int],
                 cat_ids, scores, traj_durations, bboxes, features=None, lengths=None):
        self.video_name = video_name
        self.video_len = int(video_len)
        self.video_wh = (video_wh[0], video_wh[1])
        self.cat_ids = _as_long(cat_ids)
        self.num_proposals = int(self.cat_ids.shape[0])
        if self.num_proposals == 0:
            # the reference object carries no tensor attributes at all in this case
            # (dataloader_vidvrd.py:25-28); keep empty ones so packing stays uniform
            self.scores = torch.zeros(0, dtype=torch.float32)
            self.traj_durations = torch.zeros(0, 2, dtype=torch.long)
            self.bboxes = torch.zeros(0, 4, dtype=torch.float32)
            self.features = None
            self.lengths = torch.zeros(0, dtype=torch.long)
            self.dim_feat = "[]"
            return
        self.scores = torch.as_tensor(scores, dtype=torch.float32)
        self.traj_durations = _as_long(traj_durations).reshape(-1, 2)
        self.bboxes = torch.as_tensor(bboxes, dtype=torch.float32).reshape(-1, 4)
        # fp32 like the reference; a bf16 tensor is kept as is (the opt-in bf16 feature transport of precision="bf16")
        self.features = None if features is None else (features if (torch.is_tensor(features) and features.dtype == torch.bfloat16)
                                                      else torch.as_tensor(features, dtype=torch.float32))
        if lengths is None:
            lengths = (self.traj_durations[:, 1] - self.traj_durations[:, 0] + 1).cpu()
        self.lengths = _as_long(lengths).cpu()
        assert int(self.lengths.sum()) == self.bboxes.shape[0], "packed boxes do not match lengths"
        d = self.traj_durations.cpu()
        assert torch.equal(d[:, 1] - d[:, 0] + 1, self.lengths), "L_i must equal e-s+1 (closed spans)"
        if self.features is not None:
            assert self.features.shape[0] == self.bboxes.shape[0]
        self.dim_feat = self.features.shape[1] if self.features is not None else 0

    # ---- reference-compatible list views (zero copy) -------------------------------
    @property
    def bboxes_list(self) -> List[torch.Tensor]:
        return list(torch.split(self.bboxes, self.lengths.tolist(), dim=0))

    @property
    def features_list(self) -> Optional[List[torch.Tensor]]:
        if self.features is None:
            return None
        return list(torch.split(self.features, self.lengths.tolist(), dim=0))

    @classmethod
    def from_lists(cls, video_name, video_len, video_wh, cat_ids, scores, traj_durations,
                   bboxes_list: Sequence[torch.Tensor], features_list: Optional[Sequence[torch.Tensor]] = None):
        """Build from reference-style per-tracklet lists (one concatenation)."""
        lengths = [int(b.shape[0]) for b in bboxes_list]
        bboxes = torch.cat(list(bboxes_list), 0) if lengths else torch.zeros(0, 4)
        feats = torch.cat(list(features_list), 0) if (features_list is not None and lengths) else None
        return cls(video_name, video_len, video_wh, cat_ids, scores, traj_durations, bboxes, feats, lengths)

    @classmethod
    def from_reference(cls, ref):
        """Adopt a reference ``TrajProposal`` object (duck-typed on its attributes)."""
        if ref.num_proposals == 0:
            return cls(ref.video_name, getattr(ref, "video_len", 0), getattr(ref, "video_wh", (1, 1)),
                       [], [], [], torch.zeros(0, 4))
        return cls.from_lists(ref.video_name, ref.video_len, ref.video_wh, ref.cat_ids, ref.scores,
                              ref.traj_durations, ref.bboxes_list, ref.features_list)

    def to(self, device):
        """In-place move, like the reference (dataloader_vidvrd.py:66-77) -- but 5 copies, not 2n+3."""
        if self.num_proposals == 0:
            return self
        self.cat_ids = self.cat_ids.to(device)
        self.scores = self.scores.to(device)
        self.traj_durations = self.traj_durations.to(device)
        self.bboxes = self.bboxes.to(device)
        if self.features is not None:
            self.features = self.features.to(device)
        return self

    @property
    def device(self):
        return self.bboxes.device

    def __repr__(self):
        return "TrajProposal[{},num_proposals={},feature_dim={}]".format(
            self.video_name, self.num_proposals, self.dim_feat)


class VideoGraph(object):
    """Ground-truth graph of one video (reference dataloader_vidvrd.py:84-143).

    ``traj_durations`` are closed spans; ``pred_durations`` float closed spans;
    ``adj_matrix`` is ``f32[2, num_preds, num_trajs]`` (subject plane, object plane).
    """

    def __init__(self, video_name, video_len, video_wh, traj_cat_ids, traj_durations, traj_bboxes,
                 pred_cat_ids, pred_durations, adj_matrix, lengths=None):
        self.video_name, self.video_len, self.video_wh = video_name, int(video_len), tuple(video_wh)
        self.traj_cat_ids = _as_long(traj_cat_ids)
        self.traj_durations = _as_long(traj_durations).reshape(-1, 2)
        self.num_trajs = int(self.traj_cat_ids.shape[0])
        self.bboxes = torch.as_tensor(traj_bboxes, dtype=torch.float32).reshape(-1, 4)
        if lengths is None:
            lengths = (self.traj_durations[:, 1] - self.traj_durations[:, 0] + 1).cpu()
        self.lengths = _as_long(lengths).cpu()
        assert int(self.lengths.sum()) == self.bboxes.shape[0]
        self.pred_cat_ids = _as_long(pred_cat_ids)
        self.num_preds = int(self.pred_cat_ids.shape[0])
        self.pred_durations = torch.as_tensor(pred_durations, dtype=torch.float32).reshape(-1, 2)
        self.adj_matrix = torch.as_tensor(adj_matrix, dtype=torch.float32).reshape(2, self.num_preds, self.num_trajs)
        self.n_frames_list = self.lengths.tolist()
        self.max_frames = max(self.n_frames_list) if self.n_frames_list else 0

    @property
    def traj_bboxes(self) -> List[torch.Tensor]:
        return list(torch.split(self.bboxes, self.lengths.tolist(), dim=0))

    def to(self, device):
        self.traj_cat_ids = self.traj_cat_ids.to(device)
        self.pred_cat_ids = self.pred_cat_ids.to(device)
        self.traj_durations = self.traj_durations.to(device)
        self.pred_durations = self.pred_durations.to(device)
        self.adj_matrix = self.adj_matrix.to(device)
        self.bboxes = self.bboxes.to(device)
        return self

    def __repr__(self):
        return "VideoGraph[num_trajs={},num_preds={}]".format(self.num_trajs, self.num_preds)
