"""Base-C pairwise baseline on B200: host-side mirror of the reference ``Base_C`` module in inference mode
(models/model_pairwise_baseline.py, exported by models/__init__.py:3).

    model = Base_C(config, is_train=False); model.load_state_dict(sd); model.cuda()
    triplets = model(proposal_list, topk=3)      # list[None | (quintuples i64[m,5], scores f32[m,3], spans i64[m,2], query_ids f32[m])]

Every ordered pair (s, o), s != o, of a video's tracklets is scored: per-track encoding (the BIG-C front: per-frame MLPs on the unique
frames, stretched conv / max-pool, fc_enti2enco) -> gather [clsme_s, clsme_o, feat_s, feat_o] -> Linear-ReLU-Linear + frequency bias ->
softmax / top-k -> overlap filter, lexicographic order, background removed, optional ``rt_triplets_topk`` best by mean score.
All videos of a call are batched: the pair MLP is ONE tcgen05 GEMM over every pair of every video (M up to n(n-1) per video).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch

from . import linalg
from ._cabi import VsgError, check, lib, stream_ptr
from .bigc import BIG_C, PackedVideos, _raw
from .containers import TrajProposal
from .linalg import Weight, gemm


class Base_C(BIG_C):
    """Shares the packed-batch machinery and the per-track encoding of ``BIG_C``; no encoder / decoder."""

    def __init__(self, config: dict, is_train: bool = False, precision: str = "tf32+bf16x2"):
        if is_train:
            raise NotImplementedError("vidsgg_big_b200.Base_C covers inference only (is_train=False)")
        c = config
        self.is_train = False
        self.config = dict(config)
        self.num_pred_cats, self.num_enti_cats = c["num_pred_cats"], c["num_enti_cats"]
        if self.num_enti_cats > 4096 or self.num_pred_cats > 256:      # 12-bit category / track fields of the sort key (csrc/basec.cu)
            raise VsgError("triplet sort key: num_enti_cats <= 4096 and num_pred_cats <= 256")
        self.dim_feat, self.dim_clsme, self.dim_enti, self.dim_ffn = c["dim_feat"], c["dim_clsme"], c["dim_enti"], c["dim_ffn"]
        self.enco_pool_len = c["enco_pool_len"]
        self.use_clsme = c["use_clsme"]
        self.rt_triplets_topk = int(c["rt_triplets_topk"])
        if not self.use_clsme:
            # the reference sizes fc_pred2logits for the classeme input regardless (:63-67), so its use_clsme=False forward fails
            raise VsgError("Base_C: use_clsme=False is not runnable in the reference either (fc_pred2logits expects the classeme columns)")
        self.has_entiemb = c.get("EntiNameEmb_path", None) is not None
        self.extra_width = 0 if self.has_entiemb else self.dim_clsme
        self.dim_z = 2 * self.dim_clsme + 2 * self.dim_enti
        self.precision, self.mode = precision, linalg.MODES[precision]
        self.topk = 10
        self.device = None
        self._w = None
        self._state: Dict[str, torch.Tensor] = {}
        for key, path in (("EntiNameEmb", c.get("EntiNameEmb_path")), ("bias_matrix", c.get("bias_matrix_path"))):
            if path is not None and isinstance(path, str) and path.endswith(".npy"):
                self._state[key] = torch.from_numpy(np.load(path)).float()

    def _expected_keys_static(self) -> List[str]:
        k = ["EntiNameEmb"] if self.has_entiemb else []
        k.append("bias_matrix")
        for n in ("fc_feat2enti.0", "fc_feat2enti.2", "fc_bbox2enti.0", "fc_bbox2enti.2", "conv_feat2enti", "fc_enti2enco.0",
                  "fc_enti2enco.2", "fc_pred2logits.0", "fc_pred2logits.2"):
            k += [n + ".weight", n + ".bias"]
        return k

    def _prepare(self):
        dev = self.device
        if dev.type != "cuda":
            raise VsgError("Base_C runs on a CUDA device only (no CPU fallback)")
        st = {k: v.to(dev) for k, v in self._state.items()}
        E = self.dim_enti
        split = linalg.SPLITS.get(self.mode, False)
        W = lambda name: Weight(st[name + ".weight"], st[name + ".bias"], split=split)
        w = {}
        w["bbox1_w"], w["bbox1_b"] = st["fc_bbox2enti.0.weight"].contiguous(), st["fc_bbox2enti.0.bias"].contiguous()
        w["bbox2"], w["feat1"], w["feat2"] = W("fc_bbox2enti.2"), W("fc_feat2enti.0"), W("fc_feat2enti.2")
        cw = st["conv_feat2enti.weight"]
        w["conv"] = Weight(cw.permute(2, 0, 1).reshape(3 * E, 2 * E).contiguous(), None, split=split)
        w["conv_b"] = st["conv_feat2enti.bias"].contiguous()
        w["enco1"], w["enco2"] = W("fc_enti2enco.0"), W("fc_enti2enco.2")
        w["log1"], w["log2"] = W("fc_pred2logits.0"), W("fc_pred2logits.2")
        w["bias_matrix"] = st["bias_matrix"].reshape(self.num_enti_cats * self.num_enti_cats, self.num_pred_cats).contiguous()
        if self.has_entiemb:
            w["entiemb"] = st["EntiNameEmb"].contiguous()
        self._w = w

    # ---- pairs ------------------------------------------------------------------------------------------
    def _pairs(self, pk: PackedVideos):
        counts = np.asarray(pk.counts, dtype=np.int64)
        per = counts * (counts - 1)
        pair_off = np.zeros(pk.V + 1, np.int64)
        pair_off[1:] = np.cumsum(per)
        n_pairs = int(pair_off[-1])
        dev = self.device
        so = torch.empty(max(n_pairs, 1), 2, dtype=torch.int32, device=dev)
        pair_vid = torch.empty(max(n_pairs, 1), dtype=torch.int32, device=dev)
        pair_off_d = torch.from_numpy(pair_off).to(dev)
        check(lib().vsg_pair_ids_batched(_raw(pk.seg), pk.V, _raw(pair_off_d), n_pairs, _raw(so), _raw(pair_vid), stream_ptr(dev)),
              "vsg_pair_ids_batched")
        return so[:n_pairs], pair_vid[:n_pairs], pair_off, n_pairs

    def _pair_logits(self, pk: PackedVideos, so):
        """model_pairwise_baseline.py:243-273 for every pair of the batch (one gather + two GEMMs)."""
        w, m = self._w, self.mode
        E = self.dim_enti
        enti2enco, extra = self._track_encoding(pk)
        pair_index, so_cat = self._so_cats(pk, so)
        s_idx, o_idx = so, so[:, 1:]
        if self.has_entiemb:
            cl = [(w["entiemb"], so_cat, self.dim_clsme), (w["entiemb"], so_cat[:, 1:], self.dim_clsme)]
        else:
            cl = [(extra, s_idx, self.dim_clsme), (extra, o_idx, self.dim_clsme)]
        pieces = cl + [(enti2enco, s_idx, E), (enti2enco, o_idx, E)]
        ldz = (self.dim_z + 3) // 4 * 4
        Z = self._concat(pieces, so.shape[0], ldz)
        hid = gemm(m, Z, w["log1"], relu=True, K=self.dim_z)
        return gemm(m, hid, w["log2"], rowbias=w["bias_matrix"], rb_index=pair_index)

    def forward_propagation(self, proposal, pairid2trajids=None):
        """Logits of one video's pairs (model_pairwise_baseline.py:170-196); ``pairid2trajids`` defaults to every ordered pair."""
        pk = PackedVideos([proposal], self.device)
        if pairid2trajids is None:
            so, _, _, n_pairs = self._pairs(pk)
        else:
            so = pairid2trajids.to(self.device, torch.int32).contiguous()
        if so.shape[0] == 0:
            return torch.zeros(0, self.num_pred_cats, device=self.device)
        return self._pair_logits(pk, so)[:, :self.num_pred_cats]

    @staticmethod
    def trajid2pairid(num_prop: int, device="cuda"):
        from .geometry import trajid2pairid
        return trajid2pairid(num_prop, device)

    # ---- public API -------------------------------------------------------------------------------------
    def forward(self, proposal_list, pos_id_list=None, label_list=None, topk=10, max_pairs: int = 4_000_000):
        if self._w is None:
            raise VsgError("Base_C has no weights on a CUDA device: call load_state_dict(...) and .cuda() first")
        self.topk = topk
        props = [p if isinstance(p, TrajProposal) else TrajProposal.from_reference(p) for p in proposal_list]
        results: List[Optional[tuple]] = [None] * len(props)
        live = [i for i, p in enumerate(props) if p.num_proposals > 1]      # 0 tracks -> None (:117-118); 1 track -> no pair -> None (:333-334)
        batch, pairs = [], 0

        def flush():
            nonlocal batch, pairs
            if batch:
                for i, r in zip(batch, self._forward_batch([props[i] for i in batch], topk)):
                    results[i] = r
            batch, pairs = [], 0
        for i in live:
            n = props[i].num_proposals
            if batch and pairs + n * (n - 1) > max_pairs:
                flush()
            batch.append(i)
            pairs += n * (n - 1)
        flush()
        return results

    __call__ = forward

    def _forward_batch(self, props, topk):
        dev = self.device
        pk = PackedVideos(props, dev)
        so, pair_vid, pair_off, n_pairs = self._pairs(pk)
        logits = self._pair_logits(pk, so)
        cand_off = pair_off * topk
        chunks = (cand_off[1:] - cand_off[:-1] + 255) // 256
        chunk_off = np.zeros(pk.V + 1, np.int32)
        chunk_off[1:] = np.cumsum(chunks)
        n_cand = int(cand_off[-1])
        keys = torch.empty(n_cand, dtype=torch.int64, device=dev)
        sc_ws = torch.empty(n_cand, 3, dtype=torch.float32, device=dev)
        counts = torch.empty(pk.V, 2, dtype=torch.int32, device=dev)
        quint = torch.empty(n_cand, 5, dtype=torch.long, device=dev)
        scores = torch.empty(n_cand, 3, dtype=torch.float32, device=dev)
        spans = torch.empty(n_cand, 2, dtype=torch.long, device=dev)
        cand_off_d, chunk_off_d = torch.from_numpy(cand_off).to(dev), torch.from_numpy(chunk_off).to(dev)
        check(lib().vsg_pair_construct_triplet(_raw(logits), logits.stride(0), self.num_pred_cats, topk, _raw(so), _raw(pair_vid), n_pairs,
                                               _raw(pk.seg), pk.V, _raw(pk.dura), _raw(pk.cat_ids), _raw(pk.scores), _raw(cand_off_d),
                                               _raw(chunk_off_d), int(chunk_off[-1]), self.rt_triplets_topk, _raw(keys), _raw(sc_ws),
                                               _raw(counts), _raw(quint), _raw(scores), _raw(spans), stream_ptr(dev)),
              "vsg_pair_construct_triplet")
        cnt = counts.cpu().numpy()
        out = []
        for v in range(pk.V):
            n_valid, n_bg = int(cnt[v, 0]), int(cnt[v, 1])
            if n_valid == 0:
                out.append(None)                                    # no overlapping pair (:333-334)
                continue
            n_out = n_valid - n_bg
            if self.rt_triplets_topk > 0:
                n_out = min(n_out, self.rt_triplets_topk)
            s = slice(int(cand_off[v]), int(cand_off[v]) + n_out)
            out.append((quint[s], scores[s], spans[s], torch.empty(n_out)))     # uniq_query_ids is an uninitialised float tensor (:386)
        return out

    # BIG-C-only entry points
    def forward_packed(self, *a, **k):
        raise NotImplementedError("Base_C returns per-video tuples (forward)")

    forward_debug = forward_packed
