"""BIG-C classification stage on B200: host-side mirror of the reference ``BIG_C`` modules
(models/model_0v10.py:239-785 for VidVRD, models/model_0v7.py:240-790 for VidOR; exported by the
reference as ``BIG_C_vidvrd`` / ``BIG_C_vidor``, models/__init__.py:1-4).

Same constructor (``BIG_C(config, is_train=False)``), same ``forward(proposal_list, gt_graph_list=None, topk)``
signature and return value (list of ``None | (quintuples i64[m,5], scores f32[m,3], spans i64[m,2],
query_ids i64[m])``), and it loads reference ``state_dict``s unchanged.  Everything between the packed
inputs and the triplets runs in the hand-written sm_100a kernels of libvsgb200 (csrc/gemm.cu tcgen05 GEMMs,
csrc/bigc.cu fused kernels); PyTorch only owns the device buffers.  Inference only: the training path
(``_forward_train``, Hungarian matching) is out of this hot path's scope.

Differences in *how* (never in *what*):
  * all videos passed to one ``forward`` call are batched: their rows are stacked into single GEMMs;
  * the stretched ``[n, Tmax, D]`` tensors of ``_preprocess_proposal`` are never materialised -- the per-frame
    MLPs run on the unique frames and the conv / pool / time-mean kernels apply the stretch as an index map.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import linalg
from ._cabi import VsgError, check, lib, ptr, require_cuda, stream_ptr
from .containers import TrajProposal
from .linalg import Weight, gemm

_P = C.c_void_p


def _raw(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class PackedVideos(object):
    """Device-resident packed form of a batch of ``TrajProposal``s (see csrc/bigc.cu header comment)."""

    def __init__(self, proposals: Sequence[TrajProposal], device):
        self.proposals = list(proposals)
        V = len(self.proposals)
        lens = [p.lengths for p in self.proposals]
        counts = [p.num_proposals for p in self.proposals]
        all_len = torch.cat(lens) if V else torch.zeros(0, dtype=torch.long)
        off = torch.zeros(all_len.numel() + 1, dtype=torch.long)
        off[1:] = torch.cumsum(all_len, 0)
        seg = torch.zeros(V + 1, dtype=torch.int32)
        seg[1:] = torch.cumsum(torch.tensor(counts, dtype=torch.int32), 0)
        tmax = torch.cat([torch.full((c,), int(l.max()), dtype=torch.int32) for c, l in zip(counts, lens)])
        track_vid = torch.cat([torch.full((c,), i, dtype=torch.int32) for i, c in enumerate(counts)])
        wh = torch.tensor([[float(p.video_wh[0]), float(p.video_wh[1])] for p in self.proposals], dtype=torch.float32)
        from .geometry import cat_rows as cat
        self.boxes = cat([p.bboxes.to(device, torch.float32) for p in self.proposals])
        # features: fp32 (the reference's contract) or -- opt-in, bf16 mode only -- bf16 as handed over by a loader that stores them so
        fdt = torch.bfloat16 if all(p.features is not None and p.features.dtype == torch.bfloat16 for p in self.proposals) else torch.float32
        self.feats = cat([p.features.to(device, fdt) for p in self.proposals])
        self.dura = cat([p.traj_durations.to(device, torch.long) for p in self.proposals])
        self.cat_ids = cat([p.cat_ids.to(device, torch.long) for p in self.proposals])
        self.scores = cat([p.scores.to(device, torch.float32) for p in self.proposals])
        self.off, self.seg = off.to(device), seg.to(device)
        self.seg64 = seg.long().to(device)
        self.tmax, self.track_vid, self.wh = tmax.to(device), track_vid.to(device), wh.to(device)
        self.V, self.N, self.R = V, int(all_len.numel()), int(all_len.sum())
        self.max_tracks = max(counts) if counts else 0
        self.counts = counts
        # the triplet kernels pack (pred | s_cat | o_cat | s_id | o_id) into one 64-bit sort key with 12 bits per category / track id
        # (csrc/bigc.cu construct_triplet_kernel, csrc/basec.cu): larger values would alias keys silently
        if self.max_tracks > 4096:
            raise VsgError("a video has %d tracks; the triplet sort key holds track ids < 4096" % self.max_tracks)

        self.mha_blocks = linalg.mha_block_list(counts, device)     # ragged (video, 64-token block) work list of the encoder attention


class PackedTriplets(object):
    """Triplets of a batch of videos as written by the kernel: video v owns rows [v*cap, v*cap + counts[v,0])."""

    def __init__(self, quint, scores, spans, qids, counts_host, cap, counts_dev=None):
        self.quint, self.scores, self.spans, self.qids, self._counts, self.cap = quint, scores, spans, qids, counts_host, cap
        self._counts_dev = counts_dev

    @property
    def counts(self):
        """Per-video (rows written, positive pairs) on the host.  With ``forward_packed(..., sync=False)`` the D2H read is deferred to
        the first access, so that the caller can enqueue more GPU work (or switch streams) before blocking on it."""
        if self._counts is None:
            self._counts = self._counts_dev.cpu().numpy()
        return self._counts

    def compact(self):
        """Dense rows of all videos (video-major) + host offsets: 4 gathers for the whole batch."""
        n = self.counts[:, 0].astype(np.int64)
        off = np.zeros(n.size + 1, np.int64)
        off[1:] = np.cumsum(n)
        idx = np.repeat(np.arange(n.size, dtype=np.int64) * self.cap - off[:-1], n) + np.arange(off[-1], dtype=np.int64)
        idx_d = torch.from_numpy(idx).to(self.quint.device)
        return self.quint[idx_d], self.scores[idx_d], self.spans[idx_d], self.qids[idx_d], off

    def per_video(self):
        out = []
        for v in range(self.counts.shape[0]):
            n_out, n_pos = int(self.counts[v, 0]), int(self.counts[v, 1])
            s = slice(v * self.cap, v * self.cap + n_out)
            out.append(None if n_pos == 0 else (self.quint[s], self.scores[s], self.spans[s], self.qids[s]))
        return out


class BIG_C(object):
    """Drop-in for the reference ``BIG_C`` in inference mode (see module docstring)."""

    variant = "vidvrd"

    def __init__(self, config: dict, is_train: bool = False, precision: str = "tf32+bf16x2"):
        if is_train:
            raise NotImplementedError("vidsgg_big_b200.BIG_C covers the inference hot path only (is_train=False)")
        self.is_train = False
        self.config = dict(config)
        c = config
        self.num_pred_cats, self.num_enti_cats = c["num_pred_cats"], c["num_enti_cats"]
        self.dim_feat, self.dim_clsme = c["dim_feat"], c["dim_clsme"]
        self.dim_enti, self.dim_pred, self.dim_att, self.dim_ffn = c["dim_enti"], c["dim_pred"], c["dim_att"], c["dim_ffn"]
        self.enco_pool_len, self.n_enco_layers, self.n_deco_layers = c["enco_pool_len"], c["n_enco_layers"], c["n_deco_layers"]
        self.n_att_head, self.num_querys = c["n_att_head"], c["num_querys"]
        self.num_anchors = self.num_querys
        if self.num_enti_cats > 4096 or self.num_pred_cats > 256:
            raise VsgError("triplet sort key: num_enti_cats <= 4096 and num_pred_cats <= 256 (category ids index tables of that size)")
        if not (self.dim_enti == self.dim_pred == self.dim_att):
            raise VsgError("kernels assume dim_enti == dim_pred == dim_att (true for every reference config)")
        if self.dim_enti not in (64, 128, 512):
            raise VsgError("role-attention kernel is instantiated for dim 64 / 128 / 512")
        if self.dim_enti // self.n_att_head not in (16, 32, 64):
            raise VsgError("attention kernel is instantiated for head_dim 16 / 32 / 64")
        self.precision = precision
        self.mode = linalg.MODES[precision]
        # decoder self-attention: "tc" = ONE fused tcgen05 kernel (64-wide heads; csrc/attn_tc.cu mha64_tc_kernel), else as "tc_gemm";
        # "tc_gemm" = batched tcgen05 GEMMs + softmax / transpose glue through HBM; "simt" = fp32 SIMT kernel
        self.attention = "tc"
        self.role_fold = True      # first fc_rolewise layer folded into the role attention (csrc/bigc.cu role_attention_hid_kernel)
        self.backend = "c"         # forward_packed: "c" = ONE call of vsg_bigc_forward (csrc/forward.cu), "py" = the same launches from Python
        self.topk = 10
        self.device = None
        self._w = None
        self._state: Dict[str, torch.Tensor] = {}
        self._init_variant(c)
        # the reference loads these two tables from .npy files at construction (model_0v10.py:266, :283)
        for key, path in (("EntiNameEmb", c.get("EntiNameEmb_path")), ("bias_matrix", c.get("bias_matrix_path"))):
            if path is not None and isinstance(path, str) and path.endswith(".npy"):
                self._state[key] = torch.from_numpy(np.load(path)).float()

    def _init_variant(self, c):
        self.dim_i3d = c.get("dim_i3d", None)
        self.has_entiemb = True
        self.extra_width = self.dim_i3d or 0
        self.dim_z = self.dim_pred + 2 * self.dim_clsme + (4 if self.dim_i3d else 2) * self.dim_enti

    # ---- nn.Module-like surface -------------------------------------------------------------------
    def eval(self):
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else
                                    (device.index if isinstance(device, torch.device) else int(device))))

    def to(self, device):
        self.device = torch.device(device)
        if self._state:
            self._prepare()
        return self

    def state_dict(self):
        return dict(self._state)

    def _expected_keys_static(self) -> List[str]:
        k = []
        if self.has_entiemb:
            k.append("EntiNameEmb")
        k += ["pos_embedding", "pred_query_init", "bias_matrix"]
        for n in ("fc_feat2enti.0", "fc_feat2enti.2", "fc_bbox2enti.0", "fc_bbox2enti.2", "conv_feat2enti",
                  "fc_enti2enco.0", "fc_enti2enco.2"):
            k += [n + ".weight", n + ".bias"]
        if self.variant == "vidvrd" and self.dim_i3d:
            k += ["fc_i3d.0.weight", "fc_i3d.0.bias"]
        for i in range(self.n_enco_layers):
            p = "encoder_layers.%d." % i
            k += [p + "self_attn.in_proj_weight", p + "self_attn.in_proj_bias", p + "self_attn.out_proj.weight",
                  p + "self_attn.out_proj.bias"]
            for n in ("linear1", "linear2", "norm1", "norm2"):
                k += [p + n + ".weight", p + n + ".bias"]
        for i in range(self.n_deco_layers):
            p = "decoder_layers.%d." % i
            k += [p + "self_attn.in_proj_weight", p + "self_attn.in_proj_bias", p + "self_attn.out_proj.weight",
                  p + "self_attn.out_proj.bias"]
            for n in ("fc_rolewise.0.0", "fc_rolewise.0.2", "fc_rolewise.1.0", "fc_rolewise.1.2", "fc_enti2att", "fc_pred2att",
                      "fc2.0", "fc2.3", "norm1", "norm2", "norm3"):
                k += [p + n + ".weight", p + n + ".bias"]
        if self.variant == "vidor":
            k += ["fc_pred2logits.0.weight", "fc_pred2logits.0.bias", "fc_pred2logits.2.weight", "fc_pred2logits.2.bias"]
        else:
            k += ["fc_pred2logits.weight", "fc_pred2logits.bias"]
        return k

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}   # tools/eval_vidor.py:59-68
        if strict:
            want = set(self._expected_keys_static())
            missing, unexpected = sorted(want - set(sd)), sorted(set(sd) - want)
            if missing or unexpected:
                raise RuntimeError("Error(s) in loading state_dict for BIG_C: missing %s unexpected %s" % (missing, unexpected))
        self._state = {k: v.detach().float().cpu() for k, v in sd.items()}
        if self.device is not None:
            self._prepare()
        return self

    # ---- weight staging in HBM --------------------------------------------------------------------
    def _prepare(self):
        dev = self.device
        if dev.type != "cuda":
            raise VsgError("BIG_C runs on a CUDA device only (no CPU fallback)")
        st = {k: v.to(dev) for k, v in self._state.items()}
        E, Pd, Q = self.dim_enti, self.dim_pred, self.num_querys
        split = linalg.SPLITS.get(self.mode, False)
        W = lambda name: Weight(st[name + ".weight"], st[name + ".bias"], split=split)
        w = {}
        w["bbox1_w"], w["bbox1_b"] = st["fc_bbox2enti.0.weight"].contiguous(), st["fc_bbox2enti.0.bias"].contiguous()
        w["bbox2"], w["feat1"], w["feat2"] = W("fc_bbox2enti.2"), W("fc_feat2enti.0"), W("fc_feat2enti.2")
        cw = st["conv_feat2enti.weight"]                                  # [E, 2E, 3] -> tap-major [3E, 2E]
        w["conv"] = Weight(cw.permute(2, 0, 1).reshape(3 * E, 2 * E).contiguous(), None, split=split)
        w["conv_b"] = st["conv_feat2enti.bias"].contiguous()
        w["enco1"], w["enco2"] = W("fc_enti2enco.0"), W("fc_enti2enco.2")
        if self.variant == "vidvrd" and self.dim_i3d:
            w["i3d"] = W("fc_i3d.0")
        w["enc"] = []
        for i in range(self.n_enco_layers):
            p = "encoder_layers.%d." % i
            w["enc"].append(dict(
                qkv=Weight(st[p + "self_attn.in_proj_weight"], st[p + "self_attn.in_proj_bias"], split=split),
                out=W(p + "self_attn.out_proj"), l1=W(p + "linear1"), l2=W(p + "linear2"),
                n1=(st[p + "norm1.weight"].contiguous(), st[p + "norm1.bias"].contiguous()),
                n2=(st[p + "norm2.weight"].contiguous(), st[p + "norm2.bias"].contiguous())))
        pos = st["pos_embedding"].contiguous()
        w["pos"], w["query_init"] = pos, st["pred_query_init"].contiguous()
        w["qk_init"] = (w["query_init"] + pos).contiguous()
        w["dec"] = []
        for i in range(self.n_deco_layers):
            p = "decoder_layers.%d." % i
            # q = k = (query + pos) W_qk^T, v = query W_v^T (model_0v10.py:181-183): two GEMMs on two inputs; `query + pos` comes out of
            # the previous layer's last LayerNorm as a second output (a row-periodic bias in ONE qkv GEMM made its epilogue the bottleneck)
            ipw, ipb = st[p + "self_attn.in_proj_weight"], st[p + "self_attn.in_proj_bias"]
            qk = Weight(ipw[:2 * Pd].contiguous(), ipb[:2 * Pd].contiguous(), split=split)
            vw = Weight(ipw[2 * Pd:].contiguous(), ipb[2 * Pd:].contiguous(), split=split)
            r2 = Weight(torch.cat([st[p + "fc_rolewise.0.2.weight"], st[p + "fc_rolewise.1.2.weight"]], 1).contiguous(),
                        st[p + "fc_rolewise.0.2.bias"] + st[p + "fc_rolewise.1.2.bias"], split=split)
            w["dec"].append(dict(
                qk=qk, v=vw, out=W(p + "self_attn.out_proj"),
                p2a=W(p + "fc_pred2att"), e2a=W(p + "fc_enti2att"),
                r1=(W(p + "fc_rolewise.0.0"), W(p + "fc_rolewise.1.0")), r2=r2,
                f1=W(p + "fc2.0"), f2=W(p + "fc2.3"),
                n1=(st[p + "norm1.weight"].contiguous(), st[p + "norm1.bias"].contiguous()),
                n2=(st[p + "norm2.weight"].contiguous(), st[p + "norm2.bias"].contiguous()),
                n3=(st[p + "norm3.weight"].contiguous(), st[p + "norm3.bias"].contiguous())))
        # role_fold: fc_rolewise[r].0 (att[r] @ enco) == att[r] @ (enco W_r^T): the track-side projections of every decoder layer
        # ([fc_enti2att ; fc_rolewise.0.0 ; fc_rolewise.1.0] x n_dec layers) become ONE GEMM over the encoder output
        if self.role_fold and self.dim_enti in (128, 512) and w["dec"]:
            ws, bs = [], []
            for i in range(self.n_deco_layers):
                p = "decoder_layers.%d." % i
                ws += [st[p + "fc_enti2att.weight"], st[p + "fc_rolewise.0.0.weight"], st[p + "fc_rolewise.1.0.weight"]]
                bs += [st[p + "fc_enti2att.bias"], torch.zeros(2 * Pd, device=dev)]
                w["dec"][i]["b1"] = torch.cat([st[p + "fc_rolewise.0.0.bias"], st[p + "fc_rolewise.1.0.bias"]]).contiguous()
            w["eg_all"] = Weight(torch.cat(ws, 0).contiguous(), torch.cat(bs).contiguous(), split=split)
        w["bias_matrix"] = st["bias_matrix"].reshape(self.num_enti_cats * self.num_enti_cats, self.num_pred_cats).contiguous()
        if self.has_entiemb:
            w["entiemb"] = st["EntiNameEmb"].contiguous()
        if self.variant == "vidor":
            w["log1"], w["log2"] = W("fc_pred2logits.0"), W("fc_pred2logits.2")
        else:
            w["log"] = W("fc_pred2logits")
        self._w = w
        self._cw = None            # VsgBigCWeights of the C entry point, built on first use

    # ---- the whole forward as one C call (include/vsg_b200.h: vsg_bigc_forward) ---------------------------
    def _c_weights(self):
        from ._cabi import VsgBigCWeights, VsgLinear, VsgNorm, VSG_MAX_LAYERS
        if self._cw is not None:
            return self._cw
        w = self._w
        if self.n_enco_layers > VSG_MAX_LAYERS or self.n_deco_layers > VSG_MAX_LAYERS:
            raise VsgError("vsg_bigc_forward holds at most %d layers" % VSG_MAX_LAYERS)
        addr = lambda t: None if t is None else t.data_ptr()

        def lin(W):
            l = VsgLinear()
            l.w, l.bias, l.N, l.K, l.ldw = addr(W.w), addr(W.bias), W.N, W.K, W.w.stride(0)
            l.hi, l.lo, l.w16, l.lo16, l.ld16 = addr(W.hi), addr(W.lo), addr(W.w16), addr(W.lo16), getattr(W, "ld16", 0)
            l.img, l.img_bn = addr(W.img), W.img_bn
            l.img16, l.img16_bn, l.alpha = addr(W.img16), W.img16_bn, W.alpha
            return l

        def norm(n):
            v = VsgNorm()
            v.gamma, v.beta = addr(n[0]), addr(n[1])
            return v
        c = VsgBigCWeights()
        c.variant = 1 if self.variant == "vidor" else 0
        c.dim_enti, c.dim_pred, c.dim_feat, c.dim_clsme, c.dim_i3d = self.dim_enti, self.dim_pred, self.dim_feat, self.dim_clsme, self.dim_i3d or 0
        c.num_querys, c.num_pred_cats, c.num_enti_cats = self.num_querys, self.num_pred_cats, self.num_enti_cats
        c.pool_len, c.n_enc, c.n_dec, c.n_head = self.enco_pool_len, self.n_enco_layers, self.n_deco_layers, self.n_att_head
        c.use_clsme, c.has_entiemb = int(bool(getattr(self, "use_clsme", True))), int(bool(self.has_entiemb))
        c.extra_width, c.dim_z, c.tc_attention = self.extra_width, self.dim_z, {"tc": 1, "tc_gemm": 2}.get(self.attention, 0)
        c.bbox1_w, c.bbox1_b, c.conv_b = addr(w["bbox1_w"]), addr(w["bbox1_b"]), addr(w["conv_b"])
        for k in ("bbox2", "feat1", "feat2", "conv", "enco1", "enco2"):
            setattr(c, k, lin(w[k]))
        for k in ("i3d", "log", "log1", "log2"):
            if k in w:
                setattr(c, k, lin(w[k]))
        for i, lw in enumerate(w["enc"]):
            e = c.enc[i]
            e.qkv, e.out, e.l1, e.l2, e.n1, e.n2 = lin(lw["qkv"]), lin(lw["out"]), lin(lw["l1"]), lin(lw["l2"]), norm(lw["n1"]), norm(lw["n2"])
        for i, lw in enumerate(w["dec"]):
            d = c.dec[i]
            d.qk, d.v, d.out, d.p2a, d.e2a = lin(lw["qk"]), lin(lw["v"]), lin(lw["out"]), lin(lw["p2a"]), lin(lw["e2a"])
            d.r1_0, d.r1_1, d.r2, d.f1, d.f2 = lin(lw["r1"][0]), lin(lw["r1"][1]), lin(lw["r2"]), lin(lw["f1"]), lin(lw["f2"])
            d.n1, d.n2, d.n3 = norm(lw["n1"]), norm(lw["n2"]), norm(lw["n3"])
            d.r1_bias = addr(lw.get("b1"))
        c.pos, c.query_init, c.qk_init, c.bias_matrix = addr(w["pos"]), addr(w["query_init"]), addr(w["qk_init"]), addr(w["bias_matrix"])
        c.entiemb = addr(w.get("entiemb"))
        c.role_fold = int("eg_all" in w)
        if "eg_all" in w:
            c.eg_all = lin(w["eg_all"])
        self._cw = c
        return c

    def _forward_c(self, pk: "PackedVideos", topk: int, sync: bool):
        """``_encode2decode`` + ``_construct_triplets`` through vsg_bigc_forward: one ctypes call, workspace and outputs owned here."""
        from ._cabi import VsgTripletOut, VsgVideoBatch
        dev = self.device
        if pk.feats.shape[1] < self.dim_feat + self.extra_width:
            raise VsgError("features have %d columns, model needs %d" % (pk.feats.shape[1], self.dim_feat + self.extra_width))
        cw = self._c_weights()
        cw.tc_attention = {"tc": 1, "tc_gemm": 2}.get(self.attention, 0)
        b = VsgVideoBatch()
        b.n_videos, b.n_tracks, b.max_tracks, b.n_rows = pk.V, pk.N, pk.max_tracks, pk.R
        b.boxes, b.feats, b.ld_feats = pk.boxes.data_ptr(), pk.feats.data_ptr(), pk.feats.stride(0)
        b.feats_bf16 = int(pk.feats.dtype == torch.bfloat16)
        if b.feats_bf16 and self.mode != linalg.BF16:
            raise VsgError("bf16 features are the opt-in transport of precision='bf16'; the fp32-class modes take fp32 features")
        b.off, b.seg, b.seg64, b.tmax, b.track_vid, b.wh = (t.data_ptr() for t in (pk.off, pk.seg, pk.seg64, pk.tmax, pk.track_vid, pk.wh))
        b.dura, b.cat_ids, b.scores = pk.dura.data_ptr(), pk.cat_ids.data_ptr(), pk.scores.data_ptr()
        b.mha_blk_seg, b.mha_blk_q0, b.n_mha_blk = pk.mha_blocks[0].data_ptr(), pk.mha_blocks[1].data_ptr(), pk.mha_blocks[2]
        V, cap = pk.V, self.num_querys * topk
        quint = torch.empty(V * cap, 5, dtype=torch.long, device=dev)
        scores = torch.empty(V * cap, 3, dtype=torch.float32, device=dev)
        spans = torch.empty(V * cap, 2, dtype=torch.long, device=dev)
        qids = torch.empty(V * cap, dtype=torch.long, device=dev)
        counts = torch.empty(V, 2, dtype=torch.int32, device=dev)
        o = VsgTripletOut()
        o.quint, o.scores, o.spans, o.qids, o.counts, o.cap = quint.data_ptr(), scores.data_ptr(), spans.data_ptr(), qids.data_ptr(), counts.data_ptr(), cap
        need = int(lib().vsg_bigc_workspace_bytes(C.byref(cw), C.byref(b), topk, self.mode))
        if need < 0:
            check(-1, "vsg_bigc_workspace_bytes")
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        base = (ws.data_ptr() + 255) // 256 * 256
        check(lib().vsg_bigc_forward(C.byref(cw), C.byref(b), C.byref(o), topk, self.mode, C.c_void_p(base), need - (base - ws.data_ptr()),
                                     stream_ptr(dev)), "vsg_bigc_forward")
        return PackedTriplets(quint, scores, spans, qids, counts.cpu().numpy() if sync else None, cap, counts_dev=counts)

    # ---- kernels ------------------------------------------------------------------------------------
    def _add_ln(self, x, a, norm, post=None, period=0, dual=False):
        out = torch.empty_like(x)
        if dual:        # -> (LN(x + a), LN(x + a) + post)
            out2 = torch.empty_like(x)
            check(lib().vsg_add_layernorm_dual(_raw(x), x.stride(0), _raw(a), 0 if a is None else a.stride(0), _raw(norm[0]), _raw(norm[1]),
                                               _raw(post), period, x.shape[0], x.shape[1], _raw(out), out.stride(0), _raw(out2),
                                               out2.stride(0), stream_ptr(x.device)), "vsg_add_layernorm_dual")
            return out, out2
        check(lib().vsg_add_layernorm(_raw(x), x.stride(0), _raw(a), 0 if a is None else a.stride(0), _raw(norm[0]), _raw(norm[1]),
                                      _raw(post), period, x.shape[0], x.shape[1], _raw(out), out.stride(0), stream_ptr(x.device)),
              "vsg_add_layernorm")
        return out

    def _mha(self, qkv, d, seg_off, n_seg, fixed_len, max_len, blocks=None):
        out = torch.empty(qkv.shape[0], d, dtype=torch.float32, device=qkv.device)
        ld = qkv.stride(0)
        bs, bq, nb = blocks if blocks is not None else (None, None, 0)
        check(lib().vsg_mha(_raw(qkv), ld, C.c_void_p(qkv.data_ptr() + 4 * d), ld, C.c_void_p(qkv.data_ptr() + 8 * d), ld,
                            _raw(seg_off), n_seg, fixed_len, max_len, self.n_att_head, d // self.n_att_head, _raw(out), d,
                            _raw(bs), _raw(bq), nb, stream_ptr(qkv.device)), "vsg_mha")
        return out

    def _mha_tc64(self, qkv, d, n_seg, Q):
        """Self-attention of ``n_seg`` fixed-length segments of ``Q`` rows with 64-wide heads as one fused tcgen05 launch."""
        out = torch.empty(qkv.shape[0], d, dtype=torch.float32, device=qkv.device)
        ld = qkv.stride(0)
        with linalg._Profile.span("mha", 4.0 * n_seg * Q * Q * d, (n_seg, Q, d)):
            check(lib().vsg_mha_tc64(_raw(qkv), ld, C.c_void_p(qkv.data_ptr() + 4 * d), ld, C.c_void_p(qkv.data_ptr() + 8 * d), ld, None, n_seg, Q,
                                     self.n_att_head, _raw(out), d, None, None, 0, 3 if self.mode in linalg.FP32_CLASS else 1,
                                     stream_ptr(qkv.device)), "vsg_mha_tc64")
        return out

    def _mha_tc(self, qkv, qkv_lo, n_seg, Q, d):
        """Self-attention of ``n_seg`` fixed-length segments (the Q decoder queries of every video) on the tensor cores:
        S = Q K^T and O = P V are batched tcgen05 GEMMs (one problem per (segment, head)), softmax / V^T are glue kernels."""
        m, dev = linalg.attention_mode(self.mode), qkv.device
        H = self.n_att_head
        dh = d // H
        rows = n_seg * Q
        L = lib()
        sp = stream_ptr(dev)
        S = torch.empty(n_seg * H * Q, Q, dtype=torch.float32, device=dev)
        linalg.gemm_batched(m, qkv[:, :d], qkv[:, d:2 * d], None if qkv_lo is None else qkv_lo[:, d:2 * d], Q, Q, dh, S, Q,
                            n_seg * H, H, a_off=(Q, 0, 0, dh), b_off=(Q, 0, 0, dh), c_off=(H * Q * Q, Q * Q))
        check(L.vsg_softmax_rows(_raw(S), Q, Q, S.shape[0], float(1.0 / math.sqrt(dh)), sp), "vsg_softmax_rows")
        ldt = (rows + 3) // 4 * 4
        vt_hi = torch.empty(d, ldt, dtype=torch.float32, device=dev)
        vt_lo = torch.empty(d, ldt, dtype=torch.float32, device=dev) if m == linalg.X3TF32 else None
        check(L.vsg_transpose_split(C.c_void_p(qkv.data_ptr() + 8 * d), qkv.stride(0), rows, d, _raw(vt_hi), _raw(vt_lo), ldt, sp),
              "vsg_transpose_split")
        att = torch.empty(rows, d, dtype=torch.float32, device=dev)
        linalg.gemm_batched(m, S, vt_hi[:, :rows], None if vt_lo is None else vt_lo[:, :rows], Q, dh, Q, att, d, n_seg * H, H,
                            a_off=(H * Q, Q, 0, 0), b_off=(0, dh, Q, 0), c_off=(Q * d, dh))
        return att

    def _track_encoding(self, pk: PackedVideos):
        """Per-track encoding shared by BIG-C and Base-C (model_0v10.py:446-458 / model_pairwise_baseline.py:170-185): per-frame MLPs
        on the unique frames, stretched conv / max-pool, fc_enti2enco; plus the stretched time-mean of the extra feature columns.
        Needs ``self._w`` (bbox1_w, bbox1_b, bbox2, feat1, feat2, conv, conv_b, enco1, enco2), ``mode``, ``dim_enti``, ``dim_feat``,
        ``extra_width``, ``enco_pool_len``, ``device``.  Returns (enti2enco [N, E], extra [N, extra_width] or None)."""
        w, m, dev = self._w, self.mode, self.device
        E, F_in = self.dim_enti, self.dim_feat
        R, N = pk.R, pk.N
        sp = stream_ptr(dev)
        L = lib()
        if pk.feats.shape[1] < F_in + self.extra_width:
            raise VsgError("features have %d columns, model needs %d" % (pk.feats.shape[1], F_in + self.extra_width))
        if m == linalg.BF16:
            return self._track_encoding_bf16(pk)
        if pk.feats.dtype != torch.float32:
            raise VsgError("bf16 features are the opt-in transport of precision='bf16'; the fp32-class modes take fp32 features")
        # --- per-frame MLPs on the unique frames, written into the two halves of X [R, 2E]
        X = torch.empty(R, 2 * E, dtype=torch.float32, device=dev)
        h = torch.empty(R, E, dtype=torch.float32, device=dev)
        check(L.vsg_bbox_feat_mlp1(_raw(pk.boxes), _raw(pk.off), N, R, _raw(pk.track_vid), _raw(pk.wh), _raw(w["bbox1_w"]),
                                   _raw(w["bbox1_b"]), E, _raw(h), E, None, sp), "vsg_bbox_feat_mlp1")
        dbg = getattr(self, "_dbg", None)
        if dbg is not None:
            dbg["bbox_h1"] = h.clone()
        gemm(m, h, w["bbox2"], out=X[:, :E], relu=True)
        gemm(m, pk.feats, w["feat1"], out=h, relu=True, K=F_in)
        gemm(m, h, w["feat2"], out=X[:, E:], relu=True)
        # --- conv taps (one GEMM, N = 3E) + stretched conv / max-pool
        Y = gemm(m, X, w["conv"], bias=False)
        if dbg is not None:
            dbg["X"], dbg["Y"] = X.clone(), Y.clone()
        del X, h
        pooled = torch.empty(N, E * self.enco_pool_len, dtype=torch.float32, device=dev)
        check(L.vsg_conv_pool(_raw(Y), Y.stride(0), E, _raw(w["conv_b"]), _raw(pk.off), _raw(pk.tmax), N, self.enco_pool_len,
                              _raw(pooled), sp), "vsg_conv_pool")
        del Y
        if dbg is not None:
            dbg["pooled"] = pooled.clone()
        enti2enco = gemm(m, gemm(m, pooled, w["enco1"], relu=True), w["enco2"], relu=True)
        # --- stretched time-mean of the extra columns (I3D / classeme)
        extra = None
        if self.extra_width:
            extra = torch.empty(N, self.extra_width, dtype=torch.float32, device=dev)
            check(L.vsg_stretched_mean(_raw(pk.feats), pk.feats.stride(0), F_in, self.extra_width, _raw(pk.off), _raw(pk.tmax), N,
                                       _raw(extra), self.extra_width, sp), "vsg_stretched_mean")
        return enti2enco, extra

    def _track_encoding_bf16(self, pk: PackedVideos):
        """``_track_encoding`` in the bf16 mode: the R-row chain (h, X, Y: ~12 KB of fp32 per box-frame in the fp32-class modes) lives in
        HBM as bf16 -- every GEMM's epilogue writes the bf16 operand of the next one, the raw fp32 features are cast once."""
        w, m, dev = self._w, self.mode, self.device
        E, F_in = self.dim_enti, self.dim_feat
        R, N = pk.R, pk.N
        sp = stream_ptr(dev)
        L = lib()
        bf = torch.bfloat16
        X16 = torch.empty(R, 2 * E, dtype=bf, device=dev)
        h16 = torch.empty(R, E, dtype=bf, device=dev)
        check(L.vsg_bbox_feat_mlp1_bf16(_raw(pk.boxes), _raw(pk.off), N, R, _raw(pk.track_vid), _raw(pk.wh), _raw(w["bbox1_w"]),
                                        _raw(w["bbox1_b"]), E, _raw(h16), E, sp), "vsg_bbox_feat_mlp1_bf16")
        gemm(m, h16, w["bbox2"], out16=X16[:, :E], relu=True, f32_out=False)
        F16 = pk.feats if pk.feats.dtype == bf else linalg.cast_bf16(pk.feats, K=F_in)      # bf16 transport: no cast pass
        gemm(m, F16, w["feat1"], out16=h16, relu=True, f32_out=False, K=F_in)
        del F16
        gemm(m, h16, w["feat2"], out16=X16[:, E:], relu=True, f32_out=False)
        Y16 = gemm(m, X16, w["conv"], bias=False, f32_out=False)
        del X16, h16
        pooled = torch.empty(N, E * self.enco_pool_len, dtype=torch.float32, device=dev)
        check(L.vsg_conv_pool_bf16(_raw(Y16), Y16.stride(0), E, _raw(w["conv_b"]), _raw(pk.off), _raw(pk.tmax), N, self.enco_pool_len,
                                   _raw(pooled), sp), "vsg_conv_pool_bf16")
        del Y16
        enti2enco = gemm(m, gemm(m, pooled, w["enco1"], relu=True), w["enco2"], relu=True)
        extra = None
        if self.extra_width:
            extra = torch.empty(N, self.extra_width, dtype=torch.float32, device=dev)
            fn = L.vsg_stretched_mean_bf16 if pk.feats.dtype == bf else L.vsg_stretched_mean
            check(fn(_raw(pk.feats), pk.feats.stride(0), F_in, self.extra_width, _raw(pk.off), _raw(pk.tmax), N,
                     _raw(extra), self.extra_width, sp), "vsg_stretched_mean")
        return enti2enco, extra

    def _encode2decode(self, pk: PackedVideos, want_att: bool = False):
        """model_0v10.py:434-475 for the whole batch.  Returns (logits [V*Q, P], so int32[V*Q,2], extras)."""
        w, m, dev = self._w, self.mode, self.device
        E, Pd, Q, F_in = self.dim_enti, self.dim_pred, self.num_querys, self.dim_feat
        R, N, V = pk.R, pk.N, pk.V
        sp = stream_ptr(dev)
        L = lib()
        dbg = getattr(self, "_dbg", None)
        enti2enco, extra = self._track_encoding(pk)
        # --- encoder (post-norm, tokens = tracks of a video)
        x = enti2enco
        for lw in w["enc"]:
            qkv = gemm(m, x, lw["qkv"])
            if self.attention == "tc" and m != linalg.SIMT and E // self.n_att_head == 64:
                att = torch.empty(qkv.shape[0], E, dtype=torch.float32, device=dev)       # ragged tracks-per-video sequences on the fused tcgen05 kernel
                bs, bq, nb = pk.mha_blocks
                check(L.vsg_mha_tc64(_raw(qkv), 3 * E, C.c_void_p(qkv.data_ptr() + 4 * E), 3 * E, C.c_void_p(qkv.data_ptr() + 8 * E), 3 * E,
                                     _raw(pk.seg64), V, 0, self.n_att_head, _raw(att), E, _raw(bs), _raw(bq), nb,
                                     3 if m in linalg.FP32_CLASS else 1, sp), "vsg_mha_tc64")
            else:
                att = self._mha(qkv, E, pk.seg64, V, 0, pk.max_tracks, pk.mha_blocks)
            x = self._add_ln(x, gemm(m, att, lw["out"]), lw["n1"])
            x = self._add_ln(x, gemm(m, gemm(m, x, lw["l1"], relu=True), lw["l2"]), lw["n2"])
            if dbg is not None:
                dbg.setdefault("enc_layers", []).append(x.clone())
        enco = x
        # --- decoder
        VQ = V * Q
        so = torch.empty(VQ, 2, dtype=torch.int32, device=dev)
        att_out = torch.zeros(VQ, 2, max(pk.max_tracks, 1), dtype=torch.float32, device=dev) if want_att else None
        fold = "eg_all" in w
        values = None if fold else torch.empty(VQ, 2 * E, dtype=torch.float32, device=dev)
        hid = torch.empty(VQ, 2 * Pd, dtype=torch.float32, device=dev)
        EG = gemm(m, enco, w["eg_all"]) if fold else None              # [N, n_dec * 3E]: per layer [e2a | G_subject | G_object]
        n_dec = len(w["dec"])

        def bcast(x):
            out = torch.empty(VQ, x.shape[1], dtype=torch.float32, device=dev)
            check(L.vsg_broadcast_rows(_raw(x), Q, x.shape[1], VQ, _raw(out), sp), "vsg_broadcast_rows")
            return out

        query = query_pos = None
        for li, lw in enumerate(w["dec"]):
            last = li == n_dec - 1
            # layer 0: the queries of every video are still pred_query_init, so its self-attention block and
            # fc_pred2att are video-independent -- computed once on Q rows, then broadcast
            x = w["query_init"] if li == 0 else query
            x_qk = w["qk_init"] if li == 0 else query_pos
            nv = 1 if li == 0 else V
            use_tc = self.attention in ("tc", "tc_gemm") and m != linalg.SIMT and (Pd // self.n_att_head) % 32 == 0 and Q % 32 == 0
            use_fused = self.attention == "tc" and m != linalg.SIMT and Pd // self.n_att_head == 64
            qkv = torch.empty(x.shape[0], 3 * Pd, dtype=torch.float32, device=dev)
            if use_fused:
                gemm(m, x_qk, lw["qk"], out=qkv[:, :2 * Pd])
                gemm(m, x, lw["v"], out=qkv[:, 2 * Pd:])
                att = self._mha_tc64(qkv, Pd, nv, Q)
            elif use_tc:
                qkv_lo = torch.empty_like(qkv) if linalg.attention_mode(m) == linalg.X3TF32 else None
                gemm(m, x_qk, lw["qk"], out=qkv[:, :2 * Pd], out_lo=None if qkv_lo is None else qkv_lo[:, :2 * Pd],
                     lo_cols=(Pd, 2 * Pd))                                              # only K's low part is read
                gemm(m, x, lw["v"], out=qkv[:, 2 * Pd:])
                att = self._mha_tc(qkv, qkv_lo, nv, Q, Pd)
            else:
                gemm(m, x_qk, lw["qk"], out=qkv[:, :2 * Pd])
                gemm(m, x, lw["v"], out=qkv[:, 2 * Pd:])
                att = self._mha(qkv, Pd, None, nv, Q, Q)
            x = self._add_ln(x, gemm(m, att, lw["out"]), lw["n1"], post=w["pos"], period=Q)
            p2a = gemm(m, x, lw["p2a"])
            if li == 0:
                query, p2a = bcast(x), bcast(p2a)
            else:
                query = x
            if fold:
                e2a, G = EG[:, li * 3 * E:li * 3 * E + E], EG[:, li * 3 * E + E:(li + 1) * 3 * E]
                check(L.vsg_role_attention_hid(_raw(p2a), _raw(e2a), EG.stride(0), _raw(G), EG.stride(0), _raw(lw["b1"]), _raw(pk.seg), V, Q, E,
                                               pk.max_tracks, float(1.0 / np.sqrt(self.dim_enti)), _raw(hid),
                                               _raw(att_out) if last else None, 0 if att_out is None else att_out.shape[2],
                                               _raw(so) if last else None, sp), "vsg_role_attention_hid")
            else:
                e2a = gemm(m, enco, lw["e2a"])
                check(L.vsg_role_attention(_raw(p2a), _raw(e2a), _raw(enco), _raw(pk.seg), V, Q, E, pk.max_tracks,
                                           float(1.0 / np.sqrt(self.dim_enti)), _raw(values),
                                           _raw(att_out) if last else None, 0 if att_out is None else att_out.shape[2],
                                           _raw(so) if last else None, sp), "vsg_role_attention")
                gemm(m, values[:, :E], lw["r1"][0], out=hid[:, :Pd], relu=True)
                gemm(m, values[:, E:], lw["r1"][1], out=hid[:, Pd:], relu=True)
            query = self._add_ln(query, gemm(m, hid, lw["r2"]), lw["n2"])
            ffn = gemm(m, gemm(m, query, lw["f1"], relu=True), lw["f2"])
            if last:
                query = self._add_ln(query, ffn, lw["n3"])
            else:                                                                       # also emit query + pos for the next layer's q / k
                query, query_pos = self._add_ln(query, ffn, lw["n3"], post=w["pos"], period=Q, dual=True)
            if dbg is not None:
                dbg.setdefault("dec_layers", []).append(query.clone())
        logits = self._prediction_head(pk, query, so, enti2enco, extra)
        return logits, so, dict(query=query, att=att_out, enti2enco=enti2enco, enco=enco, extra=extra)

    def _concat(self, pieces, rows, ldo):
        n = len(pieces)
        src = (_P * n)(*[C.c_void_p(p[0].data_ptr()) for p in pieces])
        idx = (_P * n)(*[None if p[1] is None else C.c_void_p(p[1].data_ptr()) for p in pieces])
        istr = (C.c_int * n)(*[2 for _ in pieces])
        ld = (C.c_int * n)(*[int(p[0].stride(0)) for p in pieces])
        wd = (C.c_int * n)(*[int(p[2]) for p in pieces])
        out = torch.empty(rows, ldo, dtype=torch.float32, device=self.device)
        check(lib().vsg_gather_concat(src, idx, istr, ld, wd, n, rows, _raw(out), ldo, stream_ptr(self.device)), "vsg_gather_concat")
        return out

    def _so_cats(self, pk, so):
        VQ = so.shape[0]
        pair_index = torch.empty(VQ, dtype=torch.int32, device=self.device)
        so_cat = torch.empty(VQ, 2, dtype=torch.int32, device=self.device)
        check(lib().vsg_so_category(_raw(so), _raw(pk.cat_ids), self.num_enti_cats, VQ, _raw(pair_index), _raw(so_cat),
                                    stream_ptr(self.device)), "vsg_so_category")
        return pair_index, so_cat

    def _prediction_head(self, pk, query, so, enti_feat, extra):
        """model_0v10.py:478-507."""
        w, m = self._w, self.mode
        E, Pd = self.dim_enti, self.dim_pred
        VQ = query.shape[0]
        pair_index, so_cat = self._so_cats(pk, so)
        s_idx, o_idx = so, so[:, 1:]                      # strided views: element r*2 (+1)
        sc_idx, oc_idx = so_cat, so_cat[:, 1:]
        emb = w["entiemb"]
        if self.dim_i3d:
            i3d = gemm(m, extra, w["i3d"], relu=True)   # fc_i3d commutes with the row gather
            pieces = [(query, None, Pd), (i3d, s_idx, E), (i3d, o_idx, E), (enti_feat, s_idx, E), (enti_feat, o_idx, E),
                      (emb, sc_idx, self.dim_clsme), (emb, oc_idx, self.dim_clsme)]
        else:
            pieces = [(query, None, Pd), (emb, sc_idx, self.dim_clsme), (emb, oc_idx, self.dim_clsme),
                      (enti_feat, s_idx, E), (enti_feat, o_idx, E)]
        ldz = (self.dim_z + 3) // 4 * 4
        Z = self._concat(pieces, VQ, ldz)
        return gemm(m, Z, w["log"], rowbias=w["bias_matrix"], rb_index=pair_index, K=self.dim_z)

    def _construct_triplets(self, pk, logits, so, topk, packed=False, sync=True):
        """model_0v10.py:707-785 for every video; one D2H read of the per-video counts."""
        V, Q = pk.V, self.num_querys
        cap = Q * topk
        dev = self.device
        quint = torch.empty(V * cap, 5, dtype=torch.long, device=dev)
        scores = torch.empty(V * cap, 3, dtype=torch.float32, device=dev)
        spans = torch.empty(V * cap, 2, dtype=torch.long, device=dev)
        qids = torch.empty(V * cap, dtype=torch.long, device=dev)
        counts = torch.empty(V, 2, dtype=torch.int32, device=dev)
        check(lib().vsg_construct_triplet(_raw(logits), logits.stride(0), self.num_pred_cats, Q, topk, _raw(so), _raw(pk.seg), V,
                                          _raw(pk.dura), _raw(pk.cat_ids), _raw(pk.scores), _raw(quint), _raw(scores), _raw(spans),
                                          _raw(qids), _raw(counts), cap, stream_ptr(dev)), "vsg_construct_triplet")
        if packed:
            return PackedTriplets(quint, scores, spans, qids, counts.cpu().numpy() if sync else None, cap, counts_dev=counts)
        cnt = counts.cpu().tolist()
        out = []
        for v in range(V):
            n_out, n_pos = cnt[v]
            if n_pos == 0:
                out.append(None)                         # no overlapping pair (model_0v10.py:733-734)
                continue
            s = slice(v * cap, v * cap + n_out)
            out.append((quint[s], scores[s], spans[s], qids[s]))
        return out

    # ---- public API -----------------------------------------------------------------------------------
    def forward(self, proposal_list, gt_graph_list=None, topk=None, max_rows: int = 2_000_000):
        if self._w is None:
            raise VsgError("BIG_C has no weights on a CUDA device: call load_state_dict(...) and .cuda() first")
        self.topk = self.default_topk if topk is None else topk
        props = [p if isinstance(p, TrajProposal) else TrajProposal.from_reference(p) for p in proposal_list]
        results: List[Optional[tuple]] = [None] * len(props)
        live = [i for i, p in enumerate(props) if p.num_proposals > 0]   # num_proposals == 0 -> None (:377-380)
        # sub-batches bounded by rows so the intermediate buffers stay within budget
        batch, rows = [], 0
        def flush():
            nonlocal batch, rows
            if not batch:
                return
            pk = PackedVideos([props[i] for i in batch], self.device)
            for i, r in zip(batch, self._forward_eager(pk, self.topk, True).per_video()):
                results[i] = r
            batch, rows = [], 0
        for i in live:
            r = int(props[i].lengths.sum())
            if batch and rows + r > max_rows:
                flush()
            batch.append(i)
            rows += r
        flush()
        return results

    __call__ = forward
    default_topk = 10

    def pack(self, proposal_list) -> "PackedVideos":
        """Index arrays / packed views of a batch (reusable across calls while the batch is resident in HBM)."""
        return PackedVideos(proposal_list, self.device)

    def forward_packed(self, proposal_list, topk=None, packed_videos=None, sync=True, graph=False):
        """Same computation as ``forward`` for a batch of non-empty videos, but the result stays packed on the device
        (``PackedTriplets``): no per-video slicing -- the fast path into ``evalapi.PackedRelations``.
        ``graph=True`` (needs ``packed_videos``): the ~115 launches of the forward are captured ONCE per packed batch into a CUDA graph
        and replayed on later calls -- same kernels, same buffers, one launch call instead of ~150 Python-issued ones (the decoder's
        launches are shorter than the time Python needs to issue them).  The batch's input buffers must stay at their addresses."""
        if self._w is None:
            raise VsgError("BIG_C has no weights on a CUDA device: call load_state_dict(...) and .cuda() first")
        self.topk = self.default_topk if topk is None else topk
        assert all(p.num_proposals > 0 for p in proposal_list)
        pk = packed_videos if packed_videos is not None else PackedVideos(proposal_list, self.device)
        if graph and packed_videos is not None:
            return self._forward_graph(pk, self.topk, sync)
        return self._forward_eager(pk, self.topk, sync)

    def _forward_eager(self, pk, topk, sync):
        if self.backend == "c" and getattr(self, "_dbg", None) is None:
            return self._forward_c(pk, topk, sync)
        logits, so, _ = self._encode2decode(pk)
        return self._construct_triplets(pk, logits, so, topk, packed=True, sync=sync)

    def _forward_graph(self, pk: PackedVideos, topk: int, sync: bool):
        key = (id(self), topk, self.mode, self.attention, self.backend)
        graphs = pk.__dict__.setdefault("_graphs", {})
        if key not in graphs:
            # one eager pass first: lazy per-device initialisation (function attributes, tensor-map cache) must not happen under capture
            self._forward_eager(pk, topk, False)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._forward_eager(pk, topk, False)
            graphs[key] = (g, out)
        g, out = graphs[key]
        g.replay()
        # the graph's output buffers are overwritten by the next replay: hand out copies (a few MB, stream-ordered)
        res = PackedTriplets(out.quint.clone(), out.scores.clone(), out.spans.clone(), out.qids.clone(), None, out.cap,
                             counts_dev=out._counts_dev.clone())
        if sync:
            res.counts
        return res

    def bipartite_cost(self, pred_logit, gt_pred, att_matrx, gt_adj_enti_align):
        """Cost matrix of ``bipartite_match`` (model_0v10.py:606-636) on the device: f32[n_querys, n_gt_pred]."""
        require_cuda(pred_logit, gt_pred, att_matrx, gt_adj_enti_align)
        Q, P = pred_logit.shape
        G, n = gt_adj_enti_align.shape[1], gt_adj_enti_align.shape[2]
        assert att_matrx.shape == (2, Q, n) and gt_pred.shape[0] == G
        cf = self.config.get("cost_coeff_dict", dict(classification=1.0, adj_matrix=30.0))
        cost = torch.empty(Q, G, dtype=torch.float32, device=pred_logit.device)
        check(lib().vsg_bipartite_cost(_raw(pred_logit.float().contiguous()), Q, P, _raw(gt_pred.long().contiguous()), G,
                                       _raw(att_matrx.float().contiguous()), _raw(gt_adj_enti_align.float().contiguous()), n,
                                       float(cf["classification"]), float(cf["adj_matrix"]), _raw(cost), stream_ptr(pred_logit.device)),
              "vsg_bipartite_cost")
        return cost

    def bipartite_match(self, pred_logit, gt_pred, att_matrx, gt_adj_enti_align):
        """model_0v10.py:606-639: the cost matrix comes from the device, the assignment from scipy on the host (as in the reference)."""
        from scipy.optimize import linear_sum_assignment
        return linear_sum_assignment(self.bipartite_cost(pred_logit, gt_pred, att_matrx, gt_adj_enti_align).cpu())

    def forward_debug(self, proposal):
        """(pred_queries, pred_logits [Q,P], att_matrx [2,Q,n]) of one video, like ``encode2decode`` (:434-475)."""
        pk = PackedVideos([proposal], self.device)
        logits, so, ex = self._encode2decode(pk, want_att=True)
        n = proposal.num_proposals
        att = ex["att"][:, :, :n].permute(1, 0, 2).contiguous()
        return ex["query"], logits[:, :self.num_pred_cats], att, so, ex


class BIG_C_vidvrd(BIG_C):
    variant = "vidvrd"


class BIG_C_vidor(BIG_C):
    """models/model_0v7.py: no I3D branch, optional classeme (from features or from EntiNameEmb), 2-layer classifier,
    default topk=3."""
    variant = "vidor"
    default_topk = 3

    def _init_variant(self, c):
        self.dim_i3d = None
        self.use_clsme = c["use_clsme"]
        self.has_entiemb = c.get("EntiNameEmb_path", None) is not None
        self.extra_width = self.dim_clsme if (self.use_clsme and not self.has_entiemb) else 0
        self.dim_z = self.dim_pred + 2 * self.dim_enti + (2 * self.dim_clsme if self.use_clsme else 0)

    def _prediction_head(self, pk, query, so, enti_feat, extra):
        """model_0v7.py:483-513."""
        w, m = self._w, self.mode
        E, Pd = self.dim_enti, self.dim_pred
        VQ = query.shape[0]
        pair_index, so_cat = self._so_cats(pk, so)
        s_idx, o_idx = so, so[:, 1:]
        if self.use_clsme:
            if self.has_entiemb:
                cl = [(w["entiemb"], so_cat, self.dim_clsme), (w["entiemb"], so_cat[:, 1:], self.dim_clsme)]
            else:
                cl = [(extra, s_idx, self.dim_clsme), (extra, o_idx, self.dim_clsme)]
            pieces = [(query, None, Pd)] + cl + [(enti_feat, s_idx, E), (enti_feat, o_idx, E)]
        else:
            pieces = [(query, None, Pd), (enti_feat, s_idx, E), (enti_feat, o_idx, E)]
        ldz = (self.dim_z + 3) // 4 * 4
        Z = self._concat(pieces, VQ, ldz)
        hid = gemm(m, Z, w["log1"], relu=True, K=self.dim_z)
        return gemm(m, hid, w["log2"], rowbias=w["bias_matrix"], rb_index=pair_index)
