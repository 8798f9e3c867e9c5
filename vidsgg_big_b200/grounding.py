"""Grounding stage on B200: host-side mirror of the reference ``DEBUG`` module (models/grd_model_v5.py:140-737,
exported as ``DEBUG`` by models/__init__.py:4) in inference mode.

    model = DEBUG(config, is_train=False); model.load_state_dict(sd); model.cuda()
    pooled_se, bins_probs, bins_mask = model(video_feature_list, data_list, score_th, tiou_th, bins_th, nms_th,
                                             with_gt_data=False)

with ``data_list[i] = (quintuples i64[m,5], spans i64[m,2], video_len)`` -- the call made by
tools/eval_vidor.py:239-242.  Returns ``f32[m,k+1,2]`` (normalised 0..1), ``f32[m,k+1]``, ``bool[m,k+1]`` or
``(None, None)``.  Unlike the reference (which asserts a batch of one, :211) any number of videos can be passed;
then a list of triples is returned.  All compute runs in libvsgb200 kernels (csrc/gemm.cu, csrc/grounding.cu,
csrc/bigc.cu MHA / LayerNorm); inference only.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import linalg
from ._cabi import VsgError, check, lib, stream_ptr
from .linalg import Weight, gemm


def _raw(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def expand_after_grounding(quintuples, cls_scores3, pooled_se, bins_probs, bins_mask, video_len):
    """Driver-side expansion of tools/eval_vidor.py:245-253 (tiny index ops on the kernel outputs):
    score = mean(cls scores) * bin prob, span = round(pooled * video_len) as int64, rows selected by the mask."""
    nb = bins_probs.shape[1]
    q = quintuples[:, None, :].repeat(1, nb, 1)[bins_mask, :]
    s = (cls_scores3.mean(-1)[:, None] * bins_probs)[bins_mask]
    sp = torch.round((pooled_se * video_len)[bins_mask, :]).type(torch.long)
    return q, s, sp


class DEBUG(object):
    def __init__(self, config: dict, is_train: bool = False, precision: str = "tf32+bf16x2"):
        if is_train:
            raise NotImplementedError("vidsgg_big_b200.DEBUG covers the inference hot path only (is_train=False)")
        self.is_train = False
        self.config = dict(config)
        self.dim_feat, self.dim_clsme = config["dim_feat"], config["dim_clsme"]
        self.dim_hidden, self.num_bins = config["dim_hidden"], config["num_bins"]
        if self.dim_hidden != 128:
            raise VsgError("context-query kernel is instantiated for dim_hidden == 128 (the reference value)")
        self.precision, self.mode = precision, linalg.MODES[precision]
        self.device = None
        self._w = None
        self._state: Dict[str, torch.Tensor] = {}
        for key in ("EntiNameEmb", "PredNameEmb"):
            path = config.get(key + "_path")
            if isinstance(path, str) and path.endswith(".npy"):
                self._state[key] = torch.from_numpy(np.load(path)).float()

    # ---- nn.Module-like surface ---------------------------------------------------------------------
    def eval(self):
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else
                                    (device.index if isinstance(device, torch.device) else int(device))))

    def to(self, device):
        self.device = torch.device(device)
        if len(self._state) > 2:
            self._prepare()
        return self

    def state_dict(self):
        return dict(self._state)

    def _expected_keys(self) -> List[str]:
        k = ["EntiNameEmb", "PredNameEmb"]
        for n in ("video_fc", "query_fc", "temp_fc", "vq_fc"):
            k += [n + ".weight", n + ".bias"]
        for enc in ("video_encoder", "query_encoder", "combined_encoder"):
            for c in range(4):
                for part in ("depth_wise", "point_wise"):
                    k += ["%s.convs.%d.%s.weight" % (enc, c, part), "%s.convs.%d.%s.bias" % (enc, c, part)]
            k += [enc + ".mh_attn.in_proj_weight", enc + ".mh_attn.in_proj_bias", enc + ".mh_attn.out_proj.weight",
                  enc + ".mh_attn.out_proj.bias", enc + ".fc.weight", enc + ".fc.bias", enc + ".normb.weight", enc + ".normb.bias"]
            for c in range(4):
                k += ["%s.norm_seq.%d.weight" % (enc, c), "%s.norm_seq.%d.bias" % (enc, c)]
            k += [enc + ".norme.weight", enc + ".norme.bias"]
        k.append("proj2sim.weight")
        for head in ("cls_head", "conf_head", "regr_head"):
            for c in range(4):
                for part in ("depth_wise", "point_wise"):
                    k += ["%s.%d.0.%s.weight" % (head, c, part), "%s.%d.0.%s.bias" % (head, c, part)]
            for part in ("depth_wise", "point_wise"):
                k += ["%s.4.%s.weight" % (head, part), "%s.4.%s.bias" % (head, part)]
        return k

    def load_state_dict(self, state_dict, strict: bool = True):
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        if strict:
            want = set(self._expected_keys())
            missing, unexpected = sorted(want - set(sd)), sorted(set(sd) - want)
            if missing or unexpected:
                raise RuntimeError("Error(s) in loading state_dict for DEBUG: missing %s unexpected %s" % (missing, unexpected))
        self._state = {k: v.detach().float().cpu() for k, v in sd.items()}
        if self.device is not None:
            self._prepare()
        return self

    # ---- weights in HBM ---------------------------------------------------------------------------------
    def _prepare(self):
        dev = self.device
        if dev.type != "cuda":
            raise VsgError("DEBUG runs on a CUDA device only (no CPU fallback)")
        st = {k: v.to(dev) for k, v in self._state.items()}
        H = self.dim_hidden
        split = linalg.SPLITS.get(self.mode, False)
        W = lambda name: Weight(st[name + ".weight"], st[name + ".bias"], split=split)
        w = {"video_fc": W("video_fc"), "vq_fc": W("vq_fc")}
        qfc = W("query_fc")
        # query_fc is linear, so project the two embedding tables once instead of every query word
        w["proj_enti"] = gemm(self.mode, st["EntiNameEmb"].contiguous(), qfc)
        w["proj_pred"] = gemm(self.mode, st["PredNameEmb"].contiguous(), qfc)
        w["temp_w"], w["temp_b"] = st["temp_fc.weight"].contiguous(), st["temp_fc.bias"].contiguous()
        w["proj2sim"] = Weight(st["proj2sim.weight"], None, split=split)

        def dws(prefix):
            dw = st[prefix + ".depth_wise.weight"]
            pw = st[prefix + ".point_wise.weight"]
            return dict(dw_w=dw.reshape(dw.shape[0], dw.shape[2]).contiguous(), dw_b=st[prefix + ".depth_wise.bias"].contiguous(),
                        k=int(dw.shape[2]), pw=Weight(pw.reshape(pw.shape[0], pw.shape[1]).contiguous(), st[prefix + ".point_wise.bias"], split=split))

        def norm(name):
            return st[name + ".weight"].contiguous(), st[name + ".bias"].contiguous()

        for enc in ("video_encoder", "query_encoder", "combined_encoder"):
            w[enc] = dict(convs=[dws("%s.convs.%d" % (enc, c)) for c in range(4)],
                          qkv=Weight(st[enc + ".mh_attn.in_proj_weight"], st[enc + ".mh_attn.in_proj_bias"], split=split),
                          out=W(enc + ".mh_attn.out_proj"), fc=W(enc + ".fc"), normb=norm(enc + ".normb"),
                          norm_seq=[norm("%s.norm_seq.%d" % (enc, c)) for c in range(4)], norme=norm(enc + ".norme"))
        for head in ("cls_head", "conf_head", "regr_head"):
            w[head] = [dws("%s.%d.0" % (head, c)) for c in range(4)] + [dws("%s.4" % head)]
        # PosEncoder tables exactly as the reference builds them (python doubles -> float32 tensor, :61-64)
        freqs = [10000 ** (-i / H) if i % 2 == 0 else -10000 ** ((1 - i) / H) for i in range(H)]
        phases = [0 if i % 2 == 0 else np.pi / 2 for i in range(H)]
        w["freq"], w["phase"] = torch.Tensor(freqs).to(dev), torch.Tensor(phases).to(dev)
        self._w = w
        self._cw = None

    # ---- the whole forward as one C call (include/vsg_b200.h: vsg_grd_forward) ----------------------------------
    def _c_weights(self):
        from ._cabi import VsgGrdConv, VsgGrdWeights, VsgLinear, VsgNorm
        if self._cw is not None:
            return self._cw
        w = self._w
        addr = lambda t: None if t is None else t.data_ptr()

        def lin(W):
            l = VsgLinear()
            l.w, l.bias, l.N, l.K, l.ldw = addr(W.w), addr(W.bias), W.N, W.K, W.w.stride(0)
            l.hi, l.lo, l.w16, l.lo16, l.ld16 = addr(W.hi), addr(W.lo), addr(W.w16), addr(W.lo16), getattr(W, "ld16", 0)
            l.img, l.img_bn = addr(W.img), W.img_bn
            l.img16, l.img16_bn, l.alpha = addr(W.img16), W.img16_bn, W.alpha
            return l

        def conv(dst, cw):
            dst.dw_w, dst.dw_b, dst.k, dst.pw = addr(cw["dw_w"]), addr(cw["dw_b"]), cw["k"], lin(cw["pw"])

        def norm(dst, n):
            dst.gamma, dst.beta = addr(n[0]), addr(n[1])
        c = VsgGrdWeights()
        c.dim_hidden, c.num_bins, c.dim_feat = self.dim_hidden, self.num_bins, self.dim_feat
        c.video_fc, c.vq_fc, c.proj2sim = lin(w["video_fc"]), lin(w["vq_fc"]), lin(w["proj2sim"])
        c.proj_enti, c.proj_pred = addr(w["proj_enti"]), addr(w["proj_pred"])
        c.temp_w, c.temp_b, c.freq, c.phase = addr(w["temp_w"]), addr(w["temp_b"]), addr(w["freq"]), addr(w["phase"])
        for name in ("video_encoder", "query_encoder", "combined_encoder"):
            e, ew = getattr(c, name), w[name]
            for i in range(4):
                conv(e.convs[i], ew["convs"][i])
                norm(e.norm_seq[i], ew["norm_seq"][i])
            e.qkv, e.out, e.fc = lin(ew["qkv"]), lin(ew["out"]), lin(ew["fc"])
            norm(e.normb, ew["normb"]); norm(e.norme, ew["norme"])
        for name in ("cls_head", "conf_head", "regr_head"):
            for i in range(5):
                conv(getattr(c, name)[i], w[name][i])
        self._cw = c
        return c

    @staticmethod
    def _c_seq(sq):
        from ._cabi import VsgGrdSeq
        q = VsgGrdSeq()
        q.off, q.n, q.rows, q.pos, q.rem, q.max_len = sq["off"].data_ptr(), sq["n"], sq["rows"], sq["pos"].data_ptr(), sq["rem"].data_ptr(), sq["max_len"]
        q.blk_seg, q.blk_q0, q.n_blk = sq["blocks"][0].data_ptr(), sq["blocks"][1].data_ptr(), sq["blocks"][2]
        if sq["tc_blocks"] is not None:
            q.tc_blk_seg, q.tc_blk_q0, q.n_tc_blk = sq["tc_blocks"][0].data_ptr(), sq["tc_blocks"][1].data_ptr(), sq["tc_blocks"][2]
        else:
            q.n_tc_blk = -1
        return q

    # ---- building blocks --------------------------------------------------------------------------------
    def _ln(self, x, norm):
        out = torch.empty_like(x)
        check(lib().vsg_add_layernorm(_raw(x), x.stride(0), None, 0, _raw(norm[0]), _raw(norm[1]), None, 0, x.shape[0], x.shape[1],
                                      _raw(out), out.stride(0), stream_ptr(x.device)), "vsg_add_layernorm")
        return out

    def _dwconv(self, x, cw, pos, rem):
        out = torch.empty_like(x)
        check(lib().vsg_dwconv(_raw(x), _raw(pos), _raw(rem), _raw(cw["dw_w"]), _raw(cw["dw_b"]), cw["k"], x.shape[0], x.shape[1],
                               _raw(out), stream_ptr(x.device)), "vsg_dwconv")
        return out

    backend = "c"            # _forward_videos: "c" = ONE call of vsg_grd_forward (csrc/forward.cu), "py" = the same launches from Python
    attention = "tc"         # mh_attn of the video / combined encoders: "tc" = tcgen05 kernel (csrc/attn_tc.cu), "simt" = fp32 SIMT kernel
    fuse_dwconv = True       # depthwise conv inside the point-wise GEMM's operand pipeline (False: separate vsg_dwconv launch)

    def _dw_pw(self, x, cw, sq, relu=False, residual=None):
        """DepthWiseSeparableConv1d (:36-56): depthwise conv over the sequence axis, then the 1x1 conv (+ ReLU, + residual)."""
        if self.fuse_dwconv and linalg.can_fuse_dwconv(self.mode, cw["pw"], k=cw["k"]):
            return gemm(self.mode, x, cw["pw"], relu=relu, residual=residual, dwconv=(cw["dw_w"], cw["dw_b"], cw["k"], sq["pos"], sq["rem"]))
        return gemm(self.mode, self._dwconv(x, cw, sq["pos"], sq["rem"]), cw["pw"], relu=relu, residual=residual)

    def _seq(self, seq_off_host: np.ndarray):
        dev = self.device
        seq_off = torch.from_numpy(np.ascontiguousarray(seq_off_host.astype(np.int64))).to(dev)
        rows = int(seq_off_host[-1])
        pos = torch.empty(max(rows, 1), dtype=torch.int32, device=dev)
        rem = torch.empty(max(rows, 1), dtype=torch.int32, device=dev)
        check(lib().vsg_seq_positions(_raw(seq_off), len(seq_off_host) - 1, rows, _raw(pos), _raw(rem), stream_ptr(dev)), "vsg_seq_positions")
        lens = np.diff(seq_off_host)
        return dict(off=seq_off, n=len(seq_off_host) - 1, rows=rows, pos=pos, rem=rem, max_len=int(lens.max()) if lens.size else 0,
                    blocks=linalg.mha_block_list(lens, dev),
                    # sequences long enough to fill tensor-core tiles (the 3-word query sequences stay on the SIMT kernel)
                    tc_blocks=linalg.mha_block_list(lens, dev, qb=128) if (lens.size and float(lens.mean()) >= 24) else None,
                    att_flops=4.0 * self.dim_hidden * float((lens.astype(np.float64) ** 2).sum()))     # QK^T + PV over all heads

    def _qanet(self, ew, x, sq):
        """QANetEncoderLayer.forward (:110-137) on rows [rows, H] of ragged sequences."""
        w, m, H = self._w, self.mode, self.dim_hidden
        res = torch.empty_like(x)
        out = torch.empty_like(x)
        check(lib().vsg_pos_add_ln(_raw(x), _raw(sq["pos"]), _raw(w["freq"]), _raw(w["phase"]), _raw(ew["normb"][0]), _raw(ew["normb"][1]),
                                   x.shape[0], H, _raw(res), _raw(out), stream_ptr(x.device)), "vsg_pos_add_ln")
        for i in range(4):
            res = self._dw_pw(out, ew["convs"][i], sq, relu=True, residual=res)      # relu(conv) + res  (:120-122)
            out = self._ln(res, ew["norm_seq"][i])
        qkv = gemm(m, out, ew["qkv"])
        att = torch.empty(x.shape[0], H, dtype=torch.float32, device=x.device)
        ld = qkv.stride(0)
        with linalg._Profile.span("mha", sq["att_flops"]):
            if self.attention == "tc" and m != linalg.SIMT and sq["tc_blocks"] is not None:
                # tcgen05 attention (csrc/attn_tc.cu): fp32-class modes split every operand 3xTF32-style, the reduced modes run one tf32 pass
                products = 3 if m in linalg.FP32_CLASS else 1
                bs, bq, nb = sq["tc_blocks"]
                check(lib().vsg_mha_tc16(_raw(qkv), ld, C.c_void_p(qkv.data_ptr() + 4 * H), ld, C.c_void_p(qkv.data_ptr() + 8 * H), ld,
                                         _raw(sq["off"]), 8, _raw(att), H, _raw(bs), _raw(bq), nb, products, stream_ptr(x.device)), "vsg_mha_tc16")
            else:
                check(lib().vsg_mha(_raw(qkv), ld, C.c_void_p(qkv.data_ptr() + 4 * H), ld, C.c_void_p(qkv.data_ptr() + 8 * H), ld, _raw(sq["off"]),
                                    sq["n"], 0, sq["max_len"], 8, H // 8, _raw(att), H, _raw(sq["blocks"][0]), _raw(sq["blocks"][1]), sq["blocks"][2],
                                    stream_ptr(x.device)), "vsg_mha")
        res = gemm(m, att, ew["out"], residual=res)                                 # attn + res (:129-130)
        out = self._ln(res, ew["norme"])
        return gemm(m, out, ew["fc"], relu=True, residual=res)                      # relu(fc(LN)) + res (:133-136)

    def _head(self, hw, x, sq):
        y = x
        for c in range(4):
            y = self._dw_pw(y, hw[c], sq, relu=True)
        return self._dw_pw(y, hw[4], sq)

    # ---- batched forward ----------------------------------------------------------------------------------
    def _forward_videos(self, feats: Sequence[torch.Tensor], datas: Sequence[tuple], th, want_net: bool = False):
        w, m, H, B, dev = self._w, self.mode, self.dim_hidden, self.num_bins, self.device
        L = lib()
        sp = stream_ptr(dev)
        T = [int(f.shape[0]) for f in feats]
        nq = [int(d[0].shape[0]) for d in datas]
        NQ = sum(nq)
        vid_off_h = np.concatenate([[0], np.cumsum(T)]).astype(np.int64)
        q_vid_h = np.repeat(np.arange(len(T)), nq).astype(np.int32)
        comb_off_h = np.concatenate([[0], np.cumsum(np.repeat(T, nq))]).astype(np.int64)
        vf = feats[0].to(dev, torch.float32).contiguous() if len(feats) == 1 else torch.cat([f.to(dev, torch.float32) for f in feats], 0)
        quint = torch.cat([d[0].to(dev, torch.long) for d in datas], 0).contiguous()
        spans = torch.cat([d[1].to(dev, torch.long) for d in datas], 0).contiguous()
        vlen = torch.tensor([float(d[2]) for d in datas], dtype=torch.float32, device=dev)
        q_vid = torch.from_numpy(q_vid_h).to(dev)
        clip = torch.cat([torch.linspace(0, 1, t) for t in T]).to(dev)          # (:705) host-evaluated table, see grounding.cu
        sv, sq3, sc = self._seq(vid_off_h), self._seq(np.arange(NQ + 1, dtype=np.int64) * 3), self._seq(comb_off_h)
        if self.backend == "c":
            return self._forward_c(vf, quint, spans, vlen, q_vid, clip, sv, sq3, sc, T, nq, NQ, th, want_net)
        # --- embeddings
        v0 = gemm(m, vf, w["video_fc"])
        q0 = torch.empty(3 * NQ, H, dtype=torch.float32, device=dev)
        so_norm = torch.empty(NQ, 2, dtype=torch.float32, device=dev)
        check(L.vsg_grd_query_init(_raw(quint), _raw(spans), _raw(vlen), _raw(q_vid), NQ, _raw(w["proj_enti"]), _raw(w["proj_pred"]),
                                   _raw(w["temp_w"]), _raw(w["temp_b"]), H, _raw(q0), _raw(so_norm), sp), "vsg_grd_query_init")
        v = self._qanet(w["video_encoder"], v0, sv)
        q = self._qanet(w["query_encoder"], q0, sq3)
        # --- context-query attention -> vq_fc -> combined encoder
        pv = gemm(m, v, w["proj2sim"], bias=False)
        comb_in = torch.empty(sc["rows"], 4 * H, dtype=torch.float32, device=dev)
        check(L.vsg_cq_attention(_raw(v), _raw(pv), _raw(q), _raw(sv["off"]), _raw(q_vid), _raw(sc["off"]), NQ, H, max(T), _raw(comb_in), sp),
              "vsg_cq_attention")
        comb = self._qanet(w["combined_encoder"], gemm(m, comb_in, w["vq_fc"]), sc)
        del comb_in
        regr, conf, cls = self._head(w["regr_head"], comb, sc), self._head(w["conf_head"], comb, sc), self._head(w["cls_head"], comb, sc)
        out = self._post(regr, conf, cls, sc["off"], so_norm, clip, sv["off"], q_vid, NQ, nq, th, regr_activated=False)
        if want_net:
            return out, (regr, conf, cls, so_norm)
        return out

    def _forward_c(self, vf, quint, spans, vlen, q_vid, clip, sv, sq3, sc, T, nq, NQ, th, want_net):
        """The network + post-processing of ``_forward_videos`` through vsg_grd_forward (one ctypes call)."""
        from ._cabi import VsgGrdBatch, VsgGrdOut
        dev, B = self.device, self.num_bins
        cw = self._c_weights()
        cw.tc_attention, cw.fuse_dwconv = int(self.attention == "tc"), int(bool(self.fuse_dwconv))
        b = VsgGrdBatch()
        b.n_videos, b.n_queries, b.max_T = len(T), NQ, max(T)
        b.video_feats, b.quint, b.spans, b.vlen = vf.data_ptr(), quint.data_ptr(), spans.data_ptr(), vlen.data_ptr()
        b.q_vid, b.clip_tab = q_vid.data_ptr(), clip.data_ptr()
        b.video, b.query, b.combined = self._c_seq(sv), self._c_seq(sq3), self._c_seq(sc)
        pooled = torch.empty(NQ, B + 1, 2, dtype=torch.float32, device=dev)
        probs = torch.empty(NQ, B + 1, dtype=torch.float32, device=dev)
        mask = torch.empty(NQ, B + 1, dtype=torch.uint8, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        o = VsgGrdOut()
        o.pooled, o.probs, o.mask, o.err_count = pooled.data_ptr(), probs.data_ptr(), mask.data_ptr(), err.data_ptr()
        net = None
        if want_net:
            rows_c = sc["rows"]
            net = (torch.empty(rows_c, 2 * B, dtype=torch.float32, device=dev), torch.empty(rows_c, B, dtype=torch.float32, device=dev),
                   torch.empty(rows_c, B, dtype=torch.float32, device=dev))
            o.regr, o.conf, o.cls = (t.data_ptr() for t in net)
        need = int(lib().vsg_grd_workspace_bytes(C.byref(cw), C.byref(b), self.mode))
        if need < 0:
            check(-1, "vsg_grd_workspace_bytes")
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        base = (ws.data_ptr() + 255) // 256 * 256
        check(lib().vsg_grd_forward(C.byref(cw), C.byref(b), C.byref(o), float(th[0]), float(th[1]), float(th[2]), float(th[3]), self.mode,
                                    C.c_void_p(base), need - (base - ws.data_ptr()), stream_ptr(dev)), "vsg_grd_forward")
        if int(err.item()) > 0:
            raise RuntimeError("temporal_pooling: %d (query, bin) pairs have an empty pooling set" % int(err.item()))
        out, r = [], 0
        for n in nq:
            out.append((pooled[r:r + n], probs[r:r + n], mask[r:r + n].bool()))
            r += n
        if want_net:
            so_norm = spans.float() / vlen[q_vid.long()][:, None]         # == the kernel's so_norm (span / video_len, fp32)
            return out, (net[0], net[1], net[2], so_norm)
        return out

    def _post(self, regr, conf, cls, comb_off, so_norm, clip, vid_off, q_vid, NQ, nq, th, regr_activated):
        """Everything after the network (:533-576) in one kernel."""
        L, dev, B = lib(), self.device, self.num_bins
        sp = stream_ptr(dev)
        pooled = torch.empty(NQ, B + 1, 2, dtype=torch.float32, device=dev)
        probs = torch.empty(NQ, B + 1, dtype=torch.float32, device=dev)
        mask = torch.empty(NQ, B + 1, dtype=torch.uint8, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        check(L.vsg_grounding_post(_raw(regr), _raw(conf), _raw(cls), _raw(comb_off), _raw(so_norm), _raw(clip), _raw(vid_off), _raw(q_vid),
                                   NQ, B, 1 if regr_activated else 0, float(th[0]), float(th[1]), float(th[2]), float(th[3]), _raw(pooled), _raw(probs), _raw(mask),
                                   _raw(err), sp), "vsg_grounding_post")
        if int(err.item()) > 0:
            # the reference raises in temporal_pooling when no clip survives (:726, min() of an empty tensor)
            raise RuntimeError("temporal_pooling: %d (query, bin) pairs have an empty pooling set" % int(err.item()))
        out, r = [], 0
        for n in nq:
            out.append((pooled[r:r + n], probs[r:r + n], mask[r:r + n].bool()))
            r += n
        return out

    def postprocess(self, regrs, conf_logits, cls_logits, so_norm, th=(0.9, 0.5, 0.2, 0.8)):
        """Post-network stage alone for one video, on ``forward_propagation``-style outputs
        (regrs f32[nq,T,2B] AFTER the sigmoid, logits f32[nq,T,B], so_norm f32[nq,2])."""
        dev = self.device
        nq, T, _ = conf_logits.shape
        comb_off = (torch.arange(nq + 1, dtype=torch.long) * T).to(dev)
        vid_off = torch.tensor([0, T], dtype=torch.long, device=dev)
        q_vid = torch.zeros(nq, dtype=torch.int32, device=dev)
        clip = torch.linspace(0, 1, T).to(dev)
        f = lambda t: t.to(dev, torch.float32).reshape(nq * T, -1).contiguous()
        return self._post(f(regrs), f(conf_logits), f(cls_logits), comb_off, so_norm.to(dev, torch.float32).contiguous(), clip, vid_off,
                          q_vid, nq, [nq], th, regr_activated=True)[0]

    def prepare_gt_data(self, gt_graph):
        """grd_model_v5.py:253-308 in inference mode: the queries of a GT graph are its unique
        ``[pred_cat, subj_cat, obj_cat, so_start, so_end]`` tags (``unique_with_idx_nd``), so_* = intersection of the subject / object
        track spans.  Returns ``((tags5, spans, video_len), target, index_map)`` -- the first item is exactly a ``data_list`` entry of the
        ``with_gt_data=False`` call (columns 3, 4 of ``tags5`` are not read by the network); ``target`` = normalised GT predicate spans
        and ``index_map`` = per unique query the GT predicate ids it stands for (the inputs of the reference's ``eval_tiou``)."""
        from . import geometry
        if gt_graph.num_trajs == 0 or gt_graph.num_preds == 0:
            return (None, None, int(gt_graph.video_len)), None, None
        dev = self.device
        video_len = int(gt_graph.video_len)
        traj_cats = gt_graph.traj_cat_ids.to(dev)
        traj_duras = gt_graph.traj_durations.to(dev)
        pred_cats = gt_graph.pred_cat_ids.to(dev)
        so_ids = torch.argmax(gt_graph.adj_matrix.to(dev), dim=-1).t()               # (n_pred, 2)
        so_cats = traj_cats[so_ids]
        inter, _ = geometry.dura_intersection_ts(traj_duras[so_ids[:, 0]], traj_duras[so_ids[:, 1]], broadcast=False)
        tags = torch.cat([pred_cats[:, None], so_cats, inter], dim=-1)               # (n_pred, 5)
        uniq, index_map = geometry.unique_with_idx_nd(tags)
        target = gt_graph.pred_durations.to(dev).float() / video_len
        return (uniq, uniq[:, 3:].contiguous(), video_len), target, index_map

    def forward(self, video_feature_list, data_list, score_th=0.5, tiou_th=0.5, bins_th=0.1, nms_th=0.5, with_gt_data=True,
                max_rows: int = 1_500_000):
        if with_gt_data:
            # grounding stage evaluated alone on the GT queries (grd_model_v5.py:198-202): data_list holds GT graphs
            self.last_gt_targets = []
            datas = []
            for gt in data_list:
                d, target, index_map = self.prepare_gt_data(gt)
                self.last_gt_targets.append((target, index_map))
                datas.append(d)
            data_list = datas
        if self._w is None:
            raise VsgError("DEBUG has no weights on a CUDA device: call load_state_dict(...) and .cuda() first")
        assert len(video_feature_list) == len(data_list)
        single = len(video_feature_list) == 1
        th = (score_th, tiou_th, bins_th, nms_th)
        self.bin_conf_th, self.score_th, self.tiou_th, self.nms_th = bins_th, score_th, tiou_th, nms_th
        results: List[Optional[tuple]] = [None] * len(data_list)
        live = [i for i, d in enumerate(data_list) if d[0] is not None and d[0].shape[0] > 0]
        batch, rows = [], 0

        def flush():
            nonlocal batch, rows
            if batch:
                outs = self._forward_videos([video_feature_list[i] for i in batch], [data_list[i] for i in batch], th)
                for i, o in zip(batch, outs):
                    results[i] = o
            batch, rows = [], 0
        for i in live:
            r = int(video_feature_list[i].shape[0]) * int(data_list[i][0].shape[0])
            if batch and rows + r > max_rows:
                flush()
            batch.append(i)
            rows += r
        flush()
        if single:
            return results[0] if results[0] is not None else (None, None)
        return results

    __call__ = forward

    def forward_packed(self, video_feature_list, data_list, score_th=0.5, tiou_th=0.5, bins_th=0.1, nms_th=0.5,
                       max_rows: int = 1_500_000):
        """Batched fast path: every video must have >= 1 query.  Returns the concatenated
        ``(pooled_se f32[NQ,k+1,2], bins_probs f32[NQ,k+1], bins_mask bool[NQ,k+1])`` in video order."""
        th = (score_th, tiou_th, bins_th, nms_th)
        outs, batch, rows = [], [], 0

        def flush():
            nonlocal batch, rows
            if batch:
                outs.extend(self._forward_videos([video_feature_list[i] for i in batch], [data_list[i] for i in batch], th))
            batch, rows = [], 0
        for i in range(len(data_list)):
            r = int(video_feature_list[i].shape[0]) * int(data_list[i][0].shape[0])
            if batch and rows + r > max_rows:
                flush()
            batch.append(i)
            rows += r
        flush()
        return (torch.cat([o[0] for o in outs], 0), torch.cat([o[1] for o in outs], 0), torch.cat([o[2] for o in outs], 0))

    def forward_propagation_debug(self, video_feature, quintuples, spans, video_len, th=(0.9, 0.5, 0.2, 0.8)):
        """(regrs after sigmoid, conf_logits, cls_logits) as ``forward_propagation`` (:331-373) returns them, for tests."""
        out, (regr, conf, cls, so_norm) = self._forward_videos([video_feature], [(quintuples, spans, video_len)], th, want_net=True)
        nq, T = quintuples.shape[0], video_feature.shape[0]
        B = self.num_bins
        return torch.sigmoid(regr).view(nq, T, 2 * B), conf.view(nq, T, B), cls.view(nq, T, B), so_norm, out[0]
