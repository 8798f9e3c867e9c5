#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (the reference is mounted read-only there):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

For every case the inputs and weights are regenerated from numpy seeds by
``vidsgg_big_b200.synth`` (so they are not stored), the reference code is imported from
``/root/reference`` and executed on them, and ONLY its outputs are stored.  While generating,
the script also runs the oracle on the same inputs and asserts agreement, so a fixture is
never written for a case where oracle and reference disagree.

Reference entry points exercised (paths relative to /root/reference):
  utils/utils_func.py      dura_intersection_ts, vIoU_ts, tIoU, generalized_tIoU, unique_with_idx_nd
  models/model_0v10.py     BIG_C (VidVRD), stack_with_repeat_2d, enti_viou_align
  models/model_0v7.py      BIG_C (VidOR)
  models/model_pairwise_baseline.py  Base_C.trajid2pairid
  models/grd_model_v5.py   DEBUG
  utils/evaluate.py        EvalFmtCvtor.to_eval_format_pr
  VidVRDhelperEvalAPIs     eval_visual_relation, evaluate_v2, common.viou
"""
import io
import os
import sys
import tempfile
import contextlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("VSG_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from vidsgg_big_b200 import synth                                    # noqa: E402
from oracle import geometry as og, evalapi as oe, bigc as ob, grounding as ogr, convert as oc, basec as obc   # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    from utils import utils_func as R                                # noqa: E402  (reference)
    from models import BIG_C_vidvrd, BIG_C_vidor, Base_C, DEBUG      # noqa: E402
    from models.model_0v10 import stack_with_repeat_2d               # noqa: E402
    from VidVRDhelperEvalAPIs import eval_visual_relation            # noqa: E402
    from VidVRDhelperEvalAPIs.visual_relation_detection import evaluate_v2   # noqa: E402
    from VidVRDhelperEvalAPIs.common import viou as ref_viou, voc_ap as ref_voc_ap  # noqa: E402
    from utils.evaluate import EvalFmtCvtor                          # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(8)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def build_ref_bigc(cfg, state, cls):
    tmp = tempfile.mkdtemp()
    cfg = dict(cfg)
    if "EntiNameEmb" in state:
        np.save(os.path.join(tmp, "emb.npy"), state["EntiNameEmb"].numpy())
        cfg["EntiNameEmb_path"] = os.path.join(tmp, "emb.npy")
    np.save(os.path.join(tmp, "bias.npy"), state["bias_matrix"].numpy())
    cfg["bias_matrix_path"] = os.path.join(tmp, "bias.npy")
    with contextlib.redirect_stdout(io.StringIO()):
        m = cls(cfg, is_train=False)
    m.load_state_dict(state, strict=True)      # reference keys / shapes must match exactly
    m.eval()
    return m


# ------------------------------------------------------------------ geometry -----------
def gen_geometry():
    out = {}
    rng = np.random.default_rng(11)
    s = rng.integers(0, 200, size=(23, 1)); d1 = np.concatenate([s, s + rng.integers(0, 120, size=(23, 1))], 1)
    s = rng.integers(0, 200, size=(17, 1)); d2 = np.concatenate([s, s + rng.integers(0, 120, size=(17, 1))], 1)
    t1, t2 = torch.from_numpy(d1), torch.from_numpy(d2)
    inter, mask = R.dura_intersection_ts(t1, t2)
    oi, om = og.dura_intersection(t1, t2)
    assert torch.equal(inter, oi) and torch.equal(mask, om)
    out.update(dura_inter=inter.numpy(), dura_mask=mask.numpy())
    inter_nb, mask_nb = R.dura_intersection_ts(t1[:17], t2, broadcast=False)
    oi, om = og.dura_intersection(t1[:17], t2, broadcast=False)
    assert torch.equal(inter_nb, oi) and torch.equal(mask_nb, om)
    out.update(dura_inter_nb=inter_nb.numpy(), dura_mask_nb=mask_nb.numpy())
    # KAT of SURVEY §8c
    k = torch.tensor([[0, 10], [5, 20], [30, 40]])
    ki, km = R.dura_intersection_ts(k, k)
    assert ki[0, 1].tolist() == [5, 10] and ki[0, 2].tolist() == [30, 10] and not bool(km[0, 2])
    out.update(kat_inter=ki.numpy(), kat_mask=km.numpy())
    # tIoU / gIoU on float spans
    f1, f2 = t1.float() / 320, t2.float() / 320
    out.update(tiou=R.tIoU(f1, f2).numpy(), gtiou=R.generalized_tIoU(f1, f2).numpy())
    assert torch.equal(R.tIoU(f1, f2), og.tiou(f1, f2)) and torch.equal(R.generalized_tIoU(f1, f2), og.generalized_tiou(f1, f2))
    # pair ids
    with contextlib.redirect_stdout(io.StringIO()):
        pid = Base_C.trajid2pairid(None, 7)
    assert torch.equal(pid, og.pair_ids(7))
    out["pair_ids_7"] = pid.numpy()
    # trajectory vIoU matrix: proposals vs proposals and proposals vs GT (reference loop, model_0v10.py:565-581)
    for tag, seed, n, vl in (("a", 101, 12, 160), ("b", 102, 20, 90)):
        P = synth.make_proposal(seed, n, vl, 8, 36, with_features=False)
        G = synth.make_gt_graph(seed, P, 133)
        for side, (bx, du) in (("pp", (P.bboxes_list, P.traj_durations)), ("pg", (G.traj_bboxes, G.traj_durations))):
            inter, mask = R.dura_intersection_ts(P.traj_durations, du)
            rp = inter - P.traj_durations[:, 0, None, None]
            rg = inter - du[None, :, 0, None]
            mat = torch.zeros_like(mask, dtype=torch.float)
            for p, g in zip(*[x.tolist() for x in mask.nonzero(as_tuple=True)]):
                mat[p, g] = R.vIoU_ts(P.bboxes_list[p], bx[g], rp[p, g], rg[p, g])
            om_, oi_, omask_ = og.traj_viou_matrix(P.bboxes_list, P.traj_durations, bx, du)
            assert torch.equal(mat, om_) and torch.equal(inter, oi_) and torch.equal(mask, omask_)
            out["viou_%s_%s" % (side, tag)] = mat.numpy()
            out["inter_%s_%s" % (side, tag)] = inter.numpy()
    # unique_with_idx_nd KAT + random
    kat = torch.tensor([[3, 1, 2, 0, 1], [1, 1, 2, 0, 1], [3, 1, 2, 0, 1], [1, 0, 2, 5, 1]])
    u, groups = R.unique_with_idx_nd(kat)
    ou, ogp = og.unique_rows_with_groups(kat)
    assert torch.equal(u, ou) and all(torch.equal(a, b) for a, b in zip(groups, ogp))
    assert [g.tolist() for g in groups] == [[3], [1], [0, 2]]
    rnd = torch.from_numpy(rng.integers(0, 3, size=(60, 5)))
    u, groups = R.unique_with_idx_nd(rnd)
    ou, ogp = og.unique_rows_with_groups(rnd)
    assert torch.equal(u, ou) and all(torch.equal(a, b) for a, b in zip(groups, ogp))
    out.update(uniq_rows=u.numpy(), uniq_first=np.array([int(g[0]) for g in groups]),
               uniq_counts=np.array([len(g) for g in groups]))
    # stretch index map
    for L, T in ((3, 7), (5, 5), (4, 13), (7, 8), (1, 6), (6, 29)):
        ref = stack_with_repeat_2d([torch.arange(L)[:, None].float(), torch.zeros(T, 1)], dim=0)[0, :, 0].long().numpy()
        assert np.array_equal(ref, og.stretch_index_map(L, T)), (L, T)
        out["stretch_%d_%d" % (L, T)] = ref
    assert out["stretch_3_7"].tolist() == [0, 0, 0, 1, 1, 2, 2]
    save("geometry", **out)


# ------------------------------------------------------------------ eval ---------------
def eval_case(seeds, n_pred=120):
    gts, prs = {}, {}
    en, pn = oc.default_names("e", 64), oc.default_names("p", 200)
    for sd in seeds:
        rng = np.random.default_rng(sd)
        P = synth.make_proposal(sd, int(rng.integers(6, 16)), int(rng.integers(60, 200)), 8, 36, with_features=False)
        G = synth.make_gt_graph(sd, P, 133, n_rel=(3, 25))
        T = synth.make_predictions(sd, P, G, 133, m=n_pred)
        gts.update(oc.to_eval_format_gt(G, en, pn))
        prs.update(oc.to_eval_format_pr(P, T, en, pn))
    return gts, prs


def gen_eval():
    out = {}
    # KATs
    v = ref_viou([[0, 0, 9, 9]] * 4, [0, 4], [[5, 5, 14, 14]] * 3, [2, 5])
    assert v == 50 / 650 and oe.viou([[0, 0, 9, 9]] * 4, [0, 4], [[5, 5, 14, 14]] * 3, [2, 5]) == v
    assert ref_viou([[0, 0, 9, 9]] * 4, [0, 4], [[5, 5, 14, 14]] * 3, [4, 7]) == 0.0
    out["kat_viou"] = np.array([v])
    seeds = list(range(500, 512))
    gts, prs = eval_case(seeds)
    gts["empty_gt_video"] = []                      # skipped by the reference (:73-74)
    prs.pop("synth_%06d" % 505)                     # a video without predictions
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        m_ap, rec, mprec = eval_visual_relation(gts, prs)
        m_ap2, rec2, mprec2, infos = evaluate_v2(gts, prs)
        m_ap3, rec3, _ = eval_visual_relation(gts, prs, viou_threshold=0.7)
    o_ap, o_rec, o_mprec, o_infos = oe.evaluate(gts, prs, with_infos=True)
    assert o_ap == m_ap == m_ap2 and all(o_rec[k] == rec[k] for k in rec) and all(o_mprec[k] == mprec[k] for k in mprec)
    for vid in infos:
        assert np.array_equal(infos[vid][0], o_infos[vid][0]) and np.array_equal(infos[vid][1], o_infos[vid][1])
    vids = sorted(infos)
    out.update(mean_ap=np.array([m_ap]), rec=np.array([rec[50], rec[100]], np.float64),
               mprec=np.array([mprec[1], mprec[5], mprec[10]], np.float64),
               mean_ap_thr07=np.array([m_ap3]), rec_thr07=np.array([rec3[50], rec3[100]], np.float64),
               n_tp=np.array([int(np.isfinite(infos[v][0]).sum()) for v in vids]))
    for v in vids:
        out["hit_" + v] = infos[v][0]
        out["g2d_" + v] = infos[v][1]
    print("eval golden: mAP %.6f R@50 %.4f R@100 %.4f TPs %d" % (m_ap, rec[50], rec[100], out["n_tp"].sum()))
    # per-pair viou values on float(pred) x int(gt) boxes
    vals = []
    for v in vids[:3]:
        for pr in prs.get(v, [])[:15]:
            for gt in gts[v][:6]:
                a = ref_viou(pr["sub_traj"], pr["duration"], gt["sub_traj"], gt["duration"])
                assert a == oe.viou(pr["sub_traj"], pr["duration"], gt["sub_traj"], gt["duration"])
                vals.append(a)
    out["viou_samples"] = np.array(vals)
    # reference convertor == oracle convertor (reference category tables are used only here)
    from utils.categories_v2 import vidvrd_CatId2name, vidvrd_PredId2name
    P = synth.make_proposal(500, 9, 120, 8, 36, with_features=False)
    G = synth.make_gt_graph(500, P, 133)
    T = synth.make_predictions(500, P, G, 133, m=40)
    P.video_name = "ILSVRC2015_train_00005015"
    a = EvalFmtCvtor("vidvrd").to_eval_format_pr(P, T)
    b = oc.to_eval_format_pr(P, T, vidvrd_CatId2name, vidvrd_PredId2name)
    assert a == b, "oracle convertor differs from reference EvalFmtCvtor"
    # voc_ap
    r = np.array([0.1, 0.1, 0.3, 0.3, 0.5]); p = np.array([1.0, 0.5, 0.66, 0.5, 0.6])
    assert ref_voc_ap(r, p) == oe.voc_ap(r, p)
    out["voc_ap"] = np.array([ref_voc_ap(r, p)])
    save("evalapi", **out)


# ------------------------------------------------------------------ BIG-C --------------
def bigc_case(tag, cfg, cls, seeds_shapes, wseed, topk):
    state = synth.make_bigc_state(wseed, cfg)
    model = build_ref_bigc(cfg, state, cls)
    out = {}
    feat_total = cfg["dim_feat"] + (cfg.get("dim_i3d") or 0 if cfg["variant"] == "vidvrd" else cfg["dim_clsme"])
    for sd, n, vl, maxlen in seeds_shapes:
        P = synth.make_proposal(sd, n, vl, feat_total, cfg["num_enti_cats"], max_len=maxlen)
        with torch.no_grad():
            pq, logits, att = model.encode2decode(P)
            model.topk = topk
            ret = model.construct_triplet(P, logits, att)
            ret_fw = model(list([P]), topk=topk)[0]
            oq, ologits, oatt = ob.encode2decode(state, cfg, P)
            oret = ob.construct_triplet(P, ologits, oatt, topk)
        assert (ret is None) == (ret_fw is None)
        err = (ologits - logits).abs().max().item()
        assert err < 2e-4 * max(1.0, logits.abs().max().item()), ("logits", tag, sd, err)
        assert (oatt - att).abs().max().item() < 1e-5
        k = "%s_%d" % (tag, sd)
        out[k + "_logits"] = logits.numpy(); out[k + "_att"] = att.numpy(); out[k + "_query"] = pq.numpy()
        if ret is None:
            out[k + "_none"] = np.array([1]); assert oret is None
            continue
        q5, sc, sp, qi = ret
        assert torch.equal(q5, oret[0]) and torch.equal(sp, oret[2]) and torch.equal(qi, oret[3]), (tag, sd)
        assert torch.allclose(sc, oret[1], atol=1e-5)
        out[k + "_quint"] = q5.numpy(); out[k + "_scores"] = sc.numpy(); out[k + "_spans"] = sp.numpy(); out[k + "_qids"] = qi.numpy()
        print("bigc golden", k, "n=%d" % n, "triplets=%d" % q5.shape[0], "max|dlogit|=%.2e" % err)
    return out


def gen_bigc():
    out = {}
    out.update(bigc_case("tinyvrd", synth.tiny_vidvrd_config(), BIG_C_vidvrd,
                         [(201, 7, 40, None), (202, 12, 64, 30), (203, 1, 20, None), (204, 5, 33, 9)], 7, topk=5))
    out.update(bigc_case("tinyvrd_noi3d", synth.tiny_vidvrd_config(dim_i3d=None), BIG_C_vidvrd,
                         [(205, 6, 50, None)], 8, topk=5))
    out.update(bigc_case("tinyvor", synth.tiny_vidor_config(), BIG_C_vidor,
                         [(211, 9, 55, None), (212, 14, 80, 40)], 9, topk=3))
    out.update(bigc_case("tinyvor_noclsme", synth.tiny_vidor_config(use_clsme=False), BIG_C_vidor,
                         [(213, 8, 45, None)], 10, topk=3))
    out.update(bigc_case("tinyvor_emb", synth.tiny_vidor_config(_force_entiemb=True, EntiNameEmb_path="x"), BIG_C_vidor,
                         [(214, 8, 45, None)], 12, topk=3))
    # full dims (exp2 / exp5)
    out.update(bigc_case("vrd", synth.vidvrd_config(), BIG_C_vidvrd, [(301, 30, 300, 150)], 1, topk=10))
    out.update(bigc_case("vor", synth.vidor_config(), BIG_C_vidor, [(311, 40, 400, 200)], 2, topk=3))
    save("bigc", **out)


# ------------------------------------------------------------------ Base-C --------------
BASEC_CASES = [   # tag, config factory name, overrides, [(seed, n, video_len, max_len)], weight seed, topk
    ("tinybc", "tiny_basec_config", {}, [(701, 7, 40, None), (702, 12, 64, 30), (703, 1, 20, None), (704, 2, 30, 9)], 31, 3),
    ("tinybc_rt", "tiny_basec_config", {"rt_triplets_topk": 25}, [(705, 10, 50, None)], 32, 3),
    ("tinybc_emb", "tiny_basec_config", {"_force_entiemb": True, "EntiNameEmb_path": "x"}, [(706, 8, 45, None)], 33, 3),
    ("bc", "basec_config", {"rt_triplets_topk": 200}, [(711, 40, 400, 200)], 3, 3),
]


def gen_basec():
    """Base_C inference (models/model_pairwise_baseline.py:113-129, 170-273, 314-395)."""
    out = {}
    for tag, mk, over, shapes, wseed, topk in BASEC_CASES:
        cfg = getattr(synth, mk)(**over)
        state = synth.make_basec_state(wseed, cfg)
        model = build_ref_bigc(cfg, state, Base_C)
        feat_total = cfg["dim_feat"] + cfg["dim_clsme"]
        for sd, n, vl, maxlen in shapes:
            P = synth.make_proposal(sd, n, vl, feat_total, cfg["num_enti_cats"], max_len=maxlen)
            with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
                pairs = model.trajid2pairid(n)
                logits = model.forward_propagation(P, pairs)
                ret = model([P], topk=topk)[0]
                ologits = obc.forward_propagation(state, cfg, P, og.pair_ids(n))
                oret = obc.forward(state, cfg, [P], topk)[0]
            k = "%s_%d" % (tag, sd)
            out[k + "_logits"] = logits.numpy()
            if n > 1:
                err = (ologits - logits).abs().max().item()
                assert err < 2e-4 * max(1.0, logits.abs().max().item()), ("basec logits", k, err)
            if ret is None:
                out[k + "_none"] = np.array([1]); assert oret is None
                print("basec golden", k, "-> None")
                continue
            q5, sc, sp, qi = ret
            assert torch.equal(q5, oret[0]) and torch.equal(sp, oret[2]) and qi.shape == oret[3].shape, k
            assert torch.allclose(sc, oret[1], atol=1e-5)
            out[k + "_quint"] = q5.numpy(); out[k + "_scores"] = sc.numpy(); out[k + "_spans"] = sp.numpy()
            print("basec golden", k, "n=%d pairs=%d triplets=%d" % (n, n * (n - 1), q5.shape[0]))
    save("basec", **out)


# ------------------------------------------------------------------ enti_viou_align ----
def gen_align():
    out = {}
    cfg = synth.tiny_vidvrd_config()
    state = synth.make_bigc_state(7, cfg)
    model = build_ref_bigc(cfg, state, BIG_C_vidvrd)
    for sd in (401, 402, 403):
        P = synth.make_proposal(sd, 14, 120, 8, cfg["num_enti_cats"], with_features=False)
        G = synth.make_gt_graph(sd, P, cfg["num_pred_cats"], n_rel=(3, 12))
        g_closed = G.traj_durations.clone()
        G.traj_durations = g_closed.clone(); G.traj_durations[:, 1] += 1   # the reference expects half-open and mutates (:567)
        with torch.no_grad():
            aligned, viou = model.enti_viou_align(G.adj_matrix, P, G)
        oa, ov = og.enti_viou_align(G.adj_matrix, P.bboxes_list, P.traj_durations, G.traj_bboxes, g_closed, cfg["positive_vIoU_th"])
        assert torch.equal(aligned, oa) and torch.equal(viou, ov)
        out["align_%d" % sd] = aligned.numpy(); out["viou_%d" % sd] = viou.numpy()
    save("align", **out)


# ------------------------------------------------------------------ bipartite_match cost matrices ----
def gen_bipartite():
    """Cost matrix of the Hungarian matching (model_0v10.py:606-639): the reference's own ``bipartite_match`` with scipy's
    ``linear_sum_assignment`` wrapped so that the matrix it receives is captured."""
    import models.model_0v10 as M
    out = {}
    cfg = synth.tiny_vidvrd_config()
    model = build_ref_bigc(cfg, synth.make_bigc_state(7, cfg), BIG_C_vidvrd)
    captured = []
    real = M.linear_sum_assignment

    def spy(cost):
        captured.append(cost.clone())
        return real(cost)
    M.linear_sum_assignment = spy
    try:
        for sd, n_gt, n_enti in ((451, 5, 9), (452, 17, 23), (453, 1, 4)):
            logit, gt_pred, att, adj = synth.make_bipartite_case(sd, cfg["num_querys"], cfg["num_pred_cats"], n_gt, n_enti)
            with torch.no_grad():
                idx = model.bipartite_match(logit, gt_pred, att, adj)
            cost = captured[-1]
            oc_ = ob.bipartite_cost(logit, gt_pred, att, adj, cfg["cost_coeff_dict"]["classification"], cfg["cost_coeff_dict"]["adj_matrix"])
            assert torch.allclose(cost, oc_, rtol=1e-6, atol=1e-6)
            out["cost_%d" % sd] = cost.numpy(); out["row_%d" % sd] = np.asarray(idx[0]); out["col_%d" % sd] = np.asarray(idx[1])
    finally:
        M.linear_sum_assignment = real
    save("bipartite", **out)


# ------------------------------------------------------------------ grounding ----------
def gen_grounding():
    out = {}
    cfg = synth.grounding_config()
    state = synth.make_grounding_state(21, cfg)
    tmp = tempfile.mkdtemp()
    np.save(os.path.join(tmp, "e.npy"), state["EntiNameEmb"].numpy()); np.save(os.path.join(tmp, "p.npy"), state["PredNameEmb"].numpy())
    rcfg = dict(cfg, EntiNameEmb_path=os.path.join(tmp, "e.npy"), PredNameEmb_path=os.path.join(tmp, "p.npy"))
    model = DEBUG(rcfg, is_train=False)
    model.load_state_dict(state, strict=True)
    model.eval()
    inf = synth.GROUNDING_INFERENCE
    for sd, n, vl, m in ((601, 10, 200, 30), (602, 16, 520, 45), (603, 6, 64, 12)):
        P = synth.make_proposal(sd, n, vl, 8, 81, min_len=15, with_features=False)
        G = synth.make_gt_graph(sd, P, 51)
        T = synth.make_predictions(sd, P, G, 51, m=m, p_from_gt=0.3)
        quint = torch.unique(T[0], dim=0)
        d = P.traj_durations
        inter, _ = og.dura_intersection(d, d)
        spans = inter[quint[:, 3], quint[:, 4]]
        vf = synth.make_video_feature(sd, vl)
        with torch.no_grad():
            words, so = model.prepare_data((quint, spans, vl))
            regrs, conf, cls = model.forward_propagation(vf, words, so)
            pooled, probs, mask = model([vf], [(quint, spans, vl)], with_gt_data=False, **inf)
            ow, oso = ogr.prepare_data(state, quint, spans, vl)
            oregrs, oconf, ocls = ogr.forward_propagation(state, vf, ow, oso)
            opooled, oprobs, omask = ogr.forward(state, cfg, [vf], [(quint, spans, vl)], **inf)
        for a, b, nm in ((regrs, oregrs, "regrs"), (conf, oconf, "conf"), (cls, ocls, "cls")):
            e = (a - b).abs().max().item()
            assert e < 1e-4 * max(1.0, a.abs().max().item()), (nm, e)
        # discrete decisions (score > th*top, gIoU > th, NMS) can flip on 1e-7 network noise, so the
        # post-processing is pinned on the REFERENCE's network outputs, where it must be exact
        ppooled, pprobs, pmask = ogr.postprocess(regrs, conf, cls, so, cfg["num_bins"], **inf)
        assert torch.equal(mask, pmask) and torch.equal(pooled, ppooled) and torch.equal(probs, pprobs), "grounding post differs"
        nflip = int(((pooled - opooled).abs() > 1e-5).any(-1).sum()) + int((mask != omask).sum())
        print("  end-to-end oracle vs reference: %d near-tie flips of %d bins" % (nflip, mask.numel()))
        out["g%d_so" % sd] = so.numpy()
        k = "g%d" % sd
        out[k + "_regrs"] = regrs.numpy(); out[k + "_conf"] = conf.numpy(); out[k + "_cls"] = cls.numpy()
        out[k + "_pooled"] = pooled.numpy(); out[k + "_probs"] = probs.numpy(); out[k + "_mask"] = mask.numpy()
        out[k + "_quint"] = quint.numpy(); out[k + "_spans"] = spans.numpy()
        print("grounding golden", k, "nq=%d T=%d" % (quint.shape[0], vf.shape[0]), "mask true=%d" % int(mask.sum()),
              "conf range", float(conf.min()), float(conf.max()))
    save("grounding", **out)


def gen_grounding_gt():
    """DEBUG.forward(with_gt_data=True) in inference mode (grd_model_v5.py:198-221, 253-308): queries built from GT graphs."""
    out = {}
    cfg = synth.grounding_config()
    state = synth.make_grounding_state(21, cfg)
    tmp = tempfile.mkdtemp()
    np.save(os.path.join(tmp, "e.npy"), state["EntiNameEmb"].numpy()); np.save(os.path.join(tmp, "p.npy"), state["PredNameEmb"].numpy())
    rcfg = dict(cfg, EntiNameEmb_path=os.path.join(tmp, "e.npy"), PredNameEmb_path=os.path.join(tmp, "p.npy"))
    model = DEBUG(rcfg, is_train=False)
    model.load_state_dict(state, strict=True)
    model.eval()
    inf = synth.GROUNDING_INFERENCE
    for sd, n, vl in ((611, 10, 200), (612, 16, 520)):
        P = synth.make_proposal(sd, n, vl, 8, 81, min_len=15, with_features=False)
        G = synth.make_gt_graph(sd, P, 51)
        vf = synth.make_video_feature(sd, vl)
        with torch.no_grad():
            words, tinfo, target, index_map = model.prepare_gt_data(G)
            pooled, probs, mask = model([vf], [G], with_gt_data=True, **inf)
        k = "gt%d" % sd
        out[k + "_tinfo"] = tinfo.numpy(); out[k + "_target"] = target.numpy()
        out[k + "_index_map"] = torch.cat(index_map).numpy(); out[k + "_counts"] = np.array([len(x) for x in index_map])
        out[k + "_pooled"] = pooled.numpy(); out[k + "_probs"] = probs.numpy(); out[k + "_mask"] = mask.numpy()
        print("grounding (GT queries) golden", k, "n_uniq=%d of %d GT predicates" % (tinfo.shape[0], G.num_preds), "mask true=%d" % int(mask.sum()))
    save("grounding_gt", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["geometry", "eval", "bigc", "align", "grounding", "grounding_gt", "basec"]
    for w in which:
        {"geometry": gen_geometry, "eval": gen_eval, "bigc": gen_bigc, "align": gen_align, "grounding": gen_grounding,
         "grounding_gt": gen_grounding_gt, "basec": gen_basec, "bipartite": gen_bipartite}[w]()
