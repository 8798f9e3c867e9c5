"""GPU parity: eval_visual_relation & friends (dict path and packed path) vs the oracle / reference goldens."""
import numpy as np
import pytest
import torch

from oracle import convert as oc, evalapi as oe
from vidsgg_big_b200 import synth
from test_oracle_golden import _eval_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _api():
    from vidsgg_big_b200 import evalapi
    return evalapi


def test_viou_kats():
    api = _api()
    assert abs(api.viou([[0, 0, 9, 9]] * 4, [0, 4], [[5, 5, 14, 14]] * 3, [2, 5]) - 50 / 650) < 1e-15
    assert api.viou([[0, 0, 9, 9]] * 4, [0, 4], [[5, 5, 14, 14]] * 3, [4, 7]) == 0.0
    assert api.viou([[1.5, 2.5, 30.25, 40.75]] * 5, [3, 8], [[1.5, 2.5, 30.25, 40.75]] * 5, [3, 8]) == 1.0


def test_eval_dict_path_matches_reference(golden):
    api = _api()
    g = golden("evalapi")
    gts, prs = _eval_case()
    m_ap, rec, mprec, infos = api.evaluate_v2(gts, prs)
    assert abs(m_ap - g["mean_ap"][0]) < 1e-12
    assert [rec[50], rec[100]] == g["rec"].tolist()                    # identical recall@50/100
    assert [mprec[1], mprec[5], mprec[10]] == g["mprec"].tolist()
    for v in infos:
        assert np.array_equal(infos[v][0], g["hit_" + v]), v           # same hits, same order
        assert np.array_equal(infos[v][1], g["g2d_" + v]), v
    m07, r07, _ = api.eval_visual_relation(gts, prs, viou_threshold=0.7)
    assert abs(m07 - g["mean_ap_thr07"][0]) < 1e-12 and [r07[50], r07[100]] == g["rec_thr07"].tolist()


def test_eval_single_video_functions():
    api = _api()
    gts, prs = _eval_case()
    vid = "synth_000503"
    p1, r1, h1 = api.eval_detection_scores(gts[vid], prs[vid], 0.5)
    p0, r0, h0 = oe.detection_scores(gts[vid], prs[vid], 0.5)
    assert np.array_equal(h1, h0) and np.array_equal(p1, p0) and np.array_equal(r1, r0)
    t1 = api.eval_tagging_scores(gts[vid], prs[vid]); t0 = oe.tagging_scores(gts[vid], prs[vid])
    assert all(np.array_equal(a, b) for a, b in zip(t1, t0))
    # empty predictions / GT without match
    p, r, h = api.eval_detection_scores(gts[vid], [], 0.5)
    assert h.size == 0


def test_ov_values_f64():
    """ov matrix entries vs common.viou restated (1e-12, north_star fp64 bar)."""
    api = _api()
    gts, prs = _eval_case()
    vids = [v for v in gts if gts[v] and v in prs][:4]
    vocab = {}
    G = api.PackedRelations.from_dicts([gts[v] for v in vids], vocab, DEV, False)
    P = api.PackedRelations.from_dicts([prs[v] for v in vids], vocab, DEV, True)
    m = api.match_relations(P, G, 0.5, keep_ov=True)
    ov = m.ov.cpu().numpy()
    checked = 0
    for i, v in enumerate(vids):
        ng = len(gts[v])
        for pi, pr in enumerate(prs[v][:40]):
            for gi, gt in enumerate(gts[v]):
                got = ov[m.ov_off[i] + pi * ng + gi]
                if tuple(pr["triplet"]) != tuple(gt["triplet"]):
                    assert got == -1.0
                    continue
                ref = min(oe.viou(pr["sub_traj"], pr["duration"], gt["sub_traj"], gt["duration"]),
                          oe.viou(pr["obj_traj"], pr["duration"], gt["obj_traj"], gt["duration"]))
                assert abs(got - ref) <= 1e-12
                checked += 1
    assert checked > 20


def test_eval_packed_path_equals_dict_path():
    api = _api()
    from vidsgg_big_b200 import geometry
    en, pn = oc.default_names("e", 64), oc.default_names("p", 200)
    props, graphs, trips, gts, prs = [], [], [], {}, {}
    for sd in range(520, 530):
        rng = np.random.default_rng(sd)
        P = synth.make_proposal(sd, int(rng.integers(6, 16)), int(rng.integers(60, 200)), 8, 36, with_features=False)
        G = synth.make_gt_graph(sd, P, 133, n_rel=(3, 25))
        T = synth.make_predictions(sd, P, G, 133, m=150) if sd != 524 else None
        gts.update(oc.to_eval_format_gt(G, en, pn)); prs.update(oc.to_eval_format_pr(P, T, en, pn))
        props.append(P.to(DEV)); graphs.append(G.to(DEV)); trips.append(T)
    ref = oe.evaluate(gts, prs, with_infos=True)
    pt = geometry.TrackTable.from_containers(props)
    gt_t = geometry.TrackTable.from_containers(graphs)
    PR = api.PackedRelations.from_triplets(pt, trips)
    GT = api.PackedRelations.from_gt_graphs(gt_t, graphs)
    m_ap, rec, mprec, infos = api.evaluate_packed(PR, GT, with_infos=True)
    assert abs(m_ap - ref[0]) < 1e-12 and rec[50] == ref[1][50] and rec[100] == ref[1][100]
    assert all(mprec[k] == ref[2][k] for k in (1, 5, 10))
    for i, P in enumerate(props):
        h_ref, g_ref = ref[3][P.video_name]
        assert np.array_equal(infos[i][0], h_ref) and np.array_equal(infos[i][1], g_ref)


def test_convertor_equals_oracle_and_driver_roundtrip():
    """Product EvalFmtCvtor == oracle convertor (itself pinned to the reference's), and the batched VidVRD driver gives the
    metrics of evaluating its own dict export through the dict path."""
    from vidsgg_big_b200 import bigc, convert, driver
    api = _api()
    en, pn = oc.default_names("e", 64), oc.default_names("p", 200)
    cv = convert.EvalFmtCvtor("vidvrd")
    P = synth.make_proposal(500, 9, 120, 8, 36, with_features=False)
    G = synth.make_gt_graph(500, P, 133)
    T = synth.make_predictions(500, P, G, 133, m=40)
    assert cv.to_eval_format_pr(P, T) == oc.to_eval_format_pr(P, T, en, pn)
    assert cv.to_eval_format_gt(G) == oc.to_eval_format_gt(G, en, pn)
    assert cv.to_eval_format_pr(P, None) == {P.video_name: []}
    # driver on a tiny model
    cfg = synth.tiny_vidvrd_config()
    model = bigc.BIG_C_vidvrd(cfg, precision="3xtf32")
    model.load_state_dict(synth.make_bigc_state(7, cfg)); model.cuda()
    props, graphs, gts = [], [], {}
    for sd in range(540, 546):
        p = synth.make_proposal(sd, 8, 60, 136, cfg["num_enti_cats"])
        g = synth.make_gt_graph(sd, p, cfg["num_pred_cats"], n_rel=(3, 10))
        gts.update(cv.to_eval_format_gt(g))
        props.append(p.to(DEV)); graphs.append(g.to(DEV))
    m_ap, rec, mprec, dicts = driver.inference_then_eval(model, props, graphs, topk=5, want_dicts=True)
    m2, r2, p2 = api.eval_visual_relation(gts, dicts)
    assert abs(m_ap - m2) < 1e-12 and rec[50] == r2[50] and rec[100] == r2[100] and all(mprec[k] == p2[k] for k in (1, 5, 10))


def test_eval_idempotence_at_vidvrd_test_size():
    """Size-independent property at the VidVRD-test size (200 videos): evaluating the GT relations against themselves (as
    predictions with distinct scores) gives AP = 1 for every video, recall@100 = 1 and recall@50 = (sum of min(n_gt, 50)) / n_gt_total;
    dropping every prediction gives zeros."""
    from vidsgg_big_b200 import evalapi, geometry
    graphs = []
    for i in range(200):
        rng = np.random.default_rng(5000 + i)
        P = synth.make_proposal(5000 + i, int(rng.integers(5, 40)), int(rng.integers(90, 600)), 8, 36, with_features=False)
        graphs.append(synth.make_gt_graph(5000 + i, P, 133, n_rel=(5, 60)).to(DEV))
    gt_t = geometry.TrackTable.from_containers(graphs, device=DEV)
    GT = evalapi.PackedRelations.from_gt_graphs(gt_t, graphs)
    n = GT.rel.shape[0]
    scores = torch.linspace(1.0, 0.5, n, dtype=torch.float64, device=DEV)
    PR = evalapi.PackedRelations(GT.boxes, GT.off, GT.tstart, GT.rel.clone(), GT.vid_off.clone(), scores, vol_full_track=GT.vol_full_track)
    m_ap, rec, mprec = evalapi.evaluate_packed(PR, GT)
    per_vid = np.diff(GT.vid_off.cpu().numpy())
    assert abs(m_ap - 1.0) < 1e-12
    assert abs(float(rec[100]) - np.minimum(per_vid, 100).sum() / per_vid.sum()) < 1e-6
    assert abs(float(rec[50]) - np.minimum(per_vid, 50).sum() / per_vid.sum()) < 1e-6
