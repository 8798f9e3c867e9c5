"""GPU parity: grounding stage DEBUG (models/grd_model_v5.py) vs the reference goldens and the oracle."""
import numpy as np
import pytest
import torch

from oracle import grounding as ogr
from vidsgg_big_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
INF = synth.GROUNDING_INFERENCE
TH = (INF["score_th"], INF["tiou_th"], INF["bins_th"], INF["nms_th"])
CASES = ((601, 10, 200, 30), (602, 16, 520, 45), (603, 6, 64, 12))


def _model(precision):
    from vidsgg_big_b200 import grounding
    cfg = synth.grounding_config()
    m = grounding.DEBUG(cfg, is_train=False, precision=precision)
    m.load_state_dict(synth.make_grounding_state(21, cfg), strict=True)
    return m.cuda().eval()


# measured on B200 (profiles/r02_parity_errors.txt): fp32_simt <= 5.1e-7, 3xtf32 <= 2.0e-6, tf32+bf16x2 <= 2.2e-6, tf32 <= 9.2e-4
@pytest.mark.parametrize("precision,tol", [("fp32_simt", 5e-6), ("3xtf32", 2e-5), ("tf32+bf16x2", 2e-5), ("fp16x3", 2e-5), ("tf32", 5e-3)])
def test_grounding_network_vs_reference(golden, precision, tol):
    g = golden("grounding")
    model = _model(precision)
    for sd, n, vl, m in CASES:
        k = "g%d" % sd
        quint, spans = torch.from_numpy(g[k + "_quint"]).to(DEV), torch.from_numpy(g[k + "_spans"]).to(DEV)
        vf = synth.make_video_feature(sd, vl).to(DEV)
        regrs, conf, cls, so_norm, _ = model.forward_propagation_debug(vf, quint, spans, vl, TH)
        assert np.array_equal(so_norm.cpu().numpy(), g[k + "_so"])
        for name, got in (("regrs", regrs), ("conf", conf), ("cls", cls)):
            ref = g[k + "_" + name]
            err = np.abs(got.cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1e-6)
            print(k, precision, name, "rel err %.2e" % err)
            assert err <= tol, (k, name, err)


def test_grounding_post_exact_on_reference_outputs(golden):
    """K7 alone: fed the reference's network outputs the post-processing kernel reproduces pooled spans, bin probabilities
    and masks (discrete decisions identical; floats to 1 ulp of the sigmoid)."""
    g = golden("grounding")
    model = _model("fp32_simt")
    for sd, n, vl, m in CASES:
        k = "g%d" % sd
        pooled, probs, mask = model.postprocess(torch.from_numpy(g[k + "_regrs"]), torch.from_numpy(g[k + "_conf"]),
                                                torch.from_numpy(g[k + "_cls"]), torch.from_numpy(g[k + "_so"]), TH)
        assert np.array_equal(mask.cpu().numpy(), g[k + "_mask"]), k
        np.testing.assert_allclose(pooled.cpu().numpy(), g[k + "_pooled"], rtol=0, atol=2e-6)
        np.testing.assert_allclose(probs.cpu().numpy(), g[k + "_probs"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("precision", ["fp32_simt", "3xtf32", "tf32+bf16x2", "fp16x3"])
def test_grounding_forward_end_to_end(golden, precision):
    g = golden("grounding")
    model = _model(precision)
    feats, datas = [], []
    for sd, n, vl, m in CASES:
        k = "g%d" % sd
        feats.append(synth.make_video_feature(sd, vl).to(DEV))
        datas.append((torch.from_numpy(g[k + "_quint"]).to(DEV), torch.from_numpy(g[k + "_spans"]).to(DEV), vl))
    batched = model(feats, datas, with_gt_data=False, **INF)               # all videos in one batch
    for (sd, n, vl, m), f, d, b in zip(CASES, feats, datas, batched):
        k = "g%d" % sd
        pooled, probs, mask = model([f], [d], with_gt_data=False, **INF)   # reference-style single-video call
        assert torch.equal(mask, b[2]) and torch.allclose(pooled, b[0], atol=1e-6) and torch.allclose(probs, b[1], atol=1e-6)
        ref_mask, ref_pooled, ref_probs = g[k + "_mask"], g[k + "_pooled"], g[k + "_probs"]
        np.testing.assert_allclose(probs.cpu().numpy(), ref_probs, atol=5e-4)
        bad_bins = (np.abs(pooled.cpu().numpy() - ref_pooled) > 1e-4).any(-1) | (mask.cpu().numpy() != ref_mask)
        print(k, precision, "bins differing from the reference (near-tie flips): %d of %d" % (bad_bins.sum(), bad_bins.size))
        assert bad_bins.mean() <= 0.01
    # expansion used by the eval driver (tools/eval_vidor.py:245-253)
    from vidsgg_big_b200.grounding import expand_after_grounding
    q, s, sp = expand_after_grounding(datas[0][0], torch.rand(datas[0][0].shape[0], 3, device=DEV), *model([feats[0]], [datas[0]], with_gt_data=False, **INF), CASES[0][2])
    assert q.shape[0] == s.shape[0] == sp.shape[0] and sp.dtype == torch.long


@pytest.mark.parametrize("precision", ["fp32_simt", "tf32+bf16x2"])
def test_grounding_on_gt_queries(golden, precision):
    """DEBUG.forward(with_gt_data=True) (grd_model_v5.py:198-202, 253-308): queries = unique (pred, subj, obj, so-span) tags of a GT
    graph.  Tags / targets / index maps exact against the reference golden; the outputs equal the with_gt_data=False call on the
    derived queries bit for bit and match the reference outputs up to near-tie flips."""
    g = golden("grounding_gt")
    model = _model(precision)
    for sd, n, vl in ((611, 10, 200), (612, 16, 520)):
        k = "gt%d" % sd
        P = synth.make_proposal(sd, n, vl, 8, 81, min_len=15, with_features=False)
        G = synth.make_gt_graph(sd, P, 51).to(DEV)
        vf = synth.make_video_feature(sd, vl).to(DEV)
        data, target, index_map = model.prepare_gt_data(G)
        np.testing.assert_array_equal((data[1].cpu().float() / vl).numpy(), g[k + "_tinfo"])      # integer spans are exact
        np.testing.assert_allclose(target.cpu().numpy(), g[k + "_target"], rtol=2e-7)
        assert np.array_equal(torch.cat(index_map).cpu().numpy(), g[k + "_index_map"])
        assert [len(x) for x in index_map] == g[k + "_counts"].tolist()
        pooled, probs, mask = model([vf], [G], with_gt_data=True, **INF)
        p2, pr2, m2 = model([vf], [data], with_gt_data=False, **INF)
        assert torch.equal(pooled, p2) and torch.equal(probs, pr2) and torch.equal(mask, m2)
        np.testing.assert_allclose(probs.cpu().numpy(), g[k + "_probs"], atol=5e-4)
        bad = (np.abs(pooled.cpu().numpy() - g[k + "_pooled"]) > 1e-4).any(-1) | (mask.cpu().numpy() != g[k + "_mask"])
        print(k, precision, "bins differing from the reference (near-tie flips): %d of %d" % (bad.sum(), bad.size))
        assert bad.mean() <= 0.01


def test_fused_dwconv_path_equals_separate_dwconv_launches(golden):
    """grounding.DEBUG with the depthwise convs fused into the point-wise GEMMs (default) vs separate vsg_dwconv launches: the whole
    network output is bit-identical (the fused operand is computed in the same operation order)."""
    g = golden("grounding")
    sd, n, vl, m = CASES[1]
    k = "g%d" % sd
    f = synth.make_video_feature(sd, vl).to(DEV)
    d = (torch.from_numpy(g[k + "_quint"]).to(DEV), torch.from_numpy(g[k + "_spans"]).to(DEV), vl)
    model = _model("tf32+bf16x2")
    assert model.fuse_dwconv
    fused = model([f], [d], with_gt_data=False, **INF)
    model.fuse_dwconv = False
    plain = model([f], [d], with_gt_data=False, **INF)
    for a, b in zip(fused, plain):
        assert torch.equal(a, b)


def test_grounding_api_errors():
    from vidsgg_big_b200 import grounding
    cfg = synth.grounding_config()
    with pytest.raises(NotImplementedError):
        grounding.DEBUG(cfg, is_train=True)
    m = _model("fp32_simt")
    empty = (torch.zeros(0, 5, dtype=torch.long), torch.zeros(0, 2, dtype=torch.long), 10)
    assert m([torch.zeros(4, 1024, device=DEV)], [empty], with_gt_data=False) == (None, None)


def test_classify_then_ground_packed_equals_driver_expansion():
    """BIG-C (VidOR) -> grounding -> relations: the packed batched path builds exactly the relations that the reference
    driver's per-video expansion (tools/eval_vidor.py:245-253) + to_eval_format_pr would evaluate."""
    from vidsgg_big_b200 import bigc, evalapi, geometry, grounding
    cfg = synth.vidor_config()
    model = bigc.BIG_C_vidor(cfg, precision="3xtf32")
    model.load_state_dict(synth.make_bigc_state(2, cfg)); model.cuda()
    grd = _model("3xtf32")
    props, feats = [], []
    for sd, n, vl in ((701, 8, 120), (702, 14, 260)):
        P = synth.make_proposal(sd, n, vl, 1324, 81, min_len=15).to(DEV)
        props.append(P); feats.append(synth.make_video_feature(sd, vl).to(DEV))
    with torch.no_grad():
        packed = model.forward_packed(props, topk=3)
        per_video = packed.per_video()
        q, s3, sp, _, off = packed.compact()
        datas = [(q[off[i]:off[i + 1]], sp[off[i]:off[i + 1]], props[i].video_len) for i in range(2)]
        pooled, probs, mask = grd.forward_packed(feats, datas, **INF)
    tt = geometry.TrackTable.from_containers(props)
    A = evalapi.PackedRelations.from_grounded(tt, packed, pooled, probs, mask, [p.video_len for p in props])
    trips, r = [], 0
    for i, pv in enumerate(per_video):
        m = pv[0].shape[0]
        trips.append(grounding.expand_after_grounding(pv[0], pv[1], pooled[r:r + m], probs[r:r + m], mask[r:r + m], props[i].video_len))
        r += m
    B = evalapi.PackedRelations.from_triplets(tt, trips)
    assert torch.equal(A.rel, B.rel) and torch.equal(A.vid_off, B.vid_off)
    assert torch.allclose(A.scores, B.scores, rtol=1e-6)
    assert A.n_rel >= sum(t[0].shape[0] for t in per_video)          # at least one bin per query


def test_vidor_combined_driver():
    """evaluate_cls_stage / evaluate_combined (tools/eval_vidor.py) on a small VidOR-shaped set: runs end to end, the packed
    metrics equal those of the dict path fed with the driver-style expansion."""
    from vidsgg_big_b200 import bigc, convert, driver, evalapi, grounding
    cfg = synth.vidor_config()
    cls_model = bigc.BIG_C_vidor(cfg, precision="3xtf32")
    cls_model.load_state_dict(synth.make_bigc_state(2, cfg)); cls_model.cuda()
    grd = _model("3xtf32")
    cv = convert.EvalFmtCvtor("vidvrd")
    props, graphs, feats, gts = [], [], [], {}
    for sd, n, vl in ((711, 8, 120), (712, 12, 200), (713, 6, 90)):
        p = synth.make_proposal(sd, n, vl, 1324, 81, min_len=15)
        g = synth.make_gt_graph(sd, p, 51, n_rel=(3, 10))
        gts.update(cv.to_eval_format_gt(g))
        props.append(p.to(DEV)); graphs.append(g.to(DEV)); feats.append(synth.make_video_feature(sd, vl).to(DEV))
    (m0, r0, p0), save = driver.evaluate_cls_stage(cls_model, props, graphs, topk=3)
    assert set(save) == {p.video_name for p in props} and all(v is None or len(v) == 4 for v in save.values())
    m_ap, rec, mprec, infos = driver.evaluate_combined(grd, cls_model, props, feats, graphs, topk=3, **INF)
    # reference-style: per video expansion -> dicts -> dict-path evaluation
    prs = {}
    with torch.no_grad():
        res = cls_model(props, topk=3)
    for p, f, r in zip(props, feats, res):
        pooled, probs, mask = grd([f], [(r[0], r[2], p.video_len)], with_gt_data=False, **INF)
        prs.update(cv.to_eval_format_pr(p, grounding.expand_after_grounding(r[0], r[1], pooled, probs, mask, p.video_len)))
    m2, r2, p2, infos2 = evalapi.evaluate_v2(gts, prs)
    assert abs(m_ap - m2) < 1e-12 and rec[50] == r2[50] and rec[100] == r2[100]
    for v in infos2:
        assert np.array_equal(infos[v][0], infos2[v][0]) and np.array_equal(infos[v][1], infos2[v][1])


def _large_case():
    """VidOR-size grounding input (north_star sizes: T <= 675 clips, nq <= 576 = 192 * 3): video_len 4900 -> T = 613 clips, ~540 queries."""
    vl = 4900
    P = synth.make_proposal(651, 40, vl, 8, 81, min_len=15, with_features=False)
    G = synth.make_gt_graph(651, P, 51)
    T = synth.make_predictions(651, P, G, 51, m=620, p_from_gt=0.05)
    quint = torch.unique(T[0], dim=0)
    d = P.traj_durations
    s = torch.maximum(d[quint[:, 3], 0], d[quint[:, 4], 0]); e = torch.minimum(d[quint[:, 3], 1], d[quint[:, 4], 1])
    spans = torch.stack([s, e], 1)
    assert quint.shape[0] >= 500 and (s <= e).all()
    return quint, spans, vl, synth.make_video_feature(651, vl)


_LARGE_REF = {}


def _large_reference():
    """The oracle (CPU) on the large case, computed once per session (~20 s)."""
    if not _LARGE_REF:
        quint, spans, vl, vf = _large_case()
        cfg = synth.grounding_config()
        st = synth.make_grounding_state(21, cfg)
        with torch.no_grad():
            words, so = ogr.prepare_data(st, quint, spans, vl)
            regrs, conf, cls = ogr.forward_propagation(st, vf, words, so)
            pooled, probs, mask = ogr.postprocess(regrs, conf, cls, so, cfg["num_bins"], *TH)
        _LARGE_REF.update(quint=quint, spans=spans, vl=vl, vf=vf, so=so, regrs=regrs, conf=conf, cls=cls, pooled=pooled, probs=probs, mask=mask)
    return _LARGE_REF


@pytest.mark.parametrize("precision,tol", [("fp32_simt", 5e-6), ("3xtf32", 2e-5), ("tf32+bf16x2", 2e-5), ("fp16x3", 2e-5)])
def test_grounding_vidor_size_vs_oracle(precision, tol):
    """Grounding parity at VidOR size (T = 613 clips >= 600, nq >= 500 queries, 10 bins): network outputs against the oracle's, the
    post-processing pinned EXACTLY on the oracle's network outputs, and the end-to-end outputs up to near-tie flips."""
    r = _large_reference()
    model = _model(precision)
    quint, spans, vf = r["quint"].to(DEV), r["spans"].to(DEV), r["vf"].to(DEV)
    assert vf.shape[0] >= 600 and quint.shape[0] >= 500
    regrs, conf, cls, so_norm, out = model.forward_propagation_debug(vf, quint, spans, r["vl"], TH)
    assert torch.equal(so_norm.cpu(), r["so"])
    for name, got in (("regrs", regrs), ("conf", conf), ("cls", cls)):
        ref = r[name]
        err = (got.cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
        print(precision, name, "rel err %.2e" % err)
        assert err <= tol, (name, err)
    # discrete stage alone, on the oracle's network outputs: identical masks, floats to 1 ulp of the sigmoid
    pooled, probs, mask = model.postprocess(r["regrs"], r["conf"], r["cls"], r["so"], TH)
    assert torch.equal(mask.cpu(), r["mask"])
    np.testing.assert_allclose(pooled.cpu().numpy(), r["pooled"].numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(probs.cpu().numpy(), r["probs"].numpy(), rtol=0, atol=2e-6)
    # end to end
    pooled, probs, mask = out
    np.testing.assert_allclose(probs.cpu().numpy(), r["probs"].numpy(), atol=5e-4)
    bad = ((pooled.cpu() - r["pooled"]).abs() > 1e-4).any(-1) | (mask.cpu() != r["mask"])
    print(precision, "bins differing from the oracle (near-tie flips): %d of %d" % (int(bad.sum()), bad.numel()))
    assert bad.float().mean().item() <= 0.01
    # frame spans the driver rounds to (tools/eval_vidor.py:248-253) on the bins both keep
    both = mask.cpu() & r["mask"]
    a, b = torch.round(pooled.cpu() * r["vl"])[both], torch.round(r["pooled"] * r["vl"])[both]
    assert (a == b).all(-1).float().mean().item() >= 0.97


@pytest.mark.parametrize("precision", ["fp32_simt", "tf32", "3xtf32", "tf32+bf16x2", "fp16x3", "bf16"])
def test_c_grounding_forward_equals_python_issued_launches(golden, precision):
    """vsg_grd_forward (ONE C call, csrc/forward.cu) against the same launches issued op by op from Python: bit-identical network
    outputs and post-processing results for a ragged batch of videos, every precision mode, with and without the fused depthwise conv
    and the tcgen05 attention."""
    g = golden("grounding")
    model = _model(precision)
    feats, datas = [], []
    for sd, n, vl, m in CASES:
        k = "g%d" % sd
        feats.append(synth.make_video_feature(sd, vl).to(DEV))
        datas.append((torch.from_numpy(g[k + "_quint"]).to(DEV), torch.from_numpy(g[k + "_spans"]).to(DEV), vl))
    for attention, fuse in (("tc", True), ("simt", False)):
        model.attention, model.fuse_dwconv = attention, fuse
        outs = {}
        for backend in ("py", "c"):
            model.backend = backend
            outs[backend] = model(feats, datas, with_gt_data=False, **INF)
            outs[backend + "_net"] = model.forward_propagation_debug(feats[1], datas[1][0], datas[1][1], datas[1][2], TH)
        for a, b in zip(outs["py"], outs["c"]):
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]), (attention, fuse)
        for a, b in zip(outs["py_net"][:4], outs["c_net"][:4]):
            assert torch.equal(a, b), (attention, fuse)
