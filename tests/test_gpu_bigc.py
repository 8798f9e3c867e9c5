"""GPU parity: BIG-C classification forward (BIG_C_vidvrd / BIG_C_vidor) vs the reference goldens and the oracle."""
import numpy as np
import pytest
import torch

from oracle import bigc as ob
from vidsgg_big_b200 import synth
from test_oracle_golden import BIGC_CASES, bigc_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
# max |dlogit| / max |logit| allowed per precision mode
# measured on B200 (profiles/r02_parity_errors.txt): fp32_simt <= 1.1e-6, 3xtf32 <= 1.2e-5, tf32+bf16x2 <= 8.9e-6, tf32 <= 3.2e-3
LOGIT_TOL = {"fp32_simt": 5e-6, "3xtf32": 4e-5, "tf32+bf16x2": 4e-5, "fp16x3": 4e-5, "tf32": 1e-2, "bf16": 6e-2}


# reduced-precision modes are REPORTED, not used for parity claims: (min share of queries keeping the (s,o) arg-max, max att error,
# min share of reference triplets reproduced)
REDUCED = {"tf32": (0.85, 5e-2, 0.7), "bf16": (0.6, 2e-1, 0.4)}


def _model(cfg, state, precision):
    from vidsgg_big_b200 import bigc
    cls = bigc.BIG_C_vidor if cfg["variant"] == "vidor" else bigc.BIG_C_vidvrd
    m = cls(cfg, is_train=False, precision=precision)
    m.load_state_dict(state, strict=True)
    return m.cuda().eval()


def _unstable_queries(logits, att, topk, tau=2e-3):
    """Queries whose discrete decisions sit on a near-tie in the reference outputs."""
    probs = torch.softmax(torch.from_numpy(logits), -1)
    sp, _ = torch.sort(probs, dim=-1, descending=True)
    tie_k = (sp[:, topk - 1] - sp[:, topk]).abs() < tau * sp[:, topk - 1] if probs.shape[1] > topk else torch.zeros(probs.shape[0], dtype=torch.bool)
    # neighbouring ranks swapping also changes which row wins a dedup group only through the score; ignore
    a = torch.from_numpy(att)
    if a.shape[-1] > 1:
        top2 = torch.topk(a, 2, dim=-1)[0]
        tie_a = ((top2[..., 0] - top2[..., 1]).abs() < tau * top2[..., 0]).any(0)
    else:
        tie_a = torch.zeros(a.shape[1], dtype=torch.bool)
    return set((tie_k | tie_a).nonzero().flatten().tolist())


@pytest.mark.parametrize("precision", ["fp32_simt", "3xtf32", "tf32+bf16x2", "fp16x3", "tf32", "bf16"])
@pytest.mark.parametrize("case", BIGC_CASES, ids=[c[0] for c in BIGC_CASES])
def test_bigc_forward_vs_reference(golden, case, precision):
    g = golden("bigc")
    tag, mk, over, shapes, wseed, topk = case
    cfg = mk(**over)
    st = synth.make_bigc_state(wseed, cfg)
    model = _model(cfg, st, precision)
    for sd, n, vl, maxlen in shapes:
        P = bigc_inputs(cfg, sd, n, vl, maxlen)
        k = "%s_%d" % (tag, sd)
        ref_logits, ref_att = g[k + "_logits"], g[k + "_att"]
        P.to(DEV)
        with torch.no_grad():
            query, logits, att, so, ex = model.forward_debug(P)
            ret = model([P], topk=topk)[0]
        lg = logits.cpu().numpy()
        scale = np.abs(ref_logits).max()
        ref_so = np.argmax(ref_att, axis=-1).T                       # [Q, 2]
        same_so = (so.cpu().numpy() == ref_so).all(axis=1)
        if precision in REDUCED:
            # a flipped subject/object arg-max changes a query's logits wholesale: compare the agreeing queries only
            print("   %s: %d of %d queries keep the reference (s,o) arg-max" % (precision, same_so.sum(), same_so.size))
            assert same_so.mean() >= REDUCED[precision][0]
        else:
            assert same_so.mean() >= 0.99
        err = np.abs(lg - ref_logits)[same_so].max() / scale
        att_err = np.abs(att.cpu().numpy() - ref_att).max()
        print("%s %s: rel logit err %.2e, att err %.2e" % (k, precision, err, att_err))
        assert err <= LOGIT_TOL[precision], (k, precision, err)
        assert att_err <= (REDUCED[precision][1] if precision in REDUCED else 1e-4)      # measured: <= 2.9e-5 (fp32-class), <= 1.5e-2 (tf32)
        if (k + "_none") in g:
            assert ret is None
            continue
        assert ret is not None
        unstable = _unstable_queries(ref_logits, ref_att, topk, tau=(5e-2 if precision in REDUCED else 2e-3))
        ref_rows = {tuple(r): (s, sp, q) for r, s, sp, q in zip(g[k + "_quint"].tolist(), g[k + "_scores"].tolist(),
                                                                  g[k + "_spans"].tolist(), g[k + "_qids"].tolist())}
        my_rows = {tuple(r): (s, sp, q) for r, s, sp, q in zip(ret[0].cpu().tolist(), ret[1].cpu().tolist(),
                                                                 ret[2].cpu().tolist(), ret[3].cpu().tolist())}
        if precision in REDUCED:
            # reduced precision: report the overlap only (near-ties flip, dedup winners change)
            common = len(set(ref_rows) & set(my_rows))
            print("   %s: %d of %d reference triplets reproduced (%d emitted)" % (precision, common, len(ref_rows), len(my_rows)))
            assert common >= REDUCED[precision][2] * len(ref_rows)
            continue
        n_flip = 0
        for key in set(ref_rows) ^ set(my_rows):
            q = (ref_rows.get(key) or my_rows.get(key))[2]
            assert q in unstable, "triplet %s differs and query %d is not a near-tie (%s, %s)" % (key, q, k, precision)
            n_flip += 1
        for key in set(ref_rows) & set(my_rows):
            rs, rsp, rq = ref_rows[key]
            ms, msp, mq = my_rows[key]
            assert rsp == msp                                           # spans bit-exact
            if mq == rq or (mq not in unstable and rq not in unstable):
                assert abs(rs[0] - ms[0]) <= 5e-4 * max(rs[0], 1e-3) + 1e-6
            assert rs[1:] == ms[1:]                                     # detector scores copied exactly
        if precision not in REDUCED:
            assert n_flip <= max(1, len(ref_rows) // 100), "too many near-tie flips: %d of %d" % (n_flip, len(ref_rows))     # measured: 0
            # emitted order is the lexicographic key order of torch.unique
            keys = [tuple(r) for r in ret[0].cpu().tolist()]
            assert keys == sorted(keys)
        print("   triplets ref %d mine %d near-tie flips %d" % (len(ref_rows), len(my_rows), n_flip))


@pytest.mark.parametrize("case", BIGC_CASES, ids=[c[0] for c in BIGC_CASES])
def test_construct_triplet_exact_on_reference_logits(golden, case):
    """K8 alone: fed the reference's own logits and attention arg-max, the kernel must reproduce the reference
    triplets exactly (ids, order, spans, query ids) and the scores to 1e-6."""
    from vidsgg_big_b200 import bigc
    g = golden("bigc")
    tag, mk, over, shapes, wseed, topk = case
    cfg = mk(**over)
    st = synth.make_bigc_state(wseed, cfg)
    model = _model(cfg, st, "fp32_simt")
    for sd, n, vl, maxlen in shapes:
        P = bigc_inputs(cfg, sd, n, vl, maxlen).to(DEV)
        k = "%s_%d" % (tag, sd)
        pk = bigc.PackedVideos([P], torch.device(DEV))
        logits = torch.from_numpy(g[k + "_logits"]).to(DEV).contiguous()
        so = torch.argmax(torch.from_numpy(g[k + "_att"]), dim=-1).t().contiguous().int().to(DEV)
        ret = model._construct_triplets(pk, logits, so, topk)[0]
        if (k + "_none") in g:
            assert ret is None
            continue
        assert np.array_equal(ret[0].cpu().numpy(), g[k + "_quint"])
        assert np.array_equal(ret[2].cpu().numpy(), g[k + "_spans"])
        assert np.array_equal(ret[3].cpu().numpy(), g[k + "_qids"])
        np.testing.assert_allclose(ret[1].cpu().numpy(), g[k + "_scores"], rtol=2e-6, atol=1e-8)


def test_bigc_batched_equals_single():
    """Several ragged videos (incl. an empty one) in ONE forward give the same triplets as one-by-one calls."""
    cfg = synth.tiny_vidvrd_config()
    st = synth.make_bigc_state(7, cfg)
    model = _model(cfg, st, "3xtf32")
    props = [bigc_inputs(cfg, 900 + i, n, vl, None).to(DEV) if n else synth.make_proposal(1, 0, 50, 136, 9)
             for i, (n, vl) in enumerate([(7, 40), (0, 10), (12, 64), (3, 25), (20, 90)])]
    with torch.no_grad():
        batched = model(props, topk=5)
        single = [model([p], topk=5)[0] for p in props]
    assert batched[1] is None and single[1] is None
    for b, s in zip(batched, single):
        assert (b is None) == (s is None)
        if b is not None:
            assert torch.equal(b[0], s[0]) and torch.equal(b[2], s[2]) and torch.equal(b[3], s[3])
            assert torch.allclose(b[1], s[1], rtol=1e-5, atol=1e-7)


def test_bigc_stages_vs_oracle():
    """Stage-wise: pooled entity encoding, encoder output and stretched means against the oracle's intermediates."""
    cfg = synth.tiny_vidvrd_config()
    st = synth.make_bigc_state(7, cfg)
    model = _model(cfg, st, "fp32_simt")
    P = bigc_inputs(cfg, 202, 12, 64, 30)
    with torch.no_grad():
        _, _, _, inter = ob.encode2decode(st, cfg, P, return_intermediates=True)
        P.to(DEV)
        _, _, _, _, ex = model.forward_debug(P)
    for name, tol in (("extra_avg", 1e-5), ("enti2enco", 1e-4), ("enco", 2e-4)):
        ref = inter[name].numpy()
        got = ex["extra" if name == "extra_avg" else name].cpu().numpy()
        err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-6)
        print(name, err)
        assert err <= tol, (name, err)


def test_bigc_state_dict_strictness():
    from vidsgg_big_b200 import bigc
    cfg = synth.tiny_vidvrd_config()
    st = synth.make_bigc_state(7, cfg)
    m = bigc.BIG_C_vidvrd(cfg)
    bad = dict(st); bad.pop("fc_i3d.0.weight")
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad)
    m.load_state_dict({"module." + k: v for k, v in st.items()})        # DataParallel prefix is stripped
    with pytest.raises(NotImplementedError):
        bigc.BIG_C_vidvrd(cfg, is_train=True)


def test_forward_packed_equals_forward():
    from vidsgg_big_b200 import evalapi, geometry
    cfg = synth.tiny_vidvrd_config()
    st = synth.make_bigc_state(7, cfg)
    model = _model(cfg, st, "3xtf32")
    props = [bigc_inputs(cfg, 950 + i, n, vl, None).to(DEV) for i, (n, vl) in enumerate([(7, 40), (12, 64), (1, 25), (20, 90)])]
    with torch.no_grad():
        ref = model(props, topk=5)
        packed = model.forward_packed(props, topk=5)
    for a, b in zip(ref, packed.per_video()):
        assert (a is None) == (b is None)
        if a is not None:
            assert all(torch.equal(x, y) for x, y in zip(a, b))
    tt = geometry.TrackTable.from_containers(props)
    A = evalapi.PackedRelations.from_packed_triplets(tt, packed)
    B = evalapi.PackedRelations.from_triplets(tt, [None if t is None else (t[0], t[1].mean(-1), t[2]) for t in ref])
    assert torch.equal(A.rel, B.rel) and torch.equal(A.vid_off, B.vid_off) and torch.allclose(A.scores, B.scores)


def test_bigc_vidor_max_proposals_vs_oracle():
    """VidOR upper end (180 proposals, tracks up to 600 frames, exp5 dims): device logits / attention vs the oracle run here,
    and identical triplets up to near-ties.  Exercises the 180-track role attention and the long-stretch conv/pool path."""
    cfg = synth.vidor_config()
    st = synth.make_bigc_state(2, cfg)
    model = _model(cfg, st, "3xtf32")
    P = synth.make_proposal(4321, 180, 900, 1324, 81, min_len=15, max_len=600)
    with torch.no_grad():
        _, ref_logits, ref_att = ob.encode2decode(st, cfg, P)
        ref = ob.construct_triplet(P, ref_logits, ref_att, 3)
        P.to(DEV)
        _, logits, att, so, _ = model.forward_debug(P)
        ret = model([P], topk=3)[0]
    same_so = (so.cpu() == torch.argmax(ref_att, -1).t()).all(1).numpy()
    err = (logits.cpu() - ref_logits).abs().numpy()[same_so].max() / ref_logits.abs().max().item()
    print("n=180: rel logit err %.2e, (s,o) agreement %d/%d" % (err, same_so.sum(), same_so.size))
    assert same_so.mean() >= 0.97 and err <= 3e-4
    assert (att.cpu() - ref_att).abs().max().item() <= 2e-4
    mine = {tuple(r) for r in ret[0].cpu().tolist()}
    theirs = {tuple(r) for r in ref[0].tolist()}
    assert len(mine ^ theirs) <= max(2, len(theirs) // 20)


def test_bigc_full_dims_batched_equals_single():
    cfg = synth.vidvrd_config()
    st = synth.make_bigc_state(1, cfg)
    model = _model(cfg, st, "3xtf32")
    props = [bigc_inputs(cfg, 960 + i, n, vl, 150).to(DEV) for i, (n, vl) in enumerate([(30, 300), (5, 90), (50, 700), (17, 240)])]
    with torch.no_grad():
        batched = model(props, topk=10)
        single = [model([p], topk=10)[0] for p in props]
    for b, s in zip(batched, single):
        assert torch.equal(b[0], s[0]) and torch.equal(b[2], s[2]) and torch.equal(b[3], s[3])
        assert torch.allclose(b[1], s[1], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("precision", ["3xtf32", "tf32"])
def test_tensor_core_attention_equals_simt_attention(precision):
    """Decoder self-attention as batched tcgen05 GEMMs (QK^T, PV) + softmax / transpose glue vs the fp32 SIMT kernel."""
    cfg = synth.vidvrd_config()
    st = synth.make_bigc_state(1, cfg)
    model = _model(cfg, st, precision)
    V, Q, d = 3, cfg["num_querys"], cfg["dim_pred"]
    g = torch.Generator(device="cpu").manual_seed(3)
    qkv = (torch.randn(V * Q, 3 * d, generator=g) * 0.7).to(DEV)
    lo = torch.empty_like(qkv)
    from vidsgg_big_b200._cabi import lib, check, ptr, stream_ptr
    hi = torch.empty_like(qkv)
    check(lib().vsg_split_tf32(ptr(qkv), ptr(hi), ptr(lo), qkv.numel(), stream_ptr(qkv.device)), "split")
    ref = model._mha(qkv, d, None, V, Q, Q)
    got = model._mha_tc(qkv, lo if precision == "3xtf32" else None, V, Q, d)
    # fp64 reference
    q, k, v = [t.double().view(V, Q, 8, d // 8).transpose(1, 2) for t in (qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:])]
    exact = (torch.softmax(q @ k.transpose(-1, -2) / (d // 8) ** 0.5, -1) @ v).transpose(1, 2).reshape(V * Q, d)
    e_ref = (ref.double() - exact).abs().max().item()
    e_got = (got.double() - exact).abs().max().item()
    fused = model._mha_tc64(qkv, d, V, Q)                              # the default path: one fused launch (fp16 pairs / fp16 hi parts)
    e_fused = (fused.double() - exact).abs().max().item()
    print("attention max err vs fp64: simt %.2e  tensor-core %s %.2e  fused %.2e" % (e_ref, precision, e_got, e_fused))
    assert e_fused < (2e-5 if precision == "3xtf32" else 5e-3)
    assert e_ref < 1e-5
    assert e_got < (2e-5 if precision == "3xtf32" else 5e-3)


@pytest.mark.parametrize("precision", ["tf32+bf16x2", "fp16x3"])
def test_vidvrd_test_size_batch_identical_triplets_and_recall(precision):
    """The benchmark's parity claim as a test: BASELINE configs[1] at full size (200 VidVRD-shaped videos, exp2 dims, both fp32-class bench modes)
    through the batched CUDA path vs the CPU oracle video by video on the SAME inputs: identical triplets (quintuples + spans) for every
    video and identical R@50 / R@100 of the packed evaluation vs the oracle's dict evaluation (north_star parity gate)."""
    import bench
    from oracle import convert as oc, evalapi as oe
    n_videos = 200
    seeds = [1000 + i for i in range(n_videos)]
    pipe = bench.Pipeline("vidvrd", precision, torch.device(DEV))
    cfg, wl, props, _, feats = bench.make_videos("vidvrd", seeds, DEV, seeds[0], with_gt=False)
    prep = bench.cpu_prepare("vidvrd", seeds, seeds[0], feats_from=(feats.cpu(), None))      # oracle triplets + GT from them
    for p in props:
        f = p.features
        p.to(DEV)
        p.features = f
    with torch.no_grad():
        trips = pipe.model(props, topk=wl["topk"])
    same, total = bench.compare_triplets(trips, prep["trips"])
    if precision == "fp16x3":
        # fp16x3 is the MORE accurate of the two (logit error 2-3x smaller, profiles/r02_parity_errors.txt), but the oracle is an fp32
        # computation with its own rounding: 19 of the 38 400 queries of this batch have a relative (s,o) arg-max margin below 1e-4 and
        # one of them falls on the other side (measured: 199 / 200 videos identical; the parity default stays tf32+bf16x2 at 200 / 200)
        assert total == n_videos and same >= n_videos - 2, "%d of %d videos have identical triplets" % (same, total)
        return
    assert total == n_videos and same == n_videos, "%d of %d videos have identical triplets" % (same, total)
    # evaluation: packed CUDA path on our triplets vs the oracle's dict evaluation on its own triplets, same GT
    import copy
    graphs = [copy.copy(g).to(DEV) for g in prep["graphs"]]
    (m_ap, r50, r100), n_rel, _ = pipe.step(props, graphs)
    en, pn = oc.default_names("e", 256), oc.default_names("p", 256)
    gts, prs = {}, {}
    for p, g, r in zip(prep["props"], prep["graphs"], prep["trips"]):
        prs.update(oc.to_eval_format_pr(p, None if r is None else (r[0], r[1].mean(-1), r[2]), en, pn))
        gts.update(oc.to_eval_format_gt(g, en, pn))
    ref = oe.evaluate(gts, prs)
    assert r50 == float(ref[1][50]) and r100 == float(ref[1][100])
    assert abs(m_ap - ref[0]) < 5e-5      # AP integrates over score ORDER: scores equal to 1e-6 can swap two near-tied predictions


@pytest.mark.parametrize("precision", ["fp32_simt", "tf32", "3xtf32", "tf32+bf16x2", "fp16x3", "bf16"])
@pytest.mark.parametrize("which", ["tiny_vidvrd", "tiny_vidor", "vidvrd", "vidor"])
def test_c_forward_entry_equals_python_issued_launches(which, precision):
    """vsg_bigc_forward (ONE C call: csrc/forward.cu, include/vsg_b200.h) against the same launches issued op by op from Python:
    bit-identical triplets / scores / spans / query ids / counts for a ragged batch, in every precision mode and both model variants;
    also through the CUDA-graph replay."""
    from vidsgg_big_b200 import bigc
    cfg = {"tiny_vidvrd": synth.tiny_vidvrd_config, "tiny_vidor": synth.tiny_vidor_config, "vidvrd": synth.vidvrd_config,
           "vidor": synth.vidor_config}[which]()
    if which == "tiny_vidor":
        cfg["_force_entiemb"] = False
    feat = cfg["dim_feat"] + (cfg.get("dim_i3d") or 0 if cfg["variant"] == "vidvrd" else cfg["dim_clsme"])
    st = synth.make_bigc_state(3, cfg)
    model = _model(cfg, st, precision)
    props = [synth.make_proposal(4000 + i, n, vl, feat, cfg["num_enti_cats"], min_len=5, max_len=60).to(DEV)
             for i, (n, vl) in enumerate([(7, 90), (1, 40), (19, 150), (12, 64)])]
    topk = 3 if cfg["variant"] == "vidor" else 10
    pk = model.pack(props)
    outs = {}
    with torch.no_grad():
        for backend in ("py", "c"):
            model.backend = backend
            outs[backend] = model.forward_packed(props, topk=topk, packed_videos=pk)
        model.backend = "c"
        outs["graph"] = model.forward_packed(props, topk=topk, packed_videos=pk, graph=True)
        outs["graph2"] = model.forward_packed(props, topk=topk, packed_videos=pk, graph=True)      # replay
    a = outs["py"]
    assert a.counts[:, 0].sum() > 0
    for k in ("c", "graph", "graph2"):
        b = outs[k]
        assert np.array_equal(a.counts, b.counts), k
        for v in range(len(props)):
            n = int(a.counts[v, 0])
            s = slice(v * a.cap, v * a.cap + n)
            assert torch.equal(a.quint[s], b.quint[s]) and torch.equal(a.scores[s], b.scores[s]), (k, v)
            assert torch.equal(a.spans[s], b.spans[s]) and torch.equal(a.qids[s], b.qids[s]), (k, v)
    # the reference-style per-video API goes through the same entry point
    model.backend = "c"
    res = model(props, topk=topk)
    for v, r in enumerate(res):
        n = int(a.counts[v, 0])
        assert (r is None) == (int(a.counts[v, 1]) == 0)
        if r is not None:
            assert torch.equal(r[0], a.quint[v * a.cap: v * a.cap + n])


def test_bipartite_cost_matrix_vs_reference(golden):
    """f3 'then' clause: the Hungarian cost matrix (model_0v10.py:606-636) on the device against the matrix the reference computed, and
    the same assignment from scipy."""
    g = golden("bipartite")
    cfg = synth.tiny_vidvrd_config()
    model = _model(cfg, synth.make_bigc_state(7, cfg), "fp32_simt")
    for sd, n_gt, n_enti in ((451, 5, 9), (452, 17, 23), (453, 1, 4)):
        logit, gt_pred, att, adj = [t.to(DEV) for t in synth.make_bipartite_case(sd, cfg["num_querys"], cfg["num_pred_cats"], n_gt, n_enti)]
        cost = model.bipartite_cost(logit, gt_pred, att, adj)
        np.testing.assert_allclose(cost.cpu().numpy(), g["cost_%d" % sd], rtol=2e-5, atol=2e-5)
        row, col = model.bipartite_match(logit, gt_pred, att, adj)
        assert np.array_equal(row, g["row_%d" % sd]) and np.array_equal(col, g["col_%d" % sd])


def test_bf16_feature_transport_equals_cast_path():
    """Opt-in bf16 feature transport of the bf16 mode: features handed over as bf16 give exactly the result of fp32 features that hold
    the same (bf16-representable) values -- through the Python-issued launches and through vsg_bigc_forward; the fp32-class modes reject
    bf16 features."""
    from vidsgg_big_b200 import bigc
    from vidsgg_big_b200._cabi import VsgError
    import copy
    cfg = synth.tiny_vidvrd_config(dim_feat=96, dim_i3d=40)          # 136 feature columns: row stride a multiple of 8
    st = synth.make_bigc_state(3, cfg)
    model = _model(cfg, st, "bf16")
    props32, props16 = [], []
    for i, (n, vl) in enumerate([(7, 90), (12, 64)]):
        p = synth.make_proposal(4100 + i, n, vl, 136, cfg["num_enti_cats"], min_len=5, max_len=60)
        p.features = p.features.to(torch.bfloat16).float()            # values a bf16 loader would deliver
        q = copy.copy(p)
        q.features = p.features.to(torch.bfloat16)
        props32.append(p.to(DEV)); props16.append(q.to(DEV))
    for backend in ("py", "c"):
        model.backend = backend
        a = model.forward_packed(props32, topk=5)
        b = model.forward_packed(props16, topk=5)
        assert np.array_equal(a.counts, b.counts) and a.counts[:, 0].sum() > 0
        for v in range(len(props32)):
            sl = slice(v * a.cap, v * a.cap + int(a.counts[v, 0]))
            assert torch.equal(a.quint[sl], b.quint[sl]) and torch.equal(a.scores[sl], b.scores[sl]) and torch.equal(a.spans[sl], b.spans[sl])
    strict = _model(cfg, st, "tf32+bf16x2")
    with pytest.raises(VsgError):
        strict.forward_packed(props16, topk=5)


@pytest.mark.parametrize("which", ["tiny_vidvrd", "mid128", "vidvrd", "vidor"])
def test_role_fold_equals_unfolded_role_attention(which):
    """fc_rolewise[r].0(att[r] @ enco) re-associated as att[r] @ (enco W_r^T) inside vsg_role_attention_hid (one GEMM over the tracks for all
    decoder layers, no V*Q-row `values`) against the reference's order of operations (vsg_role_attention + two V*Q-row GEMMs): same (s,o)
    arg-max for every query, attention matrix and logits equal to fp32 rounding, identical triplets; ragged videos incl. a 1-track one and
    more tracks than one staged chunk."""
    mid = lambda: synth.tiny_vidvrd_config(dim_enti=128, dim_pred=128, dim_att=128, dim_ffn=192, num_querys=40)     # Q % 16 != 0
    cfg = {"tiny_vidvrd": synth.tiny_vidvrd_config, "mid128": mid, "vidvrd": synth.vidvrd_config, "vidor": synth.vidor_config}[which]()
    feat = cfg["dim_feat"] + (cfg.get("dim_i3d") or 0 if cfg["variant"] == "vidvrd" else cfg["dim_clsme"])
    st = synth.make_bigc_state(5, cfg)
    props = [synth.make_proposal(4100 + i, n, vl, feat, cfg["num_enti_cats"], min_len=5, max_len=60).to(DEV)
             for i, (n, vl) in enumerate([(7, 90), (1, 40), (37, 150), (16, 64), (17, 70)])]
    outs = {}
    for fold in (True, False):
        model = _model(cfg, st, "fp32_simt")
        model.role_fold = fold
        model._prepare()
        assert ("eg_all" in model._w) == (fold and cfg["dim_enti"] in (128, 512))
        pk = model.pack(props)
        with torch.no_grad():
            logits, so, ex = model._encode2decode(pk, want_att=True)
            outs[fold] = (logits.clone(), so.clone(), ex["att"].clone(), model.forward_packed(props, topk=5, packed_videos=pk))
    if cfg["dim_enti"] not in (128, 512):
        pytest.skip("dim_enti %d: the folded kernel is instantiated for 128 / 512" % cfg["dim_enti"])
    (la, soa, atta, pa), (lb, sob, attb, pb) = outs[True], outs[False]
    assert torch.equal(soa, sob)
    assert (atta - attb).abs().max().item() <= 1e-5      # pass 1 sums the 256-d dots in another order (measured 2e-6)
    assert (la - lb).abs().max().item() <= 2e-5 * lb.abs().max().item()
    assert np.array_equal(pa.counts, pb.counts)
    for v in range(len(props)):
        s = slice(v * pa.cap, v * pa.cap + int(pa.counts[v, 0]))
        assert torch.equal(pa.quint[s], pb.quint[s]) and torch.equal(pa.spans[s], pb.spans[s])
