"""GPU parity: geometry kernels (C ABI) vs the oracle and the reference's golden outputs."""
import numpy as np
import pytest
import torch

from oracle import geometry as og
from vidsgg_big_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _geo():
    from vidsgg_big_b200 import geometry
    return geometry


def test_pair_ids_bit_exact(golden):
    g = _geo()
    assert np.array_equal(g.trajid2pairid(7, DEV).cpu().numpy(), golden("geometry")["pair_ids_7"])
    for n in (0, 1, 2, 3, 50, 181):
        assert torch.equal(g.trajid2pairid(n, DEV).cpu(), og.pair_ids(n))


def test_spans_bit_exact(golden):
    g = _geo()
    gg = golden("geometry")
    rng = np.random.default_rng(11)
    s = rng.integers(0, 200, size=(23, 1)); d1 = np.concatenate([s, s + rng.integers(0, 120, size=(23, 1))], 1)
    s = rng.integers(0, 200, size=(17, 1)); d2 = np.concatenate([s, s + rng.integers(0, 120, size=(17, 1))], 1)
    t1, t2 = torch.from_numpy(d1).to(DEV), torch.from_numpy(d2).to(DEV)
    inter, mask = g.dura_intersection_ts(t1, t2)
    assert inter.dtype == torch.long and mask.dtype == torch.bool
    assert np.array_equal(inter.cpu().numpy(), gg["dura_inter"]) and np.array_equal(mask.cpu().numpy(), gg["dura_mask"])
    inter, mask = g.dura_intersection_ts(t1[:17], t2, broadcast=False)
    assert np.array_equal(inter.cpu().numpy(), gg["dura_inter_nb"]) and np.array_equal(mask.cpu().numpy(), gg["dura_mask_nb"])
    k = torch.tensor([[0, 10], [5, 20], [30, 40]], device=DEV)
    ki, km = g.dura_intersection_ts(k, k)
    assert ki[0, 2].tolist() == [30, 10] and not bool(km[0, 2])        # inverted span kept, like the reference
    # float spans (grounding), large int64 values
    f1, f2 = (t1.float() / 320), (t2.float() / 320)
    fi, fm = g.dura_intersection_ts(f1, f2)
    oi, om = og.dura_intersection(f1.cpu(), f2.cpu())
    assert torch.equal(fi.cpu(), oi) and torch.equal(fm.cpu(), om)
    big = torch.tensor([[2**40, 2**40 + 5], [2**40 + 3, 2**41]], device=DEV)
    bi, bm = g.dura_intersection_ts(big, big)
    oi, om = og.dura_intersection(big.cpu(), big.cpu())
    assert torch.equal(bi.cpu(), oi) and torch.equal(bm.cpu(), om)


@pytest.mark.parametrize("variant", [1, 2])
def test_viou_matrix_golden(golden, variant):
    g = _geo()
    gg = golden("geometry")
    for tag, seed, n, vl in (("a", 101, 12, 160), ("b", 102, 20, 90)):
        P = synth.make_proposal(seed, n, vl, 8, 36, with_features=False).to(DEV)
        G = synth.make_gt_graph(seed, P, 133).to(DEV)
        for side, (bx, du) in (("pp", (P.bboxes_list, P.traj_durations)), ("pg", (G.traj_bboxes, G.traj_durations))):
            viou, inter, mask = g.traj_viou_matrix(P.bboxes_list, P.traj_durations, bx, du, variant=variant)
            ref = gg["viou_%s_%s" % (side, tag)]
            np.testing.assert_allclose(viou.cpu().numpy(), ref, rtol=1e-5, atol=1e-7)     # north_star: fp32 vIoU <= 1e-5 rel
            assert np.array_equal(inter.cpu().numpy(), gg["inter_%s_%s" % (side, tag)])  # spans bit-exact
            assert np.array_equal(mask.cpu().numpy(), ref_mask(gg["inter_%s_%s" % (side, tag)]))
            assert np.all(viou.cpu().numpy()[~mask.cpu().numpy()] == 0.0)


def ref_mask(inter):
    return inter[..., 0] <= inter[..., 1]


@pytest.mark.parametrize("variant", [1, 2])
def test_viou_batched_ragged(variant):
    """Several videos (ragged, incl. a 1-track video and a single-frame track) in one launch vs the oracle."""
    g = _geo()
    props, gts = [], []
    for sd, n, vl in ((1, 33, 300), (2, 1, 40), (3, 70, 200), (4, 5, 17), (5, 40, 1200)):
        P = synth.make_proposal(sd, n, vl, 8, 36, with_features=False, min_len=1)
        props.append(P)
        gts.append(synth.make_gt_graph(sd, P, 133))
    A = g.TrackTable.from_containers([p.to(DEV) for p in props])
    B = g.TrackTable.from_containers([x.to(DEV) for x in gts])
    for X, Y, ylist in ((A, A, props), (A, B, gts)):
        viou, spans, mask, seg, _ = g.traj_viou_batched(X, Y, variant=variant)
        viou, spans, mask = viou.cpu(), spans.cpu(), mask.cpu()
        for v, P in enumerate(props):
            Q = ylist[v]
            P.to("cpu"); Q.to("cpu")
            qb = Q.bboxes_list if hasattr(Q, "bboxes_list") else Q.traj_bboxes
            o_v, o_i, o_m = og.traj_viou_matrix(P.bboxes_list, P.traj_durations, qb, Q.traj_durations)
            sl = slice(seg[v], seg[v + 1])
            np.testing.assert_allclose(viou[sl].numpy(), o_v.reshape(-1).numpy(), rtol=1e-5, atol=1e-7)
            assert torch.equal(spans[sl], o_i.reshape(-1, 2)) and torch.equal(mask[sl].bool(), o_m.reshape(-1))


@pytest.mark.parametrize("variant", [1, 2])
def test_viou_properties_large(variant):
    """Stress-shaped (scaled) input: size-independent properties -- symmetry, diagonal == 1, range, agreement
    between the two kernel variants, and a sampled comparison with the numpy oracle."""
    g = _geo()
    P = synth.make_proposal(9, 96, 4000, 8, 36, with_features=False, min_len=800).to(DEV)
    T = g.TrackTable.from_containers([P])
    viou, spans, mask, seg, _ = g.traj_viou_batched(T, T, variant=variant)
    n = P.num_proposals
    V = viou.view(n, n)
    assert torch.allclose(V, V.t(), rtol=1e-6, atol=0)
    assert torch.allclose(torch.diagonal(V), torch.ones(n, device=DEV), rtol=1e-6)
    assert bool((V >= 0).all()) and bool((V <= 1 + 1e-6).all())
    other, _, _, _, _ = g.traj_viou_batched(T, T, variant=3 - variant)
    assert torch.allclose(viou, other, rtol=2e-6, atol=1e-9)
    # sampled oracle rows
    bx = P.bboxes.cpu().numpy(); du = P.traj_durations.cpu().numpy()
    off = np.concatenate([[0], np.cumsum(P.lengths.numpy())])
    rows = [0, 17, 95]
    sub_off = np.concatenate([[0], np.cumsum([off[r + 1] - off[r] for r in rows])])
    sub_bx = np.concatenate([bx[off[r]:off[r + 1]] for r in rows], 0)
    ref = og.traj_viou_matrix_np(sub_bx, sub_off, du[rows], bx, off, du)
    np.testing.assert_allclose(V[rows].cpu().numpy(), ref, rtol=1e-5, atol=1e-7)


def test_vIoU_ts_single_pair():
    g = _geo()
    P = synth.make_proposal(21, 6, 80, 8, 36, with_features=False)
    inter, mask = og.dura_intersection(P.traj_durations, P.traj_durations)
    bl = P.bboxes_list
    done = 0
    for a in range(6):
        for b in range(6):
            if a == b or not mask[a, b]:
                continue
            ra = inter[a, b] - P.traj_durations[a, 0]; rb = inter[a, b] - P.traj_durations[b, 0]
            ref = og.viou_single(bl[a], bl[b], ra, rb)
            got = g.vIoU_ts(bl[a].to(DEV), bl[b].to(DEV), ra.to(DEV), rb.to(DEV))
            assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item()) + 1e-8
            done += 1
    assert done > 0


def test_align_and_pair_labels(golden):
    g = _geo()
    ga = golden("align")
    cfg = synth.tiny_vidvrd_config()
    for sd in (401, 402, 403):
        P = synth.make_proposal(sd, 14, 120, 8, cfg["num_enti_cats"], with_features=False)
        G = synth.make_gt_graph(sd, P, cfg["num_pred_cats"], n_rel=(3, 12))
        aligned, viou = g.enti_viou_align(G.adj_matrix.to(DEV), P.to(DEV), G.to(DEV), 0.5)
        np.testing.assert_allclose(viou.cpu().numpy(), ga["viou_%d" % sd], rtol=1e-5, atol=1e-7)
        assert np.array_equal(aligned.cpu().numpy(), ga["align_%d" % sd])
        so = torch.argmax(G.adj_matrix, dim=-1).t().contiguous()
        lab = g.pair_labels(viou, so, 0.5)
        ref = og.pair_labels(torch.from_numpy(ga["viou_%d" % sd]), so.cpu(), 0.5)
        assert torch.equal(lab.cpu(), ref)


def test_errors_are_loud():
    g = _geo()
    from vidsgg_big_b200._cabi import VsgError
    with pytest.raises(VsgError):
        g.dura_intersection_ts(torch.zeros(2, 2, dtype=torch.long), torch.zeros(2, 2, dtype=torch.long))   # CPU tensors


def test_prop_pair_to_gt_pred_dataset():
    """Batched Base-C label assignment == the reference's dict loop restated in the oracle (order of pairs included)."""
    g = _geo()
    props, graphs = [], []
    for sd, n in ((801, 14), (802, 9), (803, 1), (804, 20)):
        P = synth.make_proposal(sd, n, 120, 8, 36, with_features=False)
        props.append(P); graphs.append(synth.make_gt_graph(sd, P, 133, n_rel=(3, 12), jitter_px=1.5))
    ref = {}
    for P, G in zip(props, graphs):
        viou, _, _ = og.traj_viou_matrix(P.bboxes_list, P.traj_durations, G.traj_bboxes, G.traj_durations)
        so = torch.argmax(G.adj_matrix, dim=-1).t()
        cats = G.traj_cat_ids[so]
        gt5 = torch.cat([G.pred_cat_ids[:, None], cats, so], -1)
        ref[G.video_name] = og.label_maps(viou, gt5, 0.5, 133) if P.num_proposals > 1 else None
    out, stats = g.prop_pair_to_gt_pred([p.to(DEV) for p in props], [x.to(DEV) for x in graphs], 0.5, 133)
    # like the reference (train_vidor.py:99-101) videos without GT relations are skipped before counting
    assert stats["gt_traj"] == sum(x.num_trajs for x in graphs if x.num_preds > 0) and stats["hit_gt_traj"] > 0
    n_checked = 0
    for name, r in ref.items():
        if r is None:
            assert out[name] is None
            continue
        assert torch.equal(out[name][0].cpu(), r[0]), name
        assert torch.equal(out[name][1].cpu(), r[1]), name
        n_checked += 1
    assert n_checked >= 2


def test_row_block_sharding_reproduces_the_full_matrix():
    """shard.traj_viou_row_sharded's per-rank work (rows [r0, r1) of one video against all of its tracks) for world = 3,
    concatenated: bit-identical to the single-launch matrix (the N > 1 exchange itself is covered by the gloo CPU test)."""
    from vidsgg_big_b200 import geometry, shard, synth
    P = synth.make_proposal(77, 23, 400, 8, 36, min_len=30, max_len=300, with_features=False).to("cuda:0")
    boxes, dura = P.bboxes_list, P.traj_durations
    n = len(boxes)
    T = geometry.TrackTable.from_lists(boxes, dura)
    full_v, full_sp, full_m, _, _ = geometry.traj_viou_batched(T, T)
    vs, sps, ms = [], [], []
    for r in range(3):
        r0, r1 = shard.row_block(n, r, 3)
        A = geometry.TrackTable.from_lists(boxes[r0:r1], dura[r0:r1])
        v, sp, m, _, _ = geometry.traj_viou_batched(A, T)
        vs.append(v); sps.append(sp); ms.append(m)
    assert torch.equal(torch.cat(vs), full_v) and torch.equal(torch.cat(sps), full_sp) and torch.equal(torch.cat(ms), full_m)
    v1, sp1, m1 = shard.traj_viou_row_sharded(boxes, dura)            # world size 1: no collective
    assert torch.equal(v1.reshape(-1), full_v) and torch.equal(sp1.reshape(-1, 2), full_sp)


def test_tiou_and_generalized_tiou(golden):
    """tIoU / generalized_tIoU (utils/utils_func.py:375-410) against the reference goldens (float spans) and the oracle (int64 spans,
    row-wise mode)."""
    from vidsgg_big_b200 import geometry
    g = golden("geometry")
    rng = np.random.default_rng(11)
    s = rng.integers(0, 200, size=(23, 1)); d1 = np.concatenate([s, s + rng.integers(0, 120, size=(23, 1))], 1)
    s = rng.integers(0, 200, size=(17, 1)); d2 = np.concatenate([s, s + rng.integers(0, 120, size=(17, 1))], 1)
    t1, t2 = torch.from_numpy(d1), torch.from_numpy(d2)
    f1, f2 = (t1.float() / 320).cuda(), (t2.float() / 320).cuda()
    np.testing.assert_allclose(geometry.tIoU(f1, f2).cpu().numpy(), g["tiou"], rtol=1e-6, atol=0)
    np.testing.assert_allclose(geometry.generalized_tIoU(f1, f2).cpu().numpy(), g["gtiou"], rtol=1e-6, atol=0)
    # int64 spans (true division in float32) incl. the row-wise form
    a, b = t1[:17].cuda(), t2.cuda()
    for fn, ofn in ((geometry.tIoU, og.tiou), (geometry.generalized_tIoU, og.generalized_tiou)):
        np.testing.assert_allclose(fn(a, b).cpu().numpy(), ofn(t1[:17], t2).numpy(), rtol=1e-6)
        np.testing.assert_allclose(fn(a, b, broadcast=False).cpu().numpy(), ofn(t1[:17], t2, broadcast=False).numpy(), rtol=1e-6)
    assert geometry.tIoU(a[:0], b).shape == (0, 17)


def test_unique_with_idx_nd(golden):
    """unique_with_idx_nd (utils/utils_func.py:330-345): SURVEY KAT, the reference golden, random rows vs the oracle, N-d rows."""
    from vidsgg_big_b200 import geometry
    g = golden("geometry")
    kat = torch.tensor([[3, 1, 2, 0, 1], [1, 1, 2, 0, 1], [3, 1, 2, 0, 1], [1, 0, 2, 5, 1]])
    u, groups = geometry.unique_with_idx_nd(kat.cuda())
    assert u.cpu().tolist() == [[1, 0, 2, 5, 1], [1, 1, 2, 0, 1], [3, 1, 2, 0, 1]]
    assert [x.cpu().tolist() for x in groups] == [[3], [1], [0, 2]]
    rng = np.random.default_rng(11)
    rng.integers(0, 200, size=(23, 1)); rng.integers(0, 120, size=(23, 1)); rng.integers(0, 200, size=(17, 1)); rng.integers(0, 120, size=(17, 1))
    rnd = torch.from_numpy(rng.integers(0, 3, size=(60, 5)))
    u, groups = geometry.unique_with_idx_nd(rnd.cuda())
    assert np.array_equal(u.cpu().numpy(), g["uniq_rows"])
    assert np.array_equal(np.array([int(x[0]) for x in groups]), g["uniq_first"])
    assert np.array_equal(np.array([len(x) for x in groups]), g["uniq_counts"])
    for n, d, hi in ((1, 3, 2), (777, 5, 3), (1920, 5, 4), (4097, 2, 50), (300, 1, 7)):
        t = torch.from_numpy(np.random.default_rng(n).integers(-hi, hi, size=(n, d)))
        u, groups = geometry.unique_with_idx_nd(t.cuda())
        ou, ogr = og.unique_rows_with_groups(t)
        assert torch.equal(u.cpu(), ou) and len(groups) == len(ogr)
        assert all(torch.equal(a.cpu(), b) for a, b in zip(groups, ogr))
    t3 = torch.from_numpy(np.random.default_rng(5).integers(0, 2, size=(40, 2, 2)))           # rows of shape (2, 2)
    u, groups = geometry.unique_with_idx_nd(t3.cuda())
    ou, ogr = og.unique_rows_with_groups(t3)
    assert torch.equal(u.cpu(), ou) and all(torch.equal(a.cpu(), b) for a, b in zip(groups, ogr))
    u, groups = geometry.unique_with_idx_nd(kat[:0].cuda())
    assert u.shape[0] == 0 and groups == tuple()


def test_stack_with_repeat_2d(golden):
    """stack_with_repeat_2d (models/model_0v10.py:18-46): the stretch index maps of the reference goldens and both ragged axes."""
    from vidsgg_big_b200 import geometry
    g = golden("geometry")
    for L, T in ((3, 7), (5, 5), (4, 13), (7, 8), (1, 6), (6, 29)):
        out = geometry.stack_with_repeat_2d([torch.arange(L)[:, None].float().cuda(), torch.zeros(T, 1).cuda()], dim=0)
        assert out.shape == (2, T, 1)
        assert np.array_equal(out[0, :, 0].long().cpu().numpy(), g["stretch_%d_%d" % (L, T)])
    rng = np.random.default_rng(3)
    lens = [5, 17, 9, 17, 1]
    ts = [torch.from_numpy(rng.standard_normal((l, 6)).astype(np.float32)) for l in lens]
    want = torch.stack([t[torch.from_numpy(og.stretch_index_map(l, 17))] for t, l in zip(ts, lens)], 0)
    got = geometry.stack_with_repeat_2d([t.cuda() for t in ts], dim=0)
    assert torch.equal(got.cpu(), want)
    got1 = geometry.stack_with_repeat_2d([t.cuda() for t in ts], dim=1)                       # stacked along dim 1
    assert torch.equal(got1.cpu(), want.permute(1, 0, 2))
    cols = [t.t().contiguous() for t in ts]                                                   # ragged axis = columns
    gotc = geometry.stack_with_repeat_2d([t.cuda() for t in cols], dim=0)
    assert torch.equal(gotc.cpu(), want.permute(0, 2, 1))


def test_viou_properties_full_stress_size():
    """BASELINE.json configs[4] at FULL size (256 tracklets, 10k frames, ~270 M frame pairs): symmetry, unit diagonal, range, span
    symmetry / validity, agreement of the warp-per-pair and the tiled kernels, and sampled rows against the numpy oracle."""
    g = _geo()
    P = synth.make_proposal(4242, 256, 10000, 8, 36, min_len=2000, max_len=10000, with_features=False).to(DEV)
    T = g.TrackTable.from_containers([P])
    viou, spans, mask, seg, _ = g.traj_viou_batched(T, T, variant=1)
    n = 256
    V, S, M = viou.view(n, n), spans.view(n, n, 2), mask.view(n, n).bool()
    assert torch.allclose(V, V.t(), rtol=2e-6, atol=0)
    assert torch.allclose(torch.diagonal(V), torch.ones(n, device=DEV), rtol=2e-6)
    assert bool((V >= 0).all()) and bool((V <= 1 + 1e-6).all())
    assert torch.equal(S, S.transpose(0, 1)) and torch.equal(M, M.t())
    d = P.traj_durations
    assert torch.equal(S[..., 0], torch.maximum(d[:, None, 0], d[None, :, 0])) and torch.equal(S[..., 1], torch.minimum(d[:, None, 1], d[None, :, 1]))
    assert torch.equal(M, S[..., 0] <= S[..., 1]) and bool((V[~M] == 0).all()) and bool((V[M] > 0).any())
    tiled, _, _, _, _ = g.traj_viou_batched(T, T, variant=2)
    assert torch.allclose(viou, tiled, rtol=3e-6, atol=1e-9)
    bx = P.bboxes.cpu().numpy(); du = d.cpu().numpy()
    off = np.concatenate([[0], np.cumsum(P.lengths.numpy())])
    rows = [3, 200]
    sub_off = np.concatenate([[0], np.cumsum([off[r + 1] - off[r] for r in rows])])
    sub_bx = np.concatenate([bx[off[r]:off[r + 1]] for r in rows], 0)
    ref = og.traj_viou_matrix_np(sub_bx, sub_off, du[rows], bx, off, du)
    np.testing.assert_allclose(V[rows].cpu().numpy(), ref, rtol=1e-5, atol=1e-7)
