"""CPU tier: host logic, C-ABI symbol export, packaging rules (no GPU needed, no compute calls into the library)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    from vidsgg_big_b200 import _cabi, build
    build.build()
    lib = _cabi.lib()
    header = open(os.path.join(ROOT, "include", "vsg_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(vsg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), "libvsgb200.so does not export %s" % name
        assert name in _cabi.SIGNATURES, "%s has no ctypes signature" % name
    assert set(_cabi.SIGNATURES) <= declared | {"vsg_launch_count"}
    assert lib.vsg_built_for_sm() == 100 and lib.vsg_version() >= 1


def test_library_is_sm100a_with_tcgen05_and_tma():
    from vidsgg_big_b200 import _cabi
    try:
        sass = subprocess.run(["cuobjdump", "-sass", _cabi.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
    except Exception:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass or "UTCMMA" in sass          # tcgen05.mma
    assert "UTMALDG" in sass and "LDTM" in sass           # TMA loads, tcgen05.ld


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vidsgg_big_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_no_cpu_fallback():
    from vidsgg_big_b200 import geometry, evalapi, bigc, grounding, synth
    from vidsgg_big_b200._cabi import VsgError
    with pytest.raises(VsgError):
        geometry.dura_intersection_ts(torch.zeros(1, 2, dtype=torch.long), torch.zeros(1, 2, dtype=torch.long))
    if not torch.cuda.is_available():
        with pytest.raises(VsgError):
            evalapi.viou([[0, 0, 1, 1]], [0, 1], [[0, 0, 1, 1]], [0, 1])
    cfg = synth.tiny_vidvrd_config()
    m = bigc.BIG_C_vidvrd(cfg)
    m.load_state_dict(synth.make_bigc_state(7, cfg))
    with pytest.raises(VsgError):
        m([synth.make_proposal(1, 3, 20, 136, 9)], topk=5)          # weights never moved to a CUDA device
    with pytest.raises(VsgError):
        m.to("cpu")


def test_containers_roundtrip():
    from vidsgg_big_b200 import synth
    from vidsgg_big_b200.containers import TrajProposal
    P = synth.make_proposal(5, 6, 80, 12, 36)
    assert P.num_proposals == 6 and len(P.bboxes_list) == 6 and len(P.features_list) == 6
    assert all(b.shape[0] == int(e - s + 1) for b, (s, e) in zip(P.bboxes_list, P.traj_durations.tolist()))
    assert P.bboxes_list[2].data_ptr() == P.bboxes[int(P.lengths[:2].sum()):].data_ptr()        # zero-copy views
    Q = TrajProposal.from_lists(P.video_name, P.video_len, P.video_wh, P.cat_ids, P.scores, P.traj_durations, P.bboxes_list, P.features_list)
    assert torch.equal(Q.bboxes, P.bboxes) and torch.equal(Q.features, P.features)
    assert bool((P.scores[:-1] >= P.scores[1:]).all())
    E = synth.make_proposal(5, 0, 80, 12, 36)
    assert E.num_proposals == 0 and E.to("cpu") is E


def test_lpt_sharding_and_records():
    from vidsgg_big_b200 import shard, evalapi
    costs = [100, 1, 1, 1, 50, 49, 3, 2]
    shards = shard.assign_lpt(costs, 3)
    assert sorted(i for s in shards for i in s) == list(range(8))
    loads = [sum(costs[i] for i in s) for s in shards]
    assert max(loads) == 100 and min(loads) >= 50
    assert shard.assign_lpt([], 2) == [[], []]
    # records -> metrics equals the reference-style aggregation
    rng = np.random.default_rng(0)
    vids, ngt, hits, tags = [], [], [], []
    for v in range(7):
        n = int(rng.integers(1, 140))
        sc = rng.uniform(0, 1, n)
        sc[rng.uniform(size=n) < 0.7] = -np.inf
        vids.append(v); ngt.append(int(rng.integers(1, 30))); hits.append(sc)
        tags.append(np.sort(rng.uniform(0, 1, int(rng.integers(0, 12))))[::-1].astype(np.float32))
    a = evalapi._aggregate(vids, ngt, hits, tags, [50, 100], [1, 5, 10])
    rec = evalapi.per_video_records(vids, ngt, hits, tags)
    b = evalapi.metrics_from_records(rec)
    assert a[0] == b[0] and all(a[1][k] == b[1][k] for k in (50, 100)) and all(np.isclose(a[2][k], b[2][k], rtol=1e-7) for k in (1, 5, 10))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from vidsgg_big_b200 import shard
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rec = torch.arange((rank + 1) * 3 * 8, dtype=torch.float64).reshape(-1, 8) + 1000 * rank
    out = shard.gather_records(rec)
    q.put((rank, out.shape[0], float(out.sum())))
    dist.destroy_process_group()


def test_gather_records_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    exp_rows = 3 + 6
    exp_sum = float(torch.arange(24, dtype=torch.float64).sum() + (torch.arange(48, dtype=torch.float64) + 1000).sum())
    assert all(r[1] == exp_rows and r[2] == exp_sum for r in res)


def _gloo_rows_worker(rank, world, port, q):
    import torch.distributed as dist
    from vidsgg_big_b200 import shard
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n = 7                                                    # 7 rows over 2 ranks: blocks of 4 and 3
    full = torch.arange(n * 5 * 2, dtype=torch.int64).reshape(n, 5, 2)
    r0, r1 = shard.row_block(n, rank, world)
    out = shard.gather_row_blocks(full[r0:r1].clone(), n)
    q.put((rank, (r0, r1), bool(torch.equal(out, full))))
    dist.destroy_process_group()


def test_gather_row_blocks_gloo_world2():
    """Row-block exchange of the single-video stress configuration (pair matrix split by rows over the ranks)."""
    import torch.multiprocessing as mp
    from vidsgg_big_b200 import shard
    assert [shard.row_block(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard.row_block(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_rows_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert res == [(0, (0, 4), True), (1, (4, 7), True)]


def test_native_eval_records_equal_numpy_aggregation():
    """csrc/evalhost.cu (host code of the library) restates the numpy per-video aggregation exactly."""
    import ctypes as C
    from vidsgg_big_b200 import evalapi, _cabi
    rng = np.random.default_rng(0)
    V, po, go, hits, orders, ptr_, gtr = 9, [0], [0], [], [], [], []
    for v in range(V):
        n = int(rng.integers(0, 160)); ng = int(rng.integers(0, 20)) if v != 3 else 0
        sc = np.round(rng.uniform(0, 1, n), 2)
        hits.append(np.where(rng.uniform(size=n) < 0.3, np.sort(sc)[::-1], -np.inf))
        orders.append(rng.permutation(n).astype(np.int32))
        ptr_.append(rng.integers(0, 4, size=(n, 3))); gtr.append(rng.integers(0, 4, size=(ng, 3)))
        po.append(po[-1] + n); go.append(go[-1] + ng)
    hit = np.ascontiguousarray(np.concatenate(hits)); order = np.ascontiguousarray(np.concatenate(orders))
    ptrip = np.ascontiguousarray(np.concatenate(ptr_).astype(np.int64)); gtrip = np.ascontiguousarray(np.concatenate(gtr).astype(np.int64))
    po, go = np.array(po, np.int64), np.array(go, np.int64)
    det, tag = np.array([50, 100], np.int32), np.array([1, 5, 10], np.int32)
    rec = np.zeros((V, 8))
    hp = lambda a: C.c_void_p(a.ctypes.data)
    n = _cabi.lib().vsg_eval_records_host(hp(hit), hp(order), hp(ptrip), hp(po), hp(gtrip), hp(go), V, hp(det), 2, hp(tag), 3, hp(rec))
    vids, ngt, hh, tags = [], [], [], []
    for v in range(V):
        if go[v + 1] - go[v] == 0:
            continue
        o = order[po[v]:po[v + 1]]
        trip = [tuple(t) for t in ptrip[po[v]:po[v + 1]][o].tolist()]
        tags.append(evalapi._tagging_from_ids([tuple(t) for t in gtrip[go[v]:go[v + 1]].tolist()], trip, [1.0] * len(trip))[0])
        vids.append(v); ngt.append(int(go[v + 1] - go[v])); hh.append(hit[po[v]:po[v + 1]])
    ref = evalapi.per_video_records(vids, ngt, hh, tags)
    assert n == ref.shape[0] == 8
    assert np.array_equal(rec[:n, [0, 2, 3, 4, 5, 6, 7]], ref[:, [0, 2, 3, 4, 5, 6, 7]])
    assert np.abs(rec[:n, 1] - ref[:, 1]).max() < 1e-14


def test_basec_host_surface_without_gpu():
    """Base_C mirrors the reference constructor contract on the host: training mode and the reference's unrunnable use_clsme=False
    configuration are rejected, a checkpoint with wrong keys fails strict loading, and nothing runs without a CUDA device."""
    from vidsgg_big_b200 import Base_C, synth
    from vidsgg_big_b200._cabi import VsgError
    cfg = synth.tiny_basec_config()
    with pytest.raises(NotImplementedError):
        Base_C(cfg, is_train=True)
    with pytest.raises(VsgError):
        Base_C(synth.tiny_basec_config(use_clsme=False))
    m = Base_C(cfg)
    st = synth.make_basec_state(1, cfg)
    assert sorted(st.keys()) == sorted(m._expected_keys_static())
    with pytest.raises(RuntimeError):
        m.load_state_dict({k: v for k, v in st.items() if k != "bias_matrix"})
    m.load_state_dict(st)
    with pytest.raises(VsgError):
        m.to("cpu")                                            # weights are staged in HBM only (no CPU fallback)
    with pytest.raises(VsgError):
        Base_C(cfg)([synth.make_proposal(1, 3, 20, cfg["dim_feat"] + cfg["dim_clsme"], cfg["num_enti_cats"])])


def test_bench_triplet_comparison_helper():
    """bench.py's parity_vs_cpu_oracle counts videos whose (quintuple, span) SETS agree, independent of row order."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    q = torch.tensor([[1, 2, 3, 0, 1], [4, 2, 3, 1, 0]])
    sp = torch.tensor([[0, 9], [3, 7]])
    a = (q, torch.zeros(2, 3), sp, torch.zeros(2))
    b = (q.flip(0), torch.zeros(2, 3), sp.flip(0), torch.zeros(2))
    c = (q, torch.zeros(2, 3), sp + 1, torch.zeros(2))
    assert bench.compare_triplets([a, None, a, a], [b, None, c, None]) == (2, 4)


def test_fused_dwconv_predicate():
    from vidsgg_big_b200 import linalg

    class W(object):
        def __init__(self, n, k, has16=True):
            self.N, self.K, self.w16 = n, k, object() if has16 else None
    assert linalg.can_fuse_dwconv(linalg.TF32_BF16X2, W(128, 128), k=7)
    assert linalg.can_fuse_dwconv(linalg.TF32_BF16X2, W(20, 128), k=3)
    assert not linalg.can_fuse_dwconv(linalg.X3TF32, W(128, 128), k=7)          # only the default split mode has the CONV variant
    assert not linalg.can_fuse_dwconv(linalg.TF32_BF16X2, W(256, 128), k=7)     # 128-wide tiles only
    assert not linalg.can_fuse_dwconv(linalg.TF32_BF16X2, W(128, 120), k=7)     # K must be a multiple of 16
    assert not linalg.can_fuse_dwconv(linalg.TF32_BF16X2, W(128, 128), k=4)     # odd taps only
    assert not linalg.can_fuse_dwconv(linalg.TF32_BF16X2, W(128, 512), k=7)     # depthwise weights must fit the kernel's table
    assert linalg.attention_mode(linalg.TF32_BF16X2) == linalg.X3TF32 and linalg.attention_mode(linalg.TF32) == linalg.TF32


def test_packed_relations_spread_keeps_empty_videos():
    """Videos without predictions stay in the evaluation as empty prediction segments (driver.py; tools/eval_vidor.py:100-103)."""
    from vidsgg_big_b200 import evalapi
    z = lambda *s, dt=torch.long: torch.zeros(*s, dtype=dt)
    rel = torch.arange(5 * 7).reshape(5, 7)
    pr = evalapi.PackedRelations(z(0, 4, dt=torch.float32), z(1), z(0), rel, torch.tensor([0, 2, 5]), torch.ones(5, dtype=torch.float64))
    sp = pr.spread([1, 4], 6)
    assert sp.n_vid == 6 and sp.vid_off.tolist() == [0, 0, 2, 2, 2, 5, 5] and sp.rel is pr.rel and sp.n_rel == 5
    assert pr.spread([0, 1], 2) is pr
    with pytest.raises(AssertionError):
        pr.spread([3, 1], 6)


def test_vidor_set_plan_is_deterministic_balanced_and_complete():
    """bench.py's plan of the 835-video VidOR-val-shaped set: every video in exactly one chunk of exactly one rank at every world size,
    LPT shards balanced to < 1 % in the useful-flop cost model, chunks within the row budget, the same per-video shapes for every
    world size (strong scaling shards ONE set), and the cost model agrees with SURVEY 8d's validated flop counts."""
    import bench
    from vidsgg_big_b200 import shard
    assert abs(shard.flops_grd(150, 200) / 1e9 - 28.167) < 0.01                 # 28.167 GF (re-associated) at T=150, nq=200
    ref = None
    for world in (1, 2, 4, 8):
        info, shards, plan = bench.vidor_set_plan(835, world, 2_500_000, parity_videos=6 if world == 1 else 0)
        if ref is None:
            ref = [(v["seed"], v["video_len"], v["n"], v["rows"]) for v in info]
            assert max(v["tmax"] for v in info) > 4000 and max(v["video_len"] for v in info) == 5400     # tracks up to the whole video
        assert [(v["seed"], v["video_len"], v["n"], v["rows"]) for v in info] == ref
        flat = sorted(i for chunks in plan for c in chunks for i in c)
        assert flat == list(range(835))
        for r, chunks in enumerate(plan):
            assert sorted(i for c in chunks for i in c) == sorted(shards[r])
            for c in chunks:
                rows = sum(info[i]["rows"] for i in c)
                assert rows <= 2_500_000 or len(c) == 1
        costs = [sum(info[i]["cost"] for i in s) for s in shards]
        assert max(costs) / (sum(costs) / world) < 1.01
    info, _, plan = bench.vidor_set_plan(835, 1, 2_500_000, parity_videos=6)
    assert len(plan[0][0]) == 6 and all(info[i]["n"] * info[i]["tmax"] <= 250_000 for i in plan[0][0])
    from vidsgg_big_b200 import synth
    assert len(synth.proposal_lengths(info[0]["seed"], info[0]["n"], info[0]["video_len"], 15, None)) == info[0]["n"]


def test_c_structs_match_the_header_layout():
    """ctypes mirrors of the whole-forward structs (include/vsg_b200.h): compiled sizes == ctypes sizes (a drifted field would make
    vsg_bigc_forward / vsg_grd_forward read garbage)."""
    import ctypes as C
    import shutil
    import tempfile
    from vidsgg_big_b200 import _cabi
    if shutil.which("g++") is None:
        pytest.skip("no host compiler")
    names = ["VsgLinear", "VsgNorm", "VsgBigCEncLayer", "VsgBigCDecLayer", "VsgBigCWeights", "VsgVideoBatch", "VsgTripletOut", "VsgGrdConv",
             "VsgGrdEncoder", "VsgGrdWeights", "VsgGrdSeq", "VsgGrdBatch", "VsgGrdOut", "VsgGemmArgs", "VsgRelTable"]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = '#include "%s/include/vsg_b200.h"\n#include <stdio.h>\nint main(){ %s return 0; }\n' % (
        root, " ".join('printf("%%zu\\n", sizeof(%s));' % n for n in names))
    d = tempfile.mkdtemp()
    open(os.path.join(d, "s.cpp"), "w").write(src)
    subprocess.run(["g++", os.path.join(d, "s.cpp"), "-o", os.path.join(d, "s")], check=True)
    sizes = [int(x) for x in subprocess.run([os.path.join(d, "s")], capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(getattr(_cabi, n)) for n in names]


def test_cpulist_parsing_and_numa_binding_is_best_effort():
    """shard.bind_to_gpu_numa_node never raises (no GPU / no sysfs topology -> unbound) and the sysfs cpulist parser handles ranges."""
    from vidsgg_big_b200 import shard
    assert shard._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert shard._parse_cpulist("") == []
    import os
    before = os.sched_getaffinity(0)
    info = shard.bind_to_gpu_numa_node(0)
    assert info["device"] == 0 and isinstance(info["bound"], bool)
    if not info["bound"]:
        assert os.sched_getaffinity(0) == before


def test_fp16x3_weight_image_scale():
    """linalg.fp16_image_scale_log2: max |w| * 2^s in [2^13, 2^14) for every magnitude, exact powers of two included; degenerate weights -> 0."""
    import math
    from vidsgg_big_b200 import linalg
    for wmax in (1.0, 0.5, 0.022, 3.7e-5, 1e-12, 40.0, 8191.9, 8192.0, 16384.0, 65504.0, 2.0 ** -20, 1e20):
        s = linalg.fp16_image_scale_log2(wmax)
        assert 2.0 ** 13 <= wmax * 2.0 ** s < 2.0 ** 14, (wmax, s)
        assert math.isfinite(2.0 ** s) and math.isfinite(2.0 ** -s)
    assert linalg.fp16_image_scale_log2(0.0) == 0 and linalg.fp16_image_scale_log2(float("nan")) == 0 and linalg.fp16_image_scale_log2(float("inf")) == 0
    assert linalg.fp16_image_scale_log2(1e-45) == 100          # clamped: 2^100 is still a finite fp32
    assert set(linalg.FP32_CLASS) == {linalg.X3TF32, linalg.TF32_BF16X2, linalg.FP16X3} and linalg.MODES["fp16x3"] == 5
    assert linalg.attention_mode(linalg.FP16X3) == linalg.X3TF32 and linalg.attention_mode(linalg.BF16) == linalg.TF32
