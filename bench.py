#!/usr/bin/env python
"""bench.py -- videos/sec of the per-video relation hot path (classify [+ ground] + vIoU eval) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the reference

ONE JSON line with two workloads (README / DESIGN.md section "Measurement"):

* top level -- BASELINE.json configs[1]: a VidVRD-test-shaped batch (200 videos per GPU, weak scaling).  A "step" is one pass of pair
  geometry (n x n spans + trajectory vIoU), BIG-C classification (model_0v10, exp2 dims) incl. triplet construction and
  eval_visual_relation-style vIoU matching against synthetic GT.
* "vidor" -- BASELINE.json configs[3]: ONE VidOR-val-shaped set (835 videos, tracks up to the whole video, exp5 dims) through
  classification + grd_model_v5 grounding + evaluation (tools/eval_vidor.py:141-280), LPT-sharded over the N ranks (strong scaling,
  per-rank times reported), with its own roofline legs (grounding GEMMs, grounding attention), CPU baseline and grounded-output parity.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from vidsgg_big_b200 import synth  # noqa: E402

PRECISIONS = ["3xtf32", "tf32+bf16x2", "fp16x3", "tf32", "bf16", "fp32_simt"]
VIDOR_SEED0 = 700000          # seeds of the VidOR-val-shaped set
METRIC = "videos/sec (classify+ground+vIoU)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--videos", type=int, default=200, help="videos per GPU per step (weak scaling)")
    ap.add_argument("--workload", default="vidvrd", choices=["vidvrd", "vidor"],
                    help="workload of the top-level line (vidor: a per-GPU batch of VidOR-shaped videos incl. grounding)")
    ap.add_argument("--precision", default="tf32+bf16x2", choices=PRECISIONS)
    ap.add_argument("--cpu-sample", type=int, default=None,
                    help="videos in the bounded CPU sample (default: 200 for cpu_baseline, 100 per step for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="run the K timed steps strictly one after another (no overlap of step i-1's "
                    "evaluation host work with step i's classification kernels)")
    ap.add_argument("--force-pipeline", action="store_true", help="experiment: pipeline the steps of the VidOR-shaped top-level workload too")
    ap.add_argument("--vidor-videos", type=int, default=835, help="size of the VidOR-val-shaped set of the 'vidor' leg (0: skip the leg)")
    ap.add_argument("--vidor-passes", type=int, default=2, help="timed passes over the VidOR set")
    ap.add_argument("--chunk-rows", type=int, default=2_500_000, help="feature rows per resident chunk of the VidOR set")
    ap.add_argument("--no-graph", action="store_true", help="issue the BIG-C forward's ~115 launches from Python every step instead of "
                    "replaying the CUDA graph captured for the resident batch")
    ap.add_argument("--modes", default="fp16x3,bf16", help="comma list of extra precisions timed on the top-level workload next to --precision, with their decision-flip report "
                    "('' = none)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------------
def workload_cfg(kind):
    if kind == "vidvrd":
        cfg = synth.vidvrd_config()
        return cfg, dict(feat_total=cfg["dim_feat"] + cfg["dim_i3d"], min_len=20, max_len=150, shape=synth.vidvrd_video_shape,
                         topk=10, name="VidVRD-test-shaped batch, BIG-C exp2 dims (RoI 2048 + I3D 832), 5-50 tracklets, 20-150 frames")
    cfg = synth.vidor_config()
    # n ~ U{10..180}, video_len lognormal (120..5400), tracks of 15 frames up to the WHOLE video ("long tracklets")
    return cfg, dict(feat_total=cfg["dim_feat"] + cfg["dim_clsme"], min_len=15, max_len=None, shape=synth.vidor_video_shape,
                     topk=3, name="VidOR-val-shaped videos, BIG-C exp5 dims (RoI 1024 + classeme 300), 10-180 tracklets of 15..video_len frames, "
                                  "grd_model_v5 grounding (10 bins)")


def make_videos(kind, seeds, device, feat_seed, pinned=False, feats=None, i3d=None, with_gt=True):
    """Proposals (boxes etc. from the seeded host generator, one seed per video) whose features are row views of ONE buffer, plus GT
    graphs.  ``feats`` / ``i3d``: use these feature rows / per-video clip features instead of drawing new ones."""
    cfg, wl = workload_cfg(kind)
    props, graphs = [], []
    for seed in seeds:
        rng = np.random.default_rng(seed)
        vlen, n = wl["shape"](rng)
        P = synth.make_proposal(seed, n, vlen, wl["feat_total"], cfg["num_enti_cats"], min_len=wl["min_len"],
                                max_len=wl["max_len"], with_features=False)
        if with_gt:
            graphs.append(synth.make_gt_graph(seed, P, cfg["num_pred_cats"]))
        props.append(P)
    rows = sum(int(p.lengths.sum()) for p in props)
    g = torch.Generator(device=device).manual_seed(feat_seed)
    if feats is None:
        feats = torch.empty(rows, wl["feat_total"], device=device, dtype=torch.float32)
        fill_features(feats, g)
    assert feats.shape[0] == rows
    if pinned:
        host = torch.empty(rows, wl["feat_total"], dtype=torch.float32, pin_memory=True)
        host.copy_(feats)
        feats = host
    attach_features(props, feats)
    if kind == "vidor":      # I3D clip features for the grounding stage: f32[T, 1024] * 0.05, T = ceil(video_len / 8)
        for k, p in enumerate(props):
            T = (p.video_len + 7) // 8
            p.i3d = i3d[k] if i3d is not None else (torch.randn(T, 1024, generator=g, device=device, dtype=torch.float32) * 0.05)
            if pinned:
                p.i3d = p.i3d.cpu().pin_memory()
    return cfg, wl, props, graphs, feats


def fill_features(buf, gen, block=1 << 22):
    """randn * 0.5 in place, block-wise (no second buffer of the same size)."""
    flat = buf.view(-1)
    for a in range(0, flat.numel(), block * 64):
        b = min(flat.numel(), a + block * 64)
        flat[a:b].normal_(0.0, 0.5, generator=gen)


def fill_features_per_video(buf, rows_per_video, seeds, device):
    """Each video's feature rows from its OWN seed, so that a video's inputs do not depend on how the set is chunked / sharded
    (the metrics of the VidOR set are then the same at every GPU count)."""
    g = torch.Generator(device=device)
    r = 0
    for L, seed in zip(rows_per_video, seeds):
        g.manual_seed(int(seed) * 2 + 1)
        buf[r:r + L].normal_(0.0, 0.5, generator=g)
        r += L


def attach_features(props, feats):
    r = 0
    for p in props:
        L = int(p.lengths.sum())
        p.features = feats[r:r + L]
        p.dim_feat = feats.shape[1]
        r += L


def algorithmic_bytes_geometry(props):
    """sum over overlapping (a,b) of 32*ov + 16*(sumL_A + sumL_B) + nA*nB*(16+1+4)."""
    total = 0
    for p in props:
        d = p.traj_durations.cpu().numpy()
        s = np.maximum(d[:, None, 0], d[None, :, 0]); e = np.minimum(d[:, None, 1], d[None, :, 1])
        ov = np.clip(e - s + 1, 0, None)
        total += 32 * int(ov.sum()) + 16 * 2 * int(p.lengths.sum()) + d.shape[0] * d.shape[0] * 21
    return total


def algorithmic_bytes_rel_match(PR, GT):
    """sum over (p, g) of one video with equal triplets and overlapping durations of 64*ov + 32*(sumL_pred + sumL_gt)
    + 8 * #pairs (every (p, g) entry of the ov matrix is written)."""
    pr, gr = PR.rel.cpu().numpy(), GT.rel.cpu().numpy()
    po, go = PR.vid_off_host, GT.vid_off_host
    total = 32 * int((pr[:, 6] - pr[:, 5]).sum() + (gr[:, 6] - gr[:, 5]).sum())
    for v in range(GT.n_vid):
        p, g = pr[po[v]:po[v + 1]], gr[go[v]:go[v + 1]]
        if p.shape[0] == 0 or g.shape[0] == 0:
            continue
        total += 8 * p.shape[0] * g.shape[0]
        same = (p[:, None, 0] == g[None, :, 0]) & (p[:, None, 1] == g[None, :, 1]) & (p[:, None, 2] == g[None, :, 2])
        ov = np.clip(np.minimum(p[:, None, 6], g[None, :, 6]) - np.maximum(p[:, None, 5], g[None, :, 5]), 0, None)
        total += 64 * int((ov * same).sum())
    return total


class Clocks(object):
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), float(p["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------
# the step (our arm)
# ------------------------------------------------------------------------------------------------------
class Pipeline(object):
    def __init__(self, kind, precision, device, rank=0, graph=False):
        from vidsgg_big_b200 import bigc, grounding
        self.rank, self.kind, self.graph = rank, kind, graph
        if kind == "vidor":
            gcfg = synth.grounding_config()
            self.grd = grounding.DEBUG(gcfg, is_train=False, precision=precision)
            self.grd.load_state_dict(synth.make_grounding_state(21, gcfg))
            self.grd.to(device)
        self.cfg, self.wl = workload_cfg(kind)
        cls = bigc.BIG_C_vidor if kind == "vidor" else bigc.BIG_C_vidvrd
        self.model = cls(self.cfg, is_train=False, precision=precision)
        self.model.load_state_dict(synth.make_bigc_state(1, self.cfg))
        self.model.to(device)
        self.device = device
        self._packs, self._gts = {}, {}

    def pack_gt(self, graphs):
        """GT relations packed once per batch of videos (the reference loads its GT json once, too)."""
        from vidsgg_big_b200 import evalapi, geometry
        key = id(graphs)
        if key not in self._gts:
            gt_t = geometry.TrackTable.from_containers(graphs, device=self.device)
            self._gts[key] = evalapi.PackedRelations.from_gt_graphs(gt_t, graphs)
        return self._gts[key]

    def forget(self, props):
        self._packs.pop(id(props), None)

    def launch(self, props, timers=None):
        """First half of a step -- pair geometry + BIG-C classification + triplet construction -- enqueued on the current stream
        WITHOUT a host synchronisation (the per-video triplet counts stay on the device).  Returns a handle for ``finish``."""
        from vidsgg_big_b200 import geometry, linalg
        # packed index arrays of a batch are part of its HBM-resident form: built once per batch object
        if id(props) not in self._packs:
            self._packs[id(props)] = (geometry.TrackTable.from_containers(props), self.model.pack(props))
        tt, pk = self._packs[id(props)]
        linalg._Profile.stage = "bigc"
        if timers is not None:
            timers["geo0"].record()
        viou, spans, mask, seg, _ = geometry.traj_viou_batched(tt, tt)                 # pair geometry, all videos, one launch
        if timers is not None:
            timers["geo1"].record()
        packed = self.model.forward_packed(props, topk=self.wl["topk"], packed_videos=pk, sync=False, graph=self.graph)
        done = torch.cuda.Event()
        done.record()
        return dict(tt=tt, packed=packed, viou=viou, done=done, props=props)

    def finish(self, h, graphs, stream=None, gather=True, timers=None):
        """Second half: (VidOR: grounding of the classified triplets ->) relations -> vIoU matching kernels -> hit arrays D2H ->
        per-video records (-> all_gather -> metrics when ``gather``).  On ``stream`` (a side stream) it only waits for its own step's
        kernels, so the host part overlaps the NEXT step's classification kernels."""
        from vidsgg_big_b200 import evalapi, linalg, shard
        ctx = torch.cuda.stream(stream) if stream is not None else contextlib.nullcontext()
        with ctx:
            if stream is not None:
                stream.wait_event(h["done"])
            tt, packed, props = h["tt"], h["packed"], h["props"]
            if self.kind == "vidor":
                # grounding stage on the classification output (tools/eval_vidor.py:218-257), all videos batched; a video without
                # classified triplets skips it (:227-229) and keeps an empty prediction segment
                linalg._Profile.stage = "grounding"
                if timers is not None:
                    timers["grd0"].record()
                q, s3, sp, _, off = packed.compact()
                rows = [i for i in range(len(props)) if off[i + 1] > off[i]]
                datas = [(q[off[i]:off[i + 1]], sp[off[i]:off[i + 1]], props[i].video_len) for i in rows]
                pooled, probs, mask = self.grd.forward_packed([props[i].i3d for i in rows], datas, **synth.GROUNDING_INFERENCE)
                if timers is not None:
                    timers["grd1"].record()
                    timers["n_queries"] = int(off[-1])
                    timers["grd_flops"] = sum(shard.flops_grd(int(props[i].i3d.shape[0]), int(off[i + 1] - off[i])) for i in rows)
                PR = evalapi.PackedRelations.from_grounded(tt, packed, pooled, probs, mask, [p.video_len for p in props])
            else:
                PR = evalapi.PackedRelations.from_packed_triplets(tt, packed)          # score = mean of the 3 (eval_vidvrd.py:136)
            linalg._Profile.stage = "eval"
            GT = self.pack_gt(graphs)
            if timers is not None:
                timers["PR"], timers["GT"] = PR, GT
            # vIoU matching on the device, per-video records on the host (D2H of the hit arrays), then the only cross-rank exchange
            # of the whole path: an all_gather of 64 B / video (no-op at world size 1)
            rec = evalapi.evaluate_packed(PR, GT, want_records=True)
            if not gather:
                return rec, int(PR.n_rel), h["viou"]
            rec[:, 0] += self.rank * 1_000_000
            allrec = shard.gather_records(torch.from_numpy(rec).to(self.device)).cpu().numpy()
            m_ap, r_at, mprec = evalapi.metrics_from_records(allrec)
        return (float(m_ap), float(r_at[50]), float(r_at[100])), int(PR.n_rel), h["viou"]

    def step(self, props, graphs, timers=None, gather=True):
        """One pass over a batch that is resident in HBM.  Returns (metrics | records, n_relations, viou)."""
        return self.finish(self.launch(props, timers), graphs, gather=gather, timers=timers)


def gt_from_predictions(pipe, props, cfg, seeds, device):
    """One classification pass -> GT graphs built from its predictions (synth.make_gt_from_predictions), moved to the device."""
    with torch.no_grad():
        trips = pipe.model(props, topk=pipe.wl["topk"])
    graphs = []
    for seed, p, t in zip(seeds, props, trips):
        t3 = None if t is None else (t[0], t[1].mean(-1), t[2])
        graphs.append(synth.make_gt_from_predictions(seed, p, t3, num_pred_cats=cfg["num_pred_cats"]).to(device))
    return graphs, trips


class HostBatch(object):
    """One batch held the way a data loader would hand it over: pinned host buffers (features in one buffer, the small
    per-track fields concatenated per field) plus per-video metadata.  ``upload`` copies it into preallocated device
    buffers on a given stream and returns device-side proposals that are row views of those buffers."""

    def __init__(self, props, device):
        import copy
        self.props, self.device = props, device
        total = sum(int(p.lengths.sum()) for p in props)
        f0 = props[0].features
        self.h = {"feats": f0.as_strided((total, f0.shape[1]), f0.stride())}
        pin = lambda t: t.contiguous().pin_memory()
        self.h["boxes"] = pin(torch.cat([p.bboxes for p in props], 0))
        self.h["dura"] = pin(torch.cat([p.traj_durations for p in props], 0))
        self.h["cats"] = pin(torch.cat([p.cat_ids for p in props], 0))
        self.h["scores"] = pin(torch.cat([p.scores for p in props], 0))
        if hasattr(props[0], "i3d"):
            self.h["i3d"] = pin(torch.cat([p.i3d for p in props], 0))
        self.nbytes = sum(t.numel() * t.element_size() for t in self.h.values())
        self.slots = []
        for _ in range(2):                                        # double buffering
            d = {k: torch.empty(t.shape, dtype=t.dtype, device=device) for k, t in self.h.items()}
            views, r, n0, c0 = [], 0, 0, 0
            for p in props:
                q = copy.copy(p)
                L, n = int(p.lengths.sum()), p.num_proposals
                q.features, q.bboxes = d["feats"][r:r + L], d["boxes"][r:r + L]
                q.traj_durations, q.cat_ids, q.scores = d["dura"][n0:n0 + n], d["cats"][n0:n0 + n], d["scores"][n0:n0 + n]
                if "i3d" in d:
                    T = int(p.i3d.shape[0])
                    q.i3d = d["i3d"][c0:c0 + T]
                    c0 += T
                r += L; n0 += n
                views.append(q)
            self.slots.append((d, views))

    def upload(self, slot, stream):
        d, views = self.slots[slot]
        with torch.cuda.stream(stream):
            for k, t in self.h.items():
                d[k].copy_(t, non_blocking=True)
        return views


def run_e2e(pipe, hb, graphs, n_steps, barrier, device):
    """Same step through the public API with HOST buffers: pinned -> device every step, double-buffered on a copy stream so that the
    H2D transfer of step i+1 overlaps the kernels of step i; D2H of the hit arrays inside.  -> (seconds per step, n relations)."""
    copy_stream = torch.cuda.Stream(device=device)
    graph, pipe.graph = pipe.graph, False      # every step brings a new batch: nothing resident to replay a captured graph on
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def run(n):
        for ev in consumed:
            ev.record()
        hb.upload(0, copy_stream); copied[0].record(copy_stream)
        res = None
        for i in range(n):
            b = i % 2
            if i + 1 < n:
                copy_stream.wait_event(consumed[1 - b])
                hb.upload(1 - b, copy_stream); copied[1 - b].record(copy_stream)
            torch.cuda.current_stream().wait_event(copied[b])
            pipe.forget(hb.slots[b][1])                               # packed index arrays are rebuilt for every uploaded batch
            res = pipe.step(hb.slots[b][1], graphs, gather=False)
            consumed[b].record()
        return res
    run(2)
    barrier()
    w0 = time.perf_counter()
    _, n_rel, _ = run(n_steps)
    barrier()
    pipe.graph = graph
    return (time.perf_counter() - w0) / n_steps, n_rel


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference (torch-CPU / python loops), all host threads
# ------------------------------------------------------------------------------------------------------
CPU_STAGES = ("pair_geometry", "bigc_classify", "grounding", "to_eval_format", "viou_eval")


def cpu_pass(kind, props, graphs, st, cfg, wl, gst=None, stages=None, keep=None):
    """What the reference does per video on the host: stretched BIG-C forward, per-pair vIoU loop, (grounding incl. its python pooling
    loops,) dict conversion, eval.  ``stages``: dict accumulating seconds per stage; ``keep``: list receiving per-video outputs."""
    from oracle import bigc as ob, convert as oc, evalapi as oe, geometry as og, grounding as ogr
    en, pn = oc.default_names("e", 256), oc.default_names("p", 256)
    gts, prs = {}, {}
    stages = stages if stages is not None else {}
    tick = time.perf_counter

    def add(name, t0):
        stages[name] = stages.get(name, 0.0) + tick() - t0
    with torch.no_grad():
        for p, g in zip(props, graphs):
            t0 = tick()
            og.traj_viou_matrix(p.bboxes_list, p.traj_durations, p.bboxes_list, p.traj_durations)
            add("pair_geometry", t0); t0 = tick()
            r = ob.forward(st, cfg, [p], wl["topk"])[0]
            add("bigc_classify", t0)
            t3 = None if r is None else (r[0], r[1].mean(-1), r[2])
            grd = None
            if kind == "vidor" and r is not None and r[0].shape[0] > 0:
                t0 = tick()
                pooled, probs, mask = ogr.forward(gst, synth.grounding_config(), [p.i3d], [(r[0], r[2], p.video_len)], **synth.GROUNDING_INFERENCE)
                t3 = ogr.expand_after_grounding(r[0], r[1], pooled, probs, mask, p.video_len)
                grd = (pooled, probs, mask)
                add("grounding", t0)
            t0 = tick()
            prs.update(oc.to_eval_format_pr(p, t3, en, pn))
            gts.update(oc.to_eval_format_gt(g, en, pn))
            add("to_eval_format", t0)
            if keep is not None:
                keep.append((r, grd))
    t0 = tick()
    out = oe.evaluate(gts, prs)
    add("viou_eval", t0)
    return out


def cpu_prepare(kind, seeds, feat_seed, feats_from=None):
    """Inputs of the CPU arm: the videos of ``seeds``, oracle weights, and GT graphs derived from the oracle's own predictions
    (untimed).  ``feats_from``: (features, i3d list) of the GPU arm's videos, copied to the host so that both arms see bit-identical inputs
    (the CUDA and CPU generators differ)."""
    torch.set_num_threads(os.cpu_count() or 1)
    feats, i3d = feats_from if feats_from is not None else (None, None)
    cfg, wl, props, _, _ = make_videos(kind, seeds, "cpu", feat_seed, feats=feats, i3d=i3d, with_gt=False)
    st = synth.make_bigc_state(1, cfg)
    gst = synth.make_grounding_state(21, synth.grounding_config()) if kind == "vidor" else None
    from oracle import bigc as ob
    trips, graphs = [], []
    with torch.no_grad():
        for seed, p in zip(seeds, props):
            r = ob.forward(st, cfg, [p], wl["topk"])[0]
            trips.append(r)
            graphs.append(synth.make_gt_from_predictions(seed, p, None if r is None else (r[0], r[1].mean(-1), r[2]),
                                                         num_pred_cats=cfg["num_pred_cats"]))
    if kind == "vidvrd":
        cpu_pass(kind, props[:1], graphs[:1], st, cfg, wl, gst)            # warm-up
    return dict(kind=kind, props=props, graphs=graphs, st=st, cfg=cfg, wl=wl, gst=gst, trips=trips)


def cpu_timed_pass(prep, keep=None):
    stages = {}
    t0 = time.perf_counter()
    metrics = cpu_pass(prep["kind"], prep["props"], prep["graphs"], prep["st"], prep["cfg"], prep["wl"], prep["gst"], stages, keep)
    return time.perf_counter() - t0, metrics, stages


def cpu_info():
    model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return {"cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads(), "cpu_model": model, "torch": torch.__version__}


PORT_NOTE = ("oracle port of the reference (stretched tensors, per-pair python loops, dict conversion, python eval); bit-identical to the "
             "unmodified reference on the committed goldens; the unmodified reference, timed in the build container on 6 of these videos, "
             "was 1.16x slower than the port")


def compare_triplets(gpu_trips, cpu_trips):
    """Per-video identity of the classification output (quintuples + spans, as sets) between the CUDA path and the oracle."""
    same = total = 0
    for a, b in zip(gpu_trips, cpu_trips):
        total += 1
        if (a is None) != (b is None):
            continue
        if a is None:
            same += 1
            continue
        sa = set(map(tuple, torch.cat([a[0].cpu(), a[2].cpu()], 1).tolist()))
        sb = set(map(tuple, torch.cat([b[0], b[2]], 1).tolist()))
        same += int(sa == sb)
    return same, total


def decision_flips(ref_pipe, alt_pipe, props):
    """Discrete decisions of a reduced-precision mode against the parity mode on the same batch (SURVEY 8d parity gates): how many
    queries change their (subject, object) arg-max, how many of the others change their top-k predicate set, and how close to a tie the
    flipped decisions were in the parity mode (relative margin between the best and the second-best candidate)."""
    edges = [1e-4, 1e-3, 1e-2, 1e-1]

    def hist(m):
        m = m.float().cpu().numpy()
        return {"<1e-4": int((m < edges[0]).sum()), "<1e-3": int(((m >= edges[0]) & (m < edges[1])).sum()),
                "<1e-2": int(((m >= edges[1]) & (m < edges[2])).sum()), "<1e-1": int(((m >= edges[2]) & (m < edges[3])).sum()),
                ">=1e-1": int((m >= edges[3]).sum())}
    with torch.no_grad():
        pk = ref_pipe.model.pack(props)
        la, soa, exa = ref_pipe.model._encode2decode(pk, want_att=True)
        lb, sob, _ = alt_pipe.model._encode2decode(pk)
    k = ref_pipe.wl["topk"]
    P = ref_pipe.cfg["num_pred_cats"]
    so_flip = (soa != sob).any(1)
    top2 = torch.topk(exa["att"], 2, dim=-1)[0] if exa["att"].shape[-1] > 1 else None           # [VQ, 2 roles, 2]
    att_margin = ((top2[..., 0] - top2[..., 1]) / top2[..., 0].clamp_min(1e-30)).min(1)[0] if top2 is not None else torch.ones(soa.shape[0], device=soa.device)
    pa, pb = torch.softmax(la[:, :P], -1), torch.softmax(lb[:, :P], -1)
    ta, tb = torch.topk(pa, k + 1, dim=-1), torch.topk(pb, k, dim=-1)
    same_set = (torch.sort(ta[1][:, :k], -1)[0] == torch.sort(tb[1], -1)[0]).all(1)
    topk_flip = (~same_set) & (~so_flip)
    topk_margin = (ta[0][:, k - 1] - ta[0][:, k]) / ta[0][:, k - 1].clamp_min(1e-30)
    return {"queries": int(soa.shape[0]), "so_argmax_flips": int(so_flip.sum()), "topk_set_flips_among_the_rest": int(topk_flip.sum()),
            "parity_mode_margin_of_so_flips": hist(att_margin[so_flip]), "parity_mode_margin_of_topk_flips": hist(topk_margin[topk_flip]),
            "parity_mode_margin_all_queries_so": hist(att_margin), "max_rel_logit_diff_unflipped": float(
                ((la[:, :P] - lb[:, :P]).abs().max(1)[0][~so_flip].max() / la[:, :P].abs().max()).item()) if bool((~so_flip).any()) else None}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg, wl = workload_cfg(args.workload)
    if args.cpu_sample is None:
        args.cpu_sample = 100 if args.workload == "vidvrd" else 2       # ~8 s / ~4 s of CPU work per step: K=20, W=3 ends within ~3.5 minutes
    seeds = [1000 + i for i in range(args.cpu_sample)]
    prep = cpu_prepare(args.workload, seeds, 1000)                      # inputs, weights and GT once; every step is one full CPU pass over them
    times, stage_sum = [], {}
    for i in range(args.warmup + args.steps):
        dt, _, stages = cpu_timed_pass(prep)
        if i >= args.warmup:
            times.append(dt)
            for k, v in stages.items():
                stage_sum[k] = stage_sum.get(k, 0.0) + v
    ms = 1e3 * float(np.mean(times))
    value = args.cpu_sample / (ms / 1e3)
    cores = torch.get_num_threads()
    sample = "%d videos of the workload per step (seeds 1000..); %s" % (args.cpu_sample, PORT_NOTE)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "videos/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.workload, wl),
        "run": {"videos_per_step": args.cpu_sample, "host": cpu_info(),
                "stage_seconds_per_video": {k: v / (len(times) * args.cpu_sample) for k, v in stage_sum.items()}},
        "cpu_baseline": {"value": value, "unit": "videos/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "videos/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def bench_config(kind, wl):
    """The ``config`` object -- identical in both arms (the driver compares them)."""
    return {"workload": wl["name"],
            "stages": ["pair_geometry", "bigc_classify", "triplets"] + (["grounding"] if kind == "vidor" else []) + ["viou_eval"]}


# ------------------------------------------------------------------------------------------------------
# roofline legs of one instrumented step
# ------------------------------------------------------------------------------------------------------
def instrumented_step(pipe, props, graphs, precision, ms_step):
    """One extra step with per-launch CUDA events (on the launch stream) around every GEMM, the geometry kernel, the relation matcher
    and -- VidOR -- the grounding stage / attention.  -> dict of roofline objects."""
    from vidsgg_big_b200 import linalg
    hbm_peak, tc_peak, peak_src = peaks()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    timers = {"geo0": ev(), "geo1": ev(), "grd0": ev(), "grd1": ev()}
    # The launches are issued from Python here (no graph replay) so that each can be bracketed by events.  Many of them run for less
    # time than Python needs to issue the next one; an event pair would then also measure the GPU waiting for the host.  A spin kernel in
    # front of each stage gives the host a head start, so the brackets hold kernel time only.
    graph, pipe.graph = pipe.graph, False
    backend, pipe.model.backend = pipe.model.backend, "py"          # op-by-op launches (bit-identical to the one-call C entries)
    if pipe.kind == "vidor":
        gbackend, pipe.grd.backend = pipe.grd.backend, "py"
    spin = lambda ms: torch.cuda._sleep(int(ms * 1e-3 * 1.9e9))
    linalg._Profile.begin()
    spin(60)
    h = pipe.launch(props, timers)
    if pipe.kind == "vidor":
        h["packed"].counts          # the grounding stage needs the query counts on the host (as in the timed path)
        spin(30)
    pipe.finish(h, graphs, gather=False, timers=timers)
    linalg._Profile.end()
    pipe.graph, pipe.model.backend = graph, backend
    if pipe.kind == "vidor":
        pipe.grd.backend = gbackend
    P = linalg._Profile
    slots = {"tf32+bf16x2": 4.0, "fp16x3": 3.0, "3xtf32": 6.0, "tf32": 2.0, "bf16": 1.0}.get(precision)      # bf16-equivalent tensor slots per useful MAC
    out = {}

    def tensor_leg(kernel, n, flops, ms, extra=None):
        tf = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        leg = {"kernel": kernel, "bound": "tensor", "achieved": tf, "peak": tc_peak, "unit": "TFLOP/s", "frac": tf / tc_peak,
               "traffic": None, "launches_per_step": n, "ms": ms, "share_of_step": ms / ms_step if ms_step else None,
               "issued_bf16_equiv": None if slots is None else tf * slots, "frac_issued": None if slots is None else tf * slots / tc_peak,
               "peak_source": peak_src + ", bf16 dense sustained"}
        leg.update(extra or {})
        return leg

    def hbm_leg(kernel, nbytes, ms, extra=None):
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        leg = {"kernel": kernel, "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": None,
               "ms": ms, "algorithmic_bytes": int(nbytes), "share_of_step": ms / ms_step if ms_step else None, "peak_source": peak_src + ", HBM copy"}
        leg.update(extra or {})
        return leg

    n, fl, ms = P.summary("gemm", "bigc")
    out["bigc_gemm"] = tensor_leg("gemm_tc_kernel (tcgen05 %s), BIG-C launches" % precision, n, fl, ms,
                                  {"note": "achieved = useful 2MNK flops; the fp32-class modes issue several tensor passes per useful product "
                                           "(3xtf32: 3 tf32 = 6 bf16-equivalent slots, tf32+bf16x2: 1 tf32 + 2 bf16 = 4 slots, fp16x3: 3 fp16 = 3 slots, bf16: 1)"})
    geo_ms = timers["geo0"].elapsed_time(timers["geo1"])
    out["k1_geometry"] = hbm_leg("traj_viou_warp_kernel (+ track volumes, spans)", algorithmic_bytes_geometry(props), geo_ms,
                                 {"note": "SURVEY 8d K1 bytes; a video's tracks fit in L2, so DRAM only sees the compulsory bytes"})
    n, _, ms = P.summary("rel_match")
    out["k3_rel_match"] = hbm_leg("rel_volume + rank + rel_ov + greedy_match kernels (vsg_rel_viou_match)",
                                  algorithmic_bytes_rel_match(timers["PR"], timers["GT"]), ms,
                                  {"note": "SURVEY 8d K3 bytes (fp64 accumulate); %d relations vs %d GT relations" % (timers["PR"].n_rel, timers["GT"].n_rel)})
    if pipe.kind == "vidor":
        n, fl, ms = P.summary("gemm", "grounding")
        grd_ms = timers["grd0"].elapsed_time(timers["grd1"])
        out["k6_grounding_gemm"] = tensor_leg("gemm_tc_kernel (tcgen05 %s; CONV variant = depthwise conv fused), grounding launches" % precision, n, fl, ms,
                                              {"grounding_stage_ms": grd_ms, "grounding_stage_useful_flops": timers["grd_flops"],
                                               "grounding_stage_tflops": timers["grd_flops"] / (grd_ms * 1e-3) / 1e12 if grd_ms > 0 else 0.0,
                                               "queries": timers["n_queries"],
                                               "hbm_model": (lambda b: {"algorithmic_bytes": int(b), "achieved": b / (ms * 1e-3) / 1e9 if ms > 0 else 0.0,
                                                                        "peak": hbm_peak, "unit": "GB/s", "frac": b / (ms * 1e-3) / 1e9 / hbm_peak if ms > 0 else 0.0,
                                                                        "note": "the bound that applies: compulsory bytes of these launches (A + C + residual, "
                                                                                "fp32) / their time against the measured HBM copy rate"})(P.summary_bytes("gemm", "grounding")),
                                               "note": "N = K = 128 GEMMs: 64 flop/B unfused, i.e. HBM-bound (SURVEY 8d K6), so `frac` against the TENSOR peak is "
                                                       "small by construction -- see hbm_model; stage flops = flops_grd formula"})
        n, fl, ms = P.summary("mha", "grounding")
        out["k6_grounding_attention"] = tensor_leg("grounding mh_attn (8 heads x 16)", n, fl, ms)
    return out


# ------------------------------------------------------------------------------------------------------
# the VidOR-val-shaped set (BASELINE configs[3]): strong scaling over the ranks
# ------------------------------------------------------------------------------------------------------
def vidor_set_plan(n_set, world, chunk_rows, parity_videos=0):
    """Shapes of the whole set (cheap: only the first draws of each video's generator), flop costs, LPT shards, and per rank the
    chunks (lists of set indices, <= chunk_rows feature rows each).  ``parity_videos`` > 0: rank 0's first chunk holds that many
    moderate-size videos for the CPU oracle (n * Tmax bounded so that the oracle's stretched tensors stay ~1 GB)."""
    from vidsgg_big_b200 import shard
    cfg, wl = workload_cfg("vidor")
    info = []
    for i in range(n_set):
        seed = VIDOR_SEED0 + i
        vlen, n = wl["shape"](np.random.default_rng(seed))
        lens = synth.proposal_lengths(seed, n, vlen, wl["min_len"], wl["max_len"])
        info.append(dict(seed=seed, video_len=vlen, n=n, rows=int(lens.sum()), tmax=int(lens.max()),
                         cost=shard.video_cost_flops(lens, vlen, cfg["dim_feat"])))
    shards = shard.assign_lpt([v["cost"] for v in info], world)
    plan = []
    for r in range(world):
        ids = list(shards[r])
        chunks = []
        if r == 0 and parity_videos > 0:
            ok = sorted((i for i in ids if info[i]["n"] * info[i]["tmax"] <= 250_000), key=lambda i: info[i]["cost"])
            pick = [ok[int(round(k * (len(ok) - 1) / max(parity_videos - 1, 1)))] for k in range(min(parity_videos, len(ok)))]
            pick = sorted(set(pick))
            if pick:
                chunks.append(pick)
                ids = [i for i in ids if i not in set(pick)]
        cur, rows = [], 0
        for i in ids:
            if cur and rows + info[i]["rows"] > chunk_rows:
                chunks.append(cur); cur, rows = [], 0
            cur.append(i); rows += info[i]["rows"]
        if cur:
            chunks.append(cur)
        plan.append(chunks)
    return info, shards, plan


def vidor_leg(args, rank, world, device, dist, barrier):
    from vidsgg_big_b200 import _cabi
    t_leg = time.perf_counter()
    n_set = args.vidor_videos
    do_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    info, shards, plan = vidor_set_plan(n_set, world, args.chunk_rows, parity_videos=6 if do_cpu else 0)
    my_chunks = plan[rank]
    pipe = Pipeline("vidor", args.precision, device, rank)
    cfg, wl = pipe.cfg, pipe.wl
    my_rows = sum(info[i]["rows"] for c in my_chunks for i in c)
    resident = my_rows * wl["feat_total"] * 4 <= 64e9            # all chunk feature buffers stay in HBM; else one reusable buffer, refilled (untimed)
    max_chunk_rows = max(sum(info[i]["rows"] for i in c) for c in my_chunks)
    shared_buf = None if resident else torch.empty(max_chunk_rows, wl["feat_total"], dtype=torch.float32, device=device)

    # ---- build the chunks: proposals / I3D resident, features resident or refillable, GT from the model's own predictions (untimed) ----
    chunks = []
    for ci, ids in enumerate(my_chunks):
        seeds = [info[i]["seed"] for i in ids]
        rows = sum(info[i]["rows"] for i in ids)
        vrows = [info[i]["rows"] for i in ids]
        buf = shared_buf[:rows] if shared_buf is not None else torch.empty(rows, wl["feat_total"], dtype=torch.float32, device=device)
        fill_features_per_video(buf, vrows, seeds, device)
        g = torch.Generator(device=device)
        i3d = []
        for i in ids:                                           # clip features from the video's own seed, too
            g.manual_seed(int(info[i]["seed"]) * 2)
            i3d.append(torch.randn((info[i]["video_len"] + 7) // 8, 1024, generator=g, device=device, dtype=torch.float32) * 0.05)
        _, _, props, _, _ = make_videos("vidor", seeds, device, 0, feats=buf, i3d=i3d, with_gt=False)
        for p in props:
            f = p.features
            p.to(device)
            p.features = f
        ch = dict(ids=ids, seeds=seeds, props=props, vrows=vrows, buf=buf, rows=rows)
        chunks.append(ch)

    def refill(ch):
        if shared_buf is not None:
            fill_features_per_video(ch["buf"], ch["vrows"], ch["seeds"], device)

    cpu, parity = None, None
    if do_cpu:
        # CPU oracle on the parity chunk (chunk 0): same inputs copied to the host, before any GPU timing
        ch = chunks[0]
        refill(ch)
        i3d = [p.i3d.cpu() for p in ch["props"]]
        prep = cpu_prepare("vidor", ch["seeds"], 0, feats_from=(ch["buf"].cpu(), i3d))
        keep = []
        dt, cpu_metrics, stages = cpu_timed_pass(prep, keep)
        cpu = {"value": len(ch["ids"]) / dt, "unit": "videos/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "%d videos of the set spread over its cost range below n * Tmax <= 250k (cheaper than the set's mean video), same "
                         "inputs copied to the host, %.1f s of CPU work; %s" % (len(ch["ids"]), dt, PORT_NOTE),
               "stage_seconds_per_video": {k: v / len(ch["ids"]) for k, v in stages.items()}, "host": cpu_info()}
        parity = vidor_parity(pipe, ch, prep, keep, cpu_metrics, device)

    for ch in chunks:                                           # GT from the model's own predictions (one untimed classification pass)
        refill(ch)
        ch["graphs"], _ = gt_from_predictions(pipe, ch["props"], cfg, ch["seeds"], device)
    torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def one_pass(pipe=pipe):
        """All chunks of this rank; -> (device ms inside the chunk brackets, records, relations)."""
        brackets, recs, n_rel = [], [], 0
        for ch in chunks:
            refill(ch)                                          # untimed: regenerates a chunk that cannot stay resident next to the others
            e0, e1 = ev(), ev()
            e0.record()
            rec, nr, _ = pipe.step(ch["props"], ch["graphs"], gather=False)
            e1.record()
            rec[:, 0] = np.asarray(ch["ids"], dtype=np.float64)[rec[:, 0].astype(np.int64)]
            brackets.append((e0, e1)); recs.append(rec); n_rel += nr
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in brackets), np.concatenate(recs, 0), n_rel

    from vidsgg_big_b200 import evalapi, shard
    n_warm = 1 if min(len(c) for c in plan) >= 3 else 3          # >= 3 warm-up steps on every rank
    for _ in range(n_warm):
        one_pass()
    launches0 = int(_cabi.lib().vsg_launch_count())
    barrier()
    pass_ms, compute_ms, metrics, n_rel = [], [], None, 0
    for _ in range(max(1, args.vidor_passes)):
        ms, rec, n_rel = one_pass()
        compute_ms.append(ms)
        g0, g1 = ev(), ev()
        g0.record()
        allrec = shard.gather_records(torch.from_numpy(rec).to(device)).cpu().numpy()      # the one exchange: 64 B / video
        g1.record()
        torch.cuda.synchronize()
        pass_ms.append(ms + g0.elapsed_time(g1))
        metrics = evalapi.metrics_from_records(allrec)
    barrier()
    n_launches = int(_cabi.lib().vsg_launch_count()) - launches0
    mine = torch.tensor([float(np.mean(pass_ms))], device=device)
    per_rank = [mine.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
    per_rank = [float(t.item()) for t in per_rank]
    ms_pass = max(per_rank)
    mine_c = torch.tensor([float(np.mean(compute_ms))], device=device)         # chunk brackets only: shows the load balance (the gather equalises)
    per_rank_compute = [mine_c.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank_compute, mine_c)
    per_rank_compute = [float(t.item()) for t in per_rank_compute]

    # ---- the other fp32-class mode(s) of --modes on the same set and GT: one warm-up pass, one timed pass ----
    alt_modes = {}
    for prec in [m for m in args.modes.split(",") if m in ("fp16x3", "3xtf32", "tf32+bf16x2") and m != args.precision]:
        try:                                    # an extra mode must never take the parity mode's line down with it
            alt = Pipeline("vidor", prec, device, rank)
            one_pass(alt)
        except Exception as e:                  # noqa: BLE001  (deterministic failures hit every rank at the same point)
            alt_modes[prec] = {"error": "%s: %s" % (type(e).__name__, e)}
            continue
        barrier()
        ms_a, rec_a, _ = one_pass(alt)
        allrec_a = shard.gather_records(torch.from_numpy(rec_a).to(device)).cpu().numpy()
        mine_a = torch.tensor([ms_a], device=device)
        per_a = [mine_a.clone() for _ in range(world)]
        if world > 1:
            dist.all_gather(per_a, mine_a)
        per_a = [float(t.item()) for t in per_a]
        m_a = evalapi.metrics_from_records(allrec_a)
        alt_modes[prec] = {"value": n_set / (max(per_a) / 1e3), "unit": "videos/s", "ms_per_pass": max(per_a), "passes": 1, "per_rank_ms": per_a,
                           "result": {"mAP": float(m_a[0]), "R@50": float(m_a[1][50]), "R@100": float(m_a[1][100])},
                           "note": "same set, same sharding, GT from the parity mode's predictions; the record all_gather is outside this bracket"}
        del alt
        torch.cuda.empty_cache()

    # ---- roofline legs: one instrumented step on this rank's largest chunk ----
    big = max(chunks, key=lambda c: c["rows"])
    refill(big)
    roof = instrumented_step(pipe, big["props"], big["graphs"], args.precision, None)
    for leg in roof.values():
        leg["share_of_step"] = None
    roof_ctx = {"chunk_videos": len(big["ids"]), "chunk_rows": big["rows"]}

    # ---- e2e on a bounded sample: this rank's first chunk from pinned host memory (H2D + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        ch = min(chunks, key=lambda c: abs(c["rows"] - 1_200_000))       # ~6 GB of pinned features
        refill(ch)
        host = torch.empty(ch["rows"], wl["feat_total"], dtype=torch.float32, pin_memory=True)
        host.copy_(ch["buf"])
        import copy
        hprops = []
        for p in ch["props"]:
            q = copy.copy(p)
            q.bboxes, q.traj_durations, q.cat_ids, q.scores = p.bboxes.cpu(), p.traj_durations.cpu(), p.cat_ids.cpu(), p.scores.cpu()
            q.i3d = p.i3d.cpu().pin_memory()
            hprops.append(q)
        attach_features(hprops, host)
        hb = HostBatch(hprops, device)
        n_e2e = 3
        dt, n_rel_e = run_e2e(pipe, hb, ch["graphs"], n_e2e, barrier, device)
        vids = torch.tensor([float(len(ch["ids"])), dt], device=device)
        allv = [vids.clone() for _ in range(world)]
        if world > 1:
            dist.all_gather(allv, vids)
        tot = sum(float(t[0].item()) for t in allv)
        worst = max(float(t[1].item()) for t in allv)
        e2e = {"value": tot / worst, "unit": "videos/s", "h2d_bytes_per_step": int(hb.nbytes),
               "d2h_bytes_per_step": int(n_rel_e * (8 + 4 + 24) + len(ch["ids"]) * 16), "steps": n_e2e,
               "sample": "one chunk per rank (%d videos, %.1f GB on rank 0) from pinned host memory, double-buffered H2D; the whole set "
                         "(%.0f GB of fp32 features) is not held in host memory" % (len(ch["ids"]), hb.nbytes / 1e9, sum(v["rows"] for v in info) * wl["feat_total"] * 4 / 1e9)}
        del hb, host
    if rank != 0:
        return None
    costs = [sum(info[i]["cost"] for i in s) for s in shards]
    return {
        "metric": METRIC, "value": n_set / (ms_pass / 1e3), "unit": "videos/s", "n_gpus": world, "scaling": "strong",
        "ms_per_pass": ms_pass, "passes": len(pass_ms), "warmup_passes": n_warm, "per_rank_ms": per_rank,
        "per_rank_compute_ms": per_rank_compute,
        "config": dict(bench_config("vidor", wl), videos=n_set, sharding="shard.assign_lpt on useful-flop costs (BIG-C over ragged rows + grounding)",
                       precision=args.precision),
        "set": {"videos": n_set, "feature_rows": int(sum(v["rows"] for v in info)), "feature_gb": sum(v["rows"] for v in info) * wl["feat_total"] * 4 / 1e9,
                "max_track_frames": int(max(v["tmax"] for v in info)), "max_video_len": int(max(v["video_len"] for v in info)),
                "chunks_per_rank": [len(c) for c in plan], "rows_per_rank": [int(sum(info[i]["rows"] for c in chunks_r for i in c)) for chunks_r in plan],
                "lpt_cost_imbalance": max(costs) / (sum(costs) / len(costs)),
                "residency": ("all chunks resident in HBM" if resident else
                              "features of one chunk resident at a time (the set's fp32 features exceed HBM); each chunk is regenerated in HBM "
                              "OUTSIDE its timed bracket, timed region = sum of the chunk brackets + the record all_gather")},
        "result": {"mAP": float(metrics[0]), "R@50": float(metrics[1][50]), "R@100": float(metrics[1][100]), "relations_after_grounding_rank0": n_rel},
        "roofline": dict(roof, measured_on=roof_ctx), "cpu_baseline": cpu, "parity_vs_cpu_oracle": parity, "e2e": e2e,
        "modes": alt_modes or None,
        "gpu_launches": n_launches, "leg_wall_s": time.perf_counter() - t_leg,
    }


def vidor_parity(pipe, ch, prep, keep, cpu_metrics, device):
    """Grounded-output parity on the CPU-sample videos: (a) classification triplets, (b) the grounding stage fed with the ORACLE's
    triplets (so a near-tie flip upstream cannot hide or fake a grounding difference): bin probabilities, keep masks and the frame
    spans the driver rounds to, (c) final metrics of both pipelines on the same GT."""
    from vidsgg_big_b200 import evalapi, geometry, grounding
    props = ch["props"]
    with torch.no_grad():
        trips = pipe.model(props, topk=pipe.wl["topk"])
    same, total = compare_triplets(trips, [k[0] for k in keep])
    n_q = n_bins = mask_diff = span_same = span_total = 0
    max_dp = 0.0
    for p, (r, grd) in zip(props, keep):
        if grd is None:
            continue
        q, sp = r[0].to(device), r[2].to(device)
        with torch.no_grad():
            pooled, probs, mask = pipe.grd([p.i3d], [(q, sp, p.video_len)], with_gt_data=False, **synth.GROUNDING_INFERENCE)
        rp, rprob, rmask = grd
        n_q += q.shape[0]; n_bins += rmask.numel()
        max_dp = max(max_dp, float((probs.cpu() - rprob).abs().max()))
        mask_diff += int((mask.cpu() != rmask).sum())
        both = mask.cpu() & rmask
        a = torch.round(pooled.cpu() * p.video_len)[both]
        b = torch.round(rp * p.video_len)[both]
        span_same += int((a == b).all(-1).sum()); span_total += int(both.sum())
    # (c) our whole pipeline on these videos against the oracle's GT
    import copy
    graphs = [copy.copy(g).to(device) for g in prep["graphs"]]
    rec, _, _ = pipe.step(props, graphs, gather=False)
    pipe._gts.pop(id(graphs), None)
    ours = evalapi.metrics_from_records(rec)
    return {"videos_compared": total, "videos_with_identical_triplets": same,
            "grounding_on_oracle_triplets": {"queries": n_q, "bins": n_bins, "max_abs_bin_prob_diff": max_dp, "mask_mismatches": mask_diff,
                                             "kept_bins_with_identical_frame_span": span_same, "kept_bins_compared": span_total},
            "metrics_ours": {"mAP": float(ours[0]), "R@50": float(ours[1][50]), "R@100": float(ours[1][100])},
            "metrics_cpu_oracle": {"mAP": float(cpu_metrics[0]), "R@50": float(cpu_metrics[1][50]), "R@100": float(cpu_metrics[1][100])}}


# ------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch.distributed as dist
    from vidsgg_big_b200 import linalg, _cabi
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product path has no CPU fallback)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    from vidsgg_big_b200 import shard as _shard
    numa = _shard.bind_to_gpu_numa_node(local)      # pinned buffers are then first-touched next to the GPU (no-op where sysfs shows no topology)
    if world > 1:
        # NCCL prints its version banner (and any NCCL_DEBUG output) on stdout while the communicator is created: point fd 1 at stderr for
        # that moment so that stdout carries the ONE JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            warm = torch.zeros(1, device=device)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _cabi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pipe = Pipeline(args.workload, args.precision, device, rank, graph=not args.no_graph)
    seeds = [1000 + 100000 * rank + i for i in range(args.videos)]
    cfg, wl, props, graphs, feats = make_videos(args.workload, seeds, device, seeds[0])

    # CPU baseline (rank 0, N=1 only) before any GPU timing, on the SAME videos: their features are copied to the host, so the
    # oracle's triplets / metrics can be compared with the CUDA path's at the benchmark's full size
    cpu, cpu_metrics, cpu_trips = None, None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if args.cpu_sample is None:
            args.cpu_sample = 200 if args.workload == "vidvrd" else 4
        args.cpu_sample = min(args.cpu_sample, args.videos)
        rows = sum(int(p.lengths.sum()) for p in props[:args.cpu_sample])
        i3d = [p.i3d.cpu() for p in props[:args.cpu_sample]] if args.workload == "vidor" else None
        prep = cpu_prepare(args.workload, seeds[:args.cpu_sample], seeds[0], feats_from=(feats[:rows].cpu(), i3d))
        dt, cpu_metrics, stages = cpu_timed_pass(prep)
        cpu_trips = prep["trips"]
        cpu = {"value": args.cpu_sample / dt, "unit": "videos/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "the first %d videos of the same batch (same inputs, copied to the host), %.1f s of CPU work; %s" % (args.cpu_sample, dt, PORT_NOTE),
               "stage_seconds_per_video": {k: v / args.cpu_sample for k, v in stages.items()}, "host": cpu_info()}
        del prep
    for p in props:
        f = p.features
        p.to(device)
        p.features = f
    in_bytes = feats.numel() * 4

    # GT derived from the model's own predictions (first pass), so that the evaluation stage has real matches to find
    graphs, gpu_trips = gt_from_predictions(pipe, props, cfg, seeds, device)
    parity = None
    if cpu_trips is not None:
        same, total = compare_triplets(gpu_trips[:len(cpu_trips)], cpu_trips)
        parity = {"videos_with_identical_triplets": same, "videos_compared": total,
                  "cpu_oracle_metrics": {"mAP": float(cpu_metrics[0]), "R@50": float(cpu_metrics[1][50]), "R@100": float(cpu_metrics[1][100])}
                  if args.cpu_sample == args.videos and args.workload == "vidvrd" else None}
    del gpu_trips
    pipelined = (not args.no_pipeline) and (args.workload == "vidvrd" or args.force_pipeline)     # VidOR: the grounding kernels of step i-1 on a side stream compete
    side = torch.cuda.Stream(device=device, priority=-1) if pipelined else None   # with step i's GEMMs for the SMs (measured 7 % slower)

    def run_steps(n, pl=pipe):
        """n steps; pipelined: step i's classification kernels are enqueued before step i-1's second half (matching kernels + D2H +
        host records) runs on the side stream, so its host part overlaps GPU work -- every step still does all of its work."""
        res = None
        if not pipelined:
            for _ in range(n):
                res = pl.step(props, graphs)
            return res
        prev = None
        for _ in range(n):
            cur = pl.launch(props)
            if prev is not None:
                res = pl.finish(prev, graphs, stream=side)
            prev = cur
        res = pl.finish(prev, graphs, stream=side)
        torch.cuda.current_stream().wait_stream(side)
        return res

    def timed(pl, n_steps, n_warm):
        if n_warm:
            run_steps(n_warm, pl)                      # same code path as the timed region (also warms the side stream's allocator pool)
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        res = run_steps(n_steps, pl)
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / n_steps, res

    # ---- timed region: K steps, inputs resident in HBM (they exceed L2 by far: no flush needed) ----
    if args.warmup:
        run_steps(args.warmup)
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    launches0 = int(_cabi.lib().vsg_launch_count())
    ms_step, (metrics, n_trip, _) = timed(pipe, args.steps, 0)
    clk = clocks.stop() if rank == 0 else None
    n_launches = int(_cabi.lib().vsg_launch_count()) - launches0
    value = args.videos * world / (ms_step / 1e3)

    # ---- per-kernel roofline legs (a separate, instrumented step; CUDA events on the launch stream) ----
    legs = instrumented_step(pipe, props, graphs, args.precision, ms_step)
    roofline = legs.pop("bigc_gemm")
    # DRAM bytes per launch: a CONSTANT read from the committed ncu pass of the same command (profiles/), not measured inside this run
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_gemm_traffic_v2.summary.json")))
        if args.workload == "vidvrd" and args.videos == 200 and args.precision == "tf32+bf16x2":
            roofline["traffic"] = tj["dram_bytes_per_launch"]
            roofline["traffic_source"] = "constant from profiles/r02_gemm_traffic_v2.summary.json: " + tj["source"]
    except Exception:
        pass
    roofline["also"] = legs.pop("k1_geometry")
    roofline.update({k: v for k, v in legs.items()})

    # ---- other precisions on the same batch (same GT): time, triplet identity against the default mode's output ----
    modes, alts = {}, {}
    for prec in [m for m in args.modes.split(",") if m and m != args.precision]:
        try:                                    # an extra mode must never take the parity mode's line down with it
            alt = Pipeline(args.workload, prec, device, rank, graph=not args.no_graph)
            alt._gts = pipe._gts
            with torch.no_grad():
                a = alt.model(props, topk=alt.wl["topk"])
                b = pipe.model(props, topk=pipe.wl["topk"])
        except Exception as e:                  # noqa: BLE001
            modes[prec] = {"error": "%s: %s" % (type(e).__name__, e)}
            continue
        same, total = compare_triplets(a, [None if t is None else tuple(x.cpu() for x in t) for t in b])
        ms_alt, (m_alt, n_alt, _) = timed(alt, max(3, min(args.steps, 10)), 3)
        alt_legs = instrumented_step(alt, props, graphs, prec, ms_alt)
        modes[prec] = {"value": args.videos * world / (ms_alt / 1e3), "ms_per_step": ms_alt,
                       "videos_with_triplets_identical_to_%s" % args.precision: same, "videos_compared": total,
                       "result": {"mAP": m_alt[0], "R@50": m_alt[1], "R@100": m_alt[2], "triplets": n_alt},
                       "roofline": alt_legs["bigc_gemm"]}
        if cpu_trips is not None:
            s2, t2 = compare_triplets(a[:len(cpu_trips)], cpu_trips)
            modes[prec]["videos_with_triplets_identical_to_cpu_oracle"] = s2
        modes[prec]["decision_flips_vs_%s" % args.precision] = decision_flips(pipe, alt, props)
        alts[prec] = alt

    # ---- e2e: same metric through the public API with HOST buffers (pinned), H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        # the GT relations stay resident (the reference loads its GT json once, too)
        del props, feats
        torch.cuda.empty_cache()
        _, _, hprops, _, hfeats = make_videos(args.workload, seeds, device, seeds[0], pinned=True, with_gt=False)
        hb = HostBatch(hprops, device)
        n_e2e = max(2, min(args.steps, 6))
        dt, n_trip_e = run_e2e(pipe, hb, graphs, n_e2e, barrier, device)
        tdt = torch.tensor([dt], device=device)
        if world > 1:
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
        d2h = n_trip_e * (8 + 4 + 24) + args.videos * 8 * 2         # hit scores, ranks, triplet ids; per-video counts
        # what the platform gives for the same pinned buffer with nothing else running (all ranks copy at once): the e2e ceiling
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            hb.slots[0][0]["feats"].copy_(hb.h["feats"], non_blocking=True)
        c1.record()
        barrier()
        bare = torch.tensor([3 * hb.h["feats"].numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9], device=device)
        bare_all = [bare.clone() for _ in range(world)]
        if world > 1:
            dist.all_gather(bare_all, bare)
        bare_all = [float(t.item()) for t in bare_all]
        e2e = {"value": args.videos * world / float(tdt.item()), "unit": "videos/s", "h2d_bytes_per_step": int(hb.nbytes),
               "d2h_bytes_per_step": int(d2h), "steps": n_e2e,
               "bare_h2d_gbs_per_gpu": bare_all, "bare_h2d_gbs_total": sum(bare_all), "numa_binding_rank0": numa,
               "h2d_bound_videos_per_s": args.videos * world / (hb.nbytes / (min(bare_all) * 1e9)),
               "note": "pinned host buffers, H2D double-buffered on a copy stream; GT relations resident; bare_h2d = the same pinned feature "
                       "buffer copied with every rank copying at once and no kernels running: the platform ceiling of this number"}
        del hb
        torch.cuda.empty_cache()
        if "bf16" in alts:
            # Opt-in of the bf16 mode: the loader hands the features over as bf16 (half the H2D bytes, no cast pass).  This CHANGES THE INPUT
            # CONTRACT of the reference (fp32 .npy features), so it is reported next to the fp32 number, never instead of it.
            import copy
            h16 = torch.empty(hfeats.shape, dtype=torch.bfloat16, pin_memory=True)
            h16.copy_(hfeats)
            p16 = [copy.copy(p) for p in hprops]
            attach_features(p16, h16)
            hb16 = HostBatch(p16, device)
            dt16, _ = run_e2e(alts["bf16"], hb16, graphs, n_e2e, barrier, device)
            t16 = torch.tensor([dt16], device=device)
            if world > 1:
                dist.all_reduce(t16, op=dist.ReduceOp.MAX)
            modes["bf16"]["e2e_bf16_transport"] = {"value": args.videos * world / float(t16.item()), "unit": "videos/s",
                                                   "h2d_bytes_per_step": int(hb16.nbytes), "steps": n_e2e,
                                                   "note": "opt-in: features handed over as bf16 in pinned host memory (input contract differs "
                                                           "from the reference's fp32 features); same H2D double buffering"}
            del hb16, h16, p16
        del hprops, hfeats
        torch.cuda.empty_cache()
    alts.clear()
    del pipe
    torch.cuda.empty_cache()

    vidor = None
    if args.vidor_videos > 0 and args.workload == "vidvrd":
        try:
            vidor = vidor_leg(args, rank, world, device, dist, barrier)
        except Exception as e:                                     # the top-level line must survive a failure of the second leg
            import traceback
            traceback.print_exc(file=sys.stderr)
            vidor = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "videos/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (tcgen05 %s, fp32 accumulate)" % args.precision, "data": "synthetic",
            "config": bench_config(args.workload, wl),
            "run": {"videos_per_gpu": args.videos, "precision": args.precision,
                    "grounding": "in the 'vidor' object (VidVRD has no grounding stage, tools/eval_vidvrd.py)" if args.workload == "vidvrd" else "grd_model_v5 dims, 10 bins",
                    "l2": "inputs (%.1f GB per GPU) exceed L2" % (in_bytes / 1e9), "pipelined": bool(pipelined),
                    "cuda_graph": "BIG-C forward of the resident batch replayed from a CUDA graph" if not args.no_graph else "off",
                    "result": {"mAP": metrics[0], "R@50": metrics[1], "R@100": metrics[2], "triplets": n_trip}},
            "roofline": roofline, "cpu_baseline": cpu, "parity_vs_cpu_oracle": parity, "e2e": e2e, "gpu_launches": n_launches,
            "clocks": clk, "modes": modes or None, "vidor": vidor,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
