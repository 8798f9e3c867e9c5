"""K1 at the stress shape of BASELINE.json configs[4]: one video, 256 tracklets, 10k frames, L ~ U{2000..10000}.
Times both kernel variants with CUDA events (L2 flushed between repeats) and reports algorithmic GB/s."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vidsgg_big_b200 import synth, geometry

dev = "cuda:0"
n, vlen = int(os.environ.get("N", 256)), 10000
P = synth.make_proposal(4242, n, vlen, 8, 36, min_len=2000, max_len=10000, with_features=False).to(dev)
T = geometry.TrackTable.from_containers([P])
d = P.traj_durations.cpu().numpy()
s = np.maximum(d[:, None, 0], d[None, :, 0]); e = np.minimum(d[:, None, 1], d[None, :, 1])
ov = np.clip(e - s + 1, 0, None)
sumL = int(P.lengths.sum())
alg = 32 * int(ov.sum()) + 16 * 2 * sumL + n * n * 21
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
out = {"tracks": n, "box_frames": sumL, "frame_pairs": int(ov.sum()), "algorithmic_bytes": alg, "compulsory_bytes": 16 * sumL + n * n * 21}
ref = None
for variant in (1, 2):
    for _ in range(3):
        geometry.traj_viou_batched(T, T, variant=variant)
    ts = []
    for _ in range(10):
        flush.zero_()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        v, sp, m, seg, _ = geometry.traj_viou_batched(T, T, variant=variant)
        t1.record(); torch.cuda.synchronize()
        ts.append(t0.elapsed_time(t1))
    ms = float(np.median(ts))
    if ref is None: ref = v.clone()
    out["variant%d" % variant] = {"ms": ms, "algorithmic_GBps": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / peaks["hbm_gbs"],
                                  "max_abs_diff_vs_variant1": float((v - ref).abs().max())}
print(json.dumps(out))
