"""Base-C pairwise baseline on VidOR-shaped synthetic videos: videos/s of the batched forward (all n(n-1) pairs of
every video through one pair-MLP GEMM) next to the CPU oracle port on a small sample.  usage: python scripts/bench_basec.py [videos]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vidsgg_big_b200 import Base_C, synth
from oracle import basec as obc

n_videos = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda", 0)
cfg = synth.basec_config(rt_triplets_topk=200)
st = synth.make_basec_state(3, cfg)
model = Base_C(cfg); model.load_state_dict(st); model.cuda()
_, wl, props, _, feats = bench.make_videos("vidor", n_videos, 1000, dev)
for p in props:
    f = p.features; p.to(dev); p.features = f
n_pairs = sum(p.num_proposals * (p.num_proposals - 1) for p in props)
with torch.no_grad():
    for _ in range(2):
        out = model(props, topk=3)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); out = model(props, topk=3); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
ms = 1e3 * float(np.median(ts))
# CPU oracle on the first 2 videos (same inputs)
cpu_props = bench.make_videos("vidor", 2, 1000, "cpu", feats=feats[:sum(int(p.lengths.sum()) for p in props[:2])].cpu())[2]
torch.set_num_threads(os.cpu_count() or 1)
with torch.no_grad():
    t0 = time.perf_counter(); ref = obc.forward(st, cfg, cpu_props, 3); cpu_s = time.perf_counter() - t0
same = sum(int((a is None) == (b is None) and (a is None or set(map(tuple, a[0].cpu().tolist())) == set(map(tuple, b[0].tolist()))))
           for a, b in zip(out[:2], ref))
print(json.dumps({"config": "Base-C (exp6 dims, rt_triplets_topk=200), VidOR-shaped videos", "videos": n_videos, "pairs": n_pairs,
                  "ms_per_batch": ms, "videos_per_s": n_videos / ms * 1e3, "pairs_per_s": n_pairs / ms * 1e3,
                  "cpu_oracle_videos_per_s": 2 / cpu_s, "cpu_cores": torch.get_num_threads(), "videos_with_identical_triplet_sets": "%d / 2" % same}))
