"""Debug aid: one BIG-C forward over N bench videos with optional knobs (env): VSG_TMA_STORE=0/1, VSG_ATT=simt/tc, VSG_PREC=..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vidsgg_big_b200._cabi import lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda", 0)
pipe = bench.Pipeline("vidvrd", os.environ.get("VSG_PREC", "3xtf32"), dev)
if "VSG_TMA_STORE" in os.environ:
    lib().vsg_gemm_set_tma_store(int(os.environ["VSG_TMA_STORE"]))
if "VSG_ATT" in os.environ:
    pipe.model.attention = os.environ["VSG_ATT"]
cfg, wl, props, graphs, feats = bench.make_videos("vidvrd", n, 1000, dev)
for p in props:
    f = p.features; p.to(dev); p.features = f
torch.cuda.synchronize()
try:
    with torch.no_grad():
        trips = pipe.model(props, topk=10)
    torch.cuda.synchronize()
    print("OK", n, sum(0 if t is None else t[0].shape[0] for t in trips))
except Exception as e:
    print("FAIL", n, type(e).__name__, str(e)[:300])
    print("last error:", lib().vsg_last_error())
