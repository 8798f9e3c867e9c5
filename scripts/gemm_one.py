"""One GEMM shape / mode, a few launches -- the target of `ncu --set full --import-source on` captures.
usage: python scripts/gemm_one.py {feat1|conv|feat2|dec|dec_qkv|grd} {1|2|3|4} [launches] [cluster]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from vidsgg_big_b200 import linalg          # noqa: E402

SH = {"feat1": (481000, 512, 2048), "conv": (481000, 1536, 1024), "feat2": (481000, 512, 512), "dec": (38400, 512, 512),
      "dec_qkv": (38400, 1536, 512), "grd": (828259, 128, 128)}
M, N, K = SH[sys.argv[1]]
mode = int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
g = torch.Generator(device="cuda").manual_seed(0)
A = torch.randn(M, K, generator=g, device="cuda")
W = torch.randn(N, K, generator=g, device="cuda") / K ** 0.5
b = torch.randn(N, generator=g, device="cuda")
wt = linalg.Weight(W, b, split={3: "bf16", 4: "bf16w"}.get(mode, True))
out = torch.empty(M, N, device="cuda")
if mode == 4:
    A = linalg.cast_bf16(A)
if len(sys.argv) > 4:                       # optional: cluster variant (1 / 2 / 3)
    from vidsgg_big_b200._cabi import lib
    lib().vsg_gemm_set_cluster(int(sys.argv[4]))
for _ in range(n):
    linalg.gemm(mode, A, wt, out=out, relu=True)
torch.cuda.synchronize()
print("ok", float(out[0, 0]))
