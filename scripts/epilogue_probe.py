#!/usr/bin/env python
"""Where does the epilogue of the tcgen05 GEMM spend its time?  Output-bound shapes (grounding N = K = 128, decoder K = 512) on the
single-CTA probe kernel with the loads and the MMAs switched off (flags 13) and parts of the lean epilogue switched off on top
(64 no bias shuffles, 128 TMEM read only, 256 no fence / TMA store).  Probed results are garbage, timing only."""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from vidsgg_big_b200 import linalg          # noqa: E402
from vidsgg_big_b200._cabi import lib       # noqa: E402

DEV = "cuda:0"
g = torch.Generator(device=DEV).manual_seed(0)


def t(fn, reps=7):
    fn(); torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in evs)[reps // 2] * 1e3


lib().vsg_gemm_set_cluster(1)
for (M, N, K) in ((828259, 128, 128), (38400, 512, 512)):
    A = torch.randn(M, K, generator=g, device=DEV)
    W = torch.randn(N, K, generator=g, device=DEV) / K ** 0.5
    bias = torch.randn(N, generator=g, device=DEV)
    res = torch.randn(M, N, generator=g, device=DEV)
    out = torch.empty(M, N, device=DEV)
    for mode, split in ((1, False), (3, "bf16")):
        wt = linalg.Weight(W, bias, split=split)
        for fl, name in ((0, "normal"), (13, "no loads, no MMA"), (13 + 64, "  + no bias shuffles"), (13 + 256, "  + no fence / TMA store"),
                         (13 + 128, "  + TMEM read only")):
            lib().vsg_gemm_debug_flags(fl)
            try:
                us = t(lambda: linalg.gemm(mode, A, wt, out=out, relu=True))
                us_r = t(lambda: linalg.gemm(mode, A, wt, out=out, relu=True, residual=res))
            finally:
                lib().vsg_gemm_debug_flags(0)
            print("%-20s mode %d %-26s %8.1f us   with residual %8.1f us" % ((M, N, K), mode, name, us, us_r))
lib().vsg_gemm_set_cluster(3)
