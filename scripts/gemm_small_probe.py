#!/usr/bin/env python
"""Per-tile overhead of the tcgen05 GEMM on the grounding-sized (N = K = 128) and decoder-sized (K = 512) problems: modes x cluster variants
x TMA-store on / off x probe flags (parts of the single-CTA kernel switched off; probed results are garbage, timing only)."""
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


for (M, N, K) in ((828259, 128, 128), (38400, 512, 512), (38400, 128, 128)):
    A = torch.randn(M, K, generator=g, device=DEV)
    W = torch.randn(N, K, generator=g, device=DEV) / K ** 0.5
    res = torch.randn(M, N, generator=g, device=DEV)
    out = torch.empty(M, N, device=DEV)
    tiles = (M + 127) // 128 * ((N + 127) // 128 if N <= 128 else (N + 255) // 256)
    for mode, split in ((1, False), (3, "bf16"), (4, "bf16w")):
        wt = linalg.Weight(W, None, split=split)
        A_in = linalg.cast_bf16(A) if mode == 4 else A
        for cl in (3, 2, 1):
            lib().vsg_gemm_set_cluster(cl)
            for tma in (1, 0):
                lib().vsg_gemm_set_tma_store(tma)
                for fl, fname in ((0, "normal"), (4, "no MMA"), (9, "no loads"), (13, "no loads, no MMA"), (2, "no split")):
                    if fl and (cl != 1 or mode == 4):
                        continue
                    if fl == 2 and mode == 1:
                        continue
                    lib().vsg_gemm_debug_flags(fl)
                    try:
                        us = t(lambda: linalg.gemm(mode, A_in, wt, out=out))
                        us_r = t(lambda: linalg.gemm(mode, A_in, wt, out=out, residual=res)) if fl == 0 else float("nan")
                    finally:
                        lib().vsg_gemm_debug_flags(0)
                    print("%-20s mode %d cl %d tma_store %d %-18s %8.1f us  (+residual %8.1f us)  %.2f us/tile/SM-slot" %
                          ((M, N, K), mode, cl, tma, fname, us, us_r, us / max(1.0, tiles / 148.0)))
            lib().vsg_gemm_set_tma_store(1)
    lib().vsg_gemm_set_cluster(3)
