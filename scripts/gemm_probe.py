"""Bottleneck probes of the tcgen05 GEMM kernel (run on a B200): times the bench's dominant GEMM shapes in every tensor-core mode
with parts of the kernel switched off through vsg_gemm_debug_flags (results of the probed launches are garbage; timing only).

    python scripts/gemm_probe.py [--json out.json]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from vidsgg_big_b200 import linalg          # noqa: E402
from vidsgg_big_b200._cabi import lib       # noqa: E402

DEV = "cuda:0"
SHAPES = [("feat1", 481000, 512, 2048), ("conv", 481000, 1536, 1024), ("feat2", 481000, 512, 512), ("dec", 38400, 512, 512),
          ("dec_qkv", 38400, 1536, 512)]
if "--quick" in sys.argv:
    SHAPES = [SHAPES[0], SHAPES[3]]
FLAGS = [(0, "normal"), (1, "no W loads"), (8, "no A loads"), (9, "no loads"), (2, "no split"), (4, "no MMA"), (6, "no split, no MMA"),
         (13, "no loads, no MMA")]
CLUSTERS = (3, 2, 1)


def time_gemm(mode, A, wt, out, reps=5):
    linalg.gemm(mode, A, wt, out=out)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        linalg.gemm(mode, A, wt, out=out)
        b.record()
    torch.cuda.synchronize()
    return min(a.elapsed_time(b) for a, b in evs)


def main():
    res = []
    g = torch.Generator(device=DEV).manual_seed(0)
    for name, M, N, K in SHAPES:
        A = torch.randn(M, K, generator=g, device=DEV)
        W = torch.randn(N, K, generator=g, device=DEV) / K ** 0.5
        out = torch.empty(M, N, device=DEV)
        for mode, mname in ((1, "tf32"), (2, "3xtf32"), (3, "tf32+bf16x2")):
            wt = linalg.Weight(W, None, split="bf16" if mode == 3 else True)
            for cl in CLUSTERS:
                lib().vsg_gemm_set_cluster(cl)
                for fl, fname in FLAGS:
                    if mode == 1 and (fl & 2):
                        continue
                    if fl and (cl >= 2 or "--normal" in sys.argv):      # the probe flags are only meaningful for the single-CTA kernel
                        continue
                    lib().vsg_gemm_debug_flags(fl)
                    try:
                        ms = time_gemm(mode, A, wt, out)
                    finally:
                        lib().vsg_gemm_debug_flags(0)
                    tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
                    res.append(dict(shape=name, M=M, N=N, K=K, mode=mname, cluster=cl, probe=fname, ms=ms, useful_tflops=tf))
                    print("%-8s %-12s cl%d %-18s %8.3f ms  %7.1f TF/s useful" % (name, mname, cl, fname, ms, tf), flush=True)
            lib().vsg_gemm_set_cluster(3)
        del A, W, out
    if "--json" in sys.argv:
        json.dump(res, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
