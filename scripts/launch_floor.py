#!/usr/bin/env python
"""Where do the ~65 us of a decoder-sized GEMM go?  Times small GEMMs (a) back to back, (b) interleaved with a tiny elementwise kernel,
(c) interleaved with the LayerNorm kernel, per precision mode, with a host head start (spin kernel) so that only GPU time is measured."""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from vidsgg_big_b200 import linalg  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(0)
small = torch.zeros(1024, device=dev)


def bench(fn, between=None, n=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(20e-3 * 1.9e9))
    evs = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
        if between is not None:
            between()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    return t[len(t) // 2] * 1e3


for (M, N, K) in ((192, 512, 512), (5669, 512, 512), (38400, 512, 512), (38400, 1024, 512), (38400, 512, 1024)):
    A = torch.randn(M, K, generator=g).to(dev)
    W = torch.randn(N, K, generator=g).to(dev) / K ** 0.5
    b = torch.randn(N, generator=g).to(dev)
    x = torch.randn(M, N, generator=g).to(dev)
    for mode, split in ((3, "bf16"), (4, "bf16w"), (1, False)):
        wt = linalg.Weight(W, b, split=split)
        A16 = linalg.cast_bf16(A) if mode == 4 else A
        out = torch.empty(M, N, device=dev)
        fn = lambda: linalg.gemm(mode, A16, wt, out=out)
        t0 = bench(fn)
        t1 = bench(fn, lambda: small.add_(1.0))
        t2 = bench(fn, lambda: torch.nn.functional.layer_norm(x, (N,)))
        print("%-20s mode %d: back-to-back %6.1f us | + tiny add %6.1f us | + layer_norm %6.1f us   (%.0f TFLOP/s back-to-back)"
              % ((M, N, K), mode, t0, t1, t2, 2.0 * M * N * K / t0 / 1e6))
