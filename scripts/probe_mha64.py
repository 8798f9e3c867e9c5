#!/usr/bin/env python
"""The fused head_dim-64 attention kernel (csrc/attn_tc.cu mha64_tc_kernel) alone on the BIG-C decoder shape:
   python scripts/probe_mha64.py [videos] [queries] [products]      (CUDA-event time per launch; ncu target for the source view)"""
import ctypes as C
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from vidsgg_big_b200._cabi import check, lib, stream_ptr  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 200
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 192
products = int(sys.argv[3]) if len(sys.argv) > 3 else 3
H, d = 8, 512
dev = torch.device("cuda", 0)
qkv = torch.randn(V * Q, 3 * d, device=dev)
out = torch.empty(V * Q, d, device=dev)
raw = lambda t, o=0: C.c_void_p(t.data_ptr() + o)


def run():
    check(lib().vsg_mha_tc64(raw(qkv), 3 * d, raw(qkv, 4 * d), 3 * d, raw(qkv, 8 * d), 3 * d, None, V, Q, H, raw(out), d, None, None, 0, products,
                             stream_ptr(dev)), "vsg_mha_tc64")


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("mha64 V=%d Q=%d products=%d: %.1f us per launch, %.1f useful TFLOP/s, qkv %.0f MB -> %.0f GB/s" %
      (V, Q, products, ms * 1e3, 4.0 * V * Q * Q * d / ms / 1e9, qkv.numel() * 4 / 1e6, (qkv.numel() + out.numel()) * 4 / ms / 1e6))
