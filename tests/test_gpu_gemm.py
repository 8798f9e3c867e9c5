"""GPU: the GEMM kernels (SIMT comparator, tcgen05 tf32, tcgen05 3xTF32) against an fp64 product."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [  # M, N, K
    (128, 128, 32), (128, 128, 64), (256, 128, 512), (300, 512, 2048), (37, 512, 512), (1000, 1536, 1024),
    (192, 133, 3160), (4500, 512, 832), (1, 512, 2048), (129, 51, 512), (777, 128, 128), (20000, 512, 512), (200, 256, 100),
]
# 3xTF32: the tensor core accumulates with truncation, so the error grows ~K/8 * 2^-24 (1e-5 at K=2048)
# tf32 + 2 x bf16 corrections (mode 3): per-product error <= 2^-18 on top of the same accumulation error
# fp16x3 (mode 5): per-product error ~2^-22, same accumulation error
TOL = {0: 2e-6, 1: 2e-3, 2: 2e-5, 3: 3e-5, 5: 3e-5}


def _weight(mode, W, bias=None):
    from vidsgg_big_b200 import linalg
    return linalg.Weight(W, bias, split="bf16" if mode == 3 else "fp16" if mode == 5 else True)


def _ref(A, W, bias):
    return A.double() @ W.double().t() + (bias.double() if bias is not None else 0)


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 5])
@pytest.mark.parametrize("shape", SHAPES, ids=["%dx%dx%d" % s for s in SHAPES])
def test_gemm_modes(mode, shape):
    from vidsgg_big_b200 import linalg
    M, N, K = shape
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = torch.randn(N, K, generator=g).to(DEV) / K ** 0.5
    bias = torch.randn(N, generator=g).to(DEV)
    wt = _weight(mode, W, bias)
    out = linalg.gemm(mode, A, wt)
    torch.cuda.synchronize()
    ref = _ref(A, W, bias)
    err = (out.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= TOL[mode] * scale, "mode %d shape %s: max err %.3e (scale %.3e)" % (mode, shape, err, scale)


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 5])
def test_gemm_epilogue_options(mode):
    from vidsgg_big_b200 import linalg
    g = torch.Generator(device="cpu").manual_seed(5)
    M, N, K = 384, 256, 320
    A = torch.randn(M, K + 64, generator=g).to(DEV)            # A is a column slice of a wider buffer
    W = torch.randn(N, K, generator=g).to(DEV) / K ** 0.5
    bias = torch.randn(N, generator=g).to(DEV)
    rowb = torch.randn(192, N, generator=g).to(DEV)
    idx = torch.randint(0, 192, (M,), generator=g).int().to(DEV)
    wt = _weight(mode, W, bias)
    tol = TOL[mode]
    base = _ref(A[:, 32:32 + K], W, bias)
    # periodic row bias + relu, written into the right half of a wider output
    out = torch.full((M, 2 * N), -7.0, device=DEV)
    linalg.gemm(mode, A[:, 32:32 + K], wt, out=out[:, N:], relu=True, rowbias=rowb, rb_period=192)
    ref = torch.relu(base + rowb.double()[torch.arange(M, device=DEV) % 192])
    assert (out[:, N:].double() - ref).abs().max().item() <= tol * ref.abs().max().item()
    assert bool((out[:, :N] == -7.0).all())
    # indexed row bias + accumulate, no bias
    out2 = torch.ones(M, N, device=DEV)
    linalg.gemm(mode, A[:, 32:32 + K], wt, out=out2, rowbias=rowb, rb_index=idx, accumulate=True, bias=False)
    ref2 = _ref(A[:, 32:32 + K], W, None) + rowb.double()[idx.long()] + 1.0
    assert (out2.double() - ref2).abs().max().item() <= tol * ref2.abs().max().item()


def test_gemm_3xtf32_is_fp32_class():
    """3xTF32 is fp32-class up to the tensor core's truncating accumulation: within 40x of an fp32 cuBLAS GEMM's
    error vs fp64 and >20x better than plain tf32."""
    from vidsgg_big_b200 import linalg
    g = torch.Generator(device="cpu").manual_seed(9)
    A = torch.randn(2048, 2048, generator=g).to(DEV)
    W = torch.randn(512, 2048, generator=g).to(DEV) / 45.0
    wt = linalg.Weight(W)
    out = linalg.gemm(2, A, wt)
    ref = A.double() @ W.double().t()
    torch.backends.cuda.matmul.allow_tf32 = False
    e32 = ((A @ W.t()).double() - ref).abs().max().item()
    e3x = (out.double() - ref).abs().max().item()
    e1x = (linalg.gemm(1, A, wt).double() - ref).abs().max().item()
    print("fp32 err %.3e  3xtf32 err %.3e  tf32 err %.3e" % (e32, e3x, e1x))
    assert e3x <= 40 * e32 + 1e-6 and e1x > 20 * e3x


def test_gemm_tf32_bf16x2_is_fp32_class():
    """Mode 3 (tf32 main product + two bf16 correction products): same error class as 3xTF32 (within 4x of it) and >20x better
    than plain tf32, also when the operands span several orders of magnitude (the corrections are bf16, so no range is lost)."""
    from vidsgg_big_b200 import linalg
    g = torch.Generator(device="cpu").manual_seed(10)
    for scale_spread in (0.0, 3.0):
        A = torch.randn(2048, 2048, generator=g) * torch.exp(scale_spread * torch.randn(2048, 2048, generator=g))
        W = torch.randn(512, 2048, generator=g) / 45.0 * torch.exp(scale_spread * torch.randn(512, 2048, generator=g))
        A, W = A.to(DEV), W.to(DEV)
        ref = A.double() @ W.double().t()
        e3x = (linalg.gemm(2, A, linalg.Weight(W)).double() - ref).abs().max().item()
        emix = (linalg.gemm(3, A, _weight(3, W)).double() - ref).abs().max().item()
        e1x = (linalg.gemm(1, A, linalg.Weight(W)).double() - ref).abs().max().item()
        print("spread %.0f: 3xtf32 err %.3e  tf32+bf16x2 err %.3e  tf32 err %.3e  (scale %.3e)" % (scale_spread, e3x, emix, e1x, ref.abs().max().item()))
        assert emix <= 4 * e3x + 1e-7 * ref.abs().max().item() and e1x > 20 * emix


def test_gemm_fp16x3_is_fp32_class_over_operand_ranges():
    """Mode 5 (three kind::f16 products on fp16 hi / lo pairs): same error class as 3xTF32 (within 4x) and >20x better than plain tf32 for
    O(1) activations, for weights of ANY magnitude (the weight image is scaled by a power of two), for operands that span orders of
    magnitude, and -- with the power-of-two ``a_scale`` of the weight set -- for tiny activations; without it the low part of a tiny A goes
    subnormal (absolute 2^-25) and the error degrades gracefully (still 10x better than tf32).  Compared with an emulation of the scheme too."""
    from vidsgg_big_b200 import linalg
    g = torch.Generator(device="cpu").manual_seed(12)
    for a_mag, w_mag, spread, a_scale in ((1.0, 1 / 45.0, 0.0, 1.0), (1e-3, 1e-3, 0.0, 1024.0), (300.0, 1e-5, 0.0, 1.0), (1e-2, 40.0, 0.0, 64.0),
                                          (1.0, 1 / 45.0, 1.5, 1.0), (1e-3, 1e-3, 0.0, 1.0)):
        A = torch.randn(1024, 2048, generator=g) * a_mag * torch.exp(spread * torch.randn(1024, 2048, generator=g))
        W = torch.randn(512, 2048, generator=g) * w_mag * torch.exp(spread * torch.randn(512, 2048, generator=g))
        A, W = A.to(DEV), W.to(DEV)
        assert A.abs().max().item() * a_scale < 65504
        ref = A.double() @ W.double().t()
        scale = ref.abs().max().item()
        wt5 = _weight(5, W)
        wt5.a_scale = a_scale
        e3x = (linalg.gemm(2, A, linalg.Weight(W)).double() - ref).abs().max().item()
        out5 = linalg.gemm(5, A, wt5)
        e5 = (out5.double() - ref).abs().max().item()
        e1x = (linalg.gemm(1, A, linalg.Weight(W)).double() - ref).abs().max().item()
        print("A %.0e W %.0e spread %.1f a_scale %g: 3xtf32 err %.3e  fp16x3 err %.3e  tf32 err %.3e  (scale %.3e, alpha %g)" %
              (a_mag, w_mag, spread, a_scale, e3x, e5, e1x, scale, wt5.alpha))
        if a_mag * a_scale >= 0.5:
            assert e5 <= 4 * e3x + 1e-7 * scale and e1x > 20 * e5
        else:                                         # tiny activations without a_scale: documented graceful degradation
            assert e1x > 10 * e5
        # emulation: the three products of the split operands in fp64 (the kernel only adds fp32 accumulation error)
        Ws, As = W.double() / wt5.alpha, A.double() * a_scale
        w_hi = Ws.to(torch.float16).double()
        w_lo = (Ws - w_hi).to(torch.float16).double()
        a_hi = As.to(torch.float16).double()
        a_lo = (As - a_hi).float().to(torch.float16).double()
        emu = (a_lo @ w_hi.t() + a_hi @ w_lo.t() + a_hi @ w_hi.t()) * (wt5.alpha / a_scale)
        assert (out5.double() - emu).abs().max().item() <= 2e-5 * scale


def test_tf32_mma_ignores_low_mantissa_bits():
    """The 3xTF32 kernel leaves the raw fp32 A tile in shared memory as the 'high' operand.  That is only valid if
    tcgen05 kind::tf32 ignores the low 13 mantissa bits: the result must be BIT-IDENTICAL to the variant that masks them."""
    from vidsgg_big_b200 import linalg
    from vidsgg_big_b200._cabi import lib
    g = torch.Generator(device="cpu").manual_seed(11)
    A = torch.randn(1000, 1024, generator=g).to(DEV)
    W = torch.randn(384, 1024, generator=g).to(DEV) / 32.0
    wt = linalg.Weight(W)
    old = lib().vsg_gemm_set_store_hi(1)
    try:
        masked = linalg.gemm(2, A, wt).clone()
        lib().vsg_gemm_set_store_hi(0)
        raw = linalg.gemm(2, A, wt).clone()
    finally:
        lib().vsg_gemm_set_store_hi(old)
    assert torch.equal(masked, raw)


@pytest.mark.parametrize("mode", [1, 2, 3, 5])
def test_wide_tiles_equal_narrow_tiles(mode):
    """The 128x256-tile kernel against the 128x128-tile kernel: bit-identical in tf32 mode (same accumulation order over K);
    in 3xTF32 mode the wide kernel walks K in blocks of 16 instead of 32, which reorders the three partial products."""
    from vidsgg_big_b200 import linalg
    from vidsgg_big_b200._cabi import lib
    g = torch.Generator(device="cpu").manual_seed(13)
    for (M, N, K) in ((1000, 512, 1024), (300, 1536, 512), (192, 133, 3160), (4097, 256, 96)):
        A = torch.randn(M, K, generator=g).to(DEV)
        W = torch.randn(N, K, generator=g).to(DEV) / K ** 0.5
        b = torch.randn(N, generator=g).to(DEV)
        wt = _weight(mode, W, b)
        wide = linalg.gemm(mode, A, wt, relu=True).clone()
        old = lib().vsg_gemm_force_bn(128)
        try:
            if mode == 5:       # the fp16 image is the only source of W: built for the tile width in force
                wt = _weight(mode, W, b)
            narrow = linalg.gemm(mode, A, wt, relu=True).clone()
        finally:
            lib().vsg_gemm_force_bn(old)
        if mode in (1, 3, 5):    # same K walk (modes 3 / 5 use 16-column blocks for both tile widths)
            assert torch.equal(wide, narrow), (M, N, K)
        else:
            assert (wide - narrow).abs().max().item() <= 1e-5 * narrow.abs().max().item(), (M, N, K)


@pytest.mark.parametrize("mode", [1, 2, 3, 5])
def test_tma_store_epilogue_equals_direct_stores(mode):
    """The TMA-store epilogue (32x32 slabs staged in shared memory) against the per-row store epilogue: bit-identical C, also
    for ragged M / N (edge slabs fall back to direct stores), column-slice outputs and every epilogue option."""
    from vidsgg_big_b200 import linalg
    from vidsgg_big_b200._cabi import lib
    g = torch.Generator(device="cpu").manual_seed(17)
    for (M, N, K) in ((1000, 512, 256), (333, 1536, 128), (4097, 200, 96), (95, 64, 64), (192, 133, 320), (38400, 512, 512)):
        A = torch.randn(M, K, generator=g).to(DEV)
        W = torch.randn(N, K, generator=g).to(DEV) / K ** 0.5
        b = torch.randn(N, generator=g).to(DEV)
        rowb = torch.randn(192, N, generator=g).to(DEV)
        resid = torch.randn(M, N, generator=g).to(DEV)
        wt = _weight(mode, W, b)
        outs = []
        for on in (1, 0):
            old = lib().vsg_gemm_set_tma_store(on)
            try:
                o1 = torch.full((M, N + 8), -3.0, device=DEV)
                linalg.gemm(mode, A, wt, out=o1[:, 4:4 + N] if N % 4 == 0 else o1[:, :N], relu=True, rowbias=rowb, rb_period=192)
                o2 = torch.ones(M, N, device=DEV)
                linalg.gemm(mode, A, wt, out=o2, accumulate=True, residual=resid)
                lo = torch.zeros(M, N, device=DEV)
                o3 = torch.empty(M, N, device=DEV)
                linalg.gemm(mode, A, wt, out=o3, out_lo=lo, lo_cols=(N // 16 * 4, N // 8 * 4))
                outs.append((o1.clone(), o2.clone(), o3.clone(), lo.clone()))
            finally:
                lib().vsg_gemm_set_tma_store(old)
        for x, y in zip(*outs):
            assert torch.equal(x, y), (mode, M, N, K)
        o3, lo = outs[0][2], outs[0][3]
        want = o3 - (o3.view(torch.int32) & -8192).view(torch.float32)
        c0, c1 = N // 16 * 4, N // 8 * 4                                  # window bounds are multiples of 4 (float4 granularity)
        assert torch.equal(lo[:, c0:c1], want[:, c0:c1])
        assert bool((lo[:, :c0] == 0).all()) and bool((lo[:, c1:] == 0).all())


@pytest.mark.parametrize("mode", [1, 2, 3, 5])
def test_cta_pair_multicast_equals_single_cta(mode):
    """CTA-pair MMAs (cta_group::2) and CTA pairs with W multicast against one CTA per tile: bit-identical C, including an odd
    number of M tiles (the pair's second CTA runs a dummy tile), a single pair, ragged edges and wide / narrow N tiles."""
    from vidsgg_big_b200 import linalg
    from vidsgg_big_b200._cabi import lib
    g = torch.Generator(device="cpu").manual_seed(19)
    for (M, N, K) in ((129, 512, 256), (256, 128, 64), (1000, 512, 1024), (4097, 1536, 512), (38400, 512, 512), (641, 133, 3160),
                      (100000, 64, 96)):
        A = torch.randn(M, K, generator=g).to(DEV)
        W = torch.randn(N, K, generator=g).to(DEV) / K ** 0.5
        b = torch.randn(N, generator=g).to(DEV)
        wt = _weight(mode, W, b)
        outs = []
        for cl in (3, 2, 1):
            old = lib().vsg_gemm_set_cluster(cl)
            try:
                outs.append(linalg.gemm(mode, A, wt, relu=True).clone())
            finally:
                lib().vsg_gemm_set_cluster(old)
        assert torch.equal(outs[1], outs[2]), (mode, M, N, K)
        assert torch.equal(outs[0], outs[1]), ("cta_group::2", mode, M, N, K)
        ref = torch.relu(_ref(A, W, b))
        assert (outs[0].double() - ref).abs().max().item() <= TOL[mode] * ref.abs().max().item()


def test_batched_odd_stage_count_is_race_free():
    """The PV product of the tensor-core attention (N = 64 -> 128-wide tiles, 3 pipeline stages in 3xTF32 mode) repeated many times:
    regression test for a parity-aliasing race between the two split-warp groups when stage ownership followed the iteration."""
    from vidsgg_big_b200 import linalg
    g = torch.Generator(device="cpu").manual_seed(23)
    n_seg, H, Q, dh = 150, 8, 192, 64
    d = H * dh
    S = torch.rand(n_seg * H * Q, Q, generator=g).to(DEV)
    vt = torch.randn(d, n_seg * Q, generator=g).to(DEV)
    vt_hi = (vt.view(torch.int32) & -8192).view(torch.float32)
    vt_lo = vt - vt_hi
    ref = None
    for it in range(30):
        att = torch.empty(n_seg * Q, d, device=DEV)
        linalg.gemm_batched(2, S, vt_hi, vt_lo, Q, dh, Q, att, d, n_seg * H, H, a_off=(H * Q, Q, 0, 0), b_off=(0, dh, Q, 0), c_off=(Q * d, dh))
        torch.cuda.synchronize()
        if ref is None:
            ref = att.clone()
            want = torch.einsum("shqk,shkd->sqhd", S.view(n_seg, H, Q, Q).double(), vt.double().view(H, dh, n_seg, Q).permute(2, 0, 3, 1)).reshape(n_seg * Q, d)
            assert (att.double() - want).abs().max().item() <= 2e-5 * want.abs().max().item()
        else:
            assert torch.equal(att, ref), it


def test_weight_images_equal_tensor_map_loads():
    """Mode 3 with the pre-swizzled weight-tile images (contiguous bulk loads) against the tensor-map loads of the same operands:
    bit-identical C for wide / narrow tiles, ragged N and K, single CTAs and CTA pairs."""
    from vidsgg_big_b200 import linalg
    from vidsgg_big_b200._cabi import lib
    g = torch.Generator(device="cpu").manual_seed(29)
    for (M, N, K) in ((1000, 512, 1024), (300, 1536, 512), (641, 133, 3160), (4097, 64, 96), (128, 128, 16), (777, 200, 100)):
        A = torch.randn(M, K, generator=g).to(DEV)
        W = torch.randn(N, K, generator=g).to(DEV) / K ** 0.5
        b = torch.randn(N, generator=g).to(DEV)
        wt = _weight(3, W, b)
        assert wt.img is not None and wt.img_bn in (128, 256)
        outs = []
        for on, cl in ((1, 3), (0, 3), (1, 2), (0, 2), (1, 1)):
            o1, o2 = lib().vsg_gemm_set_weight_image(on), lib().vsg_gemm_set_cluster(cl)
            try:
                outs.append(linalg.gemm(3, A, wt, relu=True).clone())
            finally:
                lib().vsg_gemm_set_weight_image(o1); lib().vsg_gemm_set_cluster(o2)
        assert all(torch.equal(outs[0], o) for o in outs[1:]), (M, N, K)
        ref = torch.relu(_ref(A, W, b))
        assert (outs[0].double() - ref).abs().max().item() <= TOL[3] * ref.abs().max().item()


@pytest.mark.parametrize("mode", [3, 5])
def test_fused_dwconv_gemm_equals_dwconv_then_gemm(mode):
    """The CONV variant (depthwise conv computed by the split warps from a raw X tile with halo rows) against vsg_dwconv followed by the
    plain mode-3 GEMM: bit-identical, for k = 7 / 3, ragged sequences (incl. length-1 / length-2 ones and sequences that straddle
    tile boundaries), M not a multiple of 128, ReLU + residual epilogues, N = 128 and narrower."""
    import ctypes as C
    from vidsgg_big_b200 import linalg
    from vidsgg_big_b200._cabi import check, lib, stream_ptr
    g = torch.Generator(device="cpu").manual_seed(31)
    H = 128
    for lens, N, k in (([130, 1, 2, 77, 300, 5, 128, 129], 128, 7), ([40] * 3 + [3], 128, 3), ([257, 255], 40, 7), ([1000], 20, 3)):
        M = sum(lens)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        pos = torch.from_numpy(np.concatenate([np.arange(l) for l in lens]).astype(np.int32)).to(DEV)
        rem = torch.from_numpy(np.concatenate([np.arange(l)[::-1] for l in lens]).astype(np.int32)).to(DEV)
        x = torch.randn(M, H, generator=g).to(DEV)
        dw_w = (torch.randn(H, k, generator=g) * 0.4).to(DEV).contiguous()
        dw_b = (torch.randn(H, generator=g) * 0.1).to(DEV).contiguous()
        W = torch.randn(N, H, generator=g).to(DEV) / H ** 0.5
        b = torch.randn(N, generator=g).to(DEV)
        res = torch.randn(M, N, generator=g).to(DEV)
        wt = _weight(mode, W, b)
        assert linalg.can_fuse_dwconv(mode, wt, k=k)
        t = torch.empty_like(x)
        check(lib().vsg_dwconv(C.c_void_p(x.data_ptr()), C.c_void_p(pos.data_ptr()), C.c_void_p(rem.data_ptr()), C.c_void_p(dw_w.data_ptr()),
                               C.c_void_p(dw_b.data_ptr()), k, M, H, C.c_void_p(t.data_ptr()), stream_ptr(x.device)), "vsg_dwconv")
        for relu, rs in ((False, None), (True, res)):
            want = linalg.gemm(mode, t, wt, relu=relu, residual=rs)
            got = linalg.gemm(mode, x, wt, relu=relu, residual=rs, dwconv=(dw_w, dw_b, k, pos, rem))
            assert torch.equal(got, want), (lens, N, k, relu)
        # and the conv itself against a plain torch restatement (zero padding inside each sequence)
        ref = torch.zeros_like(x)
        for s0, s1 in zip(off[:-1], off[1:]):
            seq = x[s0:s1].t()[None]                                   # [1, H, L]
            ref[s0:s1] = torch.nn.functional.conv1d(seq, dw_w[:, None, :], dw_b, padding=k // 2, groups=H)[0].t()
        assert (t - ref).abs().max().item() <= 1e-5


# ---- mode 4: bf16 operands, one kind::f16 pass ------------------------------------------------------------------------
def _bf16_round(x):
    return x.to(torch.bfloat16).double()


@pytest.mark.parametrize("shape", SHAPES + [(481, 1536, 1024), (38400, 512, 512), (5000, 133, 3160)], ids=lambda s: "%dx%dx%d" % s)
def test_gemm_bf16_mode(shape):
    """The bf16 mode against the fp64 product of the bf16-ROUNDED operands (what the tensor core multiplies: only the fp32 accumulation
    differs), and loosely against the unrounded product (the precision the mode actually offers: ~2^-8 per operand)."""
    from vidsgg_big_b200 import linalg
    M, N, K = shape
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = torch.randn(N, K, generator=g).to(DEV) / K ** 0.5
    bias = torch.randn(N, generator=g).to(DEV)
    wt = linalg.Weight(W, bias, split="bf16w")
    out = linalg.gemm(linalg.BF16, A, wt)
    ref16 = _bf16_round(A) @ _bf16_round(W).t() + bias.double()
    scale = ref16.abs().max().item()
    assert (out.double() - ref16).abs().max().item() <= 2e-5 * scale, shape
    assert (out.double() - _ref(A, W, bias)).abs().max().item() <= 3e-2 * scale
    # A already bf16 (no cast launch) + bf16-only output: identical values, rounded once
    A16 = linalg.cast_bf16(A)
    assert torch.equal(A16[:, :K], A.to(torch.bfloat16))
    o16 = linalg.gemm(linalg.BF16, A16, wt, f32_out=False, K=K)
    assert o16.dtype == torch.bfloat16 and torch.equal(o16[:, :N], out.to(torch.bfloat16))


def test_gemm_bf16_epilogue_and_cluster_variants():
    from vidsgg_big_b200 import linalg
    from vidsgg_big_b200._cabi import lib
    g = torch.Generator(device="cpu").manual_seed(41)
    for (M, N, K) in ((384, 256, 320), (1000, 512, 1024), (4097, 1536, 512), (641, 133, 3160), (129, 64, 72)):
        A = torch.randn(M, K + 64, generator=g).to(DEV)
        W = torch.randn(N, K, generator=g).to(DEV) / K ** 0.5
        bias = torch.randn(N, generator=g).to(DEV)
        rowb = torch.randn(192, N, generator=g).to(DEV)
        idx = torch.randint(0, 192, (M,), generator=g).int().to(DEV)
        res = torch.randn(M, N, generator=g).to(DEV)
        wt = linalg.Weight(W, bias, split="bf16w")
        a = A[:, 32:32 + K]
        base = _bf16_round(a) @ _bf16_round(W).t()
        outs = []
        for cl in (3, 2, 1):
            old = lib().vsg_gemm_set_cluster(cl)
            try:
                out = torch.full((M, 2 * N), -7.0, device=DEV)
                o16 = torch.full((M, 2 * ((N + 7) // 8 * 8)), -7.0, device=DEV, dtype=torch.bfloat16)
                linalg.gemm(linalg.BF16, a, wt, out=out[:, N:], relu=True, rowbias=rowb, rb_period=192, residual=res, out16=o16[:, :N])
                outs.append((out.clone(), o16.clone()))
            finally:
                lib().vsg_gemm_set_cluster(old)
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[1][0], outs[2][0]), (M, N, K)
        out, o16 = outs[0]
        ref = torch.relu(base + bias.double() + rowb.double()[torch.arange(M, device=DEV) % 192]) + res.double()
        assert (out[:, N:].double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item(), (M, N, K)
        assert bool((out[:, :N] == -7.0).all())
        assert torch.equal(o16[:, :N], out[:, N:].to(torch.bfloat16)) and bool((o16[:, N:] == -7.0).all())
        # indexed row bias + accumulate, no bias
        out2 = torch.ones(M, N, device=DEV)
        linalg.gemm(linalg.BF16, a, wt, out=out2, rowbias=rowb, rb_index=idx, accumulate=True, bias=False)
        ref2 = base + rowb.double()[idx.long()] + 1.0
        assert (out2.double() - ref2).abs().max().item() <= 2e-5 * ref2.abs().max().item()
