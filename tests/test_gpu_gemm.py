"""GPU: the GEMM kernels (SIMT comparator, tcgen05 tf32, tcgen05 3xTF32) against an fp64 product."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [  # M, N, K
    (128, 128, 32), (128, 128, 64), (256, 128, 512), (300, 512, 2048), (37, 512, 512), (1000, 1536, 1024),
    (192, 133, 3160), (4500, 512, 832), (1, 512, 2048), (129, 51, 512), (777, 128, 128), (20000, 512, 512),
]
# 3xTF32: the tensor core accumulates with truncation, so the error grows ~K/8 * 2^-24 (1e-5 at K=2048)
TOL = {0: 2e-6, 1: 2e-3, 2: 2e-5}


def _ref(A, W, bias):
    return A.double() @ W.double().t() + (bias.double() if bias is not None else 0)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("shape", SHAPES, ids=["%dx%dx%d" % s for s in SHAPES])
def test_gemm_modes(mode, shape):
    from vidsgg_big_b200 import linalg
    M, N, K = shape
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = torch.randn(N, K, generator=g).to(DEV) / K ** 0.5
    bias = torch.randn(N, generator=g).to(DEV)
    wt = linalg.Weight(W, bias)
    out = linalg.gemm(mode, A, wt)
    torch.cuda.synchronize()
    ref = _ref(A, W, bias)
    err = (out.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= TOL[mode] * scale, "mode %d shape %s: max err %.3e (scale %.3e)" % (mode, shape, err, scale)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gemm_epilogue_options(mode):
    from vidsgg_big_b200 import linalg
    g = torch.Generator(device="cpu").manual_seed(5)
    M, N, K = 384, 256, 320
    A = torch.randn(M, K + 64, generator=g).to(DEV)            # A is a column slice of a wider buffer
    W = torch.randn(N, K, generator=g).to(DEV) / K ** 0.5
    bias = torch.randn(N, generator=g).to(DEV)
    rowb = torch.randn(192, N, generator=g).to(DEV)
    idx = torch.randint(0, 192, (M,), generator=g).int().to(DEV)
    wt = linalg.Weight(W, bias)
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


@pytest.mark.parametrize("mode", [1, 2])
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
        wt = linalg.Weight(W, b)
        wide = linalg.gemm(mode, A, wt, relu=True).clone()
        old = lib().vsg_gemm_force_bn(128)
        try:
            narrow = linalg.gemm(mode, A, wt, relu=True).clone()
        finally:
            lib().vsg_gemm_force_bn(old)
        if mode == 1:
            assert torch.equal(wide, narrow), (M, N, K)
        else:
            assert (wide - narrow).abs().max().item() <= 1e-5 * narrow.abs().max().item(), (M, N, K)
