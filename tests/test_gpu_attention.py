"""GPU: the tcgen05 attention kernel for head_dim 16 (csrc/attn_tc.cu, the grounding network's mh_attn) against an fp64 softmax
attention and against the fp32 SIMT kernel it replaces."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(qkv, lens, products=None):
    from vidsgg_big_b200 import linalg
    from vidsgg_big_b200._cabi import check, lib, stream_ptr
    H = 128
    off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(DEV)
    out = torch.full((qkv.shape[0], H), float("nan"), device=DEV)
    ld = qkv.stride(0)
    raw = lambda t: C.c_void_p(t.data_ptr())
    q, k, v = raw(qkv), C.c_void_p(qkv.data_ptr() + 4 * H), C.c_void_p(qkv.data_ptr() + 8 * H)
    if products is None:       # SIMT comparator
        bs, bq, nb = linalg.mha_block_list(lens, DEV)
        check(lib().vsg_mha(q, ld, k, ld, v, ld, raw(off), len(lens), 0, int(max(lens)), 8, 16, raw(out), H, raw(bs), raw(bq), nb, stream_ptr(DEV)), "vsg_mha")
    else:
        bs, bq, nb = linalg.mha_block_list(lens, DEV, qb=128)
        check(lib().vsg_mha_tc16(q, ld, k, ld, v, ld, raw(off), 8, raw(out), H, raw(bs), raw(bq), nb, products, stream_ptr(DEV)), "vsg_mha_tc16")
    torch.cuda.synchronize()
    return out


def _ref(qkv, lens):
    H, dh = 128, 16
    outs, r = [], 0
    for L in lens:
        x = qkv[r:r + L].double()
        q, k, v = [t.view(L, 8, dh).transpose(0, 1) for t in (x[:, :H], x[:, H:2 * H], x[:, 2 * H:])]
        outs.append((torch.softmax(q @ k.transpose(-1, -2) / dh ** 0.5, -1) @ v).transpose(0, 1).reshape(L, H))
        r += L
    return torch.cat(outs, 0)


@pytest.fixture(params=[32, 64], ids=["kc32", "kc64"])
def kc(request):
    """Both key-block sizes of the kernel (vsg_mha_tc16_set_kc)."""
    from vidsgg_big_b200._cabi import lib
    old = lib().vsg_mha_tc16_set_kc(request.param)
    yield request.param
    lib().vsg_mha_tc16_set_kc(old)


@pytest.mark.parametrize("lens", [[1], [17, 64, 65], [130, 3, 128, 129, 300], [675, 16, 613], [63] * 40 + [200] * 9, [31, 32, 33, 95, 96, 97]],
                         ids=["one", "small", "mixed", "vidor-max", "many", "block-edges"])
def test_mha_tc16_vs_fp64(lens, kc):
    g = torch.Generator(device="cpu").manual_seed(sum(lens))
    rows = sum(lens)
    # scores with a spread of a few units (post-LayerNorm activations through in_proj), so that the softmax is far from uniform
    qkv = (torch.randn(rows, 384, generator=g) * torch.tensor([2.0] * 128 + [2.0] * 128 + [1.0] * 128)).to(DEV)
    ref = _ref(qkv, lens)
    scale = ref.abs().max().item()
    simt = _run(qkv, lens)
    tc3 = _run(qkv, lens, 3)
    tc1 = _run(qkv, lens, 1)
    e_simt = (simt.double() - ref).abs().max().item() / scale
    e3 = (tc3.double() - ref).abs().max().item() / scale
    e1 = (tc1.double() - ref).abs().max().item() / scale
    print("attention rel err vs fp64: simt %.2e  tcgen05 3xtf32 %.2e  tcgen05 tf32 %.2e" % (e_simt, e3, e1))
    assert not torch.isnan(tc3).any() and not torch.isnan(tc1).any()
    assert e_simt < 1e-5
    assert e3 < 1e-5, "3xTF32 attention is fp32-class"
    assert e1 < 2e-2


def test_mha_tc16_in_column_slices_of_a_wider_buffer(kc):
    """Q / K / V / O strides: the kernel is called on column slices (row stride 384) exactly like grounding._qanet does."""
    lens = [90, 200]
    g = torch.Generator(device="cpu").manual_seed(5)
    qkv = torch.randn(sum(lens), 384, generator=g).to(DEV)
    a = _run(qkv, lens, 3)
    b = _run(qkv.clone(), lens, 3)
    assert torch.equal(a, b)                                          # deterministic
    assert (a.double() - _ref(qkv, lens)).abs().max().item() < 1e-5


# ---- head_dim 64 (BIG-C decoder self-attention): fused tcgen05 kernel on fp16 hi / lo operand pairs --------------------------------
def _ref64(qkv, lens, H, dh):
    d = H * dh
    outs, r = [], 0
    for L in lens:
        x = qkv[r:r + L].double()
        q, k, v = [t.view(L, H, dh).transpose(0, 1) for t in (x[:, :d], x[:, d:2 * d], x[:, 2 * d:])]
        outs.append((torch.softmax(q @ k.transpose(-1, -2) / dh ** 0.5, -1) @ v).transpose(0, 1).reshape(L, d))
        r += L
    return torch.cat(outs, 0)


def _run64(qkv, lens, H, products, fixed=False):
    from vidsgg_big_b200 import linalg
    from vidsgg_big_b200._cabi import check, lib, stream_ptr
    d = H * 64
    out = torch.full((qkv.shape[0], d), float("nan"), device=DEV)
    ld = qkv.stride(0)
    raw = lambda t: C.c_void_p(t.data_ptr())
    q, k, v = raw(qkv), C.c_void_p(qkv.data_ptr() + 4 * d), C.c_void_p(qkv.data_ptr() + 8 * d)
    if fixed:
        assert len(set(lens)) == 1
        check(lib().vsg_mha_tc64(q, ld, k, ld, v, ld, None, len(lens), lens[0], H, raw(out), d, None, None, 0, products, stream_ptr(DEV)), "vsg_mha_tc64")
    else:
        off = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(DEV)
        bs, bq, nb = linalg.mha_block_list(lens, DEV, qb=128)
        check(lib().vsg_mha_tc64(q, ld, k, ld, v, ld, raw(off), len(lens), 0, H, raw(out), d, raw(bs), raw(bq), nb, products, stream_ptr(DEV)), "vsg_mha_tc64")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("lens,fixed", [([192] * 7, True), ([192], True), ([64] * 3, True), ([1], False), ([17, 64, 65, 180, 5], False),
                                        ([130, 3, 128, 129, 300], False), ([63] * 20 + [200] * 5, False), ([15, 16, 17, 47, 48, 49, 192], False)],
                         ids=["decoder", "decoder-layer0", "one-block", "one", "encoder-like", "mixed", "many", "block-edges"])
def test_mha_tc64_vs_fp64(lens, fixed):
    H, dh = 8, 64
    g = torch.Generator(device="cpu").manual_seed(sum(lens) + len(lens))
    rows = sum(lens)
    qkv = (torch.randn(rows, 3 * H * dh, generator=g) * torch.tensor([1.5] * (H * dh) + [1.5] * (H * dh) + [1.0] * (H * dh))).to(DEV)
    ref = _ref64(qkv, lens, H, dh)
    scale = ref.abs().max().item()
    tc3 = _run64(qkv, lens, H, 3, fixed)
    tc1 = _run64(qkv, lens, H, 1, fixed)
    assert not torch.isnan(tc3).any() and not torch.isnan(tc1).any()
    e3 = (tc3.double() - ref).abs().max().item() / scale
    e1 = (tc1.double() - ref).abs().max().item() / scale
    print("attention (dh 64) rel err vs fp64: fp16 pairs %.2e  fp16 hi only %.2e" % (e3, e1))
    assert e3 < 1e-5, "fp16-pair attention is fp32-class"
    assert e1 < 2e-2
    if not fixed:                                                     # the implicit work list of fixed-length sequences == the explicit one
        return
    assert torch.equal(tc3, _run64(qkv, lens, H, 3, False))
    # small-magnitude operands: the low parts go subnormal (absolute 2^-25), which is still fp32-class for O(1) softmax arguments
    small = qkv * torch.tensor([0.05] * (2 * H * dh) + [1e-3] * (H * dh), device=DEV)
    ref_s = _ref64(small, lens, H, dh)
    e_s = (_run64(small, lens, H, 3, True).double() - ref_s).abs().max().item() / ref_s.abs().max().item()
    print("   small operands: %.2e" % e_s)
    assert e_s < 1e-4
