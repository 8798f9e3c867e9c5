"""Thin host wrapper of the GEMM entry points (csrc/gemm.cu).  CUDA tensors only."""
from __future__ import annotations

from ctypes import c_void_p
from typing import Optional

import torch

from ._cabi import check, lib, ptr, require_cuda, stream_ptr

SIMT, TF32, X3TF32, TF32_BF16X2, BF16, FP16X3 = 0, 1, 2, 3, 4, 5
MODES = {"fp32_simt": SIMT, "tf32": TF32, "3xtf32": X3TF32, "tf32+bf16x2": TF32_BF16X2, "bf16": BF16, "fp16x3": FP16X3}
SPLITS = {X3TF32: "tf32", TF32_BF16X2: "bf16", BF16: "bf16w", FP16X3: "fp16"}      # mode -> what ``Weight`` must precompute (default: nothing)
FP32_CLASS = (X3TF32, TF32_BF16X2, FP16X3)      # modes whose products carry >= 19 significant bits (parity modes)


def can_fuse_dwconv(mode: int, W: "Weight", K: Optional[int] = None, k: int = 7) -> bool:
    """The depthwise conv can run inside the GEMM's operand pipeline (csrc/gemm.cu, CONV variant)."""
    K = W.K if K is None else K
    ready = (mode == TF32_BF16X2 and W.w16 is not None) or (mode == FP16X3 and W.img16 is not None)
    return ready and W.N <= 128 and K % 16 == 0 and K * (k + 1) <= 2048 and k % 2 == 1 and k <= 7


def attention_mode(mode: int) -> int:
    """Mode of the batched attention GEMMs (activation x activation): TF32_BF16X2 needs pre-computed bf16 copies of the "weight"
    operand, which only exist for real weights -- those few problems stay on the 3xTF32 kernel."""
    if mode == BF16:
        return TF32             # reduced-precision mode: single-pass tf32 attention products
    return X3TF32 if mode in (TF32_BF16X2, FP16X3) else mode


def fp16_image_scale_log2(wmax: float) -> int:
    """Exponent s of the power-of-two scale of an fp16x3 weight image: max |w| * 2^s lies in [2^13, 2^14) -- below fp16's 65504 with room for
    rounding, and high enough that the low part fp16(w 2^s - hi) stays a normal number for |w| >= 2^-16 max |w| (csrc/gemm.cu mode 5).
    0 for an all-zero / non-finite weight; clamped to +-100 so that 2^s and 2^-s stay finite fp32 values."""
    import math
    if not (wmax > 0.0) or not math.isfinite(wmax):
        return 0
    return max(-100, min(100, 14 - math.frexp(wmax)[1]))


class Weight(object):
    """A [N,K] fp32 weight resident in HBM with its (optional) tf32 hi/lo split."""

    def __init__(self, w: torch.Tensor, bias: Optional[torch.Tensor] = None, split=True):
        """``split``: True / "tf32" -> fp32 hi/lo pair of the 3xTF32 mode; "bf16" -> the bf16 copies of the tf32+bf16x2 mode
        (w16 = bf16(w), lo16 = bf16(w - trunc_tf32(w)), rows padded to a multiple of 8); False -> none."""
        require_cuda(w)
        self.w = w.detach().float().contiguous()
        self.bias = None if bias is None else bias.detach().float().contiguous()
        self.N, self.K = self.w.shape
        self.hi = self.lo = self.w16 = self.lo16 = self.img = self.img16 = None
        self.img_bn = self.img16_bn = 0
        self.alpha = 1.0
        self.a_scale = 1.0         # fp16x3: optional power of two applied to the A operand before its fp16 split (csrc/gemm.cu mode 5)
        if split == "fp16":
            # fp16x3 mode: power-of-two scale that puts max |w| into [2^13, 2^14) (exact; undone by alpha in the GEMM epilogue)
            s = fp16_image_scale_log2(float(self.w.abs().max()) if self.w.numel() else 0.0)
            self.alpha = 2.0 ** (-s)
            self.img16_bn = int(lib().vsg_gemm_tile_n(self.N))
            nbytes = int(lib().vsg_weight_image_fp16_bytes(self.N, self.K, self.img16_bn))
            self.img16 = torch.empty(nbytes, dtype=torch.uint8, device=self.w.device)
            check(lib().vsg_build_weight_image_fp16(ptr(self.w), self.w.stride(0), self.N, self.K, self.img16_bn, 2.0 ** s, ptr(self.img16),
                                                    stream_ptr(self.w.device)), "vsg_build_weight_image_fp16")
            return
        if split in ("bf16", "bf16w"):
            self.ld16 = (self.K + 7) // 8 * 8
            self.w16 = torch.empty(self.N, self.ld16, dtype=torch.bfloat16, device=self.w.device)
            self.lo16 = torch.empty_like(self.w16)
            check(lib().vsg_split_bf16(ptr(self.w), self.w.stride(0), self.N, self.K, ptr(self.w16), ptr(self.lo16), self.ld16,
                                       stream_ptr(self.w.device)), "vsg_split_bf16")
            if split == "bf16w":          # bf16 mode: only bf16(W) is read
                self.lo16 = None
                return
            # pre-swizzled tile images for the tile width the kernel will pick (contiguous bulk loads of the W operands)
            self.img_bn = int(lib().vsg_gemm_tile_n(self.N))
            nbytes = int(lib().vsg_weight_image_bytes(self.N, self.K, self.img_bn))
            self.img = torch.empty(nbytes, dtype=torch.uint8, device=self.w.device)
            check(lib().vsg_build_weight_image(ptr(self.w), self.w.stride(0), self.N, self.K, self.img_bn, ptr(self.img),
                                               stream_ptr(self.w.device)), "vsg_build_weight_image")
        elif split:
            self.hi = torch.empty_like(self.w)
            self.lo = torch.empty_like(self.w)
            check(lib().vsg_split_tf32(ptr(self.w), ptr(self.hi), ptr(self.lo), self.w.numel(), stream_ptr(self.w.device)),
                  "vsg_split_tf32")


def _raw(t):
    return None if t is None else c_void_p(t.data_ptr())


class _Profile(object):
    """Optional per-launch CUDA-event timing (bench.py's roofline legs): every GEMM launch, plus any kernel a caller brackets
    with ``span``.  ``stage`` is a free-form label set by the caller (e.g. "bigc" / "grounding") and stored with each record."""
    enabled = False
    stage = ""
    records = []          # (start_event, end_event, useful_flops, kind, stage, meta)

    @classmethod
    def begin(cls):
        cls.enabled, cls.records, cls.stage = True, [], ""

    @classmethod
    def span(cls, kind, flops=0.0, meta=None):
        """Context manager: brackets the enclosed launches with events on the current stream (no-op unless enabled)."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            if not cls.enabled:
                yield
                return
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            yield
            ev1.record()
            cls.records.append((ev0, ev1, float(flops), kind, cls.stage, meta))
        return cm()

    @classmethod
    def end(cls, kind="gemm", stage=None):
        """-> (n_launches, useful_flops, kernel_ms) of the records of ``kind`` (and ``stage``, if given) after synchronising."""
        cls.enabled = False
        torch.cuda.synchronize()
        out = cls.summary(kind, stage)
        return out

    @classmethod
    def table(cls, kind=None):
        """[(kind, stage, meta, ms, useful TFLOP/s)] of every record (after ``end``)."""
        out = []
        for r in cls.records:
            if kind is None or r[3] == kind:
                ms = r[0].elapsed_time(r[1])
                out.append((r[3], r[4], r[5], ms, r[2] / (ms * 1e-3) / 1e12 if ms > 0 else 0.0))
        return out

    @classmethod
    def summary_bytes(cls, kind="gemm", stage=None):
        """Sum of the compulsory-HBM-bytes field of the records that carry one (plain fp32 GEMM launches)."""
        return sum(r[6] for r in cls.records if r[3] == kind and (stage is None or r[4] == stage) and len(r) > 6)

    @classmethod
    def summary(cls, kind="gemm", stage=None):
        sel = [r for r in cls.records if r[3] == kind and (stage is None or r[4] == stage)]
        return len(sel), sum(r[2] for r in sel), sum(r[0].elapsed_time(r[1]) for r in sel)


LAUNCHES = [0]            # GEMM launches issued through this wrapper (bench.py's gpu_launches)


def cast_bf16(x: torch.Tensor, K: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 copy ``[rows, ceil8(K)]`` of the first K columns of an fp32 row-major matrix (zero padding columns): the A operand of
    the bf16 mode when the producer wrote fp32."""
    require_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.float32
    K = x.shape[1] if K is None else K
    ld = (K + 7) // 8 * 8
    if out is None:
        out = torch.empty(x.shape[0], ld, dtype=torch.bfloat16, device=x.device)
    assert out.dtype == torch.bfloat16 and out.stride(1) == 1 and out.stride(0) % 8 == 0 and out.shape[1] >= K
    LAUNCHES[0] += 1
    with _Profile.span("cast", 0.0, (x.shape[0], K)):
        check(lib().vsg_cast_bf16(_raw(x), x.stride(0) if x.shape[0] > 1 else max(x.stride(0), K), x.shape[0], K, _raw(out),
                                  out.stride(0) if x.shape[0] > 1 else max(out.stride(0), ld), stream_ptr(x.device)), "vsg_cast_bf16")
    return out


def gemm(mode: int, A: torch.Tensor, W: Weight, out: Optional[torch.Tensor] = None, relu: bool = False,
         rowbias: Optional[torch.Tensor] = None, rb_index: Optional[torch.Tensor] = None, rb_period: int = 0,
         accumulate: bool = False, bias: bool = True, K: Optional[int] = None,
         residual: Optional[torch.Tensor] = None, out_lo: Optional[torch.Tensor] = None, lo_cols=None, dwconv=None,
         out16: Optional[torch.Tensor] = None, f32_out: bool = True) -> torch.Tensor:
    """``out[:, :N] = act(A[:, :K] @ W.w.T + bias (+ rowbias) (+ out))``.  ``A`` / ``out`` may be column slices of
    wider row-major buffers (their row stride is passed as the leading dimension).  ``out_lo`` (same shape / strides as ``out``)
    receives ``x - trunc_tf32(x)`` of the result, restricted to the column window ``lo_cols = (begin, end)`` if given.
    ``dwconv = (dw_w [K,k], dw_b [K], k, seq_pos i32[M], seq_rem i32[M])``: A is replaced by its depthwise conv over the row axis inside
    the kernel (mode tf32+bf16x2, N <= 128; see ``can_fuse_dwconv``).
    Mode ``BF16``: ``A`` may be fp32 (cast by ``cast_bf16`` first) or already bf16; ``out16`` (bf16 ``[M, >= N]``, row stride a
    multiple of 4) also receives the result as bf16; ``f32_out=False`` skips the fp32 output and returns ``out16``."""
    require_cuda(A)
    assert A.dim() == 2 and A.stride(1) == 1
    M = A.shape[0]
    K = W.K if K is None else K
    assert A.shape[1] >= K
    if mode == BF16:
        return _gemm_bf16(A, W, out, relu, rowbias, rb_index, rb_period, accumulate, bias, K, residual, out16, f32_out)
    assert A.dtype == torch.float32 and out16 is None and f32_out
    if out is None:
        out = torch.empty(M, W.N, dtype=torch.float32, device=A.device)
    assert out.dim() == 2 and out.stride(1) == 1 and out.shape[0] == M and out.shape[1] >= W.N
    lda = A.stride(0) if M > 1 else max(A.stride(0), K)
    ldc = out.stride(0) if M > 1 else max(out.stride(0), W.N)
    w_hi = W.hi if (mode == X3TF32 and W.hi is not None) else W.w
    w_lo = W.lo if mode == X3TF32 else None
    b = W.bias if bias else None
    LAUNCHES[0] += 1
    if _Profile.enabled:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    if mode == TF32_BF16X2 and (W.w16 is None or W.lo16 is None):
        raise ValueError("gemm: mode tf32+bf16x2 needs a Weight built with split='bf16'")
    assert dwconv is None or can_fuse_dwconv(mode, W, K)
    if mode == FP16X3 and (W.img16 is None or K != W.K):
        raise ValueError("gemm: mode fp16x3 needs a Weight built with split='fp16' and K == W.K")
    if out_lo is not None or mode in (TF32_BF16X2, FP16X3):
        from ._cabi import VsgGemmArgs
        import ctypes as C
        assert out_lo is None or (out_lo.shape == out.shape and out_lo.stride() == out.stride())
        a = VsgGemmArgs()
        a.mode = mode; a.A = A.data_ptr(); a.lda = lda; a.W_hi = w_hi.data_ptr(); a.W_lo = None if w_lo is None else w_lo.data_ptr()
        a.ldw = W.w.stride(0); a.M, a.N, a.K = M, W.N, K
        a.bias = None if b is None else b.data_ptr(); a.rowbias = None if rowbias is None else rowbias.data_ptr()
        a.rb_index = None if rb_index is None else rb_index.data_ptr(); a.rb_period = int(rb_period)
        a.ld_rb = 0 if rowbias is None else rowbias.stride(0); a.relu = 1 if relu else 0; a.accumulate = 1 if accumulate else 0
        a.residual = None if residual is None else residual.data_ptr(); a.ld_res = 0 if residual is None else residual.stride(0)
        a.C = out.data_ptr(); a.C_lo = None if out_lo is None else out_lo.data_ptr(); a.ldc = ldc; a.batch = 1; a.batch_inner = 1
        if lo_cols is not None:
            a.lo_col_begin, a.lo_col_end = int(lo_cols[0]), int(lo_cols[1])
        if dwconv is not None:
            dw_w, dw_b, dw_k, seq_pos, seq_rem = dwconv
            a.dw_w, a.dw_b, a.dw_k, a.seq_pos, a.seq_rem = dw_w.data_ptr(), dw_b.data_ptr(), int(dw_k), seq_pos.data_ptr(), seq_rem.data_ptr()
        if mode == FP16X3:
            a.W_img16, a.img16_bn, a.w_alpha, a.a_scale = W.img16.data_ptr(), W.img16_bn, W.alpha, float(W.a_scale)
        if mode == TF32_BF16X2:
            a.W_b16, a.W_lo16, a.ldw16 = W.w16.data_ptr(), W.lo16.data_ptr(), W.ld16
            if W.img is not None and K == W.K:           # the image's k blocks cover the whole of W.K
                a.W_img, a.img_bn = W.img.data_ptr(), W.img_bn
        check(lib().vsg_gemm_ex(C.byref(a), stream_ptr(A.device)), "vsg_gemm_ex")
    else:
        check(lib().vsg_gemm(mode, _raw(A), lda, _raw(w_hi), _raw(w_lo), W.w.stride(0), M, W.N, K, _raw(b), _raw(rowbias),
                             _raw(rb_index), int(rb_period), 0 if rowbias is None else rowbias.stride(0), 1 if relu else 0,
                             1 if accumulate else 0, _raw(residual), 0 if residual is None else residual.stride(0),
                             _raw(out), ldc, stream_ptr(A.device)), "vsg_gemm")
    if _Profile.enabled:
        ev1.record()
        # 7th field: compulsory HBM bytes of the launch (A + C (+ residual), fp32; W is L2-resident) for the HBM-model view of narrow GEMMs
        _Profile.records.append((ev0, ev1, 2.0 * M * W.N * K, "gemm", _Profile.stage, (M, W.N, K),
                                 4.0 * M * (K + W.N * (2 if residual is not None else 1))))
    return out


def _gemm_bf16(A, W, out, relu, rowbias, rb_index, rb_period, accumulate, bias, K, residual, out16, f32_out):
    """Mode BF16 (csrc/gemm.cu, B16 kernel): bf16 operands, one kind::f16 pass, fp32 accumulate."""
    from ._cabi import VsgGemmArgs
    import ctypes as C
    if W.w16 is None:
        raise ValueError("gemm: mode bf16 needs a Weight built with split='bf16w' (or 'bf16')")
    M = A.shape[0]
    if M == 0:
        return out if f32_out else out16
    A16 = A if A.dtype == torch.bfloat16 else cast_bf16(A, K)
    assert A16.stride(1) == 1 and (A16.stride(0) % 8 == 0 or M == 1) and A16.data_ptr() % 16 == 0
    if f32_out and out is None:
        out = torch.empty(M, W.N, dtype=torch.float32, device=A.device)
    if not f32_out and out16 is None:
        out16 = torch.empty(M, (W.N + 7) // 8 * 8, dtype=torch.bfloat16, device=A.device)
    a = VsgGemmArgs()
    a.mode = BF16; a.M, a.N, a.K = M, W.N, K
    a.A16 = A16.data_ptr(); a.lda16 = A16.stride(0) if M > 1 else max(A16.stride(0), (K + 7) // 8 * 8)
    a.W_b16 = W.w16.data_ptr(); a.ldw16 = W.ld16
    b = W.bias if bias else None
    a.bias = None if b is None else b.data_ptr(); a.rowbias = None if rowbias is None else rowbias.data_ptr()
    a.rb_index = None if rb_index is None else rb_index.data_ptr(); a.rb_period = int(rb_period)
    a.ld_rb = 0 if rowbias is None else rowbias.stride(0); a.relu = 1 if relu else 0; a.accumulate = 1 if accumulate else 0
    a.residual = None if residual is None else residual.data_ptr(); a.ld_res = 0 if residual is None else residual.stride(0)
    if f32_out:
        assert out.dim() == 2 and out.stride(1) == 1 and out.shape[0] == M and out.shape[1] >= W.N and out.dtype == torch.float32
        a.C = out.data_ptr(); a.ldc = out.stride(0) if M > 1 else max(out.stride(0), W.N)
    if out16 is not None:
        assert out16.dtype == torch.bfloat16 and out16.stride(1) == 1 and out16.shape[0] == M and out16.shape[1] >= W.N
        a.C16 = out16.data_ptr(); a.ldc16 = out16.stride(0) if M > 1 else max(out16.stride(0), W.N)
    a.batch = 1; a.batch_inner = 1
    LAUNCHES[0] += 1
    with _Profile.span("gemm", 2.0 * M * W.N * K, (M, W.N, K)):
        check(lib().vsg_gemm_ex(C.byref(a), stream_ptr(A.device)), "vsg_gemm_ex")
    return out if f32_out else out16


def gemm_batched(mode: int, A: torch.Tensor, W_hi: torch.Tensor, W_lo: Optional[torch.Tensor], M: int, N: int, K: int,
                 out: torch.Tensor, ldc: int, batch: int, batch_inner: int, a_off=(0, 0, 0, 0), b_off=(0, 0, 0, 0), c_off=(0, 0)):
    """``batch`` independent M x N x K problems inside larger operands (vsg_gemm_ex batched form).  ``A`` / ``W_hi`` / ``W_lo`` are 2-D
    views (row stride = leading dimension, shape = TMA bounds) of the full operands; problem p -> outer = p // batch_inner,
    inner = p % batch_inner; ``a_off`` / ``b_off`` = (row_outer, row_inner, col_outer, col_inner) element offsets, ``c_off`` =
    (outer, inner) element offsets into ``out``."""
    from ._cabi import VsgGemmArgs
    import ctypes as C
    a = VsgGemmArgs()
    a.mode = mode
    a.A = A.data_ptr(); a.lda = A.stride(0); a.a_rows, a.a_cols = A.shape
    a.W_hi = W_hi.data_ptr(); a.W_lo = None if W_lo is None else W_lo.data_ptr()
    a.ldw = W_hi.stride(0); a.w_rows, a.w_cols = W_hi.shape
    a.M, a.N, a.K = M, N, K
    a.C = out.data_ptr(); a.C_lo = None; a.ldc = ldc
    a.batch, a.batch_inner = batch, batch_inner
    a.a_row_outer, a.a_row_inner, a.a_col_outer, a.a_col_inner = a_off
    a.b_row_outer, a.b_row_inner, a.b_col_outer, a.b_col_inner = b_off
    a.c_outer, a.c_inner = c_off
    LAUNCHES[0] += 1
    if _Profile.enabled:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    check(lib().vsg_gemm_ex(C.byref(a), stream_ptr(out.device)), "vsg_gemm_ex")
    if _Profile.enabled:
        ev1.record()
        _Profile.records.append((ev0, ev1, 2.0 * M * N * K * batch, "gemm", _Profile.stage, (M, N, K, batch)))
    return out


def mha_block_list(lengths, device, qb: int = 64):
    """(blk_seg int32[B], blk_q0 int32[B]) work list of csrc/bigc.cu mha_kernel for ragged segment lengths (host ints)."""
    import numpy as np
    lengths = np.asarray(lengths, dtype=np.int64)
    nblk = (lengths + qb - 1) // qb
    seg = np.repeat(np.arange(lengths.size, dtype=np.int32), nblk)
    first = np.concatenate([[0], np.cumsum(nblk)])[:-1]
    q0 = (np.arange(int(nblk.sum()), dtype=np.int64) - np.repeat(first, nblk)) * qb
    return (torch.from_numpy(seg.astype(np.int32)).to(device), torch.from_numpy(q0.astype(np.int32)).to(device), int(nblk.sum()))
