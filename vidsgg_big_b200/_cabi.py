"""ctypes binding of libvsgb200.so (include/vsg_b200.h).

There is deliberately NO CPU fallback: if the library cannot be loaded or a call fails, the
caller gets an exception.  Tensors are passed as raw device pointers (``tensor.data_ptr()``)
plus the current CUDA stream; PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvsgb200.so")

_lib: Optional[C.CDLL] = None

p = C.c_void_p
i32, i64, f32, f64 = C.c_int, C.c_int64, C.c_float, C.c_double


class VsgRelTable(C.Structure):
    _fields_ = [("boxes", p), ("off", p), ("tstart", p), ("rel", p), ("vid_off", p), ("n_rel", i64), ("box_f64", i32),
                ("vol_full_track", i32)]


class VsgGemmArgs(C.Structure):
    _fields_ = [("mode", i32), ("A", p), ("lda", i32), ("a_rows", i32), ("a_cols", i32),
                ("W_hi", p), ("W_lo", p), ("ldw", i32), ("w_rows", i32), ("w_cols", i32),
                ("M", i32), ("N", i32), ("K", i32),
                ("bias", p), ("rowbias", p), ("rb_index", p), ("rb_period", i32), ("ld_rb", i32), ("relu", i32), ("accumulate", i32),
                ("residual", p), ("ld_res", i32), ("C", p), ("C_lo", p), ("ldc", i32),
                ("batch", i32), ("batch_inner", i32),
                ("a_row_outer", i32), ("a_row_inner", i32), ("a_col_outer", i32), ("a_col_inner", i32),
                ("b_row_outer", i32), ("b_row_inner", i32), ("b_col_outer", i32), ("b_col_inner", i32),
                ("c_outer", C.c_longlong), ("c_inner", C.c_longlong),
                ("lo_col_begin", i32), ("lo_col_end", i32), ("W_b16", p), ("W_lo16", p), ("ldw16", i32),
                ("W_img", p), ("img_bn", i32),
                ("dw_w", p), ("dw_b", p), ("seq_pos", p), ("seq_rem", p), ("dw_k", i32),
                ("A16", p), ("lda16", i32), ("C16", p), ("ldc16", i32),
                ("W_img16", p), ("img16_bn", i32), ("w_alpha", C.c_float), ("a_scale", C.c_float)]


class VsgLinear(C.Structure):
    _fields_ = [("w", p), ("bias", p), ("N", i32), ("K", i32), ("ldw", i32), ("hi", p), ("lo", p), ("w16", p), ("lo16", p), ("ld16", i32),
                ("img", p), ("img_bn", i32), ("img16", p), ("img16_bn", i32), ("alpha", C.c_float)]


class VsgNorm(C.Structure):
    _fields_ = [("gamma", p), ("beta", p)]


class VsgBigCEncLayer(C.Structure):
    _fields_ = [("qkv", VsgLinear), ("out", VsgLinear), ("l1", VsgLinear), ("l2", VsgLinear), ("n1", VsgNorm), ("n2", VsgNorm)]


class VsgBigCDecLayer(C.Structure):
    _fields_ = [(k, VsgLinear) for k in ("qk", "v", "out", "p2a", "e2a", "r1_0", "r1_1", "r2", "f1", "f2")] + \
               [(k, VsgNorm) for k in ("n1", "n2", "n3")] + [("r1_bias", p)]


VSG_MAX_LAYERS = 12


class VsgBigCWeights(C.Structure):
    _fields_ = [(k, i32) for k in ("variant", "dim_enti", "dim_pred", "dim_feat", "dim_clsme", "dim_i3d", "num_querys", "num_pred_cats",
                                   "num_enti_cats", "pool_len", "n_enc", "n_dec", "n_head", "use_clsme", "has_entiemb", "extra_width", "dim_z",
                                   "tc_attention")] + \
               [("bbox1_w", p), ("bbox1_b", p)] + \
               [(k, VsgLinear) for k in ("bbox2", "feat1", "feat2", "conv", "enco1", "enco2", "i3d", "log", "log1", "log2")] + \
               [("conv_b", p), ("enc", VsgBigCEncLayer * VSG_MAX_LAYERS), ("dec", VsgBigCDecLayer * VSG_MAX_LAYERS),
                ("pos", p), ("query_init", p), ("qk_init", p), ("bias_matrix", p), ("entiemb", p), ("role_fold", i32), ("eg_all", VsgLinear)]


class VsgVideoBatch(C.Structure):
    _fields_ = [("n_videos", i32), ("n_tracks", i32), ("max_tracks", i32), ("n_rows", i64), ("boxes", p), ("feats", p), ("ld_feats", i32),
                ("off", p), ("seg", p), ("seg64", p), ("tmax", p), ("track_vid", p), ("wh", p), ("dura", p), ("cat_ids", p), ("scores", p),
                ("mha_blk_seg", p), ("mha_blk_q0", p), ("n_mha_blk", i32), ("feats_bf16", i32)]


class VsgTripletOut(C.Structure):
    _fields_ = [("quint", p), ("scores", p), ("spans", p), ("qids", p), ("counts", p), ("cap", i32)]


class VsgGrdConv(C.Structure):
    _fields_ = [("dw_w", p), ("dw_b", p), ("k", i32), ("pw", VsgLinear)]


class VsgGrdEncoder(C.Structure):
    _fields_ = [("convs", VsgGrdConv * 4), ("qkv", VsgLinear), ("out", VsgLinear), ("fc", VsgLinear), ("normb", VsgNorm),
                ("norm_seq", VsgNorm * 4), ("norme", VsgNorm)]


class VsgGrdWeights(C.Structure):
    _fields_ = [("dim_hidden", i32), ("num_bins", i32), ("dim_feat", i32), ("tc_attention", i32), ("fuse_dwconv", i32),
                ("video_fc", VsgLinear), ("vq_fc", VsgLinear), ("proj2sim", VsgLinear), ("proj_enti", p), ("proj_pred", p),
                ("temp_w", p), ("temp_b", p), ("freq", p), ("phase", p),
                ("video_encoder", VsgGrdEncoder), ("query_encoder", VsgGrdEncoder), ("combined_encoder", VsgGrdEncoder),
                ("cls_head", VsgGrdConv * 5), ("conf_head", VsgGrdConv * 5), ("regr_head", VsgGrdConv * 5)]


class VsgGrdSeq(C.Structure):
    _fields_ = [("off", p), ("n", i32), ("rows", i64), ("pos", p), ("rem", p), ("max_len", i32), ("blk_seg", p), ("blk_q0", p), ("n_blk", i32),
                ("tc_blk_seg", p), ("tc_blk_q0", p), ("n_tc_blk", i32)]


class VsgGrdBatch(C.Structure):
    _fields_ = [("n_videos", i32), ("n_queries", i32), ("max_T", i32), ("video_feats", p), ("quint", p), ("spans", p), ("vlen", p),
                ("q_vid", p), ("clip_tab", p), ("video", VsgGrdSeq), ("query", VsgGrdSeq), ("combined", VsgGrdSeq)]


class VsgGrdOut(C.Structure):
    _fields_ = [("pooled", p), ("probs", p), ("mask", p), ("err_count", p), ("regr", p), ("conf", p), ("cls", p)]


class VsgError(RuntimeError):
    pass


# name -> (restype, argtypes); must list every symbol declared in include/vsg_b200.h
SIGNATURES = {
    "vsg_last_error": (C.c_char_p, []),
    "vsg_version": (i32, []),
    "vsg_built_for_sm": (i32, []),
    "vsg_device_sm_count": (i32, []),
    "vsg_launch_count": (C.c_longlong, []),
    "vsg_pair_ids": (i32, [i32, p, p]),
    "vsg_dura_intersection": (i32, [p, i32, p, i32, p, p, p]),
    "vsg_track_volumes": (i32, [p, p, i32, p, p]),
    "vsg_dura_intersection_ex": (i32, [p, i32, p, i32, i32, i32, p, p, p]),
    "vsg_traj_viou_matrix": (i32, [p, p, p, i32, p, p, p, i32, p, p, p, i32, i64, p, p, p, p, p, p, i32, p]),
    "vsg_traj_viou_matrix_tiled": (i32, [p, p, p, i32, p, p, p, i32, p, p, p, p, i32, i64, i64, p, p, p, p, p, p, p, p, p]),
    "vsg_pair_labels": (i32, [p, i32, i32, p, i32, f32, p, p]),
    "vsg_rel_viou_match": (i32, [C.POINTER(VsgRelTable), p, C.POINTER(VsgRelTable), i32, p, f64, p, p, p, p, p, p, p, p]),
    "vsg_eval_records_host": (i32, [p, p, p, p, p, p, i32, p, i32, p, i32, p]),
    "vsg_viou_pairs_f64": (i32, [p, p, p, p, p, p, i32, p, p]),
    "vsg_gemm": (i32, [i32, p, i32, p, p, i32, i32, i32, i32, p, p, p, i32, i32, i32, i32, p, i32, p, i32, p]),
    "vsg_gemm_ex": (i32, [C.POINTER(VsgGemmArgs), p]),
    "vsg_bipartite_cost": (i32, [p, i32, i32, p, i32, p, p, i32, f32, f32, p, p]),
    "vsg_bigc_workspace_bytes": (i64, [C.POINTER(VsgBigCWeights), C.POINTER(VsgVideoBatch), i32, i32]),
    "vsg_bigc_forward": (i32, [C.POINTER(VsgBigCWeights), C.POINTER(VsgVideoBatch), C.POINTER(VsgTripletOut), i32, i32, p, i64, p]),
    "vsg_grd_workspace_bytes": (i64, [C.POINTER(VsgGrdWeights), C.POINTER(VsgGrdBatch), i32]),
    "vsg_grd_forward": (i32, [C.POINTER(VsgGrdWeights), C.POINTER(VsgGrdBatch), C.POINTER(VsgGrdOut), f32, f32, f32, f32, i32, p, i64, p]),
    "vsg_pair_ids_batched": (i32, [p, i32, p, i64, p, p, p]),
    "vsg_pair_construct_triplet": (i32, [p, i32, i32, i32, p, p, i64, p, i32, p, p, p, p, p, i32, i32, p, p, p, p, p, p, p]),
    "vsg_tiou": (i32, [p, i32, p, i32, i32, i32, i32, p, p]),
    "vsg_stretch_rows": (i32, [p, i32, i32, p, i32, i32, p, p]),
    "vsg_unique_rows": (i32, [p, i32, i32, p, p, p, p]),
    "vsg_split_tf32": (i32, [p, p, p, i64, p]),
    "vsg_cast_bf16": (i32, [p, i64, i64, i32, p, i64, p]),
    "vsg_gemm_debug_flags": (i32, [i32]),
    "vsg_gemm_set_tma_store": (i32, [i32]),
    "vsg_gemm_set_cluster": (i32, [i32]),
    "vsg_gemm_tile_n": (i32, [i32]),
    "vsg_weight_image_bytes": (i64, [i32, i32, i32]),
    "vsg_build_weight_image": (i32, [p, i32, i32, i32, i32, p, p]),
    "vsg_weight_image_fp16_bytes": (i64, [i32, i32, i32]),
    "vsg_build_weight_image_fp16": (i32, [p, i32, i32, i32, i32, C.c_float, p, p]),
    "vsg_gemm_set_weight_image": (i32, [i32]),
    "vsg_split_bf16": (i32, [p, i32, i32, i32, p, p, i32, p]),
    "vsg_softmax_rows": (i32, [p, i32, i32, i64, f32, p]),
    "vsg_transpose_split": (i32, [p, i32, i64, i32, p, p, i64, p]),
    "vsg_gemm_set_store_hi": (i32, [i32]),
    "vsg_gemm_force_bn": (i32, [i32]),
    "vsg_bbox_feat_mlp1": (i32, [p, p, i32, i64, p, p, p, p, i32, p, i32, p, p]),
    "vsg_bbox_feat_mlp1_bf16": (i32, [p, p, i32, i64, p, p, p, p, i32, p, i32, p]),
    "vsg_conv_pool_bf16": (i32, [p, i32, i32, p, p, p, i32, i32, p, p]),
    "vsg_stretched_mean": (i32, [p, i32, i32, i32, p, p, i32, p, i32, p]),
    "vsg_stretched_mean_bf16": (i32, [p, i32, i32, i32, p, p, i32, p, i32, p]),
    "vsg_conv_pool": (i32, [p, i32, i32, p, p, p, i32, i32, p, p]),
    "vsg_add_layernorm": (i32, [p, i32, p, i32, p, p, p, i32, i64, i32, p, i32, p]),
    "vsg_add_layernorm_dual": (i32, [p, i32, p, i32, p, p, p, i32, i64, i32, p, i32, p, i32, p]),
    "vsg_broadcast_rows": (i32, [p, i32, i32, i64, p, p]),
    "vsg_mha": (i32, [p, i32, p, i32, p, i32, p, i32, i32, i32, i32, i32, p, i32, p, p, i32, p]),
    "vsg_mha_tc16": (i32, [p, i32, p, i32, p, i32, p, i32, p, i32, p, p, i32, i32, p]),
    "vsg_mha_tc16_set_kc": (i32, [i32]),
    "vsg_mha_tc64": (i32, [p, i32, p, i32, p, i32, p, i32, i32, i32, p, i32, p, p, i32, i32, p]),
    "vsg_role_attention": (i32, [p, p, p, p, i32, i32, i32, i32, f32, p, p, i32, p, p]),
    "vsg_role_attention_hid": (i32, [p, p, i32, p, i32, p, p, i32, i32, i32, i32, f32, p, p, i32, p, p]),
    "vsg_gather_concat": (i32, [C.POINTER(p), C.POINTER(p), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), i32, i64, p, i32, p]),
    "vsg_so_category": (i32, [p, p, i32, i64, p, p, p]),
    "vsg_seq_positions": (i32, [p, i32, i64, p, p, p]),
    "vsg_grd_query_init": (i32, [p, p, p, p, i32, p, p, p, p, i32, p, p, p]),
    "vsg_pos_add_ln": (i32, [p, p, p, p, p, p, i64, i32, p, p, p]),
    "vsg_dwconv": (i32, [p, p, p, p, p, i32, i64, i32, p, p]),
    "vsg_cq_attention": (i32, [p, p, p, p, p, p, i32, i32, i32, p, p]),
    "vsg_grounding_post": (i32, [p, p, p, p, p, p, p, p, i32, i32, i32, f32, f32, f32, f32, p, p, p, p, p]),
    "vsg_construct_triplet": (i32, [p, i32, i32, i32, i32, p, p, i32, p, p, p, p, p, p, p, p, i32, p]),
}


def lib() -> C.CDLL:
    """Load (once) and return the library; raises if it is missing -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VsgError("libvsgb200.so is not built (%s); run `python -m vidsgg_big_b200.build` -- "
                           "this package has no CPU fallback" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise VsgError("%s failed (%d): %s" % (what, rc, lib().vsg_last_error().decode()))


def ptr(t: Optional[torch.Tensor]):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise VsgError("expected a CUDA tensor (no CPU fallback); got device %s" % t.device)
    if not t.is_contiguous():
        raise VsgError("expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise VsgError("vidsgg_big_b200 runs on CUDA tensors only (no CPU fallback); got %s" % t.device)
