"""vidsgg_big_b200 -- B200-native (sm_100a) implementation of the per-video relation hot path of Dawn-LX/VidSGG-BIG.

Public names mirror the reference (models/__init__.py, VidVRDhelperEvalAPIs/__init__.py, utils/utils_func.py).
Importing this package never touches CUDA; the kernels live in libvsgb200.so and are bound lazily (no CPU fallback).
"""
from .containers import TrajProposal, VideoGraph                                         # noqa: F401
from .geometry import (dura_intersection_ts, vIoU_ts, trajid2pairid, traj_viou_matrix,   # noqa: F401
                       traj_viou_batched, enti_viou_align, pair_labels, TrackTable,
                       tIoU, generalized_tIoU, unique_with_idx_nd, stack_with_repeat_2d)
from .evalapi import (eval_visual_relation, evaluate, evaluate_v2, eval_relation_with_gt,  # noqa: F401
                      eval_detection_scores, eval_detection_scores_v2, eval_tagging_scores, viou, voc_ap,
                      PackedRelations, evaluate_packed)
from .bigc import BIG_C, BIG_C_vidvrd, BIG_C_vidor                                       # noqa: F401
from .basec import Base_C                                                                # noqa: F401
from .grounding import DEBUG, expand_after_grounding                                     # noqa: F401
from .convert import EvalFmtCvtor                                                        # noqa: F401
from . import driver                                                                     # noqa: F401

__all__ = ["TrajProposal", "VideoGraph", "dura_intersection_ts", "vIoU_ts", "trajid2pairid", "traj_viou_matrix",
           "traj_viou_batched", "enti_viou_align", "pair_labels", "TrackTable", "tIoU", "generalized_tIoU", "unique_with_idx_nd",
           "stack_with_repeat_2d", "eval_visual_relation", "evaluate",
           "evaluate_v2", "eval_relation_with_gt", "eval_detection_scores", "eval_detection_scores_v2",
           "eval_tagging_scores", "viou", "voc_ap", "PackedRelations", "evaluate_packed", "BIG_C", "BIG_C_vidvrd",
           "BIG_C_vidor", "Base_C", "DEBUG", "expand_after_grounding", "EvalFmtCvtor", "driver"]
