/*
 * vsg_b200.h -- C ABI of the B200-native VidSGG-BIG relation hot path (libvsgb200.so).
 *
 * The reference (Dawn-LX/VidSGG-BIG) is pure Python/PyTorch and has no FFI boundary; its
 * contract for this path is three Python call signatures.  This header
 * is the boundary the thin Python host layer (vidsgg_big_b200/*.py, ctypes) binds: plain
 * pointers and sizes, no torch types.  Each entry point names the reference code it replaces
 * (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - the caller owns every buffer including workspaces; nothing is allocated, nothing
 *     synchronises, no host callbacks: calls are asynchronous on `stream` and re-entrant
 *     per stream (the only exception is vsg_gemm_*'s tensor-map cache, guarded by a mutex);
 *   - return value 0 = launched, <0 = error (see VSG_E_*); vsg_last_error() returns a
 *     thread-local message for the last failing call;
 *   - spans are CLOSED [s, e] int64 on the model side (dataloaders/dataloader_vidvrd.py:33-34)
 *     and HALF-OPEN [s, e) on the evaluation side (VidVRDhelperEvalAPIs/README.md:7-47);
 *   - track storage is packed/CSR: boxes[sum L][4] f32 xyxy pixels, off[n+1] int64 row offsets.
 */
#ifndef VSG_B200_H_
#define VSG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSG_OK 0
#define VSG_E_INVALID (-1) /* bad argument (null pointer, negative size, misaligned buffer) */
#define VSG_E_LAUNCH (-2)  /* CUDA reported an error at launch                               */
#define VSG_E_UNSUPPORTED (-3)

/* ---- library ---------------------------------------------------------------------------- */
const char* vsg_last_error(void);
int vsg_version(void);          /* 100*major + minor */
int vsg_built_for_sm(void);     /* 100 (sm_100a)     */
int vsg_device_sm_count(void);  /* multiprocessors of the current device, <0 on error */
long long vsg_launch_count(void); /* kernels launched by this library in this process (monotonic) */

/* ---- geometry: pair enumeration, spans, trajectory vIoU ------ */

/* Ordered pairs (s,o), s != o, row-major: replaces Base_C.trajid2pairid
 * (models/model_pairwise_baseline.py:104-111, tools/train_vidor.py:73-78).
 * pair_ids_out: int64[n*(n-1)][2]. */
int vsg_pair_ids(int n, int64_t* pair_ids_out, void* stream);

/* Closed-span intersection (broadcast): replaces utils/utils_func.py:347-373
 * dura_intersection_ts(d1, d2, broadcast=True).  d1 int64[n1][2], d2 int64[n2][2];
 * inter_out int64[n1][n2][2] = (max start, min end) -- inverted for non-overlapping pairs,
 * exactly like the reference; mask_out uint8[n1][n2] = start<=end. */
int vsg_dura_intersection(const int64_t* d1, int n1, const int64_t* d2, int n2,
                          int64_t* inter_out, uint8_t* mask_out, void* stream);

/* Same for dtype 0 = int64 / 1 = float32 spans (the grounding stage intersects normalised float spans,
 * models/grd_model_v5.py:548) and for the element-wise mode broadcast=0 (n1 == n2, outputs [n1][2], [n1]). */
int vsg_dura_intersection_ex(const void* d1, int n1, const void* d2, int n2, int broadcast, int dtype,
                             void* inter_out, uint8_t* mask_out, void* stream);

/* Per-track volume sum_f (x2-x1+1)*(y2-y1+1) over the FULL track (utils/utils_func.py:459-460).
 * vol_out f32[n_tracks]. */
int vsg_track_volumes(const float* boxes, const int64_t* off, int n_tracks, float* vol_out, void* stream);

/* Batched trajectory-vIoU matrices, one per segment (= video):
 * replaces the per-pair Python loop around vIoU_ts
 *   models/model_0v10.py:565-581 (BIG_C.enti_viou_align), tools/train_vidor.py:107-122,
 *   utils/utils_func.py:437-471 (vIoU_ts) and :347-373 (spans).
 * Segment v pairs A tracks [segA[v], segA[v+1]) with B tracks [segB[v], segB[v+1]); its
 * nA_v x nB_v outputs start at seg_out[v] (row-major).  Outputs (any may be NULL):
 *   spans_out int64[P][2], mask_out uint8[P], viou_out f32[P] (0 where no temporal overlap),
 *   inter_out f32[P] (raw intersection volume; variant 1 only),
 * P = seg_out[n_seg].  volA / volB: f32 workspaces of n_tracks_A / n_tracks_B entries
 * (filled by this call; pass the same pointer twice when A and B are the same table).
 * variant: 0 = auto, 1 = warp-per-pair streaming kernel (2 = tiled: see vsg_traj_viou_matrix_tiled). */
int vsg_traj_viou_matrix(const float* boxesA, const int64_t* offA, const int64_t* duraA, int n_tracks_A,
                         const float* boxesB, const int64_t* offB, const int64_t* duraB, int n_tracks_B,
                         const int32_t* segA, const int32_t* segB, const int64_t* seg_out, int n_seg,
                         int64_t n_pairs_total,
                         int64_t* spans_out, uint8_t* mask_out, float* viou_out, float* inter_out,
                         float* volA, float* volB, int variant, void* stream);

/* Shared-memory tiled variant of vsg_traj_viou_matrix (same outputs; deterministic).  seg_len int64[n_seg] = frames on each
 * segment's absolute frame axis (video_len; every span must lie in [0, seg_len)).  Workspaces: jobs_ws (n_seg+1)*24 bytes,
 * njobs_ws int64[1], part_ws f32[n_jobs_bound*1024] with n_jobs_bound >= sum_v ceil(nA_v/32)*ceil(nB_v/32)*ceil(seg_len_v/256). */
int vsg_traj_viou_matrix_tiled(const float* boxesA, const int64_t* offA, const int64_t* duraA, int n_tracks_A,
                               const float* boxesB, const int64_t* offB, const int64_t* duraB, int n_tracks_B,
                               const int32_t* segA, const int32_t* segB, const int64_t* seg_out, const int64_t* seg_len,
                               int n_seg, int64_t n_pairs_total, int64_t n_jobs_bound,
                               int64_t* spans_out, uint8_t* mask_out, float* viou_out, float* volA, float* volB,
                               void* jobs_ws, int64_t* njobs_ws, float* part_ws, void* stream);

/* Base-C label assignment on top of a vIoU matrix: replaces the triple Python loop of
 * tools/train_vidor.py:143-159.  viou f32[n][n_gt_traj]; gt_so int64[n_gt_pred][2] (GT subject /
 * object track ids); labels_out uint8[n_gt_pred][n*(n-1)] in vsg_pair_ids order:
 * 1 iff viou[s][gs] > th && viou[o][go] > th. */
int vsg_pair_labels(const float* viou, int n, int n_gt_traj, const int64_t* gt_so, int n_gt_pred,
                    float th, uint8_t* labels_out, void* stream);

/* ---- evaluation: relation vIoU + greedy matching ---------------- */

/* One packed evaluation problem over n_vid videos.  A "relation" is a row
 * (triplet[3], sub_track, obj_track, start, end) with a HALF-OPEN duration; its two
 * trajectories are the slices [start, end) of tracks in a track table
 * (boxes[sum L][4], off[n_tracks+1], tstart[n_tracks] = first frame of each track).
 * box_f64 = 0: boxes are f32, 1: boxes are f64 (dict path, arbitrary Python floats).
 * All accumulation is f64 (the reference is Python float, common.py:65-106). */
typedef struct VsgRelTable {
  const void* boxes;        /* f32 or f64 [sum L][4]                                    */
  const int64_t* off;       /* [n_tracks+1]                                             */
  const int64_t* tstart;    /* [n_tracks]                                               */
  const int64_t* rel;       /* [n_rel][7] = s_cat, p_cat, o_cat, sub_track, obj_track, start, end */
  const int64_t* vid_off;   /* [n_vid+1] relation ranges per video                      */
  int64_t n_rel;
  int box_f64;
  int vol_full_track;       /* 1: volumes over the relation's whole tracks (dict path: every relation owns its two
                               lists, common.py:100-105); 0: over the [start,end) slice of shared tracks          */
} VsgRelTable;

/* Replaces eval_detection_scores / eval_detection_scores_v2
 * (VidVRDhelperEvalAPIs/visual_relation_detection.py:7-34, :124-156) and common.viou
 * (common.py:65-106) for all videos in one call:
 *   1. stable descending sort of each video's predictions by score  -> order_out int32[n_pred]
 *      (position k of video v holds the prediction index, local to the video, ranked k);
 *   2. ov[p][g] = min(viou(sub), viou(obj)) for equal triplets, f64   -> ov_ws f64[sum n_pred_v*n_gt_v]
 *      (ov_off int64[n_vid+1] gives each video's offset; -1 where triplets differ);
 *   3. greedy assignment in rank order: >= thr and strictly better than the running max,
 *      first GT wins ties, each GT consumed once
 *      -> hit_out f64[n_pred] in rank order (score or -inf), gt2det_out int32[n_gt] (rank or -1).
 * scores: f64[n_pred] (prediction scores, the sort key). */
int vsg_rel_viou_match(const VsgRelTable* pred, const double* scores, const VsgRelTable* gt, int n_vid,
                       const int64_t* ov_off, double thr,
                       int32_t* order_out, double* ov_ws, double* hit_out, int32_t* gt2det_out,
                       double* vol_pred_ws /* [n_pred][2] */, double* vol_gt_ws /* [n_gt][2] */,
                       uint8_t* taken_ws /* [n_gt + n_pred]: GT-taken flags, then per-prediction candidate flags */, void* stream);

/* HOST function (CPU pointers): per-video records (video index, AP, n_gt, TP@det_n..., P@tag_n...) from the matcher's outputs,
 * restating the numpy arithmetic of visual_relation_detection.py:28-33, :37-58, :82-93 and common.py:4-37 (voc_ap).
 * hit_host f64[n_pred] rank order, order_host int32[n_pred], *_trip_host int64[n][3], *_off_host int64[n_vid+1],
 * records_host f64[n_vid][3+n_det+n_tag].  Videos without GT are skipped; returns the number of records (<0 on error). */
int vsg_eval_records_host(const double* hit_host, const int32_t* order_host, const int64_t* pred_trip_host, const int64_t* p_off_host,
                          const int64_t* gt_trip_host, const int64_t* g_off_host, int n_vid, const int* det_n_host, int n_det,
                          const int* tag_n_host, int n_tag, double* records_host);

/* common.viou for a list of independent (traj_1, dur_1, traj_2, dur_2) problems, f64 boxes:
 * boxes1/boxes2 f64 [sum len][4], off1/off2 int64[n+1], dur1/dur2 int64[n][2] half-open;
 * out f64[n]. */
int vsg_viou_pairs_f64(const double* boxes1, const int64_t* off1, const int64_t* dur1,
                       const double* boxes2, const int64_t* off2, const int64_t* dur2,
                       int n, double* out, void* stream);

/* tIoU / generalized_tIoU of closed 1-D spans (utils/utils_func.py:375-410; models/grd_model_v5.py:18-33):
 * out f32[n1*n2] (broadcast) or f32[n1] (row-wise, n1 == n2) = (min(e1,e2) - max(s1,s2)) / (max(e1,e2) - min(s1,s2)); generalized = 0
 * additionally zeroes pairs whose spans do not touch.  dtype 0: int64 spans (true-divided as float32, like torch), 1: float32. */
int vsg_tiou(const void* d1, int n1, const void* d2, int n2, int broadcast, int generalized, int dtype, float* out, void* stream);

/* The stretch of stack_with_repeat_2d (models/model_0v10.py:18-46): out f32[n_tracks][tmax][width]; frame i of an L-frame track
 * (rows off[t] .. off[t+1] of src, leading dimension ld) is repeated ceil((tmax - i) / L) times.  The hot path never materialises
 * this tensor; the entry point exists for callers of the reference helper. */
int vsg_stretch_rows(const float* src, int ld, int width, const int64_t* off, int n_tracks, int tmax, float* out, void* stream);

/* unique_with_idx_nd (utils/utils_func.py:330-345) for int64 rows[n][d], n <= 8192: order int32[n] = original row indices sorted by
 * (row lexicographic, index); group int32[n] = id of the unique row of each sorted position; *n_groups = number of unique rows. */
int vsg_unique_rows(const int64_t* rows, int n, int d, int32_t* order, int32_t* group, int32_t* n_groups, void* stream);

/* ---- scoring: GEMM (kernels K4-K6) -------------------------------- */

#define VSG_GEMM_SIMT 0   /* fp32 FFMA tiles (comparator / shapes TMA cannot describe)          */
#define VSG_GEMM_TF32 1   /* tcgen05.mma kind::tf32, fp32 accumulate in TMEM                     */
#define VSG_GEMM_3XTF32 2 /* fp32-faithful 3xTF32 split on tcgen05 (needs W_lo from vsg_split_tf32) */
#define VSG_GEMM_TF32_BF16X2 3 /* fp32-class: tf32 main product + the two correction products as bf16 MMAs (kind::f16);
                                  needs W_b16 / W_lo16 from vsg_split_bf16; vsg_gemm_ex only, plain (non-batched) problems */
#define VSG_GEMM_BF16 4   /* reduced precision: bf16 operands (A16, W_b16), ONE kind::f16 pass, fp32 accumulate; outputs fp32 C and / or bf16 C16 */
#define VSG_GEMM_FP16X3 5 /* fp32-class at 3 tensor slots per MAC (mode 3: 4): A and W split into fp16 hi / lo pairs, three kind::f16 products,
                             power-of-two range scaling (W image pre-scaled, optional a_scale); needs W_img16 / w_alpha from
                             vsg_build_weight_image_fp16; 1e-2 <~ |A a_scale| < 65504; vsg_gemm_ex only, plain problems */

/* C[M][N] (ldc) = act( A[M][K] (lda) * W[N][K]^T (ldw) + bias[N] + rowbias[idx(row)][N] (+ C) ) (+ residual[M][N]),
 * fp32 row-major; the residual is added after the activation (QANet blocks, models/grd_model_v5.py:118-135).
 * Replaces every nn.Linear / 1x1 conv / conv tap of the hot path (models/model_0v10.py:446-458, :103-117,
 * :178-225, :478-507; models/grd_model_v5.py:331-373).  idx(row) = rb_index[row] if rb_index else row % rb_period
 * (used for the positional-embedding term of the decoder and the gathered frequency bias of the head).
 * W_lo is only read in mode VSG_GEMM_3XTF32.  Requirements for the tensor-core modes: lda, ldw multiples of 4,
 * A / W 16-byte aligned; otherwise the call silently uses the SIMT kernel (same result class). */
int vsg_gemm(int mode, const float* A, int lda, const float* W_hi, const float* W_lo, int ldw, int M, int N, int K,
             const float* bias, const float* rowbias, const int32_t* rb_index, int rb_period, int ld_rb, int relu,
             int accumulate, const float* residual, int ld_res, float* C, int ldc, void* stream);

/* Full-option form of vsg_gemm.  Extras:
 *   C_lo   optional second output: x - trunc_tf32(x) of every stored value (lets the NEXT 3xTF32 GEMM use this result as its
 *          pre-split "W" operand, e.g. K and V^T of an attention layer);
 *   batch  > 1: `batch` independent problems of extent M x N x K inside larger operands (attention QK^T / PV over (video, head)):
 *          problem p -> outer = p / batch_inner, inner = p % batch_inner; A rows / cols, W rows / cols and the C pointer are offset
 *          by outer * x_outer + inner * x_inner (rows, columns: elements; C: elements).  a_rows/a_cols/w_rows/w_cols = full operand
 *          extents (TMA bounds).  Requires a tensor-core mode, K %% 32 == 0, no bias / residual. */
typedef struct VsgGemmArgs {
  int mode;
  const float* A; int lda; int a_rows; int a_cols;
  const float* W_hi; const float* W_lo; int ldw; int w_rows; int w_cols;
  int M, N, K;
  const float* bias; const float* rowbias; const int32_t* rb_index; int rb_period; int ld_rb; int relu; int accumulate;
  const float* residual; int ld_res;
  float* C; float* C_lo; int ldc;
  int batch, batch_inner;
  int a_row_outer, a_row_inner, a_col_outer, a_col_inner;
  int b_row_outer, b_row_inner, b_col_outer, b_col_inner;
  long long c_outer, c_inner;
  int lo_col_begin, lo_col_end;                       /* C_lo is written for columns in [begin, end) only (multiples of 4); 0, 0 = every column */
  const void* W_b16; const void* W_lo16; int ldw16;   /* mode VSG_GEMM_TF32_BF16X2: bf16 [N][ldw16] copies of W (W_hi = the fp32 W) */
  const void* W_img; int img_bn;                      /* optional (mode 3): pre-swizzled tile images from vsg_build_weight_image built for
                                                         tile width img_bn over exactly this N and K (same k-block count); ignored unless img_bn == vsg_gemm_tile_n(N) */
  /* optional (mode 3, N <= 128): A is replaced by dwconv(A) computed inside the kernel -- DepthWiseSeparableConv1d.depth_wise of
   * models/grd_model_v5.py:36-56 followed by its point-wise conv as ONE launch: dwconv(A)[r][c] = dw_b[c] + sum_j dw_w[c][j] *
   * A[r + j - dw_k/2][c] with zero padding at the ends of the row's sequence (seq_pos[r] rows before r, seq_rem[r] rows after r in its
   * sequence).  Bit-identical to vsg_dwconv followed by vsg_gemm. */
  const float* dw_w; const float* dw_b; const int32_t* seq_pos; const int32_t* seq_rem; int dw_k;
  /* mode VSG_GEMM_BF16 (4): the A operand as bf16 [M][lda16] (lda16 a multiple of 8; written by a previous launch's C16 or by
   * vsg_cast_bf16) -- A / W_hi are not read; W_b16 [N][ldw16] is the weight.  C16 (any tensor-core mode of a plain problem): optional
   * bf16 copy [M][ldc16] of the stored values; in mode 4, C may then be NULL (bf16 output only). */
  const void* A16; int lda16;
  void* C16; int ldc16;
  /* mode VSG_GEMM_FP16X3 (5): fp16 weight-tile image of vsg_build_weight_image_fp16 (built for tile width img16_bn == vsg_gemm_tile_n(N)
   * over exactly this N and K) and w_alpha = 1 / scale of that image; A is the fp32 operand, W_hi the fp32 weight (only its extents are used). */
  const void* W_img16; int img16_bn; float w_alpha;
  float a_scale;            /* mode 5: 0 (= 1) or a power of two applied to A before its fp16 split (undone in the epilogue): keeps |A a_scale| inside
                               fp16's comfortable range [~1e-2, 65504) for layers whose activations are known to be tiny or huge */
} VsgGemmArgs;
int vsg_gemm_ex(const VsgGemmArgs* args, void* stream);

/* Validation knob for VSG_GEMM_3XTF32: 1 = the in-kernel split also rewrites the A tile with its masked high part;
 * 0 (default) relies on tcgen05 kind::tf32 ignoring the low 13 mantissa bits (checked bit-exact by tests). Returns the old value. */
int vsg_gemm_set_store_hi(int on);
/* Validation knob: 128 forces the 128x128-tile kernel for every shape, 0 (default) picks 128x256 tiles where N allows. */
int vsg_gemm_force_bn(int bn);
/* Validation knob: 0 = the epilogue writes C with per-row 16-byte stores only; 1 (default) = full 32x32 output slabs are staged in
 * shared memory (128B swizzle) and leave through cp.async.bulk.tensor stores.  Both give bit-identical C.  Returns the old value. */
int vsg_gemm_set_tma_store(int on);
/* Validation knob: 1 = one CTA per output tile everywhere; 2 = plain problems with M > 128 run as CTA pairs (thread-block cluster
 * of 2) that share an N tile and multicast each half of the W tile to both CTAs; 3 (default) = as 2, but the fp32-class modes with
 * 256-wide tiles issue ONE tcgen05.mma.cta_group::2 (M = 256) per pair, each CTA staging only its half of W (6 instead of 4 pipeline
 * stages).  Bit-identical C across 1 / 2 / 3.  Returns the old value. */
int vsg_gemm_set_cluster(int n);
/* Timing probes for bottleneck analysis (the results become garbage): bit 0 skip the W loads, bit 1 skip the A split,
 * bit 2 skip the MMAs, bit 3 skip the A loads; 0 (default) = normal operation.  Returns the old value. */
int vsg_gemm_debug_flags(int flags);

/* hi = w with the low 13 mantissa bits cleared (exactly representable in tf32), lo = w - hi. */
int vsg_split_tf32(const float* w, float* hi, float* lo, int64_t n, void* stream);
/* out[r][c] = bf16_rn(x[r][c]) for c < cols, 0 for cols <= c < ldo (ldo a multiple of 8, out 16-byte aligned): the A operand of
 * VSG_GEMM_BF16 when its producer wrote fp32 (reference counterpart: none -- models/model_0v10.py computes in fp32 throughout). */
int vsg_cast_bf16(const float* x, int64_t ldx, int64_t rows, int cols, void* out, int64_t ldo, void* stream);
/* bf16 operands of mode VSG_GEMM_TF32_BF16X2 for a weight w[rows][cols] (ldw): w16 = bf16_rn(w), lo16 = bf16_rn(w - trunc_tf32(w)),
 * both [rows][ld16] (ld16 >= cols, a multiple of 8; the padding columns are written as zeros). */
int vsg_split_bf16(const float* w, int ldw, int rows, int cols, void* w16, void* lo16, int ld16, void* stream);
/* Weight-tile images of mode VSG_GEMM_TF32_BF16X2: for every (N tile of `bn` rows, 16-column k block) one contiguous block holding the
 * shared-memory layout of the stage's three W operands (fp32 SWIZZLE_64B, bf16 and bf16-low SWIZZLE_32B; edges zero-padded), so that the
 * kernel's producer fetches them with contiguous bulk copies instead of 32/64-byte-row tensor loads.  vsg_gemm_tile_n(N) = the tile
 * width (128 / 256) vsg_gemm_ex will use for an N-column weight; vsg_weight_image_bytes = size of the image buffer (16-byte aligned). */
int vsg_gemm_tile_n(int N);
int64_t vsg_weight_image_bytes(int N, int K, int bn);
int vsg_build_weight_image(const float* w, int ldw, int N, int K, int bn, void* img, void* stream);
/* Validation knob: 0 = ignore W_img (tensor-map loads); 1 (default).  Bit-identical C.  Returns the old value. */
int vsg_gemm_set_weight_image(int on);
/* Weight-tile images of mode VSG_GEMM_FP16X3: per (N tile of `bn` rows, 16-column k block) two fp16 sub-images (bn rows x 32 B, SWIZZLE_32B,
 * edges zero-padded): hi = fp16_rn(w * scale), lo = fp16_rn(w * scale - hi).  `scale` must be a power of two; pick it so that max |w| * scale
 * lies in [2^13, 2^14) (fp16 overflows at 65504; the low part stays normal for |w| >= 2^-16 max |w|) and pass w_alpha = 1 / scale.
 * Reference counterpart: none (models/model_0v10.py multiplies in fp32); this is how the fp32 product is rebuilt on fp16 tensor cores. */
int64_t vsg_weight_image_fp16_bytes(int N, int K, int bn);
int vsg_build_weight_image_fp16(const float* w, int ldw, int N, int K, int bn, float scale, void* img, void* stream);

/* ---- BIG-C classification stage, non-GEMM kernels -------------------------
 * Batch layout: rows = all box-frames of all tracks of all videos; off int64[N+1]; seg int32[V+1] track range
 * per video; tmax int32[N] = longest track of the track's video; queries = V*Q rows, video-major. */

/* 8-d motion features [ctx,dctx,cty,dcty,w,dw,h,dh] (models/model_0v10.py:401-420) fused with the first
 * fc_bbox2enti layer (Linear 8->E + ReLU, :296-298).  wh f32[V][2]; W1 f32[E][8]; out f32[R][ldo];
 * feat8_out optional f32[R][8]. */
int vsg_bbox_feat_mlp1(const float* boxes, const int64_t* off, int n_tracks, int64_t n_rows, const int32_t* track_vid,
                       const float* wh, const float* W1, const float* b1, int E, float* out, int ldo, float* feat8_out,
                       void* stream);
/* Same, written as bf16 [R][ldo] (the A operand of the following VSG_GEMM_BF16 launch; models/model_0v10.py:296-298 in reduced precision). */
int vsg_bbox_feat_mlp1_bf16(const float* boxes, const int64_t* off, int n_tracks, int64_t n_rows, const int32_t* track_vid,
                            const float* wh, const float* W1, const float* b1, int E, void* out16, int ldo, void* stream);

/* Time-mean over the STRETCHED sequence of feature columns [col0, col0+width) (model_0v10.py:470,
 * model_0v7.py:473; stretch = stack_with_repeat_2d :18-46): out f32[N][ldo]. */
int vsg_stretched_mean(const float* feat, int ldf, int col0, int width, const int64_t* off, const int32_t* tmax,
                       int n_tracks, float* out, int ldo, void* stream);
/* Same over features stored as bf16 [R][ldf] (the opt-in bf16 feature transport of the bf16 mode). */
int vsg_stretched_mean_bf16(const void* feat16, int ldf, int col0, int width, const int64_t* off, const int32_t* tmax,
                            int n_tracks, float* out, int ldo, void* stream);

/* conv_feat2enti (k3,s2,p1 over stretched time) + adaptive_max_pool1d (model_0v10.py:450-457) from the three tap
 * products Y f32[R][3E] (tap-major); out f32[N][E*pool] flattened channel-major (c*pool+p). */
int vsg_conv_pool(const float* Y, int ldy, int E, const float* bias, const int64_t* off, const int32_t* tmax,
                  int n_tracks, int pool, float* out, void* stream);
/* Same with the tap products Y stored as bf16 [R][ldy] (written by a VSG_GEMM_BF16 launch's C16; model_0v10.py:450-457). */
int vsg_conv_pool_bf16(const void* Y16, int ldy, int E, const float* bias, const int64_t* off, const int32_t* tmax,
                       int n_tracks, int pool, float* out, void* stream);

/* out = LayerNorm(x + a)*gamma + beta (+ post[row % post_period])  (norm1/2/3 of model_0v10.py:111-115, :186-223;
 * the "+ pos" of :189 is the post term).  a, post may be NULL.  D % 32 == 0, D <= 1024. */
int vsg_add_layernorm(const float* x, int ldx, const float* a, int lda, const float* gamma, const float* beta,
                      const float* post, int post_period, int64_t rows, int D, float* out, int ldo, void* stream);
/* Dual form: out = LayerNorm(x + a), out2 = out + post[row % post_period] -- the decoder's last norm of a layer also emits the next layer's
 * q / k input (query + pos, models/model_0v10.py:181) so that the q/k projection needs no row-periodic bias in its epilogue. */
int vsg_add_layernorm_dual(const float* x, int ldx, const float* a, int lda, const float* gamma, const float* beta,
                           const float* post, int post_period, int64_t rows, int D, float* out, int ldo, float* out2, int ldo2,
                           void* stream);

/* out[r] = x[r % period]  (pred_query_init for every video, model_0v10.py:465). */
int vsg_broadcast_rows(const float* x, int period, int D, int64_t rows, float* out, void* stream);

/* softmax(Q K^T / sqrt(head_dim)) V per segment and head on packed rows (the core of nn.MultiheadAttention,
 * model_0v10.py:109, :183; grd_model_v5.py:100-108).  Segments: seg_off int64[n_seg+1] or fixed_len rows each.
 * Optional work list for ragged segments (avoids empty CTAs): blk_seg / blk_q0 int32[n_blocks] = segment and first query of
 * every 64-query block; NULL => a dense (n_seg x ceil(max_len/64)) grid. */
int vsg_mha(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const int64_t* seg_off, int n_seg,
            int fixed_len, int max_len, int n_head, int head_dim, float* O, int ldo, const int32_t* blk_seg,
            const int32_t* blk_q0, int n_blocks, void* stream);
/* The same attention for head_dim == 16 on the tensor cores (tcgen05.mma kind::tf32 + TMEM; csrc/attn_tc.cu): the grounding
 * network's nn.MultiheadAttention(128, 8) of models/grd_model_v5.py:90, :100-108, :128-130.  Work list as for vsg_mha but with blocks of
 * 128 queries: blk_seg[b] = sequence, blk_q0[b] = first query of block b; sequences are rows [seg_off[s], seg_off[s+1]) of Q / K / V / O.
 * products = 3: every product as lo*hi + hi*lo + hi*hi of tf32 splits (fp32-class, like VSG_GEMM_3XTF32); 1: single tf32 pass.
 * Softmax scale 1/sqrt(16); exact two-pass softmax(s - max). */
int vsg_mha_tc16(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const int64_t* seg_off,
                 int n_head, float* O, int ldo, const int32_t* blk_seg, const int32_t* blk_q0, int n_blk, int products,
                 void* stream);
/* Same for 64-wide heads, with fp16 hi / lo operand pairs on kind::f16 MMAs (products = 3: fp32-class, error <= 3e-6 vs fp64; 1: hi parts
 * only): the BIG-C decoder's self-attention over the num_querys queries of every video (nn.MultiheadAttention(512, 8) of
 * models/model_0v10.py:181-186) as ONE launch instead of two batched GEMMs + softmax + V^T transpose.  seg_off != NULL: ragged sequences
 * with the (sequence, first query) work list of 128-query blocks as for vsg_mha_tc16 (entries whose first query is not a multiple of 128 are
 * skipped, so the 64-query list of vsg_mha can be passed as is); seg_off == NULL: n_seg sequences of fixed_len rows back to back (the work
 * list is implicit).  |Q|, |K|, |V| < 65504. */
int vsg_mha_tc64(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const int64_t* seg_off, int n_seg,
                 int fixed_len, int n_head, float* O, int ldo, const int32_t* blk_seg, const int32_t* blk_q0, int n_blk, int products,
                 void* stream);
/* Tuning / validation knob: keys per block of vsg_mha_tc16 (32: 4 CTAs per SM, default; 64: 2 CTAs per SM).  Returns the old value. */
int vsg_mha_tc16_set_kc(int kc);

/* Glue of the tensor-core attention path (QK^T and PV are batched vsg_gemm_ex problems, one per (video, head)):
 * in-place row softmax of scale*S over the first n (<= 256) columns, and the transpose of an activation block
 * X[rows][cols] (ld) into T_hi[cols][rows] (ld_t) plus its tf32 low part T_lo (may be NULL). */
int vsg_softmax_rows(float* S, int ld, int n, int64_t rows, float scale, void* stream);
int vsg_transpose_split(const float* X, int ld, int64_t rows, int cols, float* T_hi, float* T_lo, int64_t ld_t, void* stream);

/* Role attention (model_0v10.py:190-214): att = softmax_tracks * softmax_roles of <p2a, e2a>/sqrt(dim_enti),
 * values f32[V*Q][2E] = att[r] @ enco.  Optional: att_out f32[V*Q][2][att_ld], so_out int32[V*Q][2] = per-role
 * argmax as GLOBAL track ids (prediction_head :485). */
int vsg_role_attention(const float* p2a, const float* e2a, const float* enco, const int32_t* seg, int n_vid, int Q, int E,
                       int max_tracks, float inv_sqrt_d, float* values, float* att_out, int att_ld, int32_t* so_out,
                       void* stream);
/* Role attention with the first fc_rolewise layer folded in (model_0v10.py:190-214): hid f32[V*Q][2E] = ReLU(att[r] @ G[r] + bias[r]) with
 * G f32[N][ld_g >= 2E] = enco @ [fc_rolewise.0.0.weight ; fc_rolewise.1.0.weight]^T computed once per TRACK (one GEMM), i.e.
 * fc_rolewise[r].0(att[r] @ enco) re-associated: the V*Q x 2E `values` rows and the two V*Q-row GEMMs on them disappear.  e2a f32[N][ld_e2a].
 * bias f32[2E] = the two first-layer biases.  att_out / so_out as for vsg_role_attention.  E in {128, 512}. */
int vsg_role_attention_hid(const float* p2a, const float* e2a, int ld_e2a, const float* G, int ld_g, const float* bias, const int32_t* seg,
                           int n_vid, int Q, int E, int max_tracks, float inv_sqrt_d, float* hid, float* att_out, int att_ld, int32_t* so_out,
                           void* stream);

/* Row-wise concat of up to 8 (optionally gathered) pieces: the prediction_head input (model_0v10.py:501/503). */
int vsg_gather_concat(const float* const* src_host, const int32_t* const* idx_host, const int* idx_stride_host,
                      const int* ld_host, const int* width_host, int n_pieces, int64_t rows, float* out, int ldo,
                      void* stream);

/* (s,o) track ids -> frequency-bias row scat*C+ocat and per-role category ids (model_0v10.py:486-487). */
int vsg_so_category(const int32_t* so, const int64_t* cat_ids, int C, int64_t rows, int32_t* pair_index, int32_t* so_cat,
                    void* stream);

/* construct_triplet (model_0v10.py:707-785) for every video: outputs at video v start at row v*cap;
 * counts int32[V][2] = (rows emitted, rows that passed the overlap filter; 0 => the reference returns None). */
int vsg_construct_triplet(const float* logits, int ld_logits, int P, int Q, int topk, const int32_t* so, const int32_t* seg,
                          int n_vid, const int64_t* dura, const int64_t* cat_ids, const float* enti_scores,
                          int64_t* quint, float* scores, int64_t* spans, int64_t* qids, int32_t* counts, int cap,
                          void* stream);

/* Cost matrix of the Hungarian matching of the training forward (models/model_0v10.py:606-636, f3 "then" clause of SURVEY 8f):
 * cost f32[Q][G] = c_cls * CE(logit[q], gt_pred[g]) + c_adj * mean over (2 roles, n tracklets) of BCE(att[role][q][:], adj[role][g][:])
 * (torch's BCE log clamp at -100).  logit f32[Q][P], gt_pred int64[G], att f32[2][Q][n], adj f32[2][G][n].  The assignment itself
 * (scipy linear_sum_assignment) stays on the host. */
int vsg_bipartite_cost(const float* logit, int Q, int P, const int64_t* gt_pred, int G, const float* att, const float* adj, int n,
                       float c_cls, float c_adj, float* cost, void* stream);

/* ---- Whole-forward entry points (last row): the BIG-C classification forward as ONE call -----------------------------
 * vsg_bigc_forward runs models/model_0v10.py:434-507 (+ :707-785) / models/model_0v7.py:483-513 for a packed batch of videos: the same
 * launch sequence the Python host layer (vidsgg_big_b200/bigc.py) issues, from C, so that a non-Python host can run the path and a
 * step is one call (capturable into one CUDA graph).  Nothing is allocated: every intermediate lives in the caller's workspace
 * (vsg_bigc_workspace_bytes), the outputs in the caller's VsgTripletOut buffers.  Results are bit-identical to the op-by-op path. */

/* One nn.Linear / 1x1 conv / conv-tap weight [N][K] in the layouts the GEMM modes read (see vsg_gemm_ex): fp32 `w` always; `hi`/`lo`
 * for VSG_GEMM_3XTF32; `w16`/`lo16` (+ optional `img`) for VSG_GEMM_TF32_BF16X2; `w16` for VSG_GEMM_BF16; `img16`/`alpha` for VSG_GEMM_FP16X3.  bias may be NULL. */
typedef struct VsgLinear {
  const float* w; const float* bias; int N, K, ldw;
  const float* hi; const float* lo;
  const void* w16; const void* lo16; int ld16;
  const void* img; int img_bn;
  const void* img16; int img16_bn; float alpha;   /* VSG_GEMM_FP16X3: fp16 tile image, its tile width, 1 / its scale */
} VsgLinear;

typedef struct VsgNorm { const float* gamma; const float* beta; } VsgNorm;

typedef struct VsgBigCEncLayer { VsgLinear qkv, out, l1, l2; VsgNorm n1, n2; } VsgBigCEncLayer;                 /* model_0v10.py:103-117 */
typedef struct VsgBigCDecLayer {                                                                                /* model_0v10.py:178-225 */
  VsgLinear qk, v, out, p2a, e2a, r1_0, r1_1, r2, f1, f2;    /* qk / v: rows [0,2P) / [2P,3P) of in_proj; r2: fc_rolewise.{0,1}.2 concatenated along K */
  VsgNorm n1, n2, n3;
  const float* r1_bias;     /* role_fold: [2P] = fc_rolewise.0.0.bias | fc_rolewise.1.0.bias */
} VsgBigCDecLayer;

#define VSG_MAX_LAYERS 12
typedef struct VsgBigCWeights {
  int variant;              /* 0 = model_0v10 (VidVRD), 1 = model_0v7 (VidOR) */
  int dim_enti, dim_pred, dim_feat, dim_clsme, dim_i3d /* 0 = none */, num_querys, num_pred_cats, num_enti_cats;
  int pool_len, n_enc, n_dec, n_head, use_clsme, has_entiemb, extra_width, dim_z;
  int tc_attention;         /* 1: decoder self-attention on vsg_mha_tc64 (64-wide heads; otherwise as 2); 2: as batched tcgen05 GEMMs + softmax /
                               transpose glue (needs head_dim % 32 == 0, Q % 32 == 0); 0: fp32 SIMT kernel */
  const float* bbox1_w; const float* bbox1_b;          /* fc_bbox2enti.0 [E][8], [E] */
  VsgLinear bbox2, feat1, feat2, conv /* tap-major [3E][2E], no bias */, enco1, enco2, i3d, log, log1, log2;
  const float* conv_b;
  VsgBigCEncLayer enc[VSG_MAX_LAYERS];
  VsgBigCDecLayer dec[VSG_MAX_LAYERS];
  const float* pos; const float* query_init; const float* qk_init /* query_init + pos */;
  const float* bias_matrix /* [C*C][P] */; const float* entiemb /* [C][dim_clsme] or NULL */;
  /* role_fold = 1 (dim_enti in {128, 512}): the track-side projections of EVERY decoder layer as one GEMM over the encoder output,
   * eg_all = [fc_enti2att ; fc_rolewise.0.0 ; fc_rolewise.1.0] of layer 0, 1, ... stacked along N (N = n_dec * 3E; bias = fc_enti2att's, zeros
   * for the fc_rolewise rows), consumed by vsg_role_attention_hid (fc_rolewise[r].0(att @ enco) re-associated as att @ (enco W^T)). */
  int role_fold;
  VsgLinear eg_all;
} VsgBigCWeights;

/* A packed batch of videos (vidsgg_big_b200.bigc.PackedVideos): rows = box-frames of all tracks of all videos. */
typedef struct VsgVideoBatch {
  int n_videos, n_tracks, max_tracks; int64_t n_rows;
  const float* boxes;          /* [R][4] */
  const float* feats; int ld_feats;   /* [R][>= dim_feat + extra] */
  const int64_t* off;          /* [N+1] row range of every track */
  const int32_t* seg;          /* [V+1] track range of every video */
  const int64_t* seg64;        /* the same as int64 (segment offsets of the encoder attention) */
  const int32_t* tmax;         /* [N] longest track of the track's video */
  const int32_t* track_vid;    /* [N] */
  const float* wh;             /* [V][2] */
  const int64_t* dura;         /* [N][2] closed spans */
  const int64_t* cat_ids;      /* [N] */
  const float* scores;         /* [N] */
  const int32_t* mha_blk_seg; const int32_t* mha_blk_q0; int n_mha_blk;   /* ragged (video, 64-track block) work list of vsg_mha */
  int feats_bf16;              /* 1 (mode VSG_GEMM_BF16 only): `feats` points to bf16 [R][ld_feats] (ld_feats a multiple of 8) -- the opt-in
                                  bf16 feature transport: half the H2D bytes, no cast pass */
} VsgVideoBatch;

typedef struct VsgTripletOut {   /* video v owns rows [v*cap, v*cap + counts[v][0]); cap = num_querys * topk */
  int64_t* quint; float* scores; int64_t* spans; int64_t* qids; int32_t* counts; int cap;
} VsgTripletOut;

/* Bytes of workspace vsg_bigc_forward needs for this batch / mode (256-byte aligned carve-up, peak over the forward). */
int64_t vsg_bigc_workspace_bytes(const VsgBigCWeights* w, const VsgVideoBatch* b, int topk, int precision_mode);
/* precision_mode = a VSG_GEMM_* mode.  workspace: device memory, 256-byte aligned.  Enqueues ~115 launches on `stream`; no sync. */
int vsg_bigc_forward(const VsgBigCWeights* w, const VsgVideoBatch* b, VsgTripletOut* out, int topk, int precision_mode,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* ---- Base-C pairwise baseline (models/model_pairwise_baseline.py) -----------------------------
 * The per-track encoding and the pair MLP are the BIG-C kernels + vsg_gemm; these two entry points are what is specific to it. */

/* trajid2pairid (:104-111) for every video of a batch: pair_off int64[V+1] = prefix of n_v (n_v - 1); so int32[n_pairs][2] receives
 * GLOBAL track ids (seg[v] + local id), ordered like nonzero() of the off-diagonal mask; pair_vid int32[n_pairs] = video of each pair. */
int vsg_pair_ids_batched(const int32_t* seg, int n_vid, const int64_t* pair_off, int64_t n_pairs, int32_t* so, int32_t* pair_vid,
                         void* stream);

/* construct_triplet (:314-395) for every video: softmax + top-k of each pair's logits, pairs whose closed spans do not overlap dropped,
 * rows in lexicographic quintuple order [pred, scat, ocat, sid, oid] (sid / oid = LOCAL track ids), background (pred 0) removed; when
 * rt_topk > 0 only the rt_topk best by mean(score3) (descending, ties in lexicographic order).  Video v owns output rows
 * cand_off[v] .. cand_off[v] + n_out(v), cand_off[v] = topk * pair_off[v]; counts int32[V][2] = {#valid candidates, #background among
 * them} -> n_out = counts[v][0] - counts[v][1] (min'ed with rt_topk), "None" in the reference when counts[v][0] == 0.
 * chunk_off int32[V+1] = prefix of ceil(candidates_v / 256) (host-built work list), n_chunks = chunk_off[V].
 * keys_ws u64[n_pairs*topk], score_ws f32[n_pairs*topk*3]: caller-owned workspace. */
int vsg_pair_construct_triplet(const float* logits, int ld_logits, int P, int topk, const int32_t* so, const int32_t* pair_vid,
                               int64_t n_pairs, const int32_t* seg, int n_vid, const int64_t* dura, const int64_t* cat_ids,
                               const float* enti_scores, const int64_t* cand_off, const int32_t* chunk_off, int n_chunks, int rt_topk,
                               unsigned long long* keys_ws, float* score_ws, int32_t* counts, int64_t* quint, float* scores,
                               int64_t* spans, void* stream);

/* ---- grounding stage, non-GEMM kernels (models/grd_model_v5.py) -----------
 * Row-major [rows][H] over ragged sequences seq_off int64[n_seq+1] (video clips / 3 query words / clips per query). */

/* pos[r] = index of row r inside its sequence, rem[r] = rows left after it. */
int vsg_seq_positions(const int64_t* seq_off, int n_seq, int64_t rows, int32_t* pos, int32_t* rem, void* stream);

/* prepare_data + query_fc + temp_fc (:310-328, :337-339) with the word embeddings pre-projected:
 * out f32[3*nq][H] (sub, pred, obj word of each query), so_norm f32[nq][2] = span / video_len. */
int vsg_grd_query_init(const int64_t* quint, const int64_t* spans, const float* vlen, const int32_t* q_vid, int nq,
                       const float* proj_enti, const float* proj_pred, const float* Wt, const float* bt, int H, float* out,
                       float* so_norm, void* stream);

/* PosEncoder (:58-78) + normb (:116-118): res = x + sin(pos*freq+phase), out = LayerNorm(res). */
int vsg_pos_add_ln(const float* x, const int32_t* pos, const float* freq, const float* phase, const float* gamma,
                   const float* beta, int64_t rows, int H, float* res, float* out, void* stream);

/* Depthwise Conv1d (groups = channels, pad k/2) along each sequence (:44); w f32[H][k]. */
int vsg_dwconv(const float* in, const int32_t* pos, const int32_t* rem, const float* w, const float* b, int k, int64_t rows,
               int H, float* out, void* stream);

/* Context-query attention (:345-363), (S_r S_c^T) V re-associated through the 3 query words:
 * out f32[sum nq_v*T_v][4H] = [v, A, A*v, B*v]. */
int vsg_cq_attention(const float* v, const float* pv, const float* q, const int64_t* vid_off, const int32_t* q_vid,
                     const int64_t* comb_off, int nq, int H, int max_T, float* out, void* stream);

/* Post-network part of _forward_test_single (:533-576) incl. temporal_pooling (:697-737) and the multi-bin NMS (:667-695):
 * pooled f32[nq][B+1][2], probs f32[nq][B+1], mask u8[nq][B+1]; err_count += #(query,bin) with an empty pooling set
 * (the reference raises there).  clip_tab f32[sum T_v] = torch.linspace(0,1,T_v) per video. */
int vsg_grounding_post(const float* regr, const float* conf, const float* cls, const int64_t* comb_off, const float* so_norm,
                       const float* clip_tab, const int64_t* vid_off, const int32_t* q_vid, int nq, int num_bins,
                       int regr_activated /* 1: regr already passed through the sigmoid */, float score_th,
                       float tiou_th, float bins_th, float nms_th, float* pooled, float* probs, uint8_t* mask, int32_t* err_count,
                       void* stream);

/* ---- grounding forward as ONE call (models/grd_model_v5.py:331-373 network + :530-576 post-processing) --------------------------------
 * The launch sequence of vidsgg_big_b200/grounding.py:_forward_videos for a batch of videos with their classified queries; the
 * data-dependent index tables (sequence offsets / positions, attention work lists, the per-video linspace clip table) are built by the
 * caller (host arithmetic on the per-video T and query counts) and passed in VsgGrdBatch.  Bit-identical to the op-by-op path. */
typedef struct VsgGrdConv { const float* dw_w /* [C][k] */; const float* dw_b; int k; VsgLinear pw; } VsgGrdConv;   /* DepthWiseSeparableConv1d :36-56 */
typedef struct VsgGrdEncoder {                                                                                    /* QANetEncoderLayer :81-137 */
  VsgGrdConv convs[4]; VsgLinear qkv, out, fc; VsgNorm normb, norm_seq[4], norme;
} VsgGrdEncoder;
typedef struct VsgGrdWeights {
  int dim_hidden, num_bins, dim_feat;
  int tc_attention;           /* 1: sequences of mean length >= 24 attend on vsg_mha_tc16 */
  int fuse_dwconv;            /* 1: depthwise conv inside the point-wise GEMM (mode VSG_GEMM_TF32_BF16X2) */
  VsgLinear video_fc, vq_fc, proj2sim;
  const float* proj_enti; const float* proj_pred;      /* word embeddings pre-projected through query_fc: [81][H], [51][H] */
  const float* temp_w; const float* temp_b; const float* freq; const float* phase;
  VsgGrdEncoder video_encoder, query_encoder, combined_encoder;
  VsgGrdConv cls_head[5], conf_head[5], regr_head[5];
} VsgGrdWeights;

typedef struct VsgGrdSeq {      /* one family of ragged sequences over packed rows */
  const int64_t* off; int n; int64_t rows; const int32_t* pos; const int32_t* rem; int max_len;
  const int32_t* blk_seg; const int32_t* blk_q0; int n_blk;          /* vsg_mha work list (64-query blocks) */
  const int32_t* tc_blk_seg; const int32_t* tc_blk_q0; int n_tc_blk; /* vsg_mha_tc16 work list (128-query blocks); n_tc_blk < 0: not used */
} VsgGrdSeq;

typedef struct VsgGrdBatch {
  int n_videos, n_queries, max_T;
  const float* video_feats;     /* [sum T][dim_feat] */
  const int64_t* quint;         /* [NQ][5] */
  const int64_t* spans;         /* [NQ][2] */
  const float* vlen;            /* [n_videos] video_len as float */
  const int32_t* q_vid;         /* [NQ] video of every query */
  const float* clip_tab;        /* [sum T] torch.linspace(0, 1, T_v) per video (:705) */
  VsgGrdSeq video, query, combined;   /* sequences of clips per video / 3 words per query / clips per query */
} VsgGrdBatch;

typedef struct VsgGrdOut {
  float* pooled;    /* [NQ][B+1][2] */
  float* probs;     /* [NQ][B+1] */
  uint8_t* mask;    /* [NQ][B+1] */
  int32_t* err_count;   /* [1], zero-initialised by the caller: #(query, bin) with an empty pooling set (the reference raises there) */
  float* regr; float* conf; float* cls;   /* optional: raw head outputs [rows_c][2B], [rows_c][B], [rows_c][B] (NULL: kept in the workspace) */
} VsgGrdOut;

int64_t vsg_grd_workspace_bytes(const VsgGrdWeights* w, const VsgGrdBatch* b, int precision_mode);
int vsg_grd_forward(const VsgGrdWeights* w, const VsgGrdBatch* b, VsgGrdOut* out, float score_th, float tiou_th, float bins_th, float nms_th,
                    int precision_mode, void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VSG_B200_H_ */
