// vsg_bigc_forward: the BIG-C classification forward (models/model_0v10.py:434-507 + :707-785; VidOR twin models/model_0v7.py:483-513)
// for a packed batch of videos as ONE C call.  It issues exactly the launch sequence of the Python host layer
// (vidsgg_big_b200/bigc.py: _track_encoding, _encode2decode, _prediction_head, _construct_triplets), so its results are bit-identical to
// the op-by-op path; every intermediate lives in a caller-provided workspace carved by a bump allocator with stack discipline
// (the R-row front-chain buffers are released before the encoder / decoder buffers are placed).
#include "common.cuh"
#include <math.h>
#include <string.h>

namespace vsg {

struct Arena {
  uint8_t* base;
  int64_t cap, off, peak;
  bool dry;                      // sizing pass: hand out null pointers, launch nothing
  bool overflow;
  template <typename T> T* get(int64_t count) {
    off = (off + 255) & ~(int64_t)255;
    T* p = dry ? nullptr : reinterpret_cast<T*>(base + off);
    off += count * (int64_t)sizeof(T);
    if (off > peak) peak = off;
    if (!dry && off > cap) overflow = true;
    return p;
  }
  int64_t mark() const { return off; }
  void release(int64_t m) { off = m; }
};

struct FwdBase {
  int mode;
  Arena ar;
  void* stream;
  int rc;

  bool live() const { return !ar.dry && rc == VSG_OK && !ar.overflow; }
#define FWD_CALL(expr) do { if (live()) { int rc_ = (expr); if (rc_ != VSG_OK) rc = rc_; } } while (0)

  // C[:, :N] = act(A[:, :K] W^T + bias (+ rowbias[idx])) (+ residual) -- the argument conventions of linalg.gemm
  void gemm(const float* A, int lda, const VsgLinear& W, float* C, int ldc, int64_t M, bool relu = false, bool use_bias = true, int K = -1,
            const float* rowbias = nullptr, const int32_t* rb_index = nullptr, int ld_rb = 0, float* C_lo = nullptr, int lo0 = 0, int lo1 = 0,
            const void* A16 = nullptr, int lda16 = 0, void* C16 = nullptr, int ldc16 = 0, const float* residual = nullptr, int ld_res = 0,
            const VsgGrdConv* dw = nullptr, const int32_t* seq_pos = nullptr, const int32_t* seq_rem = nullptr) {
    VsgGemmArgs a;
    memset(&a, 0, sizeof(a));
    if (K < 0) K = W.K;
    a.mode = mode; a.A = A; a.lda = lda; a.ldw = W.ldw; a.M = (int)M; a.N = W.N; a.K = K;
    a.W_hi = (mode == VSG_GEMM_3XTF32 && W.hi) ? W.hi : W.w;
    a.W_lo = mode == VSG_GEMM_3XTF32 ? W.lo : nullptr;
    a.bias = use_bias ? W.bias : nullptr;
    a.rowbias = rowbias; a.rb_index = rb_index; a.ld_rb = ld_rb;
    a.relu = relu ? 1 : 0;
    a.residual = residual; a.ld_res = ld_res;
    if (dw) { a.dw_w = dw->dw_w; a.dw_b = dw->dw_b; a.dw_k = dw->k; a.seq_pos = seq_pos; a.seq_rem = seq_rem; }
    a.C = C; a.ldc = ldc; a.C_lo = C_lo; a.lo_col_begin = lo0; a.lo_col_end = lo1;
    a.batch = 1; a.batch_inner = 1;
    if (mode == VSG_GEMM_TF32_BF16X2) {
      a.W_b16 = W.w16; a.W_lo16 = W.lo16; a.ldw16 = W.ld16;
      if (W.img && K == W.K) { a.W_img = W.img; a.img_bn = W.img_bn; }
    }
    if (mode == VSG_GEMM_FP16X3) { a.W_img16 = W.img16; a.img16_bn = W.img16_bn; a.w_alpha = W.alpha; }
    if (mode == VSG_GEMM_BF16) {
      a.W_b16 = W.w16; a.ldw16 = W.ld16;
      if (A16) { a.A16 = A16; a.lda16 = lda16; }
      else {                                   // the producer wrote fp32: cast first (linalg.cast_bf16)
        const int ld = (K + 7) / 8 * 8;
        const int64_t m0 = ar.mark();
        void* tmp = ar.get<uint16_t>(M * ld);
        FWD_CALL(vsg_cast_bf16(A, lda, M, K, tmp, ld, stream));
        a.A16 = tmp; a.lda16 = ld;
        a.C16 = C16; a.ldc16 = ldc16;
        FWD_CALL(vsg_gemm_ex(&a, stream));
        ar.release(m0);
        return;
      }
      a.C16 = C16; a.ldc16 = ldc16;
    }
    FWD_CALL(vsg_gemm_ex(&a, stream));
  }

  void add_ln(const float* x, const float* a, const VsgNorm& n, const float* post, int period, int64_t rows, int D, float* out, float* out2 = nullptr) {
    if (out2) FWD_CALL(vsg_add_layernorm_dual(x, D, a, a ? D : 0, n.gamma, n.beta, post, period, rows, D, out, D, out2, D, stream));
    else FWD_CALL(vsg_add_layernorm(x, D, a, a ? D : 0, n.gamma, n.beta, post, period, rows, D, out, D, stream));
  }
};

struct Fwd : FwdBase {
  const VsgBigCWeights* w;
  const VsgVideoBatch* b;

  // bigc.BIG_C._mha_tc: S = Q K^T and O = P V as batched tcgen05 GEMMs (one problem per (segment, head)) + softmax / V^T glue
  void mha_tc(const float* qkv, const float* qkv_lo, int n_seg, int Q, int d, float* att) {
    const int H = w->n_head, dh = d / H;
    const int64_t rows = (int64_t)n_seg * Q;
    const int m = mode == VSG_GEMM_BF16 ? VSG_GEMM_TF32 : ((mode == VSG_GEMM_TF32_BF16X2 || mode == VSG_GEMM_FP16X3) ? VSG_GEMM_3XTF32 : mode);   // linalg.attention_mode
    float* S = ar.get<float>(rows * H * Q);
    VsgGemmArgs a;
    memset(&a, 0, sizeof(a));
    a.mode = m; a.A = qkv; a.lda = 3 * d; a.a_rows = (int)rows; a.a_cols = d;
    a.W_hi = qkv + d; a.W_lo = qkv_lo ? qkv_lo + d : nullptr; a.ldw = 3 * d; a.w_rows = (int)rows; a.w_cols = d;
    a.M = Q; a.N = Q; a.K = dh; a.C = S; a.ldc = Q; a.batch = n_seg * H; a.batch_inner = H;
    a.a_row_outer = Q; a.a_col_inner = dh; a.b_row_outer = Q; a.b_col_inner = dh;
    a.c_outer = (long long)H * Q * Q; a.c_inner = (long long)Q * Q;
    FWD_CALL(vsg_gemm_ex(&a, stream));
    FWD_CALL(vsg_softmax_rows(S, Q, Q, rows * H, 1.0f / sqrtf((float)dh), stream));
    const int64_t ldt = (rows + 3) / 4 * 4;
    float* vt_hi = ar.get<float>((int64_t)d * ldt);
    float* vt_lo = m == VSG_GEMM_3XTF32 ? ar.get<float>((int64_t)d * ldt) : nullptr;
    FWD_CALL(vsg_transpose_split(qkv + 2 * d, 3 * d, rows, d, vt_hi, vt_lo, ldt, stream));
    memset(&a, 0, sizeof(a));
    a.mode = m; a.A = S; a.lda = Q; a.a_rows = (int)(rows * H); a.a_cols = Q;
    a.W_hi = vt_hi; a.W_lo = vt_lo; a.ldw = (int)ldt; a.w_rows = d; a.w_cols = (int)rows;
    a.M = Q; a.N = dh; a.K = Q; a.C = att; a.ldc = d; a.batch = n_seg * H; a.batch_inner = H;
    a.a_row_outer = H * Q; a.a_row_inner = Q; a.b_row_inner = dh; a.b_col_outer = Q;
    a.c_outer = (long long)Q * d; a.c_inner = dh;
    FWD_CALL(vsg_gemm_ex(&a, stream));
  }

  int run(VsgTripletOut* out, int topk);
};

int Fwd::run(VsgTripletOut* out, int topk) {
  const int E = w->dim_enti, Pd = w->dim_pred, Q = w->num_querys, F_in = w->dim_feat, H = w->n_head;
  const int V = b->n_videos, N = b->n_tracks;
  const int64_t R = b->n_rows;
  const int64_t VQ = (int64_t)V * Q;
  const bool bf16 = mode == VSG_GEMM_BF16;

  // ---------------- track encoding (bigc._track_encoding / _track_encoding_bf16) ----------------
  float* enti2enco = ar.get<float>((int64_t)N * E);
  float* extra = w->extra_width ? ar.get<float>((int64_t)N * w->extra_width) : nullptr;
  float* pooled = ar.get<float>((int64_t)N * E * w->pool_len);
  {
    const int64_t m0 = ar.mark();
    if (!bf16) {
      float* X = ar.get<float>(R * 2 * E);
      float* h = ar.get<float>(R * E);
      FWD_CALL(vsg_bbox_feat_mlp1(b->boxes, b->off, N, R, b->track_vid, b->wh, w->bbox1_w, w->bbox1_b, E, h, E, nullptr, stream));
      gemm(h, E, w->bbox2, X, 2 * E, R, true);
      gemm(b->feats, b->ld_feats, w->feat1, h, E, R, true, true, F_in);
      gemm(h, E, w->feat2, X + E, 2 * E, R, true);
      float* Y = ar.get<float>(R * 3 * E);
      gemm(X, 2 * E, w->conv, Y, 3 * E, R, false, false);
      FWD_CALL(vsg_conv_pool(Y, 3 * E, E, w->conv_b, b->off, b->tmax, N, w->pool_len, pooled, stream));
    } else {
      uint16_t* X16 = ar.get<uint16_t>(R * 2 * E);
      uint16_t* h16 = ar.get<uint16_t>(R * E);
      FWD_CALL(vsg_bbox_feat_mlp1_bf16(b->boxes, b->off, N, R, b->track_vid, b->wh, w->bbox1_w, w->bbox1_b, E, h16, E, stream));
      gemm(nullptr, 0, w->bbox2, nullptr, 0, R, true, true, -1, nullptr, nullptr, 0, nullptr, 0, 0, h16, E, X16, 2 * E);
      if (b->feats_bf16) {                       // bf16 feature transport: the features ARE the A operand
        gemm(nullptr, 0, w->feat1, nullptr, 0, R, true, true, F_in, nullptr, nullptr, 0, nullptr, 0, 0, b->feats, b->ld_feats, h16, E);
      } else {
        const int ldf = (F_in + 7) / 8 * 8;
        const int64_t m1 = ar.mark();
        uint16_t* F16 = ar.get<uint16_t>(R * ldf);
        FWD_CALL(vsg_cast_bf16(b->feats, b->ld_feats, R, F_in, F16, ldf, stream));
        gemm(nullptr, 0, w->feat1, nullptr, 0, R, true, true, F_in, nullptr, nullptr, 0, nullptr, 0, 0, F16, ldf, h16, E);
        ar.release(m1);
      }
      gemm(nullptr, 0, w->feat2, nullptr, 0, R, true, true, -1, nullptr, nullptr, 0, nullptr, 0, 0, h16, E, X16 + E, 2 * E);
      const int ldy = (3 * E + 7) / 8 * 8;
      uint16_t* Y16 = ar.get<uint16_t>(R * ldy);
      gemm(nullptr, 0, w->conv, nullptr, 0, R, false, false, -1, nullptr, nullptr, 0, nullptr, 0, 0, X16, 2 * E, Y16, ldy);
      FWD_CALL(vsg_conv_pool_bf16(Y16, ldy, E, w->conv_b, b->off, b->tmax, N, w->pool_len, pooled, stream));
    }
    ar.release(m0);
  }
  {
    float* t = ar.get<float>((int64_t)N * E);
    gemm(pooled, E * w->pool_len, w->enco1, t, E, N, true);
    gemm(t, E, w->enco2, enti2enco, E, N, true);
  }
  if (w->extra_width) {
    if (b->feats_bf16) FWD_CALL(vsg_stretched_mean_bf16(b->feats, b->ld_feats, F_in, w->extra_width, b->off, b->tmax, N, extra, w->extra_width, stream));
    else FWD_CALL(vsg_stretched_mean(b->feats, b->ld_feats, F_in, w->extra_width, b->off, b->tmax, N, extra, w->extra_width, stream));
  }

  // ---------------- encoder (post-norm, tokens = tracks of a video) ----------------
  const float* x = enti2enco;
  {
    float* qkv = ar.get<float>((int64_t)N * 3 * E);
    float* att = ar.get<float>((int64_t)N * E);
    float* t1 = ar.get<float>((int64_t)N * E);
    float* t2 = w->n_enc ? ar.get<float>((int64_t)N * w->enc[0].l1.N) : nullptr;
    for (int li = 0; li < w->n_enc; ++li) {
      const VsgBigCEncLayer& lw = w->enc[li];
      float* xa = ar.get<float>((int64_t)N * E);      // per-layer outputs (N x E floats each: small next to the R-row buffers)
      float* xb = ar.get<float>((int64_t)N * E);
      gemm(x, E, lw.qkv, qkv, 3 * E, N);
      if (w->tc_attention == 1 && mode != VSG_GEMM_SIMT && E / H == 64) {      // ragged tracks-per-video sequences on the fused tcgen05 kernel
        const int products = (mode == VSG_GEMM_3XTF32 || mode == VSG_GEMM_TF32_BF16X2 || mode == VSG_GEMM_FP16X3) ? 3 : 1;
        FWD_CALL(vsg_mha_tc64(qkv, 3 * E, qkv + E, 3 * E, qkv + 2 * E, 3 * E, b->seg64, V, 0, H, att, E, b->mha_blk_seg, b->mha_blk_q0, b->n_mha_blk,
                              products, stream));
      } else {
        FWD_CALL(vsg_mha(qkv, 3 * E, qkv + E, 3 * E, qkv + 2 * E, 3 * E, b->seg64, V, 0, b->max_tracks, H, E / H, att, E, b->mha_blk_seg,
                         b->mha_blk_q0, b->n_mha_blk, stream));
      }
      gemm(att, E, lw.out, t1, E, N);
      add_ln(x, t1, lw.n1, nullptr, 0, N, E, xa);
      gemm(xa, E, lw.l1, t2, lw.l1.N, N, true);
      gemm(t2, lw.l1.N, lw.l2, t1, E, N);
      add_ln(xa, t1, lw.n2, nullptr, 0, N, E, xb);
      x = xb;
    }
  }
  const float* enco = x;

  // ---------------- decoder ----------------
  int32_t* so = ar.get<int32_t>(VQ * 2);
  const bool fold = w->role_fold && (E == 128 || E == 512) && w->n_dec > 0;
  float* values = fold ? nullptr : ar.get<float>(VQ * 2 * E);
  float* hid = ar.get<float>(VQ * 2 * Pd);
  float* EG = nullptr;                             // [N][n_dec * 3E]: per layer [e2a | G_subject | G_object] (bigc.py role_fold)
  const int ld_eg = w->n_dec * 3 * E;
  if (fold) {
    EG = ar.get<float>((int64_t)N * ld_eg);
    gemm(enco, E, w->eg_all, EG, ld_eg, N);
  }
  float* qkv = ar.get<float>(VQ * 3 * Pd);
  const bool use_fused = w->tc_attention == 1 && mode != VSG_GEMM_SIMT && Pd / H == 64;      // vsg_mha_tc64: S never leaves the SM
  const bool use_tc = !use_fused && w->tc_attention && mode != VSG_GEMM_SIMT && (Pd / H) % 32 == 0 && Q % 32 == 0;
  const bool need_lo = use_tc && (mode == VSG_GEMM_3XTF32 || mode == VSG_GEMM_TF32_BF16X2 || mode == VSG_GEMM_FP16X3);
  float* qkv_lo = need_lo ? ar.get<float>(VQ * 3 * Pd) : nullptr;
  float* att = ar.get<float>(VQ * Pd);
  float* t1 = ar.get<float>(VQ * Pd);
  float* xq = ar.get<float>(VQ * Pd);            // norm1 output (+ pos)
  float* p2a = ar.get<float>(VQ * w->dec[0].p2a.N);
  float* p2a_b = ar.get<float>(VQ * w->dec[0].p2a.N);
  float* e2a = ar.get<float>((int64_t)N * w->dec[0].e2a.N);
  float* query = ar.get<float>(VQ * Pd);
  float* query2 = ar.get<float>(VQ * Pd);
  float* query_pos = ar.get<float>(VQ * Pd);
  float* ffn1 = ar.get<float>(VQ * w->dec[0].f1.N);
  const float* cur_q = nullptr;                  // current queries [VQ][Pd]
  for (int li = 0; li < w->n_dec; ++li) {
    const VsgBigCDecLayer& lw = w->dec[li];
    const bool last = li == w->n_dec - 1;
    // layer 0: the queries of every video are still pred_query_init, so its self-attention block and fc_pred2att are
    // video-independent -- computed once on Q rows, then broadcast
    const float* xin = li == 0 ? w->query_init : cur_q;
    const float* x_qk = li == 0 ? w->qk_init : query_pos;
    const int nv = li == 0 ? 1 : V;
    const int64_t rows = (int64_t)nv * Q;
    const int64_t mk = ar.mark();
    if (use_fused) {
      gemm(x_qk, Pd, lw.qk, qkv, 3 * Pd, rows);
      gemm(xin, Pd, lw.v, qkv + 2 * Pd, 3 * Pd, rows);
      const int products = (mode == VSG_GEMM_3XTF32 || mode == VSG_GEMM_TF32_BF16X2 || mode == VSG_GEMM_FP16X3) ? 3 : 1;
      FWD_CALL(vsg_mha_tc64(qkv, 3 * Pd, qkv + Pd, 3 * Pd, qkv + 2 * Pd, 3 * Pd, nullptr, nv, Q, H, att, Pd, nullptr, nullptr, 0, products, stream));
    } else if (use_tc) {
      gemm(x_qk, Pd, lw.qk, qkv, 3 * Pd, rows, false, true, -1, nullptr, nullptr, 0, qkv_lo, qkv_lo ? Pd : 0, qkv_lo ? 2 * Pd : 0);
      gemm(xin, Pd, lw.v, qkv + 2 * Pd, 3 * Pd, rows);
      mha_tc(qkv, qkv_lo, nv, Q, Pd, att);
    } else {
      gemm(x_qk, Pd, lw.qk, qkv, 3 * Pd, rows);
      gemm(xin, Pd, lw.v, qkv + 2 * Pd, 3 * Pd, rows);
      FWD_CALL(vsg_mha(qkv, 3 * Pd, qkv + Pd, 3 * Pd, qkv + 2 * Pd, 3 * Pd, nullptr, nv, Q, Q, H, Pd / H, att, Pd, nullptr, nullptr, 0, stream));
    }
    ar.release(mk);
    gemm(att, Pd, lw.out, t1, Pd, rows);
    add_ln(xin, t1, lw.n1, w->pos, Q, rows, Pd, xq);
    gemm(xq, Pd, lw.p2a, p2a, lw.p2a.N, rows);
    const float* qcur;
    const float* p2a_use = p2a;
    if (li == 0) {
      FWD_CALL(vsg_broadcast_rows(xq, Q, Pd, VQ, query2, stream));
      FWD_CALL(vsg_broadcast_rows(p2a, Q, lw.p2a.N, VQ, p2a_b, stream));
      qcur = query2; p2a_use = p2a_b;
    } else {
      qcur = xq;
    }
    if (fold) {
      FWD_CALL(vsg_role_attention_hid(p2a_use, EG + (int64_t)li * 3 * E, ld_eg, EG + (int64_t)li * 3 * E + E, ld_eg, lw.r1_bias, b->seg, V, Q, E,
                                      b->max_tracks, 1.0f / sqrtf((float)E), hid, nullptr, 0, last ? so : nullptr, stream));
    } else {
      gemm(enco, E, lw.e2a, e2a, lw.e2a.N, N);
      FWD_CALL(vsg_role_attention(p2a_use, e2a, enco, b->seg, V, Q, E, b->max_tracks, 1.0f / sqrtf((float)E), values, nullptr, 0,
                                  last ? so : nullptr, stream));
      gemm(values, 2 * E, lw.r1_0, hid, 2 * Pd, VQ, true);
      gemm(values + E, 2 * E, lw.r1_1, hid + Pd, 2 * Pd, VQ, true);
    }
    gemm(hid, 2 * Pd, lw.r2, t1, Pd, VQ);
    float* q2 = (qcur == query2) ? query : query2;         // norm2 output must not alias its input
    add_ln(qcur, t1, lw.n2, nullptr, 0, VQ, Pd, q2);
    gemm(q2, Pd, lw.f1, ffn1, lw.f1.N, VQ, true);
    gemm(ffn1, lw.f1.N, lw.f2, t1, Pd, VQ);
    float* q3 = (q2 == query) ? query2 : query;
    if (last) add_ln(q2, t1, lw.n3, nullptr, 0, VQ, Pd, q3);
    else add_ln(q2, t1, lw.n3, w->pos, Q, VQ, Pd, q3, query_pos);    // also emit query + pos for the next layer's q / k
    cur_q = q3;
    // xq is free again; cur_q lives in `query` or `query2`, and the next layer's norm1 writes xq -- no aliasing
  }

  // ---------------- prediction head (model_0v10.py:478-507 / model_0v7.py:483-513) ----------------
  int32_t* pair_index = ar.get<int32_t>(VQ);
  int32_t* so_cat = ar.get<int32_t>(VQ * 2);
  FWD_CALL(vsg_so_category(so, b->cat_ids, w->num_enti_cats, VQ, pair_index, so_cat, stream));
  const int ldz = (w->dim_z + 3) / 4 * 4;
  float* Z = ar.get<float>(VQ * ldz);
  const float* src[8]; const int32_t* idx[8]; int istr[8], ld[8], wd[8];
  int np = 0;
  auto piece = [&](const float* s, const int32_t* i, int l, int width) { src[np] = s; idx[np] = i; istr[np] = 2; ld[np] = l; wd[np] = width; ++np; };
  const int P = w->num_pred_cats;
  const int ldl = P;                                                 // logits row stride
  float* logits = ar.get<float>(VQ * ldl);
  if (w->variant == 0) {
    float* i3d = nullptr;
    if (w->dim_i3d) {
      i3d = ar.get<float>((int64_t)N * E);
      gemm(extra, w->extra_width, w->i3d, i3d, E, N, true);          // fc_i3d commutes with the row gather
      piece(cur_q, nullptr, Pd, Pd); piece(i3d, so, E, E); piece(i3d, so + 1, E, E); piece(enti2enco, so, E, E); piece(enti2enco, so + 1, E, E);
      piece(w->entiemb, so_cat, w->dim_clsme, w->dim_clsme); piece(w->entiemb, so_cat + 1, w->dim_clsme, w->dim_clsme);
    } else {
      piece(cur_q, nullptr, Pd, Pd); piece(w->entiemb, so_cat, w->dim_clsme, w->dim_clsme); piece(w->entiemb, so_cat + 1, w->dim_clsme, w->dim_clsme);
      piece(enti2enco, so, E, E); piece(enti2enco, so + 1, E, E);
    }
    FWD_CALL(vsg_gather_concat(src, idx, istr, ld, wd, np, VQ, Z, ldz, stream));
    gemm(Z, ldz, w->log, logits, ldl, VQ, false, true, w->dim_z, w->bias_matrix, pair_index, P);
  } else {
    piece(cur_q, nullptr, Pd, Pd);
    if (w->use_clsme) {
      if (w->has_entiemb) { piece(w->entiemb, so_cat, w->dim_clsme, w->dim_clsme); piece(w->entiemb, so_cat + 1, w->dim_clsme, w->dim_clsme); }
      else { piece(extra, so, w->extra_width, w->dim_clsme); piece(extra, so + 1, w->extra_width, w->dim_clsme); }
    }
    piece(enti2enco, so, E, E); piece(enti2enco, so + 1, E, E);
    FWD_CALL(vsg_gather_concat(src, idx, istr, ld, wd, np, VQ, Z, ldz, stream));
    float* hidl = ar.get<float>(VQ * w->log1.N);
    gemm(Z, ldz, w->log1, hidl, w->log1.N, VQ, true, true, w->dim_z);
    gemm(hidl, w->log1.N, w->log2, logits, ldl, VQ, false, true, -1, w->bias_matrix, pair_index, P);
  }
  // ---------------- triplets (model_0v10.py:707-785) ----------------
  if (live())
    FWD_CALL(vsg_construct_triplet(logits, ldl, P, Q, topk, so, b->seg, V, b->dura, b->cat_ids, b->scores, out->quint, out->scores, out->spans,
                                   out->qids, out->counts, out->cap, stream));
  return rc;
}

// ------------------------------------------------------------------------------------------------------------------------------
// grounding (vidsgg_big_b200/grounding.py: _forward_videos, _qanet, _dw_pw, _head, _post)
// ------------------------------------------------------------------------------------------------------------------------------
struct GrdFwd : FwdBase {
  const VsgGrdWeights* w;
  const VsgGrdBatch* b;

  bool can_fuse(const VsgGrdConv& c) const {      // linalg.can_fuse_dwconv
    return w->fuse_dwconv && ((mode == VSG_GEMM_TF32_BF16X2 && c.pw.w16) || (mode == VSG_GEMM_FP16X3 && c.pw.img16)) && c.pw.N <= 128 && c.pw.K % 16 == 0 && c.pw.K * (c.k + 1) <= 2048 &&
           (c.k & 1) && c.k <= 7;
  }
  // DepthWiseSeparableConv1d (:36-56): depthwise conv over the sequence axis, then the 1x1 conv (+ ReLU, + residual)
  void dw_pw(const float* x, const VsgGrdConv& c, const VsgGrdSeq& sq, float* out, bool relu, const float* residual) {
    const int H = w->dim_hidden;
    if (can_fuse(c)) {
      gemm(x, H, c.pw, out, c.pw.N, sq.rows, relu, true, -1, nullptr, nullptr, 0, nullptr, 0, 0, nullptr, 0, nullptr, 0, residual, residual ? c.pw.N : 0,
           &c, sq.pos, sq.rem);
      return;
    }
    const int64_t m0 = ar.mark();
    float* t = ar.get<float>(sq.rows * H);
    FWD_CALL(vsg_dwconv(x, sq.pos, sq.rem, c.dw_w, c.dw_b, c.k, sq.rows, H, t, stream));
    gemm(t, H, c.pw, out, c.pw.N, sq.rows, relu, true, -1, nullptr, nullptr, 0, nullptr, 0, 0, nullptr, 0, nullptr, 0, residual, residual ? c.pw.N : 0);
    ar.release(m0);
  }
  // QANetEncoderLayer.forward (:110-137) on rows [rows][H] of ragged sequences; returns a buffer that outlives the call
  float* qanet(const VsgGrdEncoder& e, const float* x, const VsgGrdSeq& sq) {
    const int H = w->dim_hidden;
    const int64_t rows = sq.rows;
    float* result = ar.get<float>(rows * H);
    const int64_t m0 = ar.mark();
    float* res = ar.get<float>(rows * H);
    float* res2 = ar.get<float>(rows * H);
    float* o = ar.get<float>(rows * H);
    FWD_CALL(vsg_pos_add_ln(x, sq.pos, w->freq, w->phase, e.normb.gamma, e.normb.beta, rows, H, res, o, stream));
    for (int i = 0; i < 4; ++i) {
      dw_pw(o, e.convs[i], sq, res2, true, res);                               // relu(conv) + res (:120-122)
      float* sw = res; res = res2; res2 = sw;
      FWD_CALL(vsg_add_layernorm(res, H, nullptr, 0, e.norm_seq[i].gamma, e.norm_seq[i].beta, nullptr, 0, rows, H, o, H, stream));
    }
    float* qkv = ar.get<float>(rows * 3 * H);
    float* att = ar.get<float>(rows * H);
    gemm(o, H, e.qkv, qkv, 3 * H, rows);
    if (w->tc_attention && mode != VSG_GEMM_SIMT && sq.n_tc_blk >= 0) {
      const int products = (mode == VSG_GEMM_3XTF32 || mode == VSG_GEMM_TF32_BF16X2 || mode == VSG_GEMM_FP16X3) ? 3 : 1;
      FWD_CALL(vsg_mha_tc16(qkv, 3 * H, qkv + H, 3 * H, qkv + 2 * H, 3 * H, sq.off, 8, att, H, sq.tc_blk_seg, sq.tc_blk_q0, sq.n_tc_blk, products, stream));
    } else {
      FWD_CALL(vsg_mha(qkv, 3 * H, qkv + H, 3 * H, qkv + 2 * H, 3 * H, sq.off, sq.n, 0, sq.max_len, 8, H / 8, att, H, sq.blk_seg, sq.blk_q0,
                       sq.n_blk, stream));
    }
    gemm(att, H, e.out, res2, H, rows, false, true, -1, nullptr, nullptr, 0, nullptr, 0, 0, nullptr, 0, nullptr, 0, res, H);   // attn + res (:129-130)
    FWD_CALL(vsg_add_layernorm(res2, H, nullptr, 0, e.norme.gamma, e.norme.beta, nullptr, 0, rows, H, o, H, stream));
    gemm(o, H, e.fc, result, H, rows, true, true, -1, nullptr, nullptr, 0, nullptr, 0, 0, nullptr, 0, nullptr, 0, res2, H);    // relu(fc(LN)) + res (:133-136)
    ar.release(m0);
    return result;
  }
  void head(const VsgGrdConv* hw, const float* x, const VsgGrdSeq& sq, float* out) {
    const int H = w->dim_hidden;
    const int64_t m0 = ar.mark();
    float* a = ar.get<float>(sq.rows * H);
    float* c = ar.get<float>(sq.rows * H);
    const float* y = x;
    for (int i = 0; i < 4; ++i) {
      float* dst = (i & 1) ? c : a;
      dw_pw(y, hw[i], sq, dst, true, nullptr);
      y = dst;
    }
    dw_pw(y, hw[4], sq, out, false, nullptr);
    ar.release(m0);
  }

  int run(VsgGrdOut* out, float score_th, float tiou_th, float bins_th, float nms_th) {
    const int H = w->dim_hidden, B = w->num_bins, NQ = b->n_queries;
    const int64_t rows_v = b->video.rows, rows_c = b->combined.rows;
    float* v0 = ar.get<float>(rows_v * H);
    gemm(b->video_feats, w->dim_feat, w->video_fc, v0, H, rows_v);
    float* q0 = ar.get<float>((int64_t)3 * NQ * H);
    float* so_norm = ar.get<float>((int64_t)NQ * 2);
    FWD_CALL(vsg_grd_query_init(b->quint, b->spans, b->vlen, b->q_vid, NQ, w->proj_enti, w->proj_pred, w->temp_w, w->temp_b, H, q0, so_norm, stream));
    float* v = qanet(w->video_encoder, v0, b->video);
    float* q = qanet(w->query_encoder, q0, b->query);
    float* pv = ar.get<float>(rows_v * H);
    gemm(v, H, w->proj2sim, pv, H, rows_v, false, false);
    float* comb0 = ar.get<float>(rows_c * H);
    {
      const int64_t m0 = ar.mark();
      float* comb_in = ar.get<float>(rows_c * 4 * H);
      FWD_CALL(vsg_cq_attention(v, pv, q, b->video.off, b->q_vid, b->combined.off, NQ, H, b->max_T, comb_in, stream));
      gemm(comb_in, 4 * H, w->vq_fc, comb0, H, rows_c);
      ar.release(m0);
    }
    float* comb = qanet(w->combined_encoder, comb0, b->combined);
    float* regr = out->regr ? out->regr : ar.get<float>(rows_c * 2 * B);
    float* conf = out->conf ? out->conf : ar.get<float>(rows_c * B);
    float* cls = out->cls ? out->cls : ar.get<float>(rows_c * B);
    head(w->regr_head, comb, b->combined, regr);
    head(w->conf_head, comb, b->combined, conf);
    head(w->cls_head, comb, b->combined, cls);
    if (live())
      FWD_CALL(vsg_grounding_post(regr, conf, cls, b->combined.off, so_norm, b->clip_tab, b->video.off, b->q_vid, NQ, B, 0, score_th, tiou_th, bins_th,
                                  nms_th, out->pooled, out->probs, out->mask, out->err_count, stream));
    return rc;
  }
};

static int check_grd(const VsgGrdWeights* w, const VsgGrdBatch* b, int mode) {
  VSG_REQUIRE(w && b, "vsg_grd_forward: null weights / batch");
  VSG_REQUIRE(mode >= VSG_GEMM_SIMT && mode <= VSG_GEMM_FP16X3, "vsg_grd_forward: unknown precision mode %d", mode);
  VSG_REQUIRE(w->dim_hidden == 128, "vsg_grd_forward: dim_hidden must be 128 (the context-query kernel's instantiation)");
  VSG_REQUIRE(b->n_videos > 0 && b->n_queries > 0 && b->video.rows > 0 && b->combined.rows > 0, "vsg_grd_forward: empty batch");
  return VSG_OK;
}

static int check_args(const VsgBigCWeights* w, const VsgVideoBatch* b, int topk, int mode) {
  VSG_REQUIRE(w && b, "vsg_bigc_forward: null weights / batch");
  VSG_REQUIRE(mode >= VSG_GEMM_SIMT && mode <= VSG_GEMM_FP16X3, "vsg_bigc_forward: unknown precision mode %d", mode);
  VSG_REQUIRE(w->n_enc >= 0 && w->n_enc <= VSG_MAX_LAYERS && w->n_dec >= 1 && w->n_dec <= VSG_MAX_LAYERS, "vsg_bigc_forward: layer counts out of range");
  VSG_REQUIRE(w->dim_enti == w->dim_pred, "vsg_bigc_forward: dim_enti must equal dim_pred");
  VSG_REQUIRE(b->n_videos > 0 && b->n_tracks > 0 && b->n_rows > 0, "vsg_bigc_forward: empty batch");
  VSG_REQUIRE(topk > 0 && topk <= w->num_pred_cats, "vsg_bigc_forward: bad topk");
  VSG_REQUIRE(!b->feats_bf16 || (mode == VSG_GEMM_BF16 && b->ld_feats % 8 == 0), "vsg_bigc_forward: bf16 features need mode VSG_GEMM_BF16 and ld_feats %% 8 == 0");
  return VSG_OK;
}

}  // namespace vsg

using namespace vsg;

extern "C" int64_t vsg_bigc_workspace_bytes(const VsgBigCWeights* w, const VsgVideoBatch* b, int topk, int precision_mode) {
  if (check_args(w, b, topk, precision_mode) != VSG_OK) return -1;
  Fwd f;
  f.w = w; f.b = b; f.mode = precision_mode; f.stream = nullptr; f.rc = VSG_OK;
  f.ar.base = nullptr; f.ar.cap = 0; f.ar.off = 0; f.ar.peak = 0; f.ar.dry = true; f.ar.overflow = false;
  VsgTripletOut dummy;
  memset(&dummy, 0, sizeof(dummy));
  f.run(&dummy, topk);
  return f.ar.peak + 256;
}

extern "C" int vsg_bigc_forward(const VsgBigCWeights* w, const VsgVideoBatch* b, VsgTripletOut* out, int topk, int precision_mode,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  int rc = check_args(w, b, topk, precision_mode);
  if (rc != VSG_OK) return rc;
  VSG_REQUIRE(out && out->quint && out->scores && out->spans && out->qids && out->counts && out->cap >= w->num_querys * topk,
              "vsg_bigc_forward: output buffers missing or cap < num_querys * topk");
  VSG_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "vsg_bigc_forward: workspace must be 256-byte aligned");
  Fwd f;
  f.w = w; f.b = b; f.mode = precision_mode; f.stream = stream; f.rc = VSG_OK;
  f.ar.base = reinterpret_cast<uint8_t*>(workspace); f.ar.cap = workspace_bytes; f.ar.off = 0; f.ar.peak = 0; f.ar.dry = false; f.ar.overflow = false;
  // size check BEFORE anything is launched: the dry pass is cheap (no launches)
  const int64_t need = vsg_bigc_workspace_bytes(w, b, topk, precision_mode);
  VSG_REQUIRE(need >= 0 && workspace_bytes >= need - 256, "vsg_bigc_forward: workspace too small (%lld bytes needed, %lld given)",
              (long long)need, (long long)workspace_bytes);
  rc = f.run(out, topk);
  if (rc == VSG_OK && f.ar.overflow) { set_error("vsg_bigc_forward: workspace overflow"); return VSG_E_INVALID; }
  return rc;
}

extern "C" int64_t vsg_grd_workspace_bytes(const VsgGrdWeights* w, const VsgGrdBatch* b, int precision_mode) {
  if (check_grd(w, b, precision_mode) != VSG_OK) return -1;
  GrdFwd f;
  f.w = w; f.b = b; f.mode = precision_mode; f.stream = nullptr; f.rc = VSG_OK;
  f.ar.base = nullptr; f.ar.cap = 0; f.ar.off = 0; f.ar.peak = 0; f.ar.dry = true; f.ar.overflow = false;
  VsgGrdOut dummy;
  memset(&dummy, 0, sizeof(dummy));
  f.run(&dummy, 0.f, 0.f, 0.f, 0.f);
  return f.ar.peak + 256;
}

extern "C" int vsg_grd_forward(const VsgGrdWeights* w, const VsgGrdBatch* b, VsgGrdOut* out, float score_th, float tiou_th, float bins_th,
                               float nms_th, int precision_mode, void* workspace, int64_t workspace_bytes, void* stream) {
  int rc = check_grd(w, b, precision_mode);
  if (rc != VSG_OK) return rc;
  VSG_REQUIRE(out && out->pooled && out->probs && out->mask && out->err_count, "vsg_grd_forward: output buffers missing");
  VSG_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "vsg_grd_forward: workspace must be 256-byte aligned");
  VsgGrdOut sized = *out;
  const int64_t need = [&] {
    GrdFwd d;
    d.w = w; d.b = b; d.mode = precision_mode; d.stream = nullptr; d.rc = VSG_OK;
    d.ar.base = nullptr; d.ar.cap = 0; d.ar.off = 0; d.ar.peak = 0; d.ar.dry = true; d.ar.overflow = false;
    d.run(&sized, 0.f, 0.f, 0.f, 0.f);
    return d.ar.peak;
  }();
  VSG_REQUIRE(workspace_bytes >= need, "vsg_grd_forward: workspace too small (%lld bytes needed, %lld given)", (long long)need, (long long)workspace_bytes);
  GrdFwd f;
  f.w = w; f.b = b; f.mode = precision_mode; f.stream = stream; f.rc = VSG_OK;
  f.ar.base = reinterpret_cast<uint8_t*>(workspace); f.ar.cap = workspace_bytes; f.ar.off = 0; f.ar.peak = 0; f.ar.dry = false; f.ar.overflow = false;
  rc = f.run(out, score_th, tiou_th, bins_th, nms_th);
  if (rc == VSG_OK && f.ar.overflow) { set_error("vsg_grd_forward: workspace overflow"); return VSG_E_INVALID; }
  return rc;
}
