// Grounding stage (models/grd_model_v5.py DEBUG) non-GEMM kernels: SURVEY.md section 8a rows A10, A11.
//
// Everything is row-major [rows][H] (H = dim_hidden = 128) over RAGGED SEQUENCES described by seq_off int64[n_seq+1]:
//   video encoder     rows = sum_v T_v            one sequence per video
//   query encoder     rows = 3 * sum_v nq_v       one 3-word sequence per query
//   combined encoder  rows = sum_v nq_v * T_v     one T_v-clip sequence per (video, query)
// The reference keeps (N, C, T) tensors and transposes around every LayerNorm / Linear; here the channel dim is
// always innermost so every kernel is coalesced and the 1x1 convs / Linears are plain GEMMs (csrc/gemm.cu).
#include "common.cuh"
#include <math.h>

namespace vsg {

// position (pos) and remaining length (rem = L-1-pos) of every row inside its sequence
__global__ void seq_positions_kernel(const int64_t* __restrict__ seq_off, int n_seq, int64_t rows, int32_t* __restrict__ pos,
                                     int32_t* __restrict__ rem) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int s = find_segment(seq_off, n_seq, r);
    pos[r] = (int)(r - seq_off[s]);
    rem[r] = (int)(seq_off[s + 1] - 1 - r);
  }
}

// query_emb[(n,l)] = proj_table_l[cat] + temp_fc(span / video_len)   (grd_model_v5.py:318-323, :337-339)
// quint int64[nq][5] = [pred, scat, ocat, sid, oid]; word order (sub, pred, obj) (:321)
__global__ void query_init_kernel(const int64_t* __restrict__ quint, const int64_t* __restrict__ spans, const float* __restrict__ vlen,
                                  const int32_t* __restrict__ q_vid, int nq, const float* __restrict__ proj_enti,
                                  const float* __restrict__ proj_pred, const float* __restrict__ Wt, const float* __restrict__ bt, int H,
                                  float* __restrict__ out, float* __restrict__ so_norm /* [nq][2] */) {
  const int64_t total = (int64_t)nq * 3 * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % H);
    const int64_t row = i / H;
    const int n = (int)(row / 3), l = (int)(row % 3);
    const float len = vlen[q_vid[n]];
    const float s = (float)spans[2 * n] / len, e = (float)spans[2 * n + 1] / len;
    const float* tab = (l == 1) ? proj_pred + quint[5 * n] * H : proj_enti + quint[5 * n + (l == 0 ? 1 : 2)] * H;
    float t = bt[c];
    t = fmaf(Wt[2 * c], s, t);
    t = fmaf(Wt[2 * c + 1], e, t);
    out[i] = tab[c] + t;
    if (l == 0 && c < 2) so_norm[2 * n + c] = c == 0 ? s : e;
  }
}

// PosEncoder (:58-78) + first LayerNorm of QANetEncoderLayer (:116-118):
//   res = x + sin(pos*freq + phase);  out = LN(res)*g + b.       One warp per row, H <= 128*... (H % 32 == 0, H <= 256)
__global__ void __launch_bounds__(256)
pos_add_ln_kernel(const float* __restrict__ x, const int32_t* __restrict__ pos, const float* __restrict__ freq,
                  const float* __restrict__ phase, const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows, int H,
                  float* __restrict__ res, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int per = H / 32;
  for (int64_t r = warp; r < rows; r += n_warps) {
    const float p = (float)pos[r];
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < per) {
        const int c = lane + 32 * i;
        const float arg = __fadd_rn(__fmul_rn(p, freq[c]), phase[c]);   // torch: pos * freqs + phases (two roundings)
        v[i] = x[r * H + c] + sinf(arg);
        s += v[i];
      }
    }
    const float mean = warp_sum(s) / (float)H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < per) { const float d = v[i] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)H + 1e-5f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < per) {
        const int c = lane + 32 * i;
        res[r * H + c] = v[i];
        out[r * H + c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
      }
    }
  }
}

// Depthwise conv over the sequence axis with zero padding at sequence ends (DepthWiseSeparableConv1d.depth_wise, :44):
//   out[r][c] = b[c] + sum_j w[c][j] * in[r + j - k/2][c].   One thread per (row, 4 channels).
__global__ void __launch_bounds__(256)
dwconv_kernel(const float* __restrict__ in, const int32_t* __restrict__ pos, const int32_t* __restrict__ rem,
              const float* __restrict__ w /* [H][k] */, const float* __restrict__ b, int k, int64_t rows, int H,
              float* __restrict__ out) {
  const int h4 = H / 4;
  const int64_t total = rows * h4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / h4;
    const int c = (int)(i - r * h4) * 4;
    const int p = pos[r], q = rem[r];
    float4 acc = *reinterpret_cast<const float4*>(b + c);
    for (int j = 0; j < k; ++j) {
      const int d = j - k / 2;
      if (d < -p || d > q) continue;
      const float4 x = *reinterpret_cast<const float4*>(in + (r + d) * H + c);
      acc.x = fmaf(w[(c + 0) * k + j], x.x, acc.x);
      acc.y = fmaf(w[(c + 1) * k + j], x.y, acc.y);
      acc.z = fmaf(w[(c + 2) * k + j], x.z, acc.z);
      acc.w = fmaf(w[(c + 3) * k + j], x.w, acc.w);
    }
    *reinterpret_cast<float4*>(out + r * H + c) = acc;
  }
}

// Context-query attention (:345-363) for one (video, query) per CTA, H = 128 threads (one channel each):
//   sim[t][l] = <pv[t], q[l]>;  S_r = softmax_l, S_c = softmax_t;  A[t] = sum_l S_r[t][l] q[l];
//   B[t] = sum_l S_r[t][l] U[l],  U[l] = sum_t' S_c[t'][l] v[t']      (re-association of (S_r S_c^T) V, rank 3)
//   out[(n,t)] = [v[t], A[t], A[t]*v[t], B[t]*v[t]]                    (4H columns, input of vq_fc)
constexpr int CQ_H = 128;
__global__ void __launch_bounds__(CQ_H)
cq_attention_kernel(const float* __restrict__ v, const float* __restrict__ pv, const float* __restrict__ q,
                    const int64_t* __restrict__ vid_off /* video rows */, const int32_t* __restrict__ q_vid, const int64_t* __restrict__ comb_off
                    /* combined row offset per query */, float* __restrict__ out) {
  extern __shared__ float sm[];  // sim[T][3] then reused as S_r; S_c[T][3]
  const int n = blockIdx.x;
  const int vid = q_vid[n];
  const int64_t v0 = vid_off[vid];
  const int T = (int)(vid_off[vid + 1] - v0);
  float* sR = sm;
  float* sC = sm + 3 * T;
  __shared__ float red[3][4];
  const int c = threadIdx.x, warp = c >> 5, lane = c & 31;
  float ql[3];
#pragma unroll
  for (int l = 0; l < 3; ++l) ql[l] = q[((int64_t)n * 3 + l) * CQ_H + c];
  // phase 1: sim (warp per clip; every lane holds 4 channels)
  {
    float q4[3][4];
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
      for (int j = 0; j < 4; ++j) q4[l][j] = q[((int64_t)n * 3 + l) * CQ_H + lane * 4 + j];
    for (int t = warp; t < T; t += CQ_H / 32) {
      const float4 x = *reinterpret_cast<const float4*>(pv + (v0 + t) * CQ_H + lane * 4);
      float d[3];
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        d[l] = x.x * q4[l][0] + x.y * q4[l][1] + x.z * q4[l][2] + x.w * q4[l][3];
        d[l] = warp_sum(d[l]);
      }
      if (lane == 0) { sR[3 * t] = d[0]; sR[3 * t + 1] = d[1]; sR[3 * t + 2] = d[2]; }
    }
  }
  __syncthreads();
  // phase 2: column softmax (over t) statistics, then both softmaxes
  float cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int t = c; t < T; t += CQ_H)
#pragma unroll
    for (int l = 0; l < 3; ++l) cmax[l] = fmaxf(cmax[l], sR[3 * t + l]);
#pragma unroll
  for (int l = 0; l < 3; ++l) { cmax[l] = warp_max(cmax[l]); if (lane == 0) red[l][warp] = cmax[l]; }
  __syncthreads();
#pragma unroll
  for (int l = 0; l < 3; ++l) cmax[l] = fmaxf(fmaxf(red[l][0], red[l][1]), fmaxf(red[l][2], red[l][3]));
  __syncthreads();
  float csum[3] = {0.f, 0.f, 0.f};
  for (int t = c; t < T; t += CQ_H)
#pragma unroll
    for (int l = 0; l < 3; ++l) { const float e = expf(sR[3 * t + l] - cmax[l]); sC[3 * t + l] = e; csum[l] += e; }
#pragma unroll
  for (int l = 0; l < 3; ++l) { csum[l] = warp_sum(csum[l]); if (lane == 0) red[l][warp] = csum[l]; }
  __syncthreads();
#pragma unroll
  for (int l = 0; l < 3; ++l) csum[l] = red[l][0] + red[l][1] + red[l][2] + red[l][3];
  for (int t = c; t < T; t += CQ_H) {
    const float a0 = sR[3 * t], a1 = sR[3 * t + 1], a2 = sR[3 * t + 2];
    const float m = fmaxf(a0, fmaxf(a1, a2));
    const float e0 = expf(a0 - m), e1 = expf(a1 - m), e2 = expf(a2 - m);
    const float inv = 1.f / (e0 + e1 + e2);
    sR[3 * t] = e0 * inv; sR[3 * t + 1] = e1 * inv; sR[3 * t + 2] = e2 * inv;
#pragma unroll
    for (int l = 0; l < 3; ++l) sC[3 * t + l] = sC[3 * t + l] / csum[l];
  }
  __syncthreads();
  // phase 3: U[l][c]
  float U[3] = {0.f, 0.f, 0.f};
  for (int t = 0; t < T; ++t) {
    const float x = v[(v0 + t) * CQ_H + c];
#pragma unroll
    for (int l = 0; l < 3; ++l) U[l] = fmaf(sC[3 * t + l], x, U[l]);
  }
  // phase 4: outputs
  float* o = out + comb_off[n] * (4 * CQ_H);
  for (int t = 0; t < T; ++t) {
    const float x = v[(v0 + t) * CQ_H + c];
    const float r0 = sR[3 * t], r1 = sR[3 * t + 1], r2 = sR[3 * t + 2];
    const float A = r0 * ql[0] + r1 * ql[1] + r2 * ql[2];
    const float B = r0 * U[0] + r1 * U[1] + r2 * U[2];
    float* row = o + (int64_t)t * (4 * CQ_H);
    row[c] = x; row[CQ_H + c] = A; row[2 * CQ_H + c] = A * x; row[3 * CQ_H + c] = B * x;
  }
}

// ---------------------------------------------------------------------------------------------------
// Post-network part of _forward_test_single (:533-576): scores, temporal_pooling (:697-737), clipping to the
// subject-object span, the extra "s-o overlap" bin, 1-D NMS over the k+1 bins (:667-695), mask fix-ups.
// One warp per query.  num_bins <= 16.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// clip centres: torch.linspace(0, 1, T) (:705) is evaluated by the host once per distinct T (its vectorised CPU kernel is
// not reproducible to the last ulp with a closed form) and passed as a table clip[video rows].

__global__ void __launch_bounds__(128)
grounding_post_kernel(const float* __restrict__ regr /* [rows][2B] pre-sigmoid */, const float* __restrict__ conf, const float* __restrict__ cls,
                      const int64_t* __restrict__ comb_off, const float* __restrict__ so_norm, const float* __restrict__ clip_tab,
                      const int64_t* __restrict__ vid_off, const int32_t* __restrict__ q_vid, int nq, int B, int regr_activated, float score_th, float tiou_th,
                      float bins_th, float nms_th, float* __restrict__ pooled /* [nq][B+1][2] */, float* __restrict__ probs /* [nq][B+1] */,
                      uint8_t* __restrict__ mask /* [nq][B+1] */, int32_t* __restrict__ err_count) {
  const int lane = threadIdx.x & 31;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= nq) return;
  const int64_t r0 = comb_off[n];
  const int T = (int)(comb_off[n + 1] - r0);
  const float so0 = so_norm[2 * n], so1 = so_norm[2 * n + 1];
  const float* clip = clip_tab + vid_off[q_vid[n]];
  float bin_s[17], bin_e[17], bin_p[17];
  bool bin_ov[17];
  for (int k = 0; k < B; ++k) {
    // top score (first max)
    float best = -INFINITY;
    int best_t = 0x7fffffff;
    for (int t = lane; t < T; t += 32) {
      const float s = sigmoidf_acc(conf[(r0 + t) * B + k]) * sigmoidf_acc(cls[(r0 + t) * B + k]);
      if (s > best) { best = s; best_t = t; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int ot = __shfl_xor_sync(0xffffffffu, best_t, o);
      if (ob > best || (ob == best && ot < best_t)) { best = ob; best_t = ot; }
    }
    const float clip_top = clip[best_t];
    const float rl_top = regr[(r0 + best_t) * 2 * B + k], rr_top = regr[(r0 + best_t) * 2 * B + B + k];
    const float ts = clip_top - (regr_activated ? rl_top : sigmoidf_acc(rl_top));
    const float te = clip_top + (regr_activated ? rr_top : sigmoidf_acc(rr_top));
    const float thr = score_th * best;
    float mn = INFINITY, mx = -INFINITY;
    for (int t = lane; t < T; t += 32) {
      const float s = sigmoidf_acc(conf[(r0 + t) * B + k]) * sigmoidf_acc(cls[(r0 + t) * B + k]);
      const float cl = clip[t];
      const float rl = regr[(r0 + t) * 2 * B + k], rr = regr[(r0 + t) * 2 * B + B + k];
      const float st = cl - (regr_activated ? rl : sigmoidf_acc(rl));
      const float en = cl + (regr_activated ? rr : sigmoidf_acc(rr));
      const float g = (fminf(te, en) - fmaxf(ts, st)) / (fmaxf(te, en) - fminf(ts, st));
      if (s > thr && g > tiou_th) { mn = fminf(mn, st); mx = fmaxf(mx, en); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (mn == INFINITY && lane == 0) atomicAdd(err_count, 1);   // the reference raises here (:726, min of an empty tensor)
    // clip to the s-o span (:546-551)
    const float is = fmaxf(so0, mn), ie = fminf(so1, mx);
    const bool ok = is <= ie;
    bin_s[k] = ok ? is : so0; bin_e[k] = ok ? ie : so1; bin_ov[k] = ok; bin_p[k] = best;
  }
  bin_s[B] = so0; bin_e[B] = so1; bin_ov[B] = true; bin_p[B] = 1.0f;
  if (lane != 0) return;
  const int NB = B + 1;
  // ---- 1-D NMS (:667-681): ascending order by prob, repeatedly keep the last, drop tIoU >= nms_th ----
  int order[17];
  bool alive[17], kept[17];
  for (int i = 0; i < NB; ++i) { order[i] = i; alive[i] = true; kept[i] = false; }
  for (int i = 1; i < NB; ++i) {  // stable insertion sort ascending
    const int o = order[i];
    int j = i - 1;
    while (j >= 0 && bin_p[order[j]] > bin_p[o]) { order[j + 1] = order[j]; --j; }
    order[j + 1] = o;
  }
  for (int top_i = NB - 1; top_i >= 0; --top_i) {
    const int top = order[top_i];
    if (!alive[top]) continue;
    kept[top] = true;
    alive[top] = false;
    for (int j = 0; j < top_i; ++j) {
      const int o = order[j];
      if (!alive[o]) continue;
      const bool touch = (bin_e[top] >= bin_s[o]) && (bin_e[o] >= bin_s[top]);
      float ti = (fminf(bin_e[top], bin_e[o]) - fmaxf(bin_s[top], bin_s[o])) / (fmaxf(bin_e[top], bin_e[o]) - fminf(bin_s[top], bin_s[o]));
      if (!touch) ti = 0.f;
      if (!(ti < nms_th)) alive[o] = false;
    }
  }
  bool any = false;
  float best_p = -INFINITY;
  int best_i = 0;
  float max_real = -INFINITY;
  for (int i = 0; i < NB; ++i) {
    const bool m = (bin_p[i] > bins_th) && bin_ov[i] && kept[i];
    mask[(int64_t)n * NB + i] = m ? 1 : 0;
    any |= m;
    if (bin_p[i] > best_p) { best_p = bin_p[i]; best_i = i; }
    if (i < B) max_real = fmaxf(max_real, bin_p[i]);
  }
  if (!any) mask[(int64_t)n * NB + best_i] = 1;          // :561-564
  if (max_real <= bins_th) bin_p[B] = 0.0f;               // :571-572
  for (int i = 0; i < NB; ++i) {
    pooled[((int64_t)n * NB + i) * 2] = bin_s[i];
    pooled[((int64_t)n * NB + i) * 2 + 1] = bin_e[i];
    probs[(int64_t)n * NB + i] = bin_p[i];
  }
}

static inline int grid_cap_g(int64_t blocks, int per_sm) {
  const int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace vsg

using namespace vsg;

extern "C" int vsg_seq_positions(const int64_t* seq_off, int n_seq, int64_t rows, int32_t* pos, int32_t* rem, void* stream) {
  VSG_REQUIRE(n_seq >= 0 && rows >= 0, "vsg_seq_positions: bad size");
  if (rows == 0) return VSG_OK;
  VSG_REQUIRE(seq_off && pos && rem, "vsg_seq_positions: null pointer");
  seq_positions_kernel<<<grid_cap_g((rows + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(seq_off, n_seq, rows, pos, rem);
  return check_launch("vsg_seq_positions");
}

extern "C" int vsg_grd_query_init(const int64_t* quint, const int64_t* spans, const float* vlen, const int32_t* q_vid, int nq,
                                  const float* proj_enti, const float* proj_pred, const float* Wt, const float* bt, int H, float* out,
                                  float* so_norm, void* stream) {
  VSG_REQUIRE(nq >= 0 && H > 0, "vsg_grd_query_init: bad size");
  if (nq == 0) return VSG_OK;
  VSG_REQUIRE(quint && spans && vlen && q_vid && proj_enti && proj_pred && Wt && bt && out && so_norm, "vsg_grd_query_init: null pointer");
  query_init_kernel<<<grid_cap_g(((int64_t)nq * 3 * H + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(
      quint, spans, vlen, q_vid, nq, proj_enti, proj_pred, Wt, bt, H, out, so_norm);
  return check_launch("vsg_grd_query_init");
}

extern "C" int vsg_pos_add_ln(const float* x, const int32_t* pos, const float* freq, const float* phase, const float* gamma,
                              const float* beta, int64_t rows, int H, float* res, float* out, void* stream) {
  VSG_REQUIRE(rows >= 0 && H > 0 && H % 32 == 0 && H <= 256, "vsg_pos_add_ln: H must be a multiple of 32, <= 256");
  if (rows == 0) return VSG_OK;
  VSG_REQUIRE(x && pos && freq && phase && gamma && beta && res && out, "vsg_pos_add_ln: null pointer");
  pos_add_ln_kernel<<<grid_cap_g((rows + 7) / 8, 8), 256, 0, (cudaStream_t)stream>>>(x, pos, freq, phase, gamma, beta, rows, H, res, out);
  return check_launch("vsg_pos_add_ln");
}

extern "C" int vsg_dwconv(const float* in, const int32_t* pos, const int32_t* rem, const float* w, const float* b, int k, int64_t rows,
                          int H, float* out, void* stream) {
  VSG_REQUIRE(rows >= 0 && H > 0 && H % 4 == 0 && k >= 1 && (k & 1), "vsg_dwconv: H %% 4 == 0 and odd k required");
  if (rows == 0) return VSG_OK;
  VSG_REQUIRE(in && pos && rem && w && b && out && aligned16(in) && aligned16(out) && aligned16(b), "vsg_dwconv: null/misaligned pointer");
  dwconv_kernel<<<grid_cap_g((rows * (H / 4) + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(in, pos, rem, w, b, k, rows, H, out);
  return check_launch("vsg_dwconv");
}

extern "C" int vsg_cq_attention(const float* v, const float* pv, const float* q, const int64_t* vid_off, const int32_t* q_vid,
                                const int64_t* comb_off, int nq, int H, int max_T, float* out, void* stream) {
  VSG_REQUIRE(nq >= 0 && max_T >= 0, "vsg_cq_attention: bad size");
  VSG_REQUIRE(H == CQ_H, "vsg_cq_attention: dim_hidden must be %d", CQ_H);
  if (nq == 0) return VSG_OK;
  VSG_REQUIRE(v && pv && q && vid_off && q_vid && comb_off && out && aligned16(pv), "vsg_cq_attention: null/misaligned pointer");
  const size_t smem = (size_t)6 * max_T * sizeof(float);
  VSG_REQUIRE(smem <= 200 * 1024, "vsg_cq_attention: video too long (%d clips)", max_T);
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(cq_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("vsg_cq_attention: cannot raise dynamic shared memory to %zu", smem);
      return VSG_E_LAUNCH;
    }
  }
  cq_attention_kernel<<<nq, CQ_H, smem, (cudaStream_t)stream>>>(v, pv, q, vid_off, q_vid, comb_off, out);
  return check_launch("vsg_cq_attention");
}

extern "C" int vsg_grounding_post(const float* regr, const float* conf, const float* cls, const int64_t* comb_off, const float* so_norm,
                                  const float* clip_tab, const int64_t* vid_off, const int32_t* q_vid, int nq, int num_bins, int regr_activated, float score_th, float tiou_th, float bins_th, float nms_th, float* pooled,
                                  float* probs, uint8_t* mask, int32_t* err_count, void* stream) {
  VSG_REQUIRE(nq >= 0 && num_bins >= 1 && num_bins <= 16, "vsg_grounding_post: 1 <= num_bins <= 16");
  if (nq == 0) return VSG_OK;
  VSG_REQUIRE(regr && conf && cls && comb_off && so_norm && clip_tab && vid_off && q_vid && pooled && probs && mask && err_count,
              "vsg_grounding_post: null pointer");
  grounding_post_kernel<<<(nq * 32 + 127) / 128, 128, 0, (cudaStream_t)stream>>>(regr, conf, cls, comb_off, so_norm, clip_tab, vid_off, q_vid, nq, num_bins, regr_activated, score_th,
                                                                                 tiou_th, bins_th, nms_th, pooled, probs, mask, err_count);
  return check_launch("vsg_grounding_post");
}
