// BIG-C classification stage: every non-GEMM kernel (the GEMMs are csrc/gemm.cu).
//
// All kernels work on a BATCH of videos in packed form:
//   rows     = all box-frames of all tracks of all videos (R = sum L);  off[N+1] row offsets per track
//   tracks   = N over all videos; seg[V+1] track range per video; tmax[N] = longest track of the track's video
//   queries  = V * Q rows (Q = num_querys, video-major)
// The reference stretches each track to the video's longest track by repeating frames
// (models/model_0v10.py:18-46): frame i of an L-frame track appears ceil((Tmax-i)/L) times.  Here the stretch
// is never materialised: with q = Tmax / L, rem = Tmax % L, frame i is repeated q+1 times if i < rem else q times,
// and stretched position j maps to   src(j) = j / (q+1)                     if j < rem*(q+1)
//                                            rem + (j - rem*(q+1)) / q      otherwise.
#include "common.cuh"
#include <cuda_bf16.h>
#include <float.h>

namespace vsg {

__device__ __forceinline__ int find_track(const int64_t* __restrict__ off, int n, int64_t row) {
  return find_segment(off, n, row);
}

struct Stretch {
  int L, q, rem, split;  // split = rem*(q+1)
  __device__ __forceinline__ Stretch(int L_, int Tmax) : L(L_) { q = Tmax / L_; rem = Tmax - q * L_; split = rem * (q + 1); }
  __device__ __forceinline__ int src(int j) const { return j < split ? j / (q + 1) : rem + (j - split) / q; }
  __device__ __forceinline__ int reps(int i) const { return i < rem ? q + 1 : q; }
};

// ---------------------------------------------------------------------------------------------------
// A5: 8-d motion features (model_0v10.py:401-420) fused with the first fc_bbox2enti layer (8 -> E, ReLU).
// One warp per row; lane owns channels lane, lane+32, ...
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bbox_feat_mlp1_kernel(const float4* __restrict__ boxes, const int64_t* __restrict__ off, int n_tracks, int64_t n_rows,
                      const int32_t* __restrict__ track_vid, const float* __restrict__ wh,  // wh[V][2]
                      const float* __restrict__ W1, const float* __restrict__ b1, int E,    // W1 [E][8]
                      float* __restrict__ out, int ldo, float* __restrict__ feat8_out /* optional [R][8] */,
                      __nv_bfloat16* __restrict__ out16 = nullptr /* bf16 output instead of `out` (row stride ldo) */) {
  extern __shared__ float sw[];  // W1 transposed [8][E] + b1[E]
  for (int i = threadIdx.x; i < E * 8; i += blockDim.x) sw[(i % 8) * E + i / 8] = W1[i];
  for (int i = threadIdx.x; i < E; i += blockDim.x) sw[8 * E + i] = b1[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool vec4 = (E % 4 == 0) && (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out16) & 7) == 0);
  const int G = (E + 127) / 128;                 // groups of 128 channels
  if (vec4 && !feat8_out && (G == 1 || G == 2 || G == 4)) {
    // Register-resident weights: a warp owns ONE group of 128 channels (4 per lane: 8 weights + bias = 36 registers) and walks the rows;
    // the G warps that share a row recompute its 8 features (broadcast loads).  The shared-memory form below issues 9 LDS.128 per 4
    // outputs -- 144 shared-memory wavefronts per 2 KB row, which made this write-only kernel LDS-bound at ~1/3 of the HBM write rate.
    // Same fmaf order per channel (bias, then k = 0..7), so the outputs are bit-identical.
    const int g = (int)(warp % G), c = g * 128 + lane * 4;
    const bool on = c < E;
    float4 wr[8], bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) wr[k] = on ? *reinterpret_cast<const float4*>(sw + k * E + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) bias4 = *reinterpret_cast<const float4*>(sw + 8 * E + c);
    // rows in blocks of 32: lane l derives the 8 features of row r0 + l (one binary search per lane, all 32 in parallel), then the warp walks
    // the block and broadcasts each row's features by shuffles
    const int64_t n_blocks = (n_rows + 31) / 32;
    for (int64_t blk = warp / G; blk < n_blocks; blk += n_warps / G) {
      const int64_t r0 = blk * 32, rl = r0 + lane;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (rl < n_rows) {
        const int t = find_track(off, n_tracks, rl);
        const bool last = (rl + 1 == off[t + 1]);
        const int v = track_vid[t];
        const float w = wh[2 * v], h = wh[2 * v + 1];
        float4 a = boxes[rl];
        float4 b = last ? a : boxes[rl + 1];
        a.x = a.x / w; a.z = a.z / w; a.y = a.y / h; a.w = a.w / h;
        b.x = b.x / w; b.z = b.z / w; b.y = b.y / h; b.w = b.w / h;
        const float cx = (a.z + a.x) / 2, cy = (a.w + a.y) / 2, bw = a.z - a.x, bh = a.w - a.y;
        const float cx2 = (b.z + b.x) / 2, cy2 = (b.w + b.y) / 2, bw2 = b.z - b.x, bh2 = b.w - b.y;
        f[0] = cx; f[1] = last ? 0.f : cx2 - cx;
        f[2] = cy; f[3] = last ? 0.f : cy2 - cy;
        f[4] = bw; f[5] = last ? 0.f : bw2 - bw;
        f[6] = bh; f[7] = last ? 0.f : bh2 - bh;
      }
      const int nr = (int)min((int64_t)32, n_rows - r0);
      for (int j = 0; j < nr; ++j) {
        float4 acc = bias4;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float fk = __shfl_sync(0xffffffffu, f[k], j);
          acc.x = fmaf(fk, wr[k].x, acc.x); acc.y = fmaf(fk, wr[k].y, acc.y); acc.z = fmaf(fk, wr[k].z, acc.z); acc.w = fmaf(fk, wr[k].w, acc.w);
        }
        if (!on) continue;
        const int64_t r = r0 + j;
        if (out16) {
          const __nv_bfloat162 p0 = __floats2bfloat162_rn(fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f)), p1 = __floats2bfloat162_rn(fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f));
          *reinterpret_cast<uint2*>(out16 + r * (int64_t)ldo + c) = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
        } else {
          *reinterpret_cast<float4*>(out + r * (int64_t)ldo + c) = make_float4(fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f), fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f));
        }
      }
    }
    return;
  }
  for (int64_t r = warp; r < n_rows; r += n_warps) {
    const int t = find_track(off, n_tracks, r);
    const bool last = (r + 1 == off[t + 1]);
    const int v = track_vid[t];
    const float w = wh[2 * v], h = wh[2 * v + 1];
    float4 a = boxes[r];
    float4 b = last ? a : boxes[r + 1];
    a.x = a.x / w; a.z = a.z / w; a.y = a.y / h; a.w = a.w / h;
    b.x = b.x / w; b.z = b.z / w; b.y = b.y / h; b.w = b.w / h;
    float f[8];
    const float cx = (a.z + a.x) / 2, cy = (a.w + a.y) / 2, bw = a.z - a.x, bh = a.w - a.y;
    const float cx2 = (b.z + b.x) / 2, cy2 = (b.w + b.y) / 2, bw2 = b.z - b.x, bh2 = b.w - b.y;
    f[0] = cx; f[1] = last ? 0.f : cx2 - cx;
    f[2] = cy; f[3] = last ? 0.f : cy2 - cy;
    f[4] = bw; f[5] = last ? 0.f : bw2 - bw;
    f[6] = bh; f[7] = last ? 0.f : bh2 - bh;
    if (feat8_out && lane < 8) feat8_out[r * 8 + lane] = f[lane];
    float* o = out + r * (int64_t)ldo;
    if (vec4) {
      // 4 channels per lane and LDS.128: the scalar form issues 9 shared loads per output and is LDS-bound (same fmaf order per channel)
      for (int c = lane * 4; c < E; c += 128) {
        float4 acc = *reinterpret_cast<const float4*>(sw + 8 * E + c);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 wv = *reinterpret_cast<const float4*>(sw + k * E + c);
          acc.x = fmaf(f[k], wv.x, acc.x); acc.y = fmaf(f[k], wv.y, acc.y); acc.z = fmaf(f[k], wv.z, acc.z); acc.w = fmaf(f[k], wv.w, acc.w);
        }
        if (out16) {
          const __nv_bfloat162 p0 = __floats2bfloat162_rn(fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f)), p1 = __floats2bfloat162_rn(fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f));
          *reinterpret_cast<uint2*>(out16 + r * (int64_t)ldo + c) = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
        } else {
          *reinterpret_cast<float4*>(o + c) = make_float4(fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f), fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f));
        }
      }
    } else {
      for (int c = lane; c < E; c += 32) {
        float acc = sw[8 * E + c];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(f[k], sw[k * E + c], acc);
        if (out16) out16[r * (int64_t)ldo + c] = __float2bfloat16_rn(fmaxf(acc, 0.f)); else o[c] = fmaxf(acc, 0.f);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Time-mean of the extra feature columns over the STRETCHED sequence (model_0v10.py:470, model_0v7.py:473):
// mean_j x[src(j)] = sum_i reps(i) * x[i] / Tmax.   One CTA per track, threads over channels.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }

template <typename TF>
__global__ void __launch_bounds__(256)
stretched_mean_kernel(const TF* __restrict__ feat, int ldf, int col0, int width, const int64_t* __restrict__ off,
                      const int32_t* __restrict__ tmax, float* __restrict__ out, int ldo) {
  const int t = blockIdx.x;
  const int64_t r0 = off[t];
  const int L = (int)(off[t + 1] - r0);
  const Stretch st(L, tmax[t]);
  if constexpr (sizeof(TF) == 4) {
    // fp32 rows: 4 channels per thread (16-byte loads), 4 rows in flight per thread -- the scalar loop below keeps one 4-byte load per
    // thread in flight (measured 0.36 ms for 1.6 GB: ~2/3 of the HBM rate).  Same accumulation order per channel (rows in sequence).
    const float* base = reinterpret_cast<const float*>(feat) + r0 * (int64_t)ldf + col0;
    if ((width & 3) == 0 && (ldf & 3) == 0 && (ldo & 3) == 0 && ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
      const float inv_den = (float)tmax[t];
      for (int c4 = threadIdx.x; c4 < width / 4; c4 += blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* p = base + 4 * c4;
        int i = 0;
        for (; i + 4 <= L; i += 4) {
          float4 x[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) x[u] = ldg_stream(reinterpret_cast<const float4*>(p + (int64_t)(i + u) * ldf));
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float w = (float)st.reps(i + u);
            acc.x = fmaf(w, x[u].x, acc.x); acc.y = fmaf(w, x[u].y, acc.y); acc.z = fmaf(w, x[u].z, acc.z); acc.w = fmaf(w, x[u].w, acc.w);
          }
        }
        for (; i < L; ++i) {
          const float4 x = ldg_stream(reinterpret_cast<const float4*>(p + (int64_t)i * ldf));
          const float w = (float)st.reps(i);
          acc.x = fmaf(w, x.x, acc.x); acc.y = fmaf(w, x.y, acc.y); acc.z = fmaf(w, x.z, acc.z); acc.w = fmaf(w, x.w, acc.w);
        }
        *reinterpret_cast<float4*>(out + (int64_t)t * ldo + 4 * c4) = make_float4(acc.x / inv_den, acc.y / inv_den, acc.z / inv_den, acc.w / inv_den);
      }
      return;
    }
  }
  for (int c = threadIdx.x; c < width; c += blockDim.x) {
    float acc = 0.f;
    const TF* p = feat + r0 * (int64_t)ldf + col0 + c;
    for (int i = 0; i < L; ++i) acc = fmaf((float)st.reps(i), to_f32(p[(int64_t)i * ldf]), acc);
    out[(int64_t)t * ldo + c] = acc / (float)tmax[t];
  }
}

// ---------------------------------------------------------------------------------------------------
// conv_feat2enti (k=3, stride 2, pad 1 over stretched time) + adaptive_max_pool1d(pool) (model_0v10.py:450-457)
// from the three per-frame tap products Y[r][k*E + c] = sum_ci W[c][ci][k] * X[r][ci] (one GEMM).
// out position t:  Y0[src(2t-1)] + Y1[src(2t)] + Y2[src(2t+1)] (+ bias), zero padding outside [0, Tmax).
// One CTA per (track, pool bin); threads own 4 channels; consecutive positions that gather the same three
// source frames are skipped, so a track costs O(L) loads however far it is stretched.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename TY>
__global__ void __launch_bounds__(128)
conv_pool_kernel(const TY* __restrict__ Y, int ldy, int E, const float* __restrict__ bias, const int64_t* __restrict__ off,
                 const int32_t* __restrict__ tmax, int pool, float* __restrict__ out /* [N][E*pool] channel-major */) {
  const int t = blockIdx.x / pool, p = blockIdx.x % pool;
  const int64_t r0 = off[t];
  const int L = (int)(off[t + 1] - r0);
  const int Tmax = tmax[t];
  const Stretch st(L, Tmax);
  const int Tc = (Tmax - 1) / 2 + 1;
  const int ts = (p * Tc) / pool;
  const int te = ((p + 1) * Tc + pool - 1) / pool;
  for (int c4 = threadIdx.x; c4 < E / 4; c4 += blockDim.x) {
    float4 best = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    int pa = -2, pb = -2, pc = -2;
    for (int tt = ts; tt < te; ++tt) {
      const int ja = 2 * tt - 1, jb = 2 * tt, jc = 2 * tt + 1;
      const int a = ja >= 0 ? st.src(ja) : -1;
      const int b = st.src(jb);
      const int c = jc < Tmax ? st.src(jc) : -1;
      if (a == pa && b == pb && c == pc) continue;
      pa = a; pb = b; pc = c;
      float4 v = load4(Y + (r0 + b) * (int64_t)ldy + E + 4 * c4);
      if (a >= 0) {
        const float4 x = load4(Y + (r0 + a) * (int64_t)ldy + 4 * c4);
        v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
      }
      if (c >= 0) {
        const float4 x = load4(Y + (r0 + c) * (int64_t)ldy + 2 * E + 4 * c4);
        v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
      }
      best.x = fmaxf(best.x, v.x); best.y = fmaxf(best.y, v.y); best.z = fmaxf(best.z, v.z); best.w = fmaxf(best.w, v.w);
    }
    const float4 bb = *reinterpret_cast<const float4*>(bias + 4 * c4);
    float* o = out + (int64_t)t * E * pool;
    o[(4 * c4 + 0) * pool + p] = best.x + bb.x;
    o[(4 * c4 + 1) * pool + p] = best.y + bb.y;
    o[(4 * c4 + 2) * pool + p] = best.z + bb.z;
    o[(4 * c4 + 3) * pool + p] = best.w + bb.w;
  }
}

// ---------------------------------------------------------------------------------------------------
// out = LayerNorm(x + a) * gamma + beta (+ post[row % post_period])   (eps 1e-5, biased variance).
// One warp per row, D <= 1024, D % 32 == 0.  `a` may be null.
// ---------------------------------------------------------------------------------------------------
template <int MAXPL>
__global__ void __launch_bounds__(256)
add_layernorm_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ a, int lda, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ post, int post_period, int64_t rows, int D,
                     float* __restrict__ out, int ldo, float* __restrict__ out2, int ldo2) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int per = D / 32;
  for (int64_t r = warp; r < rows; r += n_warps) {
    float v[MAXPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXPL; ++i) {
      if (i < per) {
        const int c = lane + 32 * i;
        float t = x[r * ldx + c];
        if (a) t += a[r * lda + c];
        v[i] = t;
        s += t;
      }
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXPL; ++i)
      if (i < per) { const float d = v[i] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-5f);
    const float* prow = post ? post + (r % post_period) * (int64_t)D : nullptr;      // one 64-bit modulo per row, not per element
#pragma unroll
    for (int i = 0; i < MAXPL; ++i) {
      if (i < per) {
        const int c = lane + 32 * i;
        float y = (v[i] - mean) * rstd * gamma[c] + beta[c];
        if (out2) {                       // dual form: out = LN(.), out2 = LN(.) + post   (the next layer's q/k input, model_0v10.py:181)
          out[r * ldo + c] = y;
          out2[r * ldo2 + c] = y + prow[c];
        } else {
          if (prow) y += prow[c];
          out[r * ldo + c] = y;
        }
      }
    }
  }
}

// out[r][c] = x[r % period][c]   (query initialisation: pred_query_init broadcast to every video)
__global__ void broadcast_rows_kernel(const float* __restrict__ x, int period, int D, int64_t rows, float* __restrict__ out) {
  const int64_t total = rows * (D / 4);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (D / 4);
    const int c4 = (int)(i - r * (D / 4));
    reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(x)[(r % period) * (D / 4) + c4];
  }
}

// ---------------------------------------------------------------------------------------------------
// Multi-head self-attention core on packed rows (nn.MultiheadAttention semantics without the projections):
// for every segment s (rows seg_off[s]..seg_off[s+1]) and head h:  O = softmax(Q K^T / sqrt(dh)) V.
// Q/K/V are column blocks of (possibly the same) row-major buffers.
//
// One CTA (8 warps) per (segment, head, 64-query block); keys are walked in super-chunks of 128 with an online
// softmax, so the segment length is unbounded.  Per super-chunk:
//   phase A  warp (kg, qh): lane = one key of key group kg, its K row lives in REGISTERS (read straight from
//            global/L2); the 32 queries of half qh are broadcast from shared memory -> 1 FMA per MAC, the
//            scores land transposed in St[key][query];
//   phase B  warp w owns 8 queries: chunk max / exp / running-sum update (lanes stride the keys), then PV with
//            the 8 probabilities of a key fetched by two broadcast LDS.128 and V[key][lane(+32)] from smem:
//            16 FMAs per 4 shared loads.
// ---------------------------------------------------------------------------------------------------
constexpr int MHA_QB = 64, MHA_PITCH = 68;
template <int DH> struct MhaCfg { static constexpr int SK = DH >= 64 ? 128 : 256; };   // keys per super-chunk

template <int DH>
__global__ void __launch_bounds__(256, 2)
mha_kernel(const float* __restrict__ Qp, int ldq, const float* __restrict__ Kp, int ldk, const float* __restrict__ Vp, int ldv,
           const int64_t* __restrict__ seg_off, int fixed_len, int n_head, float scale, float* __restrict__ Op, int ldo,
           const int32_t* __restrict__ blk_seg, const int32_t* __restrict__ blk_q0) {
  constexpr int MHA_SK = MhaCfg<DH>::SK;
  constexpr int KG = MHA_SK / 32;               // key groups of 32 (one per warp in phase A)
  constexpr int QH = 8 / KG;                    // query slices in phase A
  constexpr int QPS = MHA_QB / QH;              // queries per slice
  extern __shared__ __align__(16) float mha_smem[];
  float* Qs = mha_smem;                         // [64][DH] pre-scaled
  float* Vs = Qs + MHA_QB * DH;                 // [SK][DH]
  float* St = Vs + MHA_SK * DH;                 // [SK][68] scores / probabilities, transposed (key-major)
  // work item = (segment, 64-query block): either listed explicitly (ragged segments: no empty CTAs) or grid (x, z)
  const int seg = blk_seg ? blk_seg[blockIdx.x] : blockIdx.x, head = blockIdx.y;
  const int64_t row0 = seg_off ? seg_off[seg] : (int64_t)seg * fixed_len;
  const int n = seg_off ? (int)(seg_off[seg + 1] - row0) : fixed_len;
  const int q_base = blk_seg ? blk_q0[blockIdx.x] : blockIdx.z * MHA_QB;
  if (q_base >= n) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hc = head * DH;
  constexpr int OD = (DH + 31) / 32;   // output dims per lane
  constexpr int D4 = DH / 4;
  // PV lane layout: for DH < 32 the warp is split into G = 32/DH groups, each owning QPL = 8/G of the warp's 8 queries,
  // so that all 32 lanes stay busy (lane -> dim lane % DH of the queries of group lane / DH)
  constexpr int G = DH < 32 ? 32 / DH : 1;
  constexpr int QPL = 8 / G;
  const int grp = DH < 32 ? lane / DH : 0;
  const int dlane = DH < 32 ? lane % DH : lane;
  // queries of this block (zero rows beyond n)
  for (int i = threadIdx.x; i < MHA_QB * D4; i += 256) {
    const int q = i / D4, d4 = i % D4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q_base + q < n) {
      v = *reinterpret_cast<const float4*>(Qp + (row0 + q_base + q) * ldq + hc + 4 * d4);
      v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    }
    reinterpret_cast<float4*>(Qs)[i] = v;
  }
  float m_run[8], l_run[8], o_acc[QPL][OD];
#pragma unroll
  for (int j = 0; j < 8; ++j) { m_run[j] = -INFINITY; l_run[j] = 0.f; }
#pragma unroll
  for (int j = 0; j < QPL; ++j)
#pragma unroll
    for (int d = 0; d < OD; ++d) o_acc[j][d] = 0.f;
  const int kg = warp % KG, qh = warp / KG;
  for (int k0 = 0; k0 < n; k0 += MHA_SK) {
    const int kc = min(MHA_SK, n - k0);
    __syncthreads();  // previous chunk fully consumed (and Qs visible on the first pass)
    for (int i = threadIdx.x; i < MHA_SK * D4; i += 256) {
      const int kr = i / D4, d4 = i % D4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kr < kc) v = *reinterpret_cast<const float4*>(Vp + (row0 + k0 + kr) * ldv + hc + 4 * d4);
      reinterpret_cast<float4*>(Vs)[i] = v;
    }
    // ---- phase A ----
    {
      const int key = kg * 32 + lane;
      if (kg * 32 < kc) {          // warp-uniform
        float kreg[DH];
        const bool valid = key < kc;
        const float* kp = Kp + (row0 + k0 + (valid ? key : 0)) * ldk + hc;
#pragma unroll
        for (int d4 = 0; d4 < D4; ++d4) {
          const float4 v = *reinterpret_cast<const float4*>(kp + 4 * d4);
          kreg[4 * d4] = v.x; kreg[4 * d4 + 1] = v.y; kreg[4 * d4 + 2] = v.z; kreg[4 * d4 + 3] = v.w;
        }
        float* srow = St + key * MHA_PITCH + qh * QPS;
#pragma unroll 8
        for (int q = 0; q < QPS; ++q) {
          const float4* qp = reinterpret_cast<const float4*>(Qs + (qh * QPS + q) * DH);
          float acc = 0.f;
#pragma unroll
          for (int d4 = 0; d4 < D4; ++d4) {
            const float4 x = qp[d4];
            acc = fmaf(kreg[4 * d4], x.x, acc); acc = fmaf(kreg[4 * d4 + 1], x.y, acc);
            acc = fmaf(kreg[4 * d4 + 2], x.z, acc); acc = fmaf(kreg[4 * d4 + 3], x.w, acc);
          }
          srow[q] = valid ? acc : -INFINITY;
        }
      }
    }
    __syncthreads();
    // ---- phase B: warp owns queries 8*warp .. 8*warp+7 ----
    if (q_base + 8 * warp < n) {    // warp-uniform
      const int nk_groups = (kc + 31) / 32;
      float c8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int q = 8 * warp + j;
        float sc[MHA_SK / 32];
        float cmax = -INFINITY;
#pragma unroll
        for (int i = 0; i < MHA_SK / 32; ++i) {
          sc[i] = (i < nk_groups) ? St[(lane + 32 * i) * MHA_PITCH + q] : -INFINITY;
          cmax = fmaxf(cmax, sc[i]);
        }
        cmax = warp_max(cmax);
        const float m_new = fmaxf(m_run[j], cmax);
        const float corr = expf(m_run[j] - m_new);
        float psum = 0.f;
#pragma unroll
        for (int i = 0; i < MHA_SK / 32; ++i) {
          if (i < nk_groups) {
            const float pv = expf(sc[i] - m_new);      // -inf -> 0 for keys beyond kc
            St[(lane + 32 * i) * MHA_PITCH + q] = pv;
            psum += pv;
          }
        }
        psum = warp_sum(psum);
        l_run[j] = l_run[j] * corr + psum;
        m_run[j] = m_new;
        c8[j] = corr;
      }
#pragma unroll
      for (int j = 0; j < QPL; ++j) {
        float c = c8[j];
#pragma unroll
        for (int g = 1; g < G; ++g) c = (grp == g) ? c8[g * QPL + j] : c;
#pragma unroll
        for (int d = 0; d < OD; ++d) o_acc[j][d] *= c;
      }
      __syncwarp();
      if (DH >= 32 ? (lane < DH) : true) {
        const float* prow = St + 8 * warp + grp * QPL;
#pragma unroll 4
        for (int kr = 0; kr < kc; ++kr) {
          float pv[QPL];
          if (QPL == 8) {
            const float4 p0 = *reinterpret_cast<const float4*>(prow + kr * MHA_PITCH);
            const float4 p1 = *reinterpret_cast<const float4*>(prow + kr * MHA_PITCH + 4);
            pv[0] = p0.x; pv[1] = p0.y; pv[2] = p0.z; pv[3] = p0.w;
            pv[4 % QPL] = p1.x; pv[5 % QPL] = p1.y; pv[6 % QPL] = p1.z; pv[7 % QPL] = p1.w;
          } else if (QPL == 4) {
            const float4 p0 = *reinterpret_cast<const float4*>(prow + kr * MHA_PITCH);
            pv[0] = p0.x; pv[1] = p0.y; pv[2 % QPL] = p0.z; pv[3 % QPL] = p0.w;
          } else {
#pragma unroll
            for (int j = 0; j < QPL; ++j) pv[j] = prow[kr * MHA_PITCH + j];
          }
#pragma unroll
          for (int d = 0; d < OD; ++d) {
            const float v = Vs[kr * DH + dlane + 32 * d];
#pragma unroll
            for (int j = 0; j < QPL; ++j) o_acc[j][d] = fmaf(pv[j], v, o_acc[j][d]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < QPL; ++j) {
    const int q = q_base + 8 * warp + grp * QPL + j;
    float l = l_run[j];
#pragma unroll
    for (int g = 1; g < G; ++g) l = (grp == g) ? l_run[g * QPL + j] : l;
    if (q >= n) continue;
#pragma unroll
    for (int d = 0; d < OD; ++d) {
      const int dd = dlane + 32 * d;
      if (dd < DH) Op[(row0 + q) * ldo + hc + dd] = o_acc[j][d] / l;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Tensor-core attention helpers (fixed-length segments: the 192 decoder queries of every video).  QK^T and PV run as batched
// tcgen05 GEMMs (csrc/gemm.cu, one problem per (video, head)); these two kernels are the glue:
//   softmax_rows_kernel   P = softmax(scale * S) row-wise, in place; one warp per row, n <= 256 columns
//   transpose_split_kernel  V part of the packed qkv buffer [rows][ld] -> VT_hi / VT_lo [cols][ld_t] (K-major B operand of the
//                           PV GEMM + its tf32 low part), 32x32 smem tiles
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_rows_kernel(float* __restrict__ S, int ld, int n, int64_t rows, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += n_warps) {
    float* row = S + r * ld;
    float x[8];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      x[i] = c < n ? row[c] * scale : -INFINITY;
      mx = fmaxf(mx, x[i]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      x[i] = c < n ? expf(x[i] - mx) : 0.f;
      sum += x[i];
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      if (c < n) row[c] = x[i] / sum;
    }
  }
}

__global__ void __launch_bounds__(256)
transpose_split_kernel(const float* __restrict__ X, int ld, int64_t rows, int cols, float* __restrict__ Thi, float* __restrict__ Tlo,
                       int64_t ld_t) {
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i;
    tile[i][tx] = (r < rows && c0 + tx < cols) ? X[r * ld + c0 + tx] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;
    const int64_t r = r0 + tx;
    if (c < cols && r < rows) {
      const float v = tile[tx][i];
      const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      Thi[(int64_t)c * ld_t + r] = v;
      if (Tlo) Tlo[(int64_t)c * ld_t + r] = v - h;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Role attention of RoleAttnDecoderLayer (model_0v10.py:190-214) for every video:
//   logits[r][q][e] = <p2a[q][r-half], e2a[e][r-half]> / sqrt(dim_enti)
//   att = softmax_e(logits) * softmax_r(logits);   values[r][q] = sum_e att[r][q][e] * enco[e]
// One CTA per (video, 8-query block), one warp per query; lanes 0-15 hold the subject half of the 512-d
// projections, lanes 16-31 the object half.  Writes values as [row][r*E + c] (the input layout of the two
// fc_rolewise first layers), and on request the attention matrix and its per-role argmax (prediction_head :485).
// ---------------------------------------------------------------------------------------------------
constexpr int RA_MAX_TRACKS = 256;

template <int PER>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float (&x)[PER]) {
  if constexpr (PER % 4 == 0) {
#pragma unroll
    for (int i = 0; i < PER / 4; ++i) {
      const float4 v = reinterpret_cast<const float4*>(p)[i];
      x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < PER; ++i) x[i] = p[i];
  }
}

template <int E>
__global__ void __launch_bounds__(256)
role_attention_kernel(const float* __restrict__ p2a, const float* __restrict__ e2a, const float* __restrict__ enco,
                      const int32_t* __restrict__ seg, int Q, float inv_sqrt_d, float* __restrict__ values,
                      float* __restrict__ att_out /* [V*Q][2][max_n] or null */, int att_ld,
                      int32_t* __restrict__ so_out /* [V*Q][2] global track ids, or null */) {
  constexpr int PER = E / 32;  // 16 dims per lane
  __shared__ float sL[8][2][RA_MAX_TRACKS];
  const int v = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.y * 8 + warp;
  if (q >= Q) return;
  const int t0 = seg[v], n = seg[v + 1] - t0;
  const int64_t qrow = (int64_t)v * Q + q;
  float pq[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) pq[i] = p2a[qrow * E + lane * PER + i];
  // pass 1: logits; 4 tracks per iteration so that 16 independent 128-bit loads are in flight per lane
  for (int e0 = 0; e0 < n; e0 += 4) {
    float xr[4][PER];
#pragma unroll
    for (int u = 0; u < 4; ++u) load_row<PER>(e2a + (int64_t)(t0 + min(e0 + u, n - 1)) * E + lane * PER, xr[u]);
    float acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc[u] = 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i) acc[u] = fmaf(pq[i], xr[u][i], acc[u]);
    }
    // reduce within each half-warp (role)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
    }
    if ((lane & 15) == 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (e0 + u < n) sL[warp][lane >> 4][e0 + u] = acc[u] * inv_sqrt_d;
    }
  }
  __syncwarp();
  // softmax over tracks per role (lanes stride e), then softmax over roles per track
  float mx[2] = {-INFINITY, -INFINITY};
  for (int e = lane; e < n; e += 32) { mx[0] = fmaxf(mx[0], sL[warp][0][e]); mx[1] = fmaxf(mx[1], sL[warp][1][e]); }
  mx[0] = warp_max(mx[0]); mx[1] = warp_max(mx[1]);
  float sm[2] = {0.f, 0.f};
  for (int e = lane; e < n; e += 32) { sm[0] += expf(sL[warp][0][e] - mx[0]); sm[1] += expf(sL[warp][1][e] - mx[1]); }
  sm[0] = warp_sum(sm[0]); sm[1] = warp_sum(sm[1]);
  float best[2] = {-INFINITY, -INFINITY};
  int best_e[2] = {0x7fffffff, 0x7fffffff};
  for (int e = lane; e < n; e += 32) {
    const float l0 = sL[warp][0][e], l1 = sL[warp][1][e];
    const float m = fmaxf(l0, l1);
    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
    const float a0 = (expf(l0 - mx[0]) / sm[0]) * (e0 / (e0 + e1));
    const float a1 = (expf(l1 - mx[1]) / sm[1]) * (e1 / (e0 + e1));
    sL[warp][0][e] = a0; sL[warp][1][e] = a1;
    if (a0 > best[0]) { best[0] = a0; best_e[0] = e; }
    if (a1 > best[1]) { best[1] = a1; best_e[1] = e; }
    if (att_out) { att_out[(qrow * 2 + 0) * att_ld + e] = a0; att_out[(qrow * 2 + 1) * att_ld + e] = a1; }
  }
  if (so_out) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best[r], o);
        const int oe = __shfl_xor_sync(0xffffffffu, best_e[r], o);
        if (ob > best[r] || (ob == best[r] && oe < best_e[r])) { best[r] = ob; best_e[r] = oe; }
      }
    }
    if (lane == 0) { so_out[qrow * 2] = t0 + best_e[0]; so_out[qrow * 2 + 1] = t0 + best_e[1]; }
  }
  __syncwarp();
  // pass 2: values[r] = sum_e att[r][e] * enco[e]; every lane owns 16 dims of BOTH roles
  float acc0[PER], acc1[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }
  for (int e0 = 0; e0 < n; e0 += 4) {
    float xr[4][PER];
#pragma unroll
    for (int u = 0; u < 4; ++u) load_row<PER>(enco + (int64_t)(t0 + min(e0 + u, n - 1)) * E + lane * PER, xr[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool ok = e0 + u < n;
      const float a0 = ok ? sL[warp][0][e0 + u] : 0.f, a1 = ok ? sL[warp][1][e0 + u] : 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i) { acc0[i] = fmaf(a0, xr[u][i], acc0[i]); acc1[i] = fmaf(a1, xr[u][i], acc1[i]); }
    }
  }
  float* o0 = values + qrow * (2 * E) + lane * PER;
  float* o1 = o0 + E;
#pragma unroll
  for (int i = 0; i < PER; ++i) { o0[i] = acc0[i]; o1[i] = acc1[i]; }
}

// ---------------------------------------------------------------------------------------------------
// Shared-memory staged role attention for E = 128 / 512 (the production dims).  Same maths as role_attention_kernel, but the
// e2a / enco rows of the video are staged chunk-wise in shared memory once per CTA (8 queries) instead of being re-read
// from L2 by every warp, and lanes own interleaved float4 columns (j*32 + lane) so every LDS.128 is conflict-free.
// ---------------------------------------------------------------------------------------------------
constexpr int RA_CH = 16;  // tracks per staged chunk

template <int E>
__global__ void __launch_bounds__(256)
role_attention_smem_kernel(const float* __restrict__ p2a, const float* __restrict__ e2a, const float* __restrict__ enco,
                           const int32_t* __restrict__ seg, int Q, float inv_sqrt_d, float* __restrict__ values,
                           float* __restrict__ att_out, int att_ld, int32_t* __restrict__ so_out) {
  constexpr int NJ = E / 128;            // float4 columns per lane
  constexpr int E4 = E / 4;
  extern __shared__ __align__(16) float ra_smem[];
  float4* sRows = reinterpret_cast<float4*>(ra_smem);                              // [RA_CH][E4]
  float (*sL)[2][RA_MAX_TRACKS] = reinterpret_cast<float (*)[2][RA_MAX_TRACKS]>(ra_smem + RA_CH * E);
  const int v = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.y * 8 + warp;
  const bool active = q < Q;
  const int t0 = seg[v], n = seg[v + 1] - t0;
  const int64_t qrow = (int64_t)v * Q + (active ? q : 0);
  float4 pq[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) pq[j] = reinterpret_cast<const float4*>(p2a + qrow * E)[j * 32 + lane];
  // ---- pass 1: logits ----
  for (int c0 = 0; c0 < n; c0 += RA_CH) {
    const int cn = min(RA_CH, n - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < cn * E4; i += 256)
      sRows[i] = reinterpret_cast<const float4*>(e2a + (int64_t)(t0 + c0) * E)[i];
    __syncthreads();
    if (active) {
      for (int e = 0; e < cn; ++e) {
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const float4 x = sRows[e * E4 + j * 32 + lane];
          const float t = pq[j].x * x.x + pq[j].y * x.y + pq[j].z * x.z + pq[j].w * x.w;
          const bool obj = (j * 32 + lane) * 4 >= E / 2;
          d0 += obj ? 0.f : t;
          d1 += obj ? t : 0.f;
        }
        d0 = warp_sum(d0); d1 = warp_sum(d1);
        if (lane == 0) { sL[warp][0][c0 + e] = d0 * inv_sqrt_d; sL[warp][1][c0 + e] = d1 * inv_sqrt_d; }
      }
    }
  }
  __syncwarp();
  float best[2] = {-INFINITY, -INFINITY};
  int best_e[2] = {0x7fffffff, 0x7fffffff};
  if (active) {
    float mx[2] = {-INFINITY, -INFINITY};
    for (int e = lane; e < n; e += 32) { mx[0] = fmaxf(mx[0], sL[warp][0][e]); mx[1] = fmaxf(mx[1], sL[warp][1][e]); }
    mx[0] = warp_max(mx[0]); mx[1] = warp_max(mx[1]);
    float sm[2] = {0.f, 0.f};
    for (int e = lane; e < n; e += 32) { sm[0] += expf(sL[warp][0][e] - mx[0]); sm[1] += expf(sL[warp][1][e] - mx[1]); }
    sm[0] = warp_sum(sm[0]); sm[1] = warp_sum(sm[1]);
    for (int e = lane; e < n; e += 32) {
      const float l0 = sL[warp][0][e], l1 = sL[warp][1][e];
      const float m = fmaxf(l0, l1);
      const float e0 = expf(l0 - m), e1 = expf(l1 - m);
      const float a0 = (expf(l0 - mx[0]) / sm[0]) * (e0 / (e0 + e1));
      const float a1 = (expf(l1 - mx[1]) / sm[1]) * (e1 / (e0 + e1));
      sL[warp][0][e] = a0; sL[warp][1][e] = a1;
      if (a0 > best[0]) { best[0] = a0; best_e[0] = e; }
      if (a1 > best[1]) { best[1] = a1; best_e[1] = e; }
      if (att_out) { att_out[(qrow * 2 + 0) * att_ld + e] = a0; att_out[(qrow * 2 + 1) * att_ld + e] = a1; }
    }
    if (so_out) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best[r], o);
          const int oe = __shfl_xor_sync(0xffffffffu, best_e[r], o);
          if (ob > best[r] || (ob == best[r] && oe < best_e[r])) { best[r] = ob; best_e[r] = oe; }
        }
      }
      if (lane == 0) { so_out[qrow * 2] = t0 + best_e[0]; so_out[qrow * 2 + 1] = t0 + best_e[1]; }
    }
  }
  __syncwarp();
  // ---- pass 2: values ----
  float4 acc0[NJ], acc1[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) { acc0[j] = make_float4(0.f, 0.f, 0.f, 0.f); acc1[j] = acc0[j]; }
  for (int c0 = 0; c0 < n; c0 += RA_CH) {
    const int cn = min(RA_CH, n - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < cn * E4; i += 256)
      sRows[i] = reinterpret_cast<const float4*>(enco + (int64_t)(t0 + c0) * E)[i];
    __syncthreads();
    if (active) {
      for (int e = 0; e < cn; ++e) {
        const float a0 = sL[warp][0][c0 + e], a1 = sL[warp][1][c0 + e];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const float4 x = sRows[e * E4 + j * 32 + lane];
          acc0[j].x = fmaf(a0, x.x, acc0[j].x); acc0[j].y = fmaf(a0, x.y, acc0[j].y);
          acc0[j].z = fmaf(a0, x.z, acc0[j].z); acc0[j].w = fmaf(a0, x.w, acc0[j].w);
          acc1[j].x = fmaf(a1, x.x, acc1[j].x); acc1[j].y = fmaf(a1, x.y, acc1[j].y);
          acc1[j].z = fmaf(a1, x.z, acc1[j].z); acc1[j].w = fmaf(a1, x.w, acc1[j].w);
        }
      }
    }
  }
  if (active) {
    float4* o0 = reinterpret_cast<float4*>(values + qrow * (2 * E));
    float4* o1 = o0 + E4;
#pragma unroll
    for (int j = 0; j < NJ; ++j) { o0[j * 32 + lane] = acc0[j]; o1[j * 32 + lane] = acc1[j]; }
  }
}

// ---------------------------------------------------------------------------------------------------
// Role attention with the first fc_rolewise layer folded in (model_0v10.py:190-214):
//   hid[r][q] = ReLU( fc_rolewise[r].0 (att[r][q] @ enco) ) = ReLU( sum_e att[r][q][e] * G[r][e] + b[r] ),   G[r] = enco W_r^T (one GEMM over
// the TRACKS, shared by every query) -- the per-query [2E] `values` rows and the two V*Q-row GEMMs that consumed them disappear.
// One CTA per (video, 16 queries), 8 warps:
//   pass 1  logits: a warp owns 2 queries, a lane E/32 dims (lanes 0-15: subject half, 16-31: object half); the e2a rows of 16 tracks are
//           staged in shared memory, each lane accumulates its partial dot for the 16 tracks, and a 4-step transposing butterfly leaves the
//           complete dot of track i in lane i of each half-warp (15 shuffles per 16 tracks instead of 10 per track)
//   softmax over tracks x softmax over roles, arg-max / attention output as in role_attention_kernel; att is stored transposed
//           [role][track][16 queries] for pass 2
//   pass 2  a warp owns E/4 output columns of one role for ALL 16 queries (lane: E/128 columns x 16 accumulators): every staged G value is
//           read once per CTA, the 16 attention weights of a track come from four broadcast 16-byte loads
// ---------------------------------------------------------------------------------------------------
constexpr int RA2_QB = 16;      // queries per CTA
constexpr int RA2_CH = 16;      // tracks per staged chunk

template <int E>
__global__ void __launch_bounds__(256)
role_attention_hid_kernel(const float* __restrict__ p2a, const float* __restrict__ e2a, int ld_e2a, const float* __restrict__ G, int ld_g,
                          const float* __restrict__ bias, const int32_t* __restrict__ seg, int Q, int ld_t, float inv_sqrt_d,
                          float* __restrict__ hid, float* __restrict__ att_out, int att_ld, int32_t* __restrict__ so_out) {
  constexpr int PER = E / 32;            // dims per lane in pass 1
  constexpr int VW = E / 128;            // output columns per lane in pass 2 (a warp owns E / 4 columns of one role)
  extern __shared__ __align__(16) float ra2_smem[];
  float* sRows = ra2_smem;                                         // pass 1: [RA2_CH][E] e2a rows; pass 2: [RA2_CH][2E] G rows
  float* sL = ra2_smem + RA2_CH * 2 * E;                           // [RA2_QB][2][ld_t] logits
  float* sAT = sL + RA2_QB * 2 * ld_t;                             // [2][ld_t][RA2_QB] attention, transposed
  const int v = blockIdx.x, q0 = blockIdx.y * RA2_QB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = seg[v], n = seg[v + 1] - t0;
  if (n > ld_t) __trap();                                          // the caller's max_tracks sized the shared-memory tables: fail loudly, never overrun
  const int hl = lane & 15, role1 = lane >> 4;
  // ---- pass 1: logits of this warp's two queries ----
  // a lane's dims: float4 number j * 16 + hl (j < PER / 4) of its role half -- a quarter-warp then reads 128 contiguous bytes of a staged row
  float pq[2][PER];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int q = min(q0 + 2 * warp + u, Q - 1);
    const float4* src = reinterpret_cast<const float4*>(p2a + ((int64_t)v * Q + q) * E) + role1 * (E / 8) + hl;
#pragma unroll
    for (int j = 0; j < PER / 4; ++j) {
      const float4 t = src[j * 16];
      pq[u][4 * j] = t.x; pq[u][4 * j + 1] = t.y; pq[u][4 * j + 2] = t.z; pq[u][4 * j + 3] = t.w;
    }
  }
  for (int c0 = 0; c0 < n; c0 += RA2_CH) {
    const int cn = min(RA2_CH, n - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < cn * (E / 4); i += 256) {
      const int e = i / (E / 4), c = i - e * (E / 4);
      reinterpret_cast<float4*>(sRows)[i] = reinterpret_cast<const float4*>(e2a + (int64_t)(t0 + c0 + e) * ld_e2a)[c];
    }
    __syncthreads();
    float acc[2][RA2_CH];
#pragma unroll
    for (int e = 0; e < RA2_CH; ++e) {
      float x[PER];
      const float4* xr = reinterpret_cast<const float4*>(sRows + (e < cn ? e : 0) * E) + role1 * (E / 8) + hl;
#pragma unroll
      for (int j = 0; j < PER / 4; ++j) {
        const float4 t = xr[j * 16];
        x[4 * j] = t.x; x[4 * j + 1] = t.y; x[4 * j + 2] = t.z; x[4 * j + 3] = t.w;
      }
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i) { a0 = fmaf(pq[0][i], x[i], a0); a1 = fmaf(pq[1][i], x[i], a1); }
      acc[0][e] = a0; acc[1][e] = a1;
    }
    // transposing butterfly inside each half-warp: after the step with offset o a lane keeps the tracks whose bit o matches its own
#pragma unroll
    for (int u = 0; u < 2; ++u) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const bool up = (hl & o) != 0;
#pragma unroll
        for (int j = 0; j < o; ++j) {
          const float send = up ? acc[u][j] : acc[u][j + o];
          const float keep = up ? acc[u][j + o] : acc[u][j];
          acc[u][j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      if (hl < cn) sL[((2 * warp + u) * 2 + role1) * ld_t + c0 + hl] = acc[u][0] * inv_sqrt_d;
    }
  }
  __syncwarp();
  // ---- softmax over tracks (per role) x softmax over roles (per track), arg-max; same operation order as role_attention_kernel ----
#pragma unroll 1
  for (int u = 0; u < 2; ++u) {
    const int ql = 2 * warp + u, q = q0 + ql;
    float* l0p = sL + (ql * 2 + 0) * ld_t;
    float* l1p = sL + (ql * 2 + 1) * ld_t;
    float mx[2] = {-INFINITY, -INFINITY};
    for (int e = lane; e < n; e += 32) { mx[0] = fmaxf(mx[0], l0p[e]); mx[1] = fmaxf(mx[1], l1p[e]); }
    mx[0] = warp_max(mx[0]); mx[1] = warp_max(mx[1]);
    float sm[2] = {0.f, 0.f};
    for (int e = lane; e < n; e += 32) { sm[0] += expf(l0p[e] - mx[0]); sm[1] += expf(l1p[e] - mx[1]); }
    sm[0] = warp_sum(sm[0]); sm[1] = warp_sum(sm[1]);
    float best[2] = {-INFINITY, -INFINITY};
    int best_e[2] = {0x7fffffff, 0x7fffffff};
    const bool active = q < Q;
    const int64_t qrow = (int64_t)v * Q + (active ? q : 0);
    for (int e = lane; e < n; e += 32) {
      const float l0 = l0p[e], l1 = l1p[e];
      const float m = fmaxf(l0, l1);
      const float e0 = expf(l0 - m), e1 = expf(l1 - m);
      const float a0 = (expf(l0 - mx[0]) / sm[0]) * (e0 / (e0 + e1));
      const float a1 = (expf(l1 - mx[1]) / sm[1]) * (e1 / (e0 + e1));
      sAT[(0 * ld_t + e) * RA2_QB + ql] = a0;
      sAT[(1 * ld_t + e) * RA2_QB + ql] = a1;
      if (a0 > best[0]) { best[0] = a0; best_e[0] = e; }
      if (a1 > best[1]) { best[1] = a1; best_e[1] = e; }
      if (att_out && active) { att_out[(qrow * 2 + 0) * att_ld + e] = a0; att_out[(qrow * 2 + 1) * att_ld + e] = a1; }
    }
    if (so_out) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best[r], o);
          const int oe = __shfl_xor_sync(0xffffffffu, best_e[r], o);
          if (ob > best[r] || (ob == best[r] && oe < best_e[r])) { best[r] = ob; best_e[r] = oe; }
        }
      }
      if (lane == 0 && active) { so_out[qrow * 2] = t0 + best_e[0]; so_out[qrow * 2 + 1] = t0 + best_e[1]; }
    }
  }
  // ---- pass 2: hid[q][r*E + c] = ReLU(sum_e att[r][q][e] G[e][r*E + c] + bias[r*E + c]) ----
  const int role = warp >> 2;
  const int colb = role * E + (warp & 3) * (E / 4) + lane * VW;    // this lane's first output column (of 2E)
  float acc2[RA2_QB][VW];
#pragma unroll
  for (int qi = 0; qi < RA2_QB; ++qi)
#pragma unroll
    for (int j = 0; j < VW; ++j) acc2[qi][j] = 0.f;
  for (int c0 = 0; c0 < n; c0 += RA2_CH) {
    const int cn = min(RA2_CH, n - c0);
    __syncthreads();                                               // also orders the sAT writes of every warp before the first read
    for (int i = threadIdx.x; i < cn * (2 * E / 4); i += 256) {
      const int e = i / (2 * E / 4), c = i - e * (2 * E / 4);
      reinterpret_cast<float4*>(sRows)[i] = reinterpret_cast<const float4*>(G + (int64_t)(t0 + c0 + e) * ld_g)[c];
    }
    __syncthreads();
    for (int e = 0; e < cn; ++e) {
      float x[VW];
      load_row<VW>(sRows + e * 2 * E + colb, x);
      float a[RA2_QB];
      load_row<RA2_QB>(sAT + (role * ld_t + c0 + e) * RA2_QB, a);
#pragma unroll
      for (int qi = 0; qi < RA2_QB; ++qi)
#pragma unroll
        for (int j = 0; j < VW; ++j) acc2[qi][j] = fmaf(a[qi], x[j], acc2[qi][j]);
    }
  }
  float b[VW];
  load_row<VW>(bias + colb, b);
#pragma unroll
  for (int qi = 0; qi < RA2_QB; ++qi) {
    if (q0 + qi < Q) {
      float o[VW];
#pragma unroll
      for (int j = 0; j < VW; ++j) o[j] = fmaxf(acc2[qi][j] + b[j], 0.f);
      float* dst = hid + ((int64_t)v * Q + q0 + qi) * (2 * E) + colb;
      if constexpr (VW == 4) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      else {
#pragma unroll
        for (int j = 0; j < VW; ++j) dst[j] = o[j];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Head input: Z[row] = concat of up to 8 pieces, each a (possibly gathered) row of a source matrix
// (prediction_head concat, model_0v10.py:501/503, model_0v7.py:506/508).  idx < 0 => identity row.
// ---------------------------------------------------------------------------------------------------
struct ConcatPiece { const float* src; const int32_t* idx; int idx_stride; int ld; int width; int col0; };
struct ConcatArgs { ConcatPiece p[8]; int n; };

__global__ void gather_concat_kernel(ConcatArgs a, int64_t rows, float* __restrict__ out, int ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += n_warps) {
    for (int k = 0; k < a.n; ++k) {
      const ConcatPiece& pc = a.p[k];
      const int64_t sr = pc.idx ? (int64_t)pc.idx[r * pc.idx_stride] : r;
      const float* s = pc.src + sr * pc.ld;
      float* d = out + r * ldo + pc.col0;
      for (int c = lane; c < pc.width; c += 32) d[c] = s[c];
    }
  }
}

// so (global track ids [rows][2]) -> cat pair index scat*C+ocat and the per-role category ids
__global__ void so_category_kernel(const int32_t* __restrict__ so, const int64_t* __restrict__ cat_ids, int C, int64_t rows,
                                   int32_t* __restrict__ pair_index, int32_t* __restrict__ so_cat) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int sc = (int)cat_ids[so[2 * r]], oc = (int)cat_ids[so[2 * r + 1]];
    pair_index[r] = sc * C + oc;
    so_cat[2 * r] = sc; so_cat[2 * r + 1] = oc;
  }
}

// ---------------------------------------------------------------------------------------------------
// A8 construct_triplet (model_0v10.py:707-785), one CTA per video:
//   softmax over predicate classes, top-k, drop (s,o) without temporal overlap or s == o, dedup the quintuple
//   [pred, scat, ocat, sid, oid] keeping the max predicate probability (first on ties), emit in lexicographic
//   key order (== torch.unique(dim=0)), drop background (pred == 0) last.
// ---------------------------------------------------------------------------------------------------
constexpr int TR_MAX = 2048;  // >= Q * topk

struct TripletOut {
  int64_t* quint;   // [V*cap][5]
  float* scores;    // [V*cap][3]
  int64_t* spans;   // [V*cap][2]
  int64_t* qids;    // [V*cap]
  int32_t* counts;  // [V][2] = (rows emitted, rows that passed the overlap filter)
  int cap;
};

__device__ __forceinline__ bool trip_less(unsigned long long ka, float sa, unsigned ja, unsigned long long kb, float sb, unsigned jb) {
  if (ka != kb) return ka < kb;
  if (sa != sb) return sa > sb;
  return ja < jb;
}

__global__ void __launch_bounds__(512)
construct_triplet_kernel(const float* __restrict__ logits, int ld_logits, int P, int Q, int topk,
                         const int32_t* __restrict__ so, const int32_t* __restrict__ seg, const int64_t* __restrict__ dura,
                         const int64_t* __restrict__ cat_ids, const float* __restrict__ enti_scores, TripletOut out) {
  __shared__ unsigned long long sKey[TR_MAX];
  __shared__ float sScore[TR_MAX];
  __shared__ unsigned sIdx[TR_MAX];
  __shared__ int sFlag[TR_MAX];
  __shared__ int sCount[2];
  const int v = blockIdx.x;
  const int t0 = seg[v];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  const int m = Q * topk;
  if (threadIdx.x < 2) sCount[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < TR_MAX; i += blockDim.x) { sKey[i] = ~0ull; sScore[i] = 0.f; sIdx[i] = 0xffffffffu; }
  __syncthreads();
  // ---- softmax + top-k per query (warp per query, up to 8 classes per lane) ----
  for (int q = warp; q < Q; q += n_warps) {
    const int64_t row = (int64_t)v * Q + q;
    float x[8];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      x[i] = c < P ? logits[row * ld_logits + c] : -INFINITY;
      mx = fmaxf(mx, x[i]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      x[i] = c < P ? expf(x[i] - mx) : -1.f;
      if (c < P) sum += x[i];
    }
    sum = warp_sum(sum);
    const int s_id = so[2 * row] - t0, o_id = so[2 * row + 1] - t0;
    const longlong2 ds = reinterpret_cast<const longlong2*>(dura)[t0 + s_id];
    const longlong2 dz = reinterpret_cast<const longlong2*>(dura)[t0 + o_id];
    const bool overlap = (s_id != o_id) && (max(ds.x, dz.x) <= min(ds.y, dz.y));
    const unsigned long long tail = ((unsigned long long)cat_ids[t0 + s_id] << 36) | ((unsigned long long)cat_ids[t0 + o_id] << 24) |
                                    ((unsigned long long)s_id << 12) | (unsigned long long)o_id;
    for (int k = 0; k < topk; ++k) {
      float bv = -2.f;
      int bc = 0x7fffffff;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = lane + 32 * i;
        if (x[i] > bv) { bv = x[i]; bc = c; }   // ascending c within a lane => first max kept
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
        if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (lane + 32 * i == bc) x[i] = -1.f;
      if (lane == 0 && overlap) {
        const int j = q * topk + k;
        sKey[j] = ((unsigned long long)bc << 48) | tail;
        sScore[j] = bv / sum;
        sIdx[j] = (unsigned)j;
      }
    }
  }
  __syncthreads();
  // ---- bitonic sort of TR_MAX entries by (key asc, score desc, original index asc) ----
  for (int size = 2; size <= TR_MAX; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < TR_MAX / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool asc = ((lo & size) == 0);
        const unsigned long long ka = sKey[lo], kb = sKey[hi];
        const float sa = sScore[lo], sb = sScore[hi];
        const unsigned ja = sIdx[lo], jb = sIdx[hi];
        const bool a_first = trip_less(ka, sa, ja, kb, sb, jb);
        if (a_first != asc) {
          sKey[lo] = kb; sKey[hi] = ka; sScore[lo] = sb; sScore[hi] = sa; sIdx[lo] = jb; sIdx[hi] = ja;
        }
      }
      __syncthreads();
    }
  }
  // ---- run heads (= group winners), background filter, ordered compaction ----
  for (int i = threadIdx.x; i < TR_MAX; i += blockDim.x) {
    const bool valid = sKey[i] != ~0ull;
    const bool head = valid && (i == 0 || sKey[i - 1] != sKey[i]);
    sFlag[i] = (head && (sKey[i] >> 48) != 0) ? 1 : 0;
    if (valid) atomicAdd(&sCount[1], 1);
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of TR_MAX flags by one warp (64 per lane)
    constexpr int PER = TR_MAX / 32;
    int local = 0;
    for (int i = 0; i < PER; ++i) local += sFlag[lane * PER + i];
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    int run = incl - local;
    for (int i = 0; i < PER; ++i) {
      const int f = sFlag[lane * PER + i];
      sFlag[lane * PER + i] = f ? run : -1;
      run += f;
    }
    if (lane == 31) sCount[0] = incl;
  }
  __syncthreads();
  const int64_t base = (int64_t)v * out.cap;
  for (int i = threadIdx.x; i < TR_MAX; i += blockDim.x) {
    const int pos = sFlag[i];
    if (pos < 0) continue;
    const unsigned long long key = sKey[i];
    const int pred = (int)(key >> 48), sc = (int)((key >> 36) & 0xfff), oc = (int)((key >> 24) & 0xfff);
    const int sid = (int)((key >> 12) & 0xfff), oid = (int)(key & 0xfff);
    int64_t* qd = out.quint + (base + pos) * 5;
    qd[0] = pred; qd[1] = sc; qd[2] = oc; qd[3] = sid; qd[4] = oid;
    float* sd = out.scores + (base + pos) * 3;
    sd[0] = sScore[i]; sd[1] = enti_scores[t0 + sid]; sd[2] = enti_scores[t0 + oid];
    const longlong2 ds = reinterpret_cast<const longlong2*>(dura)[t0 + sid];
    const longlong2 dz = reinterpret_cast<const longlong2*>(dura)[t0 + oid];
    out.spans[(base + pos) * 2] = max(ds.x, dz.x);
    out.spans[(base + pos) * 2 + 1] = min(ds.y, dz.y);
    out.qids[base + pos] = sIdx[i] / topk;
  }
  if (threadIdx.x == 0) { out.counts[2 * v] = sCount[0]; out.counts[2 * v + 1] = sCount[1]; }
  (void)m;
}

// ---------------------------------------------------------------------------------------------------
// Cost matrix of the Hungarian matching of the training forward (model_0v10.py:606-636): for query q and GT predicate g
//   cost[q][g] = c_cls * (logsumexp(logit[q]) - logit[q][gt_pred[g]]) + c_adj * mean_{role, tracklet} BCE(att[role][q][e], adj[role][g][e])
// with torch's BCE log clamp at -100.  One CTA per query: block-wide logsumexp of its logit row, then threads over the GT predicates.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
bipartite_cost_kernel(const float* __restrict__ logit, int P, const int64_t* __restrict__ gt_pred, int G, const float* __restrict__ att,
                      const float* __restrict__ adj, int Q, int n, float c_cls, float c_adj, float* __restrict__ cost) {
  __shared__ float red[4];
  const int q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* row = logit + (int64_t)q * P;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < P; c += blockDim.x) m = fmaxf(m, row[c]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float s = 0.f;
  for (int c = threadIdx.x; c < P; c += blockDim.x) s += expf(row[c] - m);
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  const float lse = m + logf((red[0] + red[1]) + (red[2] + red[3]));
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < 2; ++r) {
      const float* a = att + ((int64_t)r * Q + q) * n;
      const float* t = adj + ((int64_t)r * G + g) * n;
      for (int e = 0; e < n; ++e) {
        const float p = a[e], y = t[e];
        acc -= y * fmaxf(logf(p), -100.f) + (1.f - y) * fmaxf(logf(1.f - p), -100.f);
      }
    }
    cost[(int64_t)q * G + g] = c_cls * (lse - row[gt_pred[g]]) + c_adj * (acc / (float)(2 * n));
  }
}

static inline int grid_cap(int64_t blocks, int per_sm) {
  const int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace vsg

using namespace vsg;

extern "C" int vsg_bbox_feat_mlp1(const float* boxes, const int64_t* off, int n_tracks, int64_t n_rows, const int32_t* track_vid,
                                  const float* wh, const float* W1, const float* b1, int E, float* out, int ldo, float* feat8_out,
                                  void* stream) {
  VSG_REQUIRE(n_tracks >= 0 && n_rows >= 0 && E > 0, "vsg_bbox_feat_mlp1: bad size");
  if (n_rows == 0) return VSG_OK;
  VSG_REQUIRE(boxes && off && track_vid && wh && W1 && b1 && out && aligned16(boxes), "vsg_bbox_feat_mlp1: null/misaligned pointer");
  const size_t smem = (size_t)9 * E * sizeof(float);
  VSG_REQUIRE(smem <= 48 * 1024, "vsg_bbox_feat_mlp1: E too large");
  bbox_feat_mlp1_kernel<<<grid_cap((n_rows + 7) / 8, 8), 256, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(boxes), off, n_tracks, n_rows, track_vid, wh, W1, b1, E, out, ldo, feat8_out);
  return check_launch("vsg_bbox_feat_mlp1");
}

extern "C" int vsg_stretched_mean(const float* feat, int ldf, int col0, int width, const int64_t* off, const int32_t* tmax,
                                  int n_tracks, float* out, int ldo, void* stream) {
  VSG_REQUIRE(n_tracks >= 0 && width >= 0, "vsg_stretched_mean: bad size");
  if (n_tracks == 0 || width == 0) return VSG_OK;
  VSG_REQUIRE(feat && off && tmax && out, "vsg_stretched_mean: null pointer");
  stretched_mean_kernel<float><<<n_tracks, 256, 0, (cudaStream_t)stream>>>(feat, ldf, col0, width, off, tmax, out, ldo);
  return check_launch("vsg_stretched_mean");
}

extern "C" int vsg_stretched_mean_bf16(const void* feat16, int ldf, int col0, int width, const int64_t* off, const int32_t* tmax,
                                       int n_tracks, float* out, int ldo, void* stream) {
  VSG_REQUIRE(n_tracks >= 0 && width >= 0, "vsg_stretched_mean_bf16: bad size");
  if (n_tracks == 0 || width == 0) return VSG_OK;
  VSG_REQUIRE(feat16 && off && tmax && out, "vsg_stretched_mean_bf16: null pointer");
  stretched_mean_kernel<__nv_bfloat16><<<n_tracks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)feat16, ldf, col0, width, off, tmax, out, ldo);
  return check_launch("vsg_stretched_mean_bf16");
}

extern "C" int vsg_conv_pool(const float* Y, int ldy, int E, const float* bias, const int64_t* off, const int32_t* tmax,
                             int n_tracks, int pool, float* out, void* stream) {
  VSG_REQUIRE(n_tracks >= 0 && pool > 0 && E > 0 && E % 4 == 0 && ldy % 4 == 0, "vsg_conv_pool: bad size");
  if (n_tracks == 0) return VSG_OK;
  VSG_REQUIRE(Y && bias && off && tmax && out && aligned16(Y) && aligned16(bias), "vsg_conv_pool: null/misaligned pointer");
  conv_pool_kernel<float><<<n_tracks * pool, 128, 0, (cudaStream_t)stream>>>(Y, ldy, E, bias, off, tmax, pool, out);
  return check_launch("vsg_conv_pool");
}

extern "C" int vsg_conv_pool_bf16(const void* Y16, int ldy, int E, const float* bias, const int64_t* off, const int32_t* tmax,
                                  int n_tracks, int pool, float* out, void* stream) {
  VSG_REQUIRE(n_tracks >= 0 && pool > 0 && E > 0 && E % 4 == 0 && ldy % 4 == 0, "vsg_conv_pool_bf16: bad size");
  if (n_tracks == 0) return VSG_OK;
  VSG_REQUIRE(Y16 && bias && off && tmax && out && (reinterpret_cast<uintptr_t>(Y16) & 7) == 0 && aligned16(bias), "vsg_conv_pool_bf16: null/misaligned pointer");
  conv_pool_kernel<__nv_bfloat16><<<n_tracks * pool, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)Y16, ldy, E, bias, off, tmax, pool, out);
  return check_launch("vsg_conv_pool_bf16");
}

extern "C" int vsg_bbox_feat_mlp1_bf16(const float* boxes, const int64_t* off, int n_tracks, int64_t n_rows, const int32_t* track_vid,
                                       const float* wh, const float* W1, const float* b1, int E, void* out16, int ldo, void* stream) {
  VSG_REQUIRE(n_rows >= 0 && n_tracks >= 0 && E > 0 && ldo >= E, "vsg_bbox_feat_mlp1_bf16: bad size");
  if (n_rows == 0) return VSG_OK;
  VSG_REQUIRE(boxes && off && track_vid && wh && W1 && b1 && out16 && aligned16(boxes), "vsg_bbox_feat_mlp1_bf16: null/misaligned pointer");
  const size_t smem = (size_t)9 * E * sizeof(float);
  VSG_REQUIRE(smem <= 48 * 1024, "vsg_bbox_feat_mlp1_bf16: E too large");
  bbox_feat_mlp1_kernel<<<grid_cap((n_rows + 7) / 8, 8), 256, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(boxes), off, n_tracks, n_rows, track_vid, wh, W1, b1, E, nullptr, ldo, nullptr, (__nv_bfloat16*)out16);
  return check_launch("vsg_bbox_feat_mlp1_bf16");
}

static int add_layernorm_impl(const float* x, int ldx, const float* a, int lda, const float* gamma, const float* beta, const float* post,
                              int post_period, int64_t rows, int D, float* out, int ldo, float* out2, int ldo2, void* stream);

extern "C" int vsg_add_layernorm(const float* x, int ldx, const float* a, int lda, const float* gamma, const float* beta,
                                 const float* post, int post_period, int64_t rows, int D, float* out, int ldo, void* stream) {
  return add_layernorm_impl(x, ldx, a, lda, gamma, beta, post, post_period, rows, D, out, ldo, nullptr, 0, stream);
}

extern "C" int vsg_add_layernorm_dual(const float* x, int ldx, const float* a, int lda, const float* gamma, const float* beta,
                                      const float* post, int post_period, int64_t rows, int D, float* out, int ldo, float* out2, int ldo2,
                                      void* stream) {
  VSG_REQUIRE(rows == 0 || (post && out2), "vsg_add_layernorm_dual: post and out2 are required");
  return add_layernorm_impl(x, ldx, a, lda, gamma, beta, post, post_period, rows, D, out, ldo, out2, ldo2, stream);
}

static int add_layernorm_impl(const float* x, int ldx, const float* a, int lda, const float* gamma, const float* beta, const float* post,
                              int post_period, int64_t rows, int D, float* out, int ldo, float* out2, int ldo2, void* stream) {
  VSG_REQUIRE(rows >= 0 && D > 0 && D % 32 == 0 && D <= 1024, "vsg_add_layernorm: D must be a multiple of 32, <= 1024");
  if (rows == 0) return VSG_OK;
  VSG_REQUIRE(x && gamma && beta && out, "vsg_add_layernorm: null pointer");
  VSG_REQUIRE(post == nullptr || post_period > 0, "vsg_add_layernorm: post needs a period");
  const int g = grid_cap((rows + 7) / 8, 8);
  if (D <= 128)
    add_layernorm_kernel<4><<<g, 256, 0, (cudaStream_t)stream>>>(x, ldx, a, lda, gamma, beta, post, post_period, rows, D, out, ldo, out2, ldo2);
  else if (D <= 512)
    add_layernorm_kernel<16><<<g, 256, 0, (cudaStream_t)stream>>>(x, ldx, a, lda, gamma, beta, post, post_period, rows, D, out, ldo, out2, ldo2);
  else
    add_layernorm_kernel<32><<<g, 256, 0, (cudaStream_t)stream>>>(x, ldx, a, lda, gamma, beta, post, post_period, rows, D, out, ldo, out2, ldo2);
  return check_launch("vsg_add_layernorm");
}

extern "C" int vsg_broadcast_rows(const float* x, int period, int D, int64_t rows, float* out, void* stream) {
  VSG_REQUIRE(rows >= 0 && period > 0 && D > 0 && D % 4 == 0, "vsg_broadcast_rows: bad size");
  if (rows == 0) return VSG_OK;
  VSG_REQUIRE(x && out && aligned16(x) && aligned16(out), "vsg_broadcast_rows: null/misaligned pointer");
  broadcast_rows_kernel<<<grid_cap((rows * (D / 4) + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(x, period, D, rows, out);
  return check_launch("vsg_broadcast_rows");
}

// Very short sequences (the grounding query encoder attends over the 3 words of a query, grd_model_v5.py:342): one THREAD per
// (sequence, query position, head) -- scores, softmax and the weighted sum of <= 8 keys in registers.  The tiled kernel would stage
// 90 KB of shared memory per 3-token sequence (measured: 0.53 ms for 19k sequences at 25 % occupancy).
__global__ void __launch_bounds__(256)
mha_tiny16_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ K, int ldk, const float* __restrict__ V, int ldv,
                  const int64_t* __restrict__ seg_off, int n_seg, int max_len, int n_head, float scale, float* __restrict__ O, int ldo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int h = (int)(i % n_head);
  const int64_t t = i / n_head;
  const int qi = (int)(t % max_len);
  const int64_t seg = t / max_len;
  if (seg >= n_seg) return;
  const int64_t r0 = seg_off[seg];
  const int L = (int)(seg_off[seg + 1] - r0);
  if (qi >= L) return;
  float q[16];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 x = *reinterpret_cast<const float4*>(Q + (r0 + qi) * (int64_t)ldq + h * 16 + 4 * c);
    q[4 * c] = x.x; q[4 * c + 1] = x.y; q[4 * c + 2] = x.z; q[4 * c + 3] = x.w;
  }
  float sc[8];
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = -INFINITY;
    if (j < L) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 x = *reinterpret_cast<const float4*>(K + (r0 + j) * (int64_t)ldk + h * 16 + 4 * c);
        d = fmaf(q[4 * c], x.x, d); d = fmaf(q[4 * c + 1], x.y, d); d = fmaf(q[4 * c + 2], x.z, d); d = fmaf(q[4 * c + 3], x.w, d);
      }
      sc[j] = d * scale;
      m = fmaxf(m, sc[j]);
    }
  }
  float sum = 0.f;
  float o[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) o[c] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j < L) {
      const float p = expf(sc[j] - m);
      sum += p;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 x = *reinterpret_cast<const float4*>(V + (r0 + j) * (int64_t)ldv + h * 16 + 4 * c);
        o[4 * c] = fmaf(p, x.x, o[4 * c]); o[4 * c + 1] = fmaf(p, x.y, o[4 * c + 1]);
        o[4 * c + 2] = fmaf(p, x.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(p, x.w, o[4 * c + 3]);
      }
    }
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<float4*>(O + (r0 + qi) * (int64_t)ldo + h * 16 + 4 * c) =
        make_float4(o[4 * c] * inv, o[4 * c + 1] * inv, o[4 * c + 2] * inv, o[4 * c + 3] * inv);
}

extern "C" int vsg_mha(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, const int64_t* seg_off, int n_seg,
                       int fixed_len, int max_len, int n_head, int head_dim, float* O, int ldo, const int32_t* blk_seg,
                       const int32_t* blk_q0, int n_blocks, void* stream) {
  VSG_REQUIRE(n_seg >= 0 && n_head > 0 && max_len >= 0, "vsg_mha: bad size");
  if (n_seg == 0 || max_len == 0) return VSG_OK;
  VSG_REQUIRE(Q && K && V && O, "vsg_mha: null pointer");
  VSG_REQUIRE(seg_off != nullptr || fixed_len > 0, "vsg_mha: need seg_off or fixed_len");
  const float scale = 1.0f / sqrtf((float)head_dim);
  VSG_REQUIRE((ldq % 4) == 0 && (ldk % 4) == 0 && (ldv % 4) == 0 && aligned16(Q) && aligned16(K) && aligned16(V),
              "vsg_mha: Q/K/V must be 16-byte aligned with leading dimensions that are multiples of 4");
  VSG_REQUIRE((blk_seg == nullptr) == (blk_q0 == nullptr), "vsg_mha: blk_seg and blk_q0 go together");
  if (blk_seg && n_blocks == 0) return VSG_OK;
  if (head_dim == 16 && seg_off != nullptr && max_len <= 8 && (ldo % 4) == 0 && aligned16(O)) {
    const int64_t threads = (int64_t)n_seg * max_len * n_head;
    mha_tiny16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(Q, ldq, K, ldk, V, ldv, seg_off, n_seg, max_len,
                                                                                          n_head, scale, O, ldo);
    return check_launch("vsg_mha(tiny)");
  }
  dim3 grid(blk_seg ? n_blocks : n_seg, n_head, blk_seg ? 1 : (max_len + MHA_QB - 1) / MHA_QB);
  const int sk = head_dim >= 64 ? 128 : 256;
  const size_t smem = (size_t)(MHA_QB * head_dim + sk * head_dim + sk * MHA_PITCH) * sizeof(float);
#define VSG_MHA_LAUNCH(DH_)                                                                                              \
  do {                                                                                                                   \
    static PerDeviceFlag attr_done;                                                                                      \
    const int dev_ = current_device();                                                                                   \
    if (!attr_done.is_set(dev_)) {                                                                                       \
      if (cudaFuncSetAttribute(mha_kernel<DH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {  \
        set_error("vsg_mha: cannot raise dynamic shared memory to %zu", smem);                                           \
        return VSG_E_LAUNCH;                                                                                             \
      }                                                                                                                  \
      attr_done.set(dev_);                                                                                               \
    }                                                                                                                    \
    mha_kernel<DH_><<<grid, 256, smem, (cudaStream_t)stream>>>(Q, ldq, K, ldk, V, ldv, seg_off, fixed_len, n_head, scale, O, ldo, blk_seg, blk_q0); \
  } while (0)
  if (head_dim == 64) VSG_MHA_LAUNCH(64);
  else if (head_dim == 32) VSG_MHA_LAUNCH(32);
  else if (head_dim == 16) VSG_MHA_LAUNCH(16);
  else { set_error("vsg_mha: head_dim %d unsupported (16, 32, 64)", head_dim); return VSG_E_UNSUPPORTED; }
#undef VSG_MHA_LAUNCH
  return check_launch("vsg_mha");
}

extern "C" int vsg_softmax_rows(float* S, int ld, int n, int64_t rows, float scale, void* stream) {
  VSG_REQUIRE(rows >= 0 && n >= 1 && n <= 256 && ld >= n, "vsg_softmax_rows: 1 <= n <= 256 <= ld required");
  if (rows == 0) return VSG_OK;
  VSG_REQUIRE(S != nullptr, "vsg_softmax_rows: null pointer");
  softmax_rows_kernel<<<grid_cap((rows + 7) / 8, 8), 256, 0, (cudaStream_t)stream>>>(S, ld, n, rows, scale);
  return check_launch("vsg_softmax_rows");
}

extern "C" int vsg_transpose_split(const float* X, int ld, int64_t rows, int cols, float* T_hi, float* T_lo, int64_t ld_t, void* stream) {
  VSG_REQUIRE(rows >= 0 && cols >= 0 && ld >= cols && ld_t >= rows, "vsg_transpose_split: bad sizes");
  if (rows == 0 || cols == 0) return VSG_OK;
  VSG_REQUIRE(X && T_hi, "vsg_transpose_split: null pointer");
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
  transpose_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ld, rows, cols, T_hi, T_lo, ld_t);
  return check_launch("vsg_transpose_split");
}

extern "C" int vsg_role_attention_hid(const float* p2a, const float* e2a, int ld_e2a, const float* G, int ld_g, const float* bias,
                                      const int32_t* seg, int n_vid, int Q, int E, int max_tracks, float inv_sqrt_d, float* hid, float* att_out,
                                      int att_ld, int32_t* so_out, void* stream) {
  VSG_REQUIRE(n_vid >= 0 && Q > 0, "vsg_role_attention_hid: bad size");
  if (n_vid == 0) return VSG_OK;
  VSG_REQUIRE(p2a && e2a && G && bias && seg && hid, "vsg_role_attention_hid: null pointer");
  VSG_REQUIRE(E == 128 || E == 512, "vsg_role_attention_hid: dim %d unsupported (128, 512)", E);
  VSG_REQUIRE(max_tracks >= 0 && max_tracks <= RA_MAX_TRACKS, "vsg_role_attention_hid: more than %d tracks in a video", RA_MAX_TRACKS);
  VSG_REQUIRE(aligned16(p2a) && aligned16(e2a) && aligned16(G) && aligned16(bias) && aligned16(hid) && ld_e2a % 4 == 0 && ld_g % 4 == 0 &&
              ld_e2a >= E && ld_g >= 2 * E, "vsg_role_attention_hid: misaligned pointer or leading dimension");
  const int ld_t = ((max_tracks > 0 ? max_tracks : 1) + 3) / 4 * 4;
  const size_t smem = (size_t)(RA2_CH * 2 * E + 2 * RA2_QB * 2 * ld_t) * sizeof(float);
  dim3 grid(n_vid, (Q + RA2_QB - 1) / RA2_QB);
  static PerDeviceFlag attr_done[2];
  const int dev_ = current_device();
  auto raise = [&](const void* fn, PerDeviceFlag& flag) -> int {
    if (!flag.is_set(dev_)) {
      if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((RA2_CH * 2 * 512 + 2 * RA2_QB * 2 * RA_MAX_TRACKS) * sizeof(float))) != cudaSuccess) {
        set_error("vsg_role_attention_hid: cannot raise dynamic shared memory");
        return VSG_E_LAUNCH;
      }
      flag.set(dev_);
    }
    return VSG_OK;
  };
  if (E == 512) {
    if (raise((const void*)role_attention_hid_kernel<512>, attr_done[0]) != VSG_OK) return VSG_E_LAUNCH;
    role_attention_hid_kernel<512><<<grid, 256, smem, (cudaStream_t)stream>>>(p2a, e2a, ld_e2a, G, ld_g, bias, seg, Q, ld_t, inv_sqrt_d, hid, att_out,
                                                                             att_ld, so_out);
  } else {
    if (raise((const void*)role_attention_hid_kernel<128>, attr_done[1]) != VSG_OK) return VSG_E_LAUNCH;
    role_attention_hid_kernel<128><<<grid, 256, smem, (cudaStream_t)stream>>>(p2a, e2a, ld_e2a, G, ld_g, bias, seg, Q, ld_t, inv_sqrt_d, hid, att_out,
                                                                             att_ld, so_out);
  }
  return check_launch("vsg_role_attention_hid");
}

extern "C" int vsg_role_attention(const float* p2a, const float* e2a, const float* enco, const int32_t* seg, int n_vid, int Q, int E,
                                  int max_tracks, float inv_sqrt_d, float* values, float* att_out, int att_ld, int32_t* so_out,
                                  void* stream) {
  VSG_REQUIRE(n_vid >= 0 && Q > 0, "vsg_role_attention: bad size");
  if (n_vid == 0) return VSG_OK;
  VSG_REQUIRE(p2a && e2a && enco && seg && values, "vsg_role_attention: null pointer");
  VSG_REQUIRE(max_tracks <= RA_MAX_TRACKS, "vsg_role_attention: more than %d tracks in a video", RA_MAX_TRACKS);
  VSG_REQUIRE(aligned16(e2a) && aligned16(enco) && aligned16(values), "vsg_role_attention: misaligned pointer");
  dim3 grid(n_vid, (Q + 7) / 8);
  VSG_REQUIRE(aligned16(p2a), "vsg_role_attention: misaligned p2a");
  const size_t smem = (size_t)(RA_CH * E + 8 * 2 * RA_MAX_TRACKS) * sizeof(float);
  if (E == 512) {
    static PerDeviceFlag attr_done;
    const int dev_ = current_device();
    if (!attr_done.is_set(dev_)) {
      if (cudaFuncSetAttribute(role_attention_smem_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        set_error("vsg_role_attention: cannot raise dynamic shared memory to %zu", smem);
        return VSG_E_LAUNCH;
      }
      attr_done.set(dev_);
    }
    role_attention_smem_kernel<512><<<grid, 256, smem, (cudaStream_t)stream>>>(p2a, e2a, enco, seg, Q, inv_sqrt_d, values, att_out, att_ld, so_out);
  } else if (E == 128) role_attention_smem_kernel<128><<<grid, 256, smem, (cudaStream_t)stream>>>(p2a, e2a, enco, seg, Q, inv_sqrt_d, values, att_out, att_ld, so_out);
  else if (E == 64) role_attention_kernel<64><<<grid, 256, 0, (cudaStream_t)stream>>>(p2a, e2a, enco, seg, Q, inv_sqrt_d, values, att_out, att_ld, so_out);
  else { set_error("vsg_role_attention: dim %d unsupported (64, 128, 512)", E); return VSG_E_UNSUPPORTED; }
  return check_launch("vsg_role_attention");
}

extern "C" int vsg_gather_concat(const float* const* src, const int32_t* const* idx, const int* idx_stride, const int* ld,
                                 const int* width, int n_pieces, int64_t rows, float* out, int ldo, void* stream) {
  VSG_REQUIRE(n_pieces >= 1 && n_pieces <= 8 && rows >= 0, "vsg_gather_concat: 1..8 pieces");
  if (rows == 0) return VSG_OK;
  ConcatArgs a;
  a.n = n_pieces;
  int col = 0;
  for (int k = 0; k < n_pieces; ++k) {
    VSG_REQUIRE(src[k] != nullptr, "vsg_gather_concat: null source");
    a.p[k].src = src[k]; a.p[k].idx = idx[k]; a.p[k].idx_stride = idx_stride[k]; a.p[k].ld = ld[k]; a.p[k].width = width[k];
    a.p[k].col0 = col;
    col += width[k];
  }
  VSG_REQUIRE(col <= ldo, "vsg_gather_concat: pieces wider than the output row");
  gather_concat_kernel<<<grid_cap((rows + 7) / 8, 8), 256, 0, (cudaStream_t)stream>>>(a, rows, out, ldo);
  return check_launch("vsg_gather_concat");
}

extern "C" int vsg_so_category(const int32_t* so, const int64_t* cat_ids, int C, int64_t rows, int32_t* pair_index, int32_t* so_cat,
                               void* stream) {
  VSG_REQUIRE(rows >= 0 && C > 0, "vsg_so_category: bad size");
  if (rows == 0) return VSG_OK;
  VSG_REQUIRE(so && cat_ids && pair_index && so_cat, "vsg_so_category: null pointer");
  so_category_kernel<<<grid_cap((rows + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(so, cat_ids, C, rows, pair_index, so_cat);
  return check_launch("vsg_so_category");
}

extern "C" int vsg_construct_triplet(const float* logits, int ld_logits, int P, int Q, int topk, const int32_t* so, const int32_t* seg,
                                     int n_vid, const int64_t* dura, const int64_t* cat_ids, const float* enti_scores,
                                     int64_t* quint, float* scores, int64_t* spans, int64_t* qids, int32_t* counts, int cap,
                                     void* stream) {
  VSG_REQUIRE(n_vid >= 0 && P > 0 && P <= 256 && Q > 0 && topk > 0 && topk <= P, "vsg_construct_triplet: bad size (P <= 256)");
  VSG_REQUIRE(Q * topk <= TR_MAX && cap >= Q * topk, "vsg_construct_triplet: Q*topk must be <= %d and <= cap", TR_MAX);
  if (n_vid == 0) return VSG_OK;
  VSG_REQUIRE(logits && so && seg && dura && cat_ids && enti_scores && quint && scores && spans && qids && counts,
              "vsg_construct_triplet: null pointer");
  VSG_REQUIRE(aligned16(dura), "vsg_construct_triplet: spans misaligned");
  TripletOut o{quint, scores, spans, qids, counts, cap};
  construct_triplet_kernel<<<n_vid, 512, 0, (cudaStream_t)stream>>>(logits, ld_logits, P, Q, topk, so, seg, dura, cat_ids,
                                                                   enti_scores, o);
  return check_launch("vsg_construct_triplet");
}

extern "C" int vsg_bipartite_cost(const float* logit, int Q, int P, const int64_t* gt_pred, int G, const float* att, const float* adj, int n,
                                  float c_cls, float c_adj, float* cost, void* stream) {
  VSG_REQUIRE(Q >= 0 && P > 0 && G >= 0 && n > 0, "vsg_bipartite_cost: bad size");
  if (Q == 0 || G == 0) return VSG_OK;
  VSG_REQUIRE(logit && gt_pred && att && adj && cost, "vsg_bipartite_cost: null pointer");
  bipartite_cost_kernel<<<Q, 128, 0, (cudaStream_t)stream>>>(logit, P, gt_pred, G, att, adj, Q, n, c_cls, c_adj, cost);
  return check_launch("vsg_bipartite_cost");
}
