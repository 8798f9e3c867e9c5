// Base-C pairwise baseline (reference models/model_pairwise_baseline.py) -- the kernels that are specific to it; the per-track
// encoding and the pair-MLP GEMMs are the BIG-C kernels (bigc.cu, gemm.cu).
//   pair_ids_batched_kernel   trajid2pairid (:104-111) for every video of a batch, global track ids
//   pair_topk_kernel          construct_triplet (:314-352): softmax + top-k per pair, temporal-overlap filter, candidate keys
//   pair_rank_kernel          construct_triplet (:353-395): lexicographic order of the (unique) quintuples by rank counting, background
//                             removed, optionally the rt_triplets_topk best by mean score
#include "common.cuh"
#include <math.h>

namespace vsg {

__global__ void pair_ids_batched_kernel(const int32_t* __restrict__ seg, int V, const int64_t* __restrict__ pair_off,
                                        int32_t* __restrict__ so, int32_t* __restrict__ pair_vid) {
  const int64_t total = pair_off[V];
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
    const int v = find_segment(pair_off, V, p);
    const int n = seg[v + 1] - seg[v];
    const int q = (int)(p - pair_off[v]);
    const int s = q / (n - 1), r = q % (n - 1);
    const int o = r + (r >= s ? 1 : 0);                 // row-major nonzero() of the off-diagonal mask
    so[2 * p] = seg[v] + s;
    so[2 * p + 1] = seg[v] + o;
    if (pair_vid) pair_vid[p] = v;
  }
}

// One warp per pair.  Candidate j of pair p lives at p * topk + j: key = pred<<48 | scat<<36 | ocat<<24 | sid<<12 | oid (local track ids)
// or ~0 when the pair's spans do not overlap; score3 = (pred prob, subject score, object score).
// counts[v] = {#valid candidates, #valid candidates with pred == 0}.
__global__ void __launch_bounds__(256)
pair_topk_kernel(const float* __restrict__ logits, int ld, int P, int topk, const int32_t* __restrict__ so,
                 const int32_t* __restrict__ pair_vid, const int32_t* __restrict__ seg, const int64_t* __restrict__ dura,
                 const int64_t* __restrict__ cat_ids, const float* __restrict__ enti_scores, int64_t n_pairs,
                 unsigned long long* __restrict__ keys, float* __restrict__ score3, int32_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < n_pairs; p += n_warps) {
    const int s = so[2 * p], o = so[2 * p + 1], v = pair_vid[p], t0 = seg[v];
    const longlong2 ds = reinterpret_cast<const longlong2*>(dura)[s];
    const longlong2 dz = reinterpret_cast<const longlong2*>(dura)[o];
    const bool overlap = max(ds.x, dz.x) <= min(ds.y, dz.y);
    if (!overlap) {
      if (lane < topk) keys[p * topk + lane] = ~0ull;
      continue;
    }
    float x[8];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      x[i] = c < P ? logits[p * ld + c] : -INFINITY;
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
    const unsigned long long tail = ((unsigned long long)cat_ids[s] << 36) | ((unsigned long long)cat_ids[o] << 24) |
                                    ((unsigned long long)(s - t0) << 12) | (unsigned long long)(o - t0);
    int n_bg = 0;
    for (int k = 0; k < topk; ++k) {
      float bv = -2.f;
      int bc = 0x7fffffff;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = lane + 32 * i;
        if (x[i] > bv) { bv = x[i]; bc = c; }   // ascending c within a lane => first max kept
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const int oc = __shfl_xor_sync(0xffffffffu, bc, off);
        if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (lane + 32 * i == bc) x[i] = -1.f;
      if (lane == 0) {
        const int64_t j = p * topk + k;
        keys[j] = ((unsigned long long)bc << 48) | tail;
        score3[3 * j] = bv / sum;
        score3[3 * j + 1] = enti_scores[s];
        score3[3 * j + 2] = enti_scores[o];
        n_bg += bc == 0;
      }
    }
    if (lane == 0) {
      atomicAdd(&counts[2 * v], topk);
      if (n_bg) atomicAdd(&counts[2 * v + 1], n_bg);
    }
  }
}

// Block = 256 candidates of one video (chunk list: chunk_off[v] .. chunk_off[v+1]).  Keys of valid candidates are unique, so the
// lexicographic position of a candidate is the number of smaller keys; background (pred 0) keys are the smallest and are dropped.
// rt_topk > 0: position = rank by (mean score descending, key ascending) among the foreground candidates, kept when < rt_topk.
constexpr int RK_TILE = 1024;
__global__ void __launch_bounds__(256)
pair_rank_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ score3, const int64_t* __restrict__ cand_off,
                 const int32_t* __restrict__ chunk_off, int V, const int32_t* __restrict__ counts, const int32_t* __restrict__ seg,
                 const int64_t* __restrict__ dura, int rt_topk, int64_t* __restrict__ quint, float* __restrict__ scores_out,
                 int64_t* __restrict__ spans_out) {
  __shared__ unsigned long long sKey[RK_TILE];
  __shared__ float sMean[RK_TILE];
  // video of this chunk
  int lo = 0, hi = V;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (chunk_off[mid] <= (int)blockIdx.x) lo = mid; else hi = mid; }
  const int v = lo;
  const int64_t c0 = cand_off[v], m = cand_off[v + 1] - c0;
  const int64_t i = (int64_t)(blockIdx.x - chunk_off[v]) * 256 + threadIdx.x;
  const bool have = i < m;
  const unsigned long long my = have ? keys[c0 + i] : ~0ull;
  const bool valid = my != ~0ull;
  float3 sc = make_float3(0.f, 0.f, 0.f);
  if (valid) sc = make_float3(score3[3 * (c0 + i)], score3[3 * (c0 + i) + 1], score3[3 * (c0 + i) + 2]);
  const float my_mean = (sc.x + sc.y + sc.z) / 3.0f;
  const bool fg = valid && (my >> 48) != 0;
  int rank = 0;
  for (int64_t t = 0; t < m; t += RK_TILE) {
    __syncthreads();
    for (int j = threadIdx.x; j < RK_TILE; j += 256) {
      const int64_t g = t + j;
      const unsigned long long k = g < m ? keys[c0 + g] : ~0ull;
      sKey[j] = k;
      if (rt_topk > 0) {
        float mean = -INFINITY;
        if (k != ~0ull && (k >> 48) != 0) mean = (score3[3 * (c0 + g)] + score3[3 * (c0 + g) + 1] + score3[3 * (c0 + g) + 2]) / 3.0f;
        sMean[j] = mean;
      }
    }
    __syncthreads();
    if (fg) {
      if (rt_topk > 0) {
#pragma unroll 8
        for (int j = 0; j < RK_TILE; ++j) {
          const float mj = sMean[j];
          rank += (mj > my_mean) || (mj == my_mean && sKey[j] < my);
        }
      } else {
#pragma unroll 8
        for (int j = 0; j < RK_TILE; ++j) rank += sKey[j] < my;
      }
    }
  }
  if (!fg) return;
  int pos;
  if (rt_topk > 0) {
    if (rank >= rt_topk) return;
    pos = rank;
  } else {
    pos = rank - counts[2 * v + 1];          // skip the background rows, which sort first
  }
  const int64_t o = c0 + pos;
  const int sid = (int)((my >> 12) & 0xFFF), oid = (int)(my & 0xFFF);
  quint[5 * o] = (int64_t)(my >> 48);
  quint[5 * o + 1] = (int64_t)((my >> 36) & 0xFFF);
  quint[5 * o + 2] = (int64_t)((my >> 24) & 0xFFF);
  quint[5 * o + 3] = sid;
  quint[5 * o + 4] = oid;
  scores_out[3 * o] = sc.x; scores_out[3 * o + 1] = sc.y; scores_out[3 * o + 2] = sc.z;
  const int t0 = seg[v];
  const longlong2 ds = reinterpret_cast<const longlong2*>(dura)[t0 + sid];
  const longlong2 dz = reinterpret_cast<const longlong2*>(dura)[t0 + oid];
  spans_out[2 * o] = max(ds.x, dz.x);
  spans_out[2 * o + 1] = min(ds.y, dz.y);
}

}  // namespace vsg

using namespace vsg;

extern "C" int vsg_pair_ids_batched(const int32_t* seg, int n_vid, const int64_t* pair_off, int64_t n_pairs, int32_t* so,
                                    int32_t* pair_vid, void* stream) {
  VSG_REQUIRE(n_vid >= 0 && n_pairs >= 0, "vsg_pair_ids_batched: negative size");
  if (n_pairs == 0) return VSG_OK;
  VSG_REQUIRE(seg && pair_off && so, "vsg_pair_ids_batched: null pointer");
  int64_t blocks = (n_pairs + 255) / 256;
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  pair_ids_batched_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(seg, n_vid, pair_off, so, pair_vid);
  return check_launch("vsg_pair_ids_batched");
}

extern "C" int vsg_pair_construct_triplet(const float* logits, int ld_logits, int P, int topk, const int32_t* so, const int32_t* pair_vid,
                                          int64_t n_pairs, const int32_t* seg, int n_vid, const int64_t* dura, const int64_t* cat_ids,
                                          const float* enti_scores, const int64_t* cand_off, const int32_t* chunk_off, int n_chunks,
                                          int rt_topk, unsigned long long* keys_ws, float* score_ws, int32_t* counts, int64_t* quint,
                                          float* scores, int64_t* spans, void* stream) {
  VSG_REQUIRE(P >= 1 && P <= 256, "vsg_pair_construct_triplet: 1 <= num_pred_cats <= 256 (got %d)", P);
  VSG_REQUIRE(topk >= 1 && topk <= 32 && topk <= P, "vsg_pair_construct_triplet: 1 <= topk <= min(32, P) (got %d)", topk);
  VSG_REQUIRE(n_pairs >= 0 && n_vid >= 0, "vsg_pair_construct_triplet: negative size");
  VSG_REQUIRE(counts != nullptr || n_vid == 0, "vsg_pair_construct_triplet: null counts");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_vid > 0) cudaMemsetAsync(counts, 0, sizeof(int32_t) * 2 * (size_t)n_vid, st);
  if (n_pairs == 0) return VSG_OK;
  VSG_REQUIRE(logits && so && pair_vid && seg && dura && cat_ids && enti_scores && cand_off && chunk_off && keys_ws && score_ws && quint &&
              scores && spans, "vsg_pair_construct_triplet: null pointer");
  VSG_REQUIRE(aligned16(dura), "vsg_pair_construct_triplet: spans misaligned");
  int64_t blocks = (n_pairs + 7) / 8;
  if (blocks > (int64_t)sm_count() * 32) blocks = (int64_t)sm_count() * 32;
  pair_topk_kernel<<<(int)blocks, 256, 0, st>>>(logits, ld_logits, P, topk, so, pair_vid, seg, dura, cat_ids, enti_scores, n_pairs, keys_ws,
                                                score_ws, counts);
  if (n_chunks > 0)
    pair_rank_kernel<<<n_chunks, 256, 0, st>>>(keys_ws, score_ws, cand_off, chunk_off, n_vid, counts, seg, dura, rt_topk, quint, scores, spans);
  return check_launch("vsg_pair_construct_triplet", 2);
}
