// Evaluation kernels: relation volume IoU and greedy prediction->GT matching for all videos at once.
// SURVEY.md section 8a rows A12 (VidVRDhelperEvalAPIs/common.py:65-106 viou) and A13
// (visual_relation_detection.py:7-34 eval_detection_scores, :124-156 _v2).
//
// Layout: a relation row is int64[7] = (s_cat, p_cat, o_cat, sub_track, obj_track, start, end) with a
// half-open duration; its trajectories are slices of a CSR track table, so predictions that share
// proposal tracks never duplicate boxes in HBM.  All sums are fp64 (the reference is Python float).
#include "common.cuh"
#include <math.h>

namespace vsg {

template <typename TB> struct Box4;
template <> struct Box4<float> { using type = float4; };
template <> struct Box4<double> { using type = double4; };

template <typename TB>
__device__ __forceinline__ void load_box(const TB* base, int64_t row, double& x1, double& y1, double& x2, double& y2);
template <>
__device__ __forceinline__ void load_box<float>(const float* base, int64_t row, double& x1, double& y1, double& x2, double& y2) {
  const float4 b = reinterpret_cast<const float4*>(base)[row];
  x1 = b.x; y1 = b.y; x2 = b.z; y2 = b.w;
}
template <>
__device__ __forceinline__ void load_box<double>(const double* base, int64_t row, double& x1, double& y1, double& x2, double& y2) {
  const double2 lo = reinterpret_cast<const double2*>(base)[2 * row];
  const double2 hi = reinterpret_cast<const double2*>(base)[2 * row + 1];
  x1 = lo.x; y1 = lo.y; x2 = hi.x; y2 = hi.y;
}

// volume of rows [r0, r1) of a track table, one warp
template <typename TB>
__device__ double warp_volume(const TB* boxes, int64_t r0, int64_t r1, int lane) {
  double acc = 0.0;
  for (int64_t r = r0 + lane; r < r1; r += 32) {
    double x1, y1, x2, y2;
    load_box<TB>(boxes, r, x1, y1, x2, y2);
    acc += (x2 - x1 + 1.0) * (y2 - y1 + 1.0);
  }
  return warp_sum(acc);
}

// overlap volume of two trajectories over absolute frames [lo, hi), one warp
template <typename TA, typename TG>
__device__ double warp_overlap(const TA* ba, int64_t rowa, const TG* bg, int64_t rowg, int64_t n, int lane) {
  double acc = 0.0;
  for (int64_t i = lane; i < n; i += 32) {
    double ax1, ay1, ax2, ay2, gx1, gy1, gx2, gy2;
    load_box<TA>(ba, rowa + i, ax1, ay1, ax2, ay2);
    load_box<TG>(bg, rowg + i, gx1, gy1, gx2, gy2);
    const double w = fmin(ax2, gx2) - fmax(ax1, gx1) + 1.0;
    const double h = fmin(ay2, gy2) - fmax(ay1, gy1) + 1.0;
    acc += fmax(w, 0.0) * fmax(h, 0.0);
  }
  return warp_sum(acc);
}

// ---- 1. per-relation volumes (sub, obj) over the relation's duration -------------------------
template <typename TB>
__global__ void rel_volume_kernel(VsgRelTable t, double* __restrict__ vol) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const TB* boxes = reinterpret_cast<const TB*>(t.boxes);
  for (int64_t job = warp; job < 2 * t.n_rel; job += n_warps) {
    const int64_t r = job >> 1;
    const int role = (int)(job & 1);
    const int64_t* row = t.rel + 7 * r;
    const int64_t trk = row[3 + role], s = row[5], e = row[6];
    const int64_t r0 = t.vol_full_track ? t.off[trk] : t.off[trk] + (s - t.tstart[trk]);
    const int64_t r1 = t.vol_full_track ? t.off[trk + 1] : r0 + (e - s);
    const double v = warp_volume<TB>(boxes, r0, r1, lane);
    if (lane == 0) vol[job] = v;
  }
}

// ---- 2. stable descending rank of each video's predictions ------------------------------------
__global__ void rank_kernel(const double* __restrict__ scores, const int64_t* __restrict__ vid_off, int32_t* __restrict__ order) {
  const int v = blockIdx.x;
  const int64_t b = vid_off[v];
  const int n = (int)(vid_off[v + 1] - b);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    // total order: NaN ranks as -inf (a NaN compares false with everything, which would hand several predictions the same rank and
    // leave order[] slots unwritten)
    double si = scores[b + i];
    if (si != si) si = -INFINITY;
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      double sj = scores[b + j];
      if (sj != sj) sj = -INFINITY;
      rank += (sj > si) || (sj == si && j < i);
    }
    order[b + rank] = i;
  }
}

// ---- 3. ov[p][g] for equal triplets ---------------------------------------------------------------
// One warp per PREDICTION: lanes scan the video's GT relations (triplet compare + temporal overlap test, one GT per lane),
// non-candidates are written directly (-1 / 0), candidates are collected with a ballot and then processed by the whole warp
// (coalesced box loads over the overlap).  Most (pred, GT) pairs differ in their triplet, so this avoids spending a warp on each.
template <typename TP, typename TG>
__global__ void rel_ov_kernel(VsgRelTable pr, VsgRelTable gt, int n_vid, const int64_t* __restrict__ ov_off,
                              const double* __restrict__ vol_p, const double* __restrict__ vol_g, double* __restrict__ ov,
                              double thr, uint8_t* __restrict__ cand_flag) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const TP* pb = reinterpret_cast<const TP*>(pr.boxes);
  const TG* gb = reinterpret_cast<const TG*>(gt.boxes);
  for (int64_t p = warp; p < pr.n_rel; p += n_warps) {
    const int v = find_segment(pr.vid_off, n_vid, p);
    const int64_t g0 = gt.vid_off[v];
    const int ng = (int)(gt.vid_off[v + 1] - g0);
    if (ng == 0) { if (lane == 0) cand_flag[p] = 0; continue; }
    bool any_hit = false;                     // lane 0: some GT of this video can be matched by this prediction (ov >= thr)
    const int64_t* rp = pr.rel + 7 * p;
    const int64_t t0 = rp[0], t1 = rp[1], t2 = rp[2], s1 = rp[5], e1 = rp[6];
    double* out = ov + ov_off[v] + (p - pr.vid_off[v]) * ng;
    for (int gbase = 0; gbase < ng; gbase += 32) {
      const int gi = gbase + lane;
      bool cand = false;
      if (gi < ng) {
        const int64_t* rg = gt.rel + 7 * (g0 + gi);
        if (rg[0] == t0 && rg[1] == t1 && rg[2] == t2) {
          if (s1 >= rg[6] || e1 <= rg[5]) out[gi] = 0.0; else cand = true;
        } else {
          out[gi] = -1.0;
        }
      }
      unsigned todo = __ballot_sync(0xffffffffu, cand);
      while (todo) {
        const int l = __ffs(todo) - 1;
        todo &= todo - 1;
        const int64_t g = g0 + gbase + l;
        const int64_t* rg = gt.rel + 7 * g;
        const int64_t s2 = rg[5], e2 = rg[6];
        const int64_t lo = s1 > s2 ? s1 : s2, hi = e1 < e2 ? e1 : e2;
        double role_iou[2];
#pragma unroll
        for (int role = 0; role < 2; ++role) {
          const int64_t tp = rp[3 + role], tg = rg[3 + role];
          const int64_t rowp = pr.off[tp] + (lo - pr.tstart[tp]);
          const int64_t rowg = gt.off[tg] + (lo - gt.tstart[tg]);
          const double o = warp_overlap<TP, TG>(pb, rowp, gb, rowg, hi - lo, lane);
          role_iou[role] = o / (vol_p[2 * p + role] + vol_g[2 * g + role] - o);
        }
        const double m = fmin(role_iou[0], role_iou[1]);
        if (lane == 0) { out[gbase + l] = m; any_hit |= (m >= thr); }
      }
    }
    if (lane == 0) cand_flag[p] = any_hit ? 1 : 0;
  }
}

// ---- 4. greedy assignment: one warp per video, predictions in rank order --------------------------
// Only predictions that CAN take a GT (cand_flag from rel_ov_kernel: some equal-triplet GT with ov >= thr) enter the sequential
// part; they are found 32 ranks at a time with a ballot.  Everything else is a miss, written in parallel.  (The sequential loop
// over all ~10^3 ranked predictions of a video was 0.35 ms of latency for ~10^1 real candidates.)
__global__ void greedy_match_kernel(const int64_t* __restrict__ p_vid_off, const int64_t* __restrict__ g_vid_off, int n_vid,
                                    const int64_t* __restrict__ ov_off, const double* __restrict__ ov,
                                    const int32_t* __restrict__ order, const double* __restrict__ scores, double thr,
                                    double* __restrict__ hit, int32_t* __restrict__ gt2det, uint8_t* __restrict__ taken_ws,
                                    const uint8_t* __restrict__ cand_flag) {
  const int lane = threadIdx.x & 31;
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (v >= n_vid) return;
  const int64_t p0 = p_vid_off[v], g0 = g_vid_off[v];
  const int np = (int)(p_vid_off[v + 1] - p0), ng = (int)(g_vid_off[v + 1] - g0);
  for (int g = lane; g < ng; g += 32) { gt2det[g0 + g] = -1; taken_ws[g0 + g] = 0; }
  for (int k = lane; k < np; k += 32) hit[p0 + k] = -INFINITY;
  if (ng == 0 || np == 0) return;
  __syncwarp();
  const double* ovv = ov + ov_off[v];
  for (int k0 = 0; k0 < np; k0 += 32) {
    const int kl = k0 + lane;
    const int pl = kl < np ? order[p0 + kl] : 0;
    unsigned todo = __ballot_sync(0xffffffffu, kl < np && cand_flag[p0 + pl] != 0);
    while (todo) {
      const int l = __ffs(todo) - 1;
      todo &= todo - 1;
      const int p = __shfl_sync(0xffffffffu, pl, l), k = k0 + l;
      // best = max ov over untaken GT with ov >= thr; ties -> smallest g (strict '>' in the reference)
      double best = -INFINITY;
      int best_g = 0x7fffffff;
      for (int g = lane; g < ng; g += 32) {
        const double o = ovv[(int64_t)p * ng + g];
        if (!taken_ws[g0 + g] && o >= 0.0 && o >= thr && o > best) { best = o; best_g = g; }   // o = -1 for other triplets
      }
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, s);
        const int og = __shfl_xor_sync(0xffffffffu, best_g, s);
        if (ob > best || (ob == best && og < best_g)) { best = ob; best_g = og; }
      }
      if (lane == 0 && best_g != 0x7fffffff) {
        hit[p0 + k] = scores[p0 + p];
        taken_ws[g0 + best_g] = 1;
        gt2det[g0 + best_g] = k;
      }
      __syncwarp();
    }
  }
}

template <typename TB>
__global__ void viou_pairs_kernel(const TB* __restrict__ b1, const int64_t* __restrict__ off1, const int64_t* __restrict__ dur1,
                                  const TB* __restrict__ b2, const int64_t* __restrict__ off2, const int64_t* __restrict__ dur2,
                                  int n, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n; i += n_warps) {
    const int64_t s1 = dur1[2 * i], e1 = dur1[2 * i + 1], s2 = dur2[2 * i], e2 = dur2[2 * i + 1];
    double res = 0.0;
    if (!(s1 >= e2 || e1 <= s2)) {
      const int64_t lo = s1 > s2 ? s1 : s2, hi = e1 < e2 ? e1 : e2;
      const double o = warp_overlap<TB, TB>(b1, off1[i] + (lo - s1), b2, off2[i] + (lo - s2), hi - lo, lane);
      const double v1 = warp_volume<TB>(b1, off1[i], off1[i + 1], lane);
      const double v2 = warp_volume<TB>(b2, off2[i], off2[i + 1], lane);
      res = o / (v1 + v2 - o);
    }
    if (lane == 0) out[i] = res;
  }
}

static inline int grid_warps(int64_t n_warp_jobs, int per_sm_blocks) {
  const int64_t blocks = (n_warp_jobs + 7) / 8;   // 256 threads = 8 warps
  const int64_t cap = (int64_t)sm_count() * per_sm_blocks;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace vsg

using namespace vsg;

extern "C" int vsg_rel_viou_match(const VsgRelTable* pred, const double* scores, const VsgRelTable* gt, int n_vid,
                                  const int64_t* ov_off, double thr, int32_t* order, double* ov_ws, double* hit,
                                  int32_t* gt2det, double* vol_pred_ws, double* vol_gt_ws, uint8_t* taken_ws, void* stream) {
  VSG_REQUIRE(pred && gt && n_vid >= 0, "vsg_rel_viou_match: null table or negative n_vid");
  if (n_vid == 0) return VSG_OK;
  VSG_REQUIRE(pred->n_rel >= 0 && gt->n_rel >= 0, "vsg_rel_viou_match: negative relation count");
  VSG_REQUIRE(pred->vid_off && gt->vid_off && ov_off, "vsg_rel_viou_match: null offsets");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t np = pred->n_rel, ng = gt->n_rel;
  VSG_REQUIRE(np + ng == 0 || taken_ws, "vsg_rel_viou_match: null taken / candidate workspace (n_gt + n_pred bytes)");
  if (np > 0) {
    VSG_REQUIRE(pred->boxes && pred->off && pred->tstart && pred->rel && scores && order && hit && vol_pred_ws,
                "vsg_rel_viou_match: null prediction pointer");
    VSG_REQUIRE(aligned16(pred->boxes), "vsg_rel_viou_match: prediction boxes misaligned");
    if (pred->box_f64) rel_volume_kernel<double><<<grid_warps(2 * np, 8), 256, 0, st>>>(*pred, vol_pred_ws);
    else rel_volume_kernel<float><<<grid_warps(2 * np, 8), 256, 0, st>>>(*pred, vol_pred_ws);
    rank_kernel<<<n_vid, 256, 0, st>>>(scores, pred->vid_off, order);
  }
  if (ng > 0) {
    VSG_REQUIRE(gt->boxes && gt->off && gt->tstart && gt->rel && gt2det && vol_gt_ws, "vsg_rel_viou_match: null GT pointer");
    VSG_REQUIRE(aligned16(gt->boxes), "vsg_rel_viou_match: GT boxes misaligned");
    if (gt->box_f64) rel_volume_kernel<double><<<grid_warps(2 * ng, 8), 256, 0, st>>>(*gt, vol_gt_ws);
    else rel_volume_kernel<float><<<grid_warps(2 * ng, 8), 256, 0, st>>>(*gt, vol_gt_ws);
  }
  if (np > 0 && ng > 0) {
    VSG_REQUIRE(ov_ws, "vsg_rel_viou_match: null ov workspace");
    const int g = grid_warps(np, 8);   // one warp per prediction
    if (pred->box_f64 && gt->box_f64) rel_ov_kernel<double, double><<<g, 256, 0, st>>>(*pred, *gt, n_vid, ov_off, vol_pred_ws, vol_gt_ws, ov_ws, thr, taken_ws + ng);
    else if (pred->box_f64) rel_ov_kernel<double, float><<<g, 256, 0, st>>>(*pred, *gt, n_vid, ov_off, vol_pred_ws, vol_gt_ws, ov_ws, thr, taken_ws + ng);
    else if (gt->box_f64) rel_ov_kernel<float, double><<<g, 256, 0, st>>>(*pred, *gt, n_vid, ov_off, vol_pred_ws, vol_gt_ws, ov_ws, thr, taken_ws + ng);
    else rel_ov_kernel<float, float><<<g, 256, 0, st>>>(*pred, *gt, n_vid, ov_off, vol_pred_ws, vol_gt_ws, ov_ws, thr, taken_ws + ng);
  }
  // greedy pass; also initialises gt2det / hit for videos with no predictions or no GT
  greedy_match_kernel<<<(n_vid * 32 + 127) / 128, 128, 0, st>>>(pred->vid_off, gt->vid_off, n_vid, ov_off, ov_ws, order, scores,
                                                               thr, hit, gt2det, taken_ws, taken_ws + ng);
  return check_launch("vsg_rel_viou_match", 1 + (np > 0 ? 2 : 0) + (ng > 0 ? 1 : 0) + ((np > 0 && ng > 0) ? 1 : 0));
}

extern "C" int vsg_viou_pairs_f64(const double* boxes1, const int64_t* off1, const int64_t* dur1, const double* boxes2,
                                  const int64_t* off2, const int64_t* dur2, int n, double* out, void* stream) {
  VSG_REQUIRE(n >= 0, "vsg_viou_pairs_f64: n < 0");
  if (n == 0) return VSG_OK;
  VSG_REQUIRE(boxes1 && off1 && dur1 && boxes2 && off2 && dur2 && out, "vsg_viou_pairs_f64: null pointer");
  VSG_REQUIRE(aligned16(boxes1) && aligned16(boxes2), "vsg_viou_pairs_f64: boxes misaligned");
  viou_pairs_kernel<double><<<grid_warps(n, 8), 256, 0, (cudaStream_t)stream>>>(boxes1, off1, dur1, boxes2, off2, dur2, n, out);
  return check_launch("vsg_viou_pairs_f64");
}
