// Geometry kernels: pair enumeration, span intersection, trajectory volume IoU.
// SURVEY.md section 8a rows A2 (utils/utils_func.py:347-373), A3 (model_pairwise_baseline.py:104-111),
// A4 (utils/utils_func.py:437-471 + the per-pair loop models/model_0v10.py:565-581),
// A9 (tools/train_vidor.py:143-159).
//
// Data layout in HBM: tracks are packed CSR -- boxes[sum L][4] f32 (one float4 per frame, 16-byte
// aligned), off[n+1] int64 -- so one warp reads 32 consecutive frames of a track as one 512-byte
// coalesced request.  Every track of a video lives on the same absolute frame axis, so the overlap of
// tracks a and b is frames [max(sa,sb), min(ea,eb)] at rows off[a] + f - sa and off[b] + f - sb.
#include "common.cuh"

namespace vsg {

// ------------------------------------------------------------------------------------------
__global__ void pair_ids_kernel(int n, int64_t* __restrict__ out) {
  const int64_t total = (int64_t)n * (n - 1);
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = k / (n - 1);
    const int64_t r = k - s * (n - 1);
    const int64_t o = r + (r >= s ? 1 : 0);
    reinterpret_cast<longlong2*>(out)[k] = make_longlong2(s, o);
  }
}

__global__ void dura_intersection_kernel(const int64_t* __restrict__ d1, int n1, const int64_t* __restrict__ d2, int n2,
                                         int64_t* __restrict__ inter, uint8_t* __restrict__ mask) {
  const int64_t total = (int64_t)n1 * n2;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(k / n2), j = (int)(k - (int64_t)i * n2);
    const longlong2 a = reinterpret_cast<const longlong2*>(d1)[i];
    const longlong2 b = reinterpret_cast<const longlong2*>(d2)[j];
    const int64_t s = a.x > b.x ? a.x : b.x;
    const int64_t e = a.y < b.y ? a.y : b.y;
    if (inter) reinterpret_cast<longlong2*>(inter)[k] = make_longlong2(s, e);
    if (mask) mask[k] = (s <= e) ? 1 : 0;
  }
}

template <typename T>
__global__ void dura_intersection_generic_kernel(const T* __restrict__ d1, int n1, const T* __restrict__ d2, int n2, int broadcast,
                                                 T* __restrict__ inter, uint8_t* __restrict__ mask) {
  const int64_t total = broadcast ? (int64_t)n1 * n2 : n1;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int i = broadcast ? (int)(k / n2) : (int)k, j = broadcast ? (int)(k - (int64_t)i * n2) : (int)k;
    const T s = d1[2 * i] > d2[2 * j] ? d1[2 * i] : d2[2 * j];
    const T e = d1[2 * i + 1] < d2[2 * j + 1] ? d1[2 * i + 1] : d2[2 * j + 1];
    if (inter) { inter[2 * k] = s; inter[2 * k + 1] = e; }
    if (mask) mask[k] = (s <= e) ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------
// Track volumes: one warp per track, float4 loads, chunked fp32 partials folded into fp64.
__global__ void track_volume_kernel(const float4* __restrict__ boxes, const int64_t* __restrict__ off, int n_tracks,
                                    float* __restrict__ vol) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int t = warp; t < n_tracks; t += n_warps) {
    const int64_t b = off[t], e = off[t + 1];
    double acc = 0.0;
    for (int64_t base = b; base < e; base += 32 * 8) {
      float part = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t r = base + u * 32 + lane;
        if (r < e) {
          const float4 x = ldg_stream(boxes + r);
          part += (x.z - x.x + 1.0f) * (x.w - x.y + 1.0f);
        }
      }
      acc += (double)part;
    }
    acc = warp_sum(acc);
    if (lane == 0) vol[t] = (float)acc;
  }
}

// ------------------------------------------------------------------------------------------
// Variant 1: one warp per (a, b) pair; lanes stride the overlap.  Reads 32 bytes per overlapping
// frame-pair (the "algorithmic bytes" of SURVEY 8d); tracks of a video are re-read from L2.
__device__ __forceinline__ float inter_area(const float4 a, const float4 b) {
  const float w = fmaxf((fminf(a.z, b.z) - fmaxf(a.x, b.x)) + 1.0f, 0.0f);
  const float h = fmaxf((fminf(a.w, b.w) - fmaxf(a.y, b.y)) + 1.0f, 0.0f);
  return w * h;
}

__global__ void __launch_bounds__(256)
traj_viou_warp_kernel(const float4* __restrict__ boxesA, const int64_t* __restrict__ offA, const int64_t* __restrict__ duraA,
                      const float4* __restrict__ boxesB, const int64_t* __restrict__ offB, const int64_t* __restrict__ duraB,
                      const int32_t* __restrict__ segA, const int32_t* __restrict__ segB,
                      const int64_t* __restrict__ seg_out, int n_seg, int64_t n_pairs,
                      const float* __restrict__ volA, const float* __restrict__ volB,
                      int64_t* __restrict__ spans, uint8_t* __restrict__ mask, float* __restrict__ viou,
                      float* __restrict__ inter_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < n_pairs; p += n_warps) {
    const int v = find_segment(seg_out, n_seg, p);
    const int nB = segB[v + 1] - segB[v];
    const int64_t local = p - seg_out[v];
    const int ia = segA[v] + (int)(local / nB);
    const int ib = segB[v] + (int)(local % nB);
    const longlong2 da = reinterpret_cast<const longlong2*>(duraA)[ia];
    const longlong2 db = reinterpret_cast<const longlong2*>(duraB)[ib];
    const int64_t s = da.x > db.x ? da.x : db.x;
    const int64_t e = da.y < db.y ? da.y : db.y;
    if (lane == 0) {
      if (spans) reinterpret_cast<longlong2*>(spans)[p] = make_longlong2(s, e);
      if (mask) mask[p] = (s <= e) ? 1 : 0;
    }
    if (!viou && !inter_out) continue;
    if (s > e) {
      if (lane == 0) {
        if (viou) viou[p] = 0.0f;
        if (inter_out) inter_out[p] = 0.0f;
      }
      continue;
    }
    const float4* pa = boxesA + offA[ia] + (s - da.x);
    const float4* pb = boxesB + offB[ib] + (s - db.x);
    const int64_t len = e - s + 1;
    double acc = 0.0;
    for (int64_t base = 0; base < len; base += 32 * 4) {
      float4 a[4], b[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t f = base + u * 32 + lane;
        ok[u] = f < len;
        if (ok[u]) {
          a[u] = ldg_stream(pa + f);
          b[u] = ldg_stream(pb + f);
        }
      }
      float part = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (ok[u]) part += inter_area(a[u], b[u]);
      acc += (double)part;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float it = (float)acc;
      if (viou) viou[p] = it / (volA[ia] + volB[ib] - it);
      if (inter_out) inter_out[p] = it;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Variant 2: shared-memory tiled kernel for long, densely overlapping tracks (VidOR / stress shapes).
// Job = (segment, 32x32 block of its pair matrix, chunk of T2_CH absolute frames).  A CTA stages the TA + TB track
// slices of the chunk window by window (W frames; frames where a track is absent are filled with an empty box
// that intersects nothing), and every thread accumulates a 4x4 register block of pairs: per frame 8 128-bit shared
// loads feed 16 pairs, so L2/HBM see each box once per tile row/column instead of once per pair.  The fp32 chunk
// partials go to a workspace and a second kernel sums them in a fixed order (deterministic) in fp64.
constexpr int T2_TA = 32, T2_TB = 32, T2_RA = 4, T2_RB = 4, T2_W = 32, T2_CH = 256;
constexpr int T2_THREADS = (T2_TA / T2_RA) * (T2_TB / T2_RB);  // 64
constexpr int T2_PITCH = T2_W + 1;                             // float4 pitch: conflict-free 128-bit reads

// per-segment job table: tiles_b, chunks, first job; built by one small kernel
struct SegJobs { int tiles_a, tiles_b, chunks; long long first_job; };

__global__ void seg_jobs_kernel(const int32_t* __restrict__ segA, const int32_t* __restrict__ segB, const int64_t* __restrict__ seg_len,
                                int n_seg, SegJobs* __restrict__ jobs, int64_t* __restrict__ n_jobs) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    long long acc = 0;
    for (int v = 0; v < n_seg; ++v) {
      const int nA = segA[v + 1] - segA[v], nB = segB[v + 1] - segB[v];
      SegJobs j;
      j.tiles_a = (nA + T2_TA - 1) / T2_TA; j.tiles_b = (nB + T2_TB - 1) / T2_TB;
      j.chunks = (int)((seg_len[v] + T2_CH - 1) / T2_CH);
      if (j.chunks < 1) j.chunks = 1;
      j.first_job = acc;
      jobs[v] = j;
      acc += (long long)j.tiles_a * j.tiles_b * j.chunks;
    }
    jobs[n_seg].first_job = acc;
    *n_jobs = acc;
  }
}

__device__ __forceinline__ int find_seg_job(const SegJobs* __restrict__ jobs, int n_seg, long long job) {
  int lo = 0, hi = n_seg;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (jobs[mid].first_job <= job) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(T2_THREADS)
traj_viou_tile_kernel(const float4* __restrict__ boxesA, const int64_t* __restrict__ offA, const int64_t* __restrict__ duraA,
                      const float4* __restrict__ boxesB, const int64_t* __restrict__ offB, const int64_t* __restrict__ duraB,
                      const int32_t* __restrict__ segA, const int32_t* __restrict__ segB, const SegJobs* __restrict__ jobs, int n_seg,
                      const int64_t* __restrict__ n_jobs_ptr, float* __restrict__ part /* [jobs][TA*TB] */) {
  __shared__ float4 sA[T2_TA * T2_PITCH];
  __shared__ float4 sB[T2_TB * T2_PITCH];
  __shared__ long long sStart[T2_TA + T2_TB], sEnd[T2_TA + T2_TB], sRow[T2_TA + T2_TB];
  __shared__ long long sRange[2];
  const int tid = threadIdx.x;
  const int ta = tid / (T2_TB / T2_RB), tb = tid % (T2_TB / T2_RB);
  const float4 EMPTY = make_float4(1e30f, 1e30f, -1e30f, -1e30f);
  const long long n_jobs = *n_jobs_ptr;

  for (long long job = blockIdx.x; job < n_jobs; job += gridDim.x) {
    const int v = find_seg_job(jobs, n_seg, job);
    const SegJobs sj = jobs[v];
    const long long lj = job - sj.first_job;
    const int chunk = (int)(lj % sj.chunks);
    const int tile = (int)(lj / sj.chunks);
    const int a0 = (tile / sj.tiles_b) * T2_TA, b0 = (tile % sj.tiles_b) * T2_TB;
    const int nA = segA[v + 1] - segA[v], nB = segB[v + 1] - segB[v];
    __syncthreads();  // previous job finished with smem
    if (tid < T2_TA + T2_TB) {
      const bool isA = tid < T2_TA;
      const int loc = isA ? a0 + tid : b0 + (tid - T2_TA);
      const bool valid = loc < (isA ? nA : nB);
      long long s = 1, e = 0, row = 0;
      if (valid) {
        const int g = (isA ? segA[v] : segB[v]) + loc;
        const longlong2 d = reinterpret_cast<const longlong2*>(isA ? duraA : duraB)[g];
        s = d.x; e = d.y; row = (isA ? offA : offB)[g];
      }
      sStart[tid] = s; sEnd[tid] = e; sRow[tid] = row;
    }
    __syncthreads();
    if (tid == 0) {
      // frames of this chunk where at least one A track and one B track of the tile exist
      long long as = LLONG_MAX, ae = LLONG_MIN, bs = LLONG_MAX, be = LLONG_MIN;
      for (int i = 0; i < T2_TA; ++i) if (sStart[i] <= sEnd[i]) { as = min(as, sStart[i]); ae = max(ae, sEnd[i]); }
      for (int i = T2_TA; i < T2_TA + T2_TB; ++i) if (sStart[i] <= sEnd[i]) { bs = min(bs, sStart[i]); be = max(be, sEnd[i]); }
      // chunks tile the absolute frame axis from the segment's first frame; seg_len bounds it (host-provided)
      sRange[0] = max(max(as, bs), (long long)chunk * T2_CH);
      sRange[1] = min(min(ae, be), (long long)(chunk + 1) * T2_CH - 1);
    }
    __syncthreads();
    const long long f_lo = sRange[0], f_hi = sRange[1];
    float acc[T2_RA][T2_RB];
#pragma unroll
    for (int i = 0; i < T2_RA; ++i)
#pragma unroll
      for (int j = 0; j < T2_RB; ++j) acc[i][j] = 0.f;

    for (long long f0 = f_lo; f0 <= f_hi; f0 += T2_W) {
      __syncthreads();
      // stage: (TA+TB) tracks x W frames; consecutive threads take consecutive frames of one track
      for (int idx = tid; idx < (T2_TA + T2_TB) * T2_W; idx += T2_THREADS) {
        const int trk = idx / T2_W, fo = idx % T2_W;
        const long long f = f0 + fo;
        float4 val = EMPTY;
        if (f <= f_hi && f >= sStart[trk] && f <= sEnd[trk]) {
          const float4* src = (trk < T2_TA ? boxesA : boxesB) + sRow[trk] + (f - sStart[trk]);
          val = ldg_stream(src);
        }
        if (trk < T2_TA) sA[trk * T2_PITCH + fo] = val; else sB[(trk - T2_TA) * T2_PITCH + fo] = val;
      }
      __syncthreads();
#pragma unroll 4
      for (int fo = 0; fo < T2_W; ++fo) {
        float4 a[T2_RA], b[T2_RB];
#pragma unroll
        for (int i = 0; i < T2_RA; ++i) a[i] = sA[(ta * T2_RA + i) * T2_PITCH + fo];
#pragma unroll
        for (int j = 0; j < T2_RB; ++j) b[j] = sB[(tb * T2_RB + j) * T2_PITCH + fo];
#pragma unroll
        for (int i = 0; i < T2_RA; ++i)
#pragma unroll
          for (int j = 0; j < T2_RB; ++j) acc[i][j] += inter_area(a[i], b[j]);
      }
    }
    float* o = part + job * (T2_TA * T2_TB);
#pragma unroll
    for (int i = 0; i < T2_RA; ++i)
      *reinterpret_cast<float4*>(o + (ta * T2_RA + i) * T2_TB + tb * T2_RB) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

// sums the chunk partials of every pair in chunk order (fp64) and writes spans / mask / vIoU
__global__ void traj_viou_tile_finalize_kernel(const int64_t* __restrict__ duraA, const int64_t* __restrict__ duraB,
                                               const int32_t* __restrict__ segA, const int32_t* __restrict__ segB,
                                               const int64_t* __restrict__ seg_out, const SegJobs* __restrict__ jobs, int n_seg,
                                               int64_t n_pairs, const float* __restrict__ part, const float* __restrict__ volA,
                                               const float* __restrict__ volB, int64_t* __restrict__ spans, uint8_t* __restrict__ mask,
                                               float* __restrict__ viou) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_pairs; p += (int64_t)gridDim.x * blockDim.x) {
    const int v = find_segment(seg_out, n_seg, p);
    const int nB = segB[v + 1] - segB[v];
    const int64_t local = p - seg_out[v];
    const int la = (int)(local / nB), lb = (int)(local % nB);
    const int ia = segA[v] + la, ib = segB[v] + lb;
    const longlong2 da = reinterpret_cast<const longlong2*>(duraA)[ia];
    const longlong2 db = reinterpret_cast<const longlong2*>(duraB)[ib];
    const int64_t s = da.x > db.x ? da.x : db.x, e = da.y < db.y ? da.y : db.y;
    if (spans) reinterpret_cast<longlong2*>(spans)[p] = make_longlong2(s, e);
    if (mask) mask[p] = (s <= e) ? 1 : 0;
    if (!viou) continue;
    float r = 0.f;
    if (s <= e) {
      const SegJobs sj = jobs[v];
      const int tile = (la / T2_TA) * sj.tiles_b + (lb / T2_TB);
      const float* src = part + (sj.first_job + (long long)tile * sj.chunks) * (T2_TA * T2_TB) + (la % T2_TA) * T2_TB + (lb % T2_TB);
      double acc = 0.0;
      const int c0 = (int)(s / T2_CH), c1 = (int)(e / T2_CH);      // only chunks that intersect the overlap hold non-zeros
      for (int c = c0; c <= c1 && c < sj.chunks; ++c) acc += (double)src[(long long)c * (T2_TA * T2_TB)];
      const float it = (float)acc;
      r = it / (volA[ia] + volB[ib] - it);
    }
    viou[p] = r;
  }
}

__global__ void pair_labels_kernel(const float* __restrict__ viou, int n, int n_gt_traj, const int64_t* __restrict__ gt_so,
                                   int n_gt_pred, float th, uint8_t* __restrict__ out) {
  const int64_t n_pairs = (int64_t)n * (n - 1);
  const int64_t total = n_pairs * n_gt_pred;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(k / n_pairs);
    const int64_t pr = k - (int64_t)g * n_pairs;
    const int s = (int)(pr / (n - 1));
    const int r = (int)(pr - (int64_t)s * (n - 1));
    const int o = r + (r >= s ? 1 : 0);
    const int gs = (int)gt_so[2 * g], go = (int)gt_so[2 * g + 1];
    out[k] = (viou[(int64_t)s * n_gt_traj + gs] > th && viou[(int64_t)o * n_gt_traj + go] > th) ? 1 : 0;
  }
}

static inline int grid_for(int64_t work_items, int threads, int per_sm) {
  const int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace vsg

using namespace vsg;

extern "C" int vsg_pair_ids(int n, int64_t* out, void* stream) {
  VSG_REQUIRE(n >= 0, "vsg_pair_ids: n < 0");
  if (n < 2) return VSG_OK;
  VSG_REQUIRE(out != nullptr && aligned16(out), "vsg_pair_ids: out must be a 16-byte aligned device pointer");
  pair_ids_kernel<<<grid_for((int64_t)n * (n - 1), 256, 8), 256, 0, (cudaStream_t)stream>>>(n, out);
  return check_launch("vsg_pair_ids");
}

extern "C" int vsg_dura_intersection(const int64_t* d1, int n1, const int64_t* d2, int n2, int64_t* inter, uint8_t* mask,
                                     void* stream) {
  VSG_REQUIRE(n1 >= 0 && n2 >= 0, "vsg_dura_intersection: negative size");
  if (n1 == 0 || n2 == 0) return VSG_OK;
  VSG_REQUIRE(d1 && d2 && aligned16(d1) && aligned16(d2), "vsg_dura_intersection: spans must be 16-byte aligned");
  VSG_REQUIRE(inter == nullptr || aligned16(inter), "vsg_dura_intersection: inter_out misaligned");
  dura_intersection_kernel<<<grid_for((int64_t)n1 * n2, 256, 8), 256, 0, (cudaStream_t)stream>>>(d1, n1, d2, n2, inter, mask);
  return check_launch("vsg_dura_intersection");
}

extern "C" int vsg_dura_intersection_ex(const void* d1, int n1, const void* d2, int n2, int broadcast, int dtype, void* inter,
                                        uint8_t* mask, void* stream) {
  VSG_REQUIRE(n1 >= 0 && n2 >= 0, "vsg_dura_intersection_ex: negative size");
  VSG_REQUIRE(broadcast || n1 == n2, "vsg_dura_intersection_ex: row-wise mode needs n1 == n2");
  VSG_REQUIRE(dtype == 0 || dtype == 1, "vsg_dura_intersection_ex: dtype must be 0 (int64) or 1 (float32)");
  if (n1 == 0 || n2 == 0) return VSG_OK;
  VSG_REQUIRE(d1 && d2, "vsg_dura_intersection_ex: null spans");
  const int g = grid_for(broadcast ? (int64_t)n1 * n2 : n1, 256, 8);
  if (dtype == 0)
    dura_intersection_generic_kernel<int64_t><<<g, 256, 0, (cudaStream_t)stream>>>((const int64_t*)d1, n1, (const int64_t*)d2, n2,
                                                                                    broadcast, (int64_t*)inter, mask);
  else
    dura_intersection_generic_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)d1, n1, (const float*)d2, n2, broadcast,
                                                                                  (float*)inter, mask);
  return check_launch("vsg_dura_intersection_ex");
}

extern "C" int vsg_track_volumes(const float* boxes, const int64_t* off, int n_tracks, float* vol, void* stream) {
  VSG_REQUIRE(n_tracks >= 0, "vsg_track_volumes: n_tracks < 0");
  if (n_tracks == 0) return VSG_OK;
  VSG_REQUIRE(boxes && off && vol && aligned16(boxes), "vsg_track_volumes: null or misaligned pointer");
  track_volume_kernel<<<grid_for((int64_t)n_tracks * 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(boxes), off, n_tracks, vol);
  return check_launch("vsg_track_volumes");
}

extern "C" int vsg_traj_viou_matrix(const float* boxesA, const int64_t* offA, const int64_t* duraA, int nTA,
                                    const float* boxesB, const int64_t* offB, const int64_t* duraB, int nTB,
                                    const int32_t* segA, const int32_t* segB, const int64_t* seg_out, int n_seg,
                                    int64_t n_pairs, int64_t* spans, uint8_t* mask, float* viou, float* inter_out,
                                    float* volA, float* volB, int variant, void* stream) {
  VSG_REQUIRE(nTA >= 0 && nTB >= 0 && n_seg >= 0 && n_pairs >= 0, "vsg_traj_viou_matrix: negative size");
  if (n_pairs == 0 || n_seg == 0) return VSG_OK;
  VSG_REQUIRE(boxesA && offA && duraA && boxesB && offB && duraB && segA && segB && seg_out,
              "vsg_traj_viou_matrix: null input pointer");
  VSG_REQUIRE(aligned16(boxesA) && aligned16(boxesB) && aligned16(duraA) && aligned16(duraB),
              "vsg_traj_viou_matrix: boxes / spans must be 16-byte aligned");
  VSG_REQUIRE(spans == nullptr || aligned16(spans), "vsg_traj_viou_matrix: spans_out misaligned");
  VSG_REQUIRE(viou == nullptr || (volA && volB), "vsg_traj_viou_matrix: volume workspaces required with viou_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (viou) {
    int rc = vsg_track_volumes(boxesA, offA, nTA, volA, stream);
    if (rc) return rc;
    if (!(volB == volA && boxesB == boxesA)) {
      rc = vsg_track_volumes(boxesB, offB, nTB, volB, stream);
      if (rc) return rc;
    }
  }
  if (variant == 0) variant = 1;
  if (variant == 1) {
    traj_viou_warp_kernel<<<grid_for(n_pairs * 32, 256, 8), 256, 0, st>>>(
        reinterpret_cast<const float4*>(boxesA), offA, duraA, reinterpret_cast<const float4*>(boxesB), offB, duraB, segA, segB,
        seg_out, n_seg, n_pairs, volA, volB, spans, mask, viou, inter_out);
    return check_launch("vsg_traj_viou_matrix(warp)");
  }
  set_error("vsg_traj_viou_matrix: variant 2 needs vsg_traj_viou_matrix_tiled (tile-offset workspace)");
  return VSG_E_UNSUPPORTED;
}

// Tiled variant.  seg_len int64[n_seg]: frames spanned by each segment's absolute frame axis (video_len; spans must lie in
// [0, seg_len)).  Workspaces: jobs_ws >= (n_seg+1)*24 bytes, njobs_ws int64[1], part_ws f32[n_jobs_bound*1024] where
// n_jobs_bound >= sum_v ceil(nA_v/32)*ceil(nB_v/32)*ceil(seg_len_v/256) (computed by the host wrapper).
extern "C" int vsg_traj_viou_matrix_tiled(const float* boxesA, const int64_t* offA, const int64_t* duraA, int nTA,
                                          const float* boxesB, const int64_t* offB, const int64_t* duraB, int nTB,
                                          const int32_t* segA, const int32_t* segB, const int64_t* seg_out, const int64_t* seg_len,
                                          int n_seg, int64_t n_pairs, int64_t n_jobs_bound, int64_t* spans, uint8_t* mask,
                                          float* viou, float* volA, float* volB, void* jobs_ws, int64_t* njobs_ws, float* part_ws,
                                          void* stream) {
  VSG_REQUIRE(nTA >= 0 && nTB >= 0 && n_seg >= 0 && n_pairs >= 0 && n_jobs_bound >= 0, "vsg_traj_viou_matrix_tiled: negative size");
  if (n_pairs == 0 || n_seg == 0) return VSG_OK;
  VSG_REQUIRE(boxesA && offA && duraA && boxesB && offB && duraB && segA && segB && seg_out && seg_len && jobs_ws && njobs_ws && part_ws,
              "vsg_traj_viou_matrix_tiled: null input pointer");
  VSG_REQUIRE(aligned16(boxesA) && aligned16(boxesB) && aligned16(duraA) && aligned16(duraB) && aligned16(part_ws),
              "vsg_traj_viou_matrix_tiled: boxes / spans / workspace must be 16-byte aligned");
  VSG_REQUIRE(spans == nullptr || aligned16(spans), "vsg_traj_viou_matrix_tiled: spans_out misaligned");
  VSG_REQUIRE(viou == nullptr || (volA && volB), "vsg_traj_viou_matrix_tiled: volume workspaces required");
  cudaStream_t st = (cudaStream_t)stream;
  int launches = 2;
  if (viou) {
    int rc = vsg_track_volumes(boxesA, offA, nTA, volA, stream);
    if (rc) return rc;
    if (!(volB == volA && boxesB == boxesA)) {
      rc = vsg_track_volumes(boxesB, offB, nTB, volB, stream);
      if (rc) return rc;
    }
  }
  SegJobs* jobs = reinterpret_cast<SegJobs*>(jobs_ws);
  seg_jobs_kernel<<<1, 32, 0, st>>>(segA, segB, seg_len, n_seg, jobs, njobs_ws);
  if (viou) {
    const int64_t cap = (int64_t)sm_count() * 6;
    const int grid = (int)(n_jobs_bound < cap ? (n_jobs_bound < 1 ? 1 : n_jobs_bound) : cap);
    traj_viou_tile_kernel<<<grid, T2_THREADS, 0, st>>>(reinterpret_cast<const float4*>(boxesA), offA, duraA,
                                                       reinterpret_cast<const float4*>(boxesB), offB, duraB, segA, segB, jobs, n_seg,
                                                       njobs_ws, part_ws);
    ++launches;
  }
  traj_viou_tile_finalize_kernel<<<grid_for(n_pairs, 256, 8), 256, 0, st>>>(duraA, duraB, segA, segB, seg_out, jobs, n_seg, n_pairs, part_ws,
                                                                           volA, volB, spans, mask, viou);
  return check_launch("vsg_traj_viou_matrix_tiled", launches);
}

extern "C" int vsg_pair_labels(const float* viou, int n, int n_gt_traj, const int64_t* gt_so, int n_gt_pred, float th,
                               uint8_t* out, void* stream) {
  VSG_REQUIRE(n >= 0 && n_gt_traj >= 0 && n_gt_pred >= 0, "vsg_pair_labels: negative size");
  if (n < 2 || n_gt_pred == 0) return VSG_OK;
  VSG_REQUIRE(viou && gt_so && out, "vsg_pair_labels: null pointer");
  pair_labels_kernel<<<grid_for((int64_t)n * (n - 1) * n_gt_pred, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      viou, n, n_gt_traj, gt_so, n_gt_pred, th, out);
  return check_launch("vsg_pair_labels");
}

// ---------------------------------------------------------------------------------------------------
// Remaining public helpers of utils/utils_func.py on the device: tIoU / generalized_tIoU (:375-410), the stretch of
// stack_with_repeat_2d (models/model_0v10.py:18-46) and unique_with_idx_nd (utils_func.py:330-345)
// ---------------------------------------------------------------------------------------------------
namespace vsg {

// (min(e1,e2) - max(s1,s2)) / (max(e1,e2) - min(s1,s2)) in fp32 (torch true-divides int64 spans as float32);
// tIoU additionally zeroes pairs whose closed spans do not touch.  0/0 stays NaN like the reference.
template <typename T>
__global__ void tiou_kernel(const T* __restrict__ d1, int n1, const T* __restrict__ d2, int n2, int broadcast, int generalized,
                            float* __restrict__ out) {
  const int64_t total = broadcast ? (int64_t)n1 * n2 : n1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = broadcast ? (int)(i / n2) : (int)i, b = broadcast ? (int)(i % n2) : (int)i;
    const T s1 = d1[2 * a], e1 = d1[2 * a + 1], s2 = d2[2 * b], e2 = d2[2 * b + 1];
    const T num = (e1 < e2 ? e1 : e2) - (s1 > s2 ? s1 : s2);
    const T den = (e1 > e2 ? e1 : e2) - (s1 < s2 ? s1 : s2);
    float v = (float)num / (float)den;
    if (!generalized && !((e1 >= s2) && (e2 >= s1))) v = 0.f;
    out[i] = v;
  }
}

// out[t][j][:] = src[off[t] + map(j)][:], j < tmax: frame i of an L-frame track repeated ceil((tmax - i) / L) times
__global__ void stretch_rows_kernel(const float* __restrict__ src, int ld, int width, const int64_t* __restrict__ off, int n_tracks,
                                    int tmax, float* __restrict__ out) {
  const int64_t rows = (int64_t)n_tracks * tmax;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const int t = (int)(r / tmax), j = (int)(r % tmax);
    const int L = (int)(off[t + 1] - off[t]);
    if (L <= 0) continue;
    const int q = tmax / L, rem = tmax % L;
    const int s = j < rem * (q + 1) ? j / (q + 1) : rem + (j - rem * (q + 1)) / max(q, 1);
    const float* p = src + (off[t] + s) * (int64_t)ld;
    float* o = out + r * (int64_t)width;
    for (int c = threadIdx.x; c < width; c += blockDim.x) o[c] = p[c];
  }
}

// One CTA: bitonic sort of the row indices by (row lexicographic, index), run heads, group id per sorted position.
__device__ __forceinline__ int row_cmp(const int64_t* __restrict__ rows, int d, int a, int b) {
  for (int c = 0; c < d; ++c) {
    const int64_t x = rows[(int64_t)a * d + c], y = rows[(int64_t)b * d + c];
    if (x != y) return x < y ? -1 : 1;
  }
  return 0;
}
__global__ void unique_rows_kernel(const int64_t* __restrict__ rows, int n, int d, int npow2, int32_t* __restrict__ order,
                                   int32_t* __restrict__ group, int32_t* __restrict__ n_groups) {
  extern __shared__ int32_t idx[];
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) idx[i] = i < n ? i : -1;      // -1 = +infinity padding
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const int a = idx[i], b = idx[p];
          bool a_gt_b;                                                                   // strict total order: padding last, ties by index
          if (a < 0 || b < 0) a_gt_b = (a < 0) && (b >= 0);
          else { const int c = row_cmp(rows, d, a, b); a_gt_b = c > 0 || (c == 0 && a > b); }
          const bool up = (i & k) == 0;
          if (a_gt_b == up) { idx[i] = b; idx[p] = a; }
        }
      }
      __syncthreads();
    }
  }
  // heads + inclusive scan of the head flags (single thread per chunk would do; n <= 4096, so a simple serial pass by thread 0)
  if (threadIdx.x == 0) {
    int g = -1;
    for (int i = 0; i < n; ++i) {
      if (i == 0 || row_cmp(rows, d, idx[i - 1], idx[i]) != 0) ++g;
      order[i] = idx[i];
      group[i] = g;
    }
    *n_groups = g + 1;
  }
}

}  // namespace vsg

extern "C" int vsg_tiou(const void* d1, int n1, const void* d2, int n2, int broadcast, int generalized, int dtype, float* out,
                        void* stream) {
  VSG_REQUIRE(n1 >= 0 && n2 >= 0, "vsg_tiou: negative size");
  VSG_REQUIRE(broadcast || n1 == n2, "vsg_tiou: row-wise mode needs n1 == n2");
  VSG_REQUIRE(dtype == 0 || dtype == 1, "vsg_tiou: dtype must be 0 (int64) or 1 (float32)");
  if (n1 == 0 || n2 == 0) return VSG_OK;
  VSG_REQUIRE(d1 && d2 && out, "vsg_tiou: null pointer");
  const int g = grid_for(broadcast ? (int64_t)n1 * n2 : n1, 256, 8);
  if (dtype == 0)
    tiou_kernel<int64_t><<<g, 256, 0, (cudaStream_t)stream>>>((const int64_t*)d1, n1, (const int64_t*)d2, n2, broadcast, generalized, out);
  else
    tiou_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)d1, n1, (const float*)d2, n2, broadcast, generalized, out);
  return check_launch("vsg_tiou");
}

extern "C" int vsg_stretch_rows(const float* src, int ld, int width, const int64_t* off, int n_tracks, int tmax, float* out,
                                void* stream) {
  VSG_REQUIRE(n_tracks >= 0 && tmax >= 0 && width >= 0 && ld >= width, "vsg_stretch_rows: bad extents");
  if (n_tracks == 0 || tmax == 0 || width == 0) return VSG_OK;
  VSG_REQUIRE(src && off && out, "vsg_stretch_rows: null pointer");
  int64_t rows = (int64_t)n_tracks * tmax;
  const int grid = (int)(rows < (int64_t)sm_count() * 32 ? rows : (int64_t)sm_count() * 32);
  stretch_rows_kernel<<<grid, width >= 256 ? 256 : 64, 0, (cudaStream_t)stream>>>(src, ld, width, off, n_tracks, tmax, out);
  return check_launch("vsg_stretch_rows");
}

extern "C" int vsg_unique_rows(const int64_t* rows, int n, int d, int32_t* order, int32_t* group, int32_t* n_groups, void* stream) {
  VSG_REQUIRE(n >= 0 && d >= 1, "vsg_unique_rows: bad extents");
  VSG_REQUIRE(n <= 8192, "vsg_unique_rows: at most 8192 rows (one CTA sorts them in shared memory), got %d", n);
  VSG_REQUIRE(n_groups != nullptr, "vsg_unique_rows: null n_groups");
  if (n == 0) { cudaMemsetAsync(n_groups, 0, sizeof(int32_t), (cudaStream_t)stream); return VSG_OK; }
  VSG_REQUIRE(rows && order && group, "vsg_unique_rows: null pointer");
  int p2 = 1;
  while (p2 < n) p2 <<= 1;
  unique_rows_kernel<<<1, 1024, p2 * sizeof(int32_t), (cudaStream_t)stream>>>(rows, n, d, p2, order, group, n_groups);
  return check_launch("vsg_unique_rows");
}
