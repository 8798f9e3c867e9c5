// Host-side (CPU, native) finishing pass of the relation evaluation: per-video AP / TP@K / tagging P@K records from the
// device matcher's outputs.  Restates the numpy arithmetic of VidVRDhelperEvalAPIs/visual_relation_detection.py:28-33
// (float32 cumulative precision / recall), common.py:4-37 (voc_ap in float64), :37-58 (tagging) and :82-93 so that
// the Python layer does not loop over videos (SURVEY.md section 8a row A14 stays host work, just not interpreted).
#include "common.cuh"
#include <math.h>
#include <float.h>
#include <unordered_set>
#include <vector>

namespace {

struct TripletKey {
  int64_t s, p, o;
  bool operator==(const TripletKey& k) const { return s == k.s && p == k.p && o == k.o; }
};
struct TripletHash {
  size_t operator()(const TripletKey& k) const {
    uint64_t h = (uint64_t)k.s * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)k.p + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= (uint64_t)k.o + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    return (size_t)h;
  }
};

// voc_ap(rec, prec), non-07 variant (common.py:22-36): float64 on float32-valued inputs
double voc_ap(const std::vector<float>& rec, const std::vector<float>& prec) {
  const size_t n = rec.size();
  std::vector<double> r(n + 2), p(n + 2);
  r[0] = 0.0; p[0] = 0.0;
  for (size_t i = 0; i < n; ++i) { r[i + 1] = (double)rec[i]; p[i + 1] = (double)prec[i]; }
  r[n + 1] = 1.0; p[n + 1] = 0.0;
  for (size_t i = n + 1; i > 0; --i) p[i - 1] = p[i - 1] > p[i] ? p[i - 1] : p[i];
  double ap = 0.0;
  for (size_t i = 0; i + 1 < n + 2; ++i)
    if (r[i + 1] != r[i]) ap += (r[i + 1] - r[i]) * p[i + 1];
  return ap;
}

}  // namespace

// hit f64[n_pred] in RANK order per video (score or -inf); order int32[n_pred] (rank -> local prediction index);
// pred_trip / gt_trip int64[n][3] (s_cat, p_cat, o_cat); p_off / g_off int64[n_vid+1].
// records f64[n_vid][3 + n_det + n_tag] = (video index, AP, n_gt, TP@det_n..., P@tag_n...); videos without GT are skipped
// (visual_relation_detection.py:73-74).  Returns the number of records written.
extern "C" int vsg_eval_records_host(const double* hit, const int32_t* order, const int64_t* pred_trip, const int64_t* p_off,
                                     const int64_t* gt_trip, const int64_t* g_off, int n_vid, const int* det_n, int n_det,
                                     const int* tag_n, int n_tag, double* records) {
  if (n_vid < 0 || n_det < 0 || n_tag < 0) { vsg::set_error("vsg_eval_records_host: negative size"); return VSG_E_INVALID; }
  if (n_vid == 0) return 0;
  if (!hit || !order || !pred_trip || !p_off || !gt_trip || !g_off || !records) { vsg::set_error("vsg_eval_records_host: null pointer"); return VSG_E_INVALID; }
  const int width = 3 + n_det + n_tag;
  const float eps = FLT_EPSILON;
  int n_rec = 0;
  std::vector<float> prec, rec, tprec;
  for (int v = 0; v < n_vid; ++v) {
    const int64_t n_gt = g_off[v + 1] - g_off[v];
    if (n_gt == 0) continue;
    const int64_t p0 = p_off[v], n = p_off[v + 1] - p0;
    double* out = records + (size_t)n_rec * width;
    out[0] = (double)v;
    out[2] = (double)n_gt;
    // detection: float32 cumulative curves (:28-33)
    prec.resize(n); rec.resize(n);
    int64_t ctp = 0, cfp = 0;
    const float denom_gt = fmaxf((float)n_gt, eps);
    for (int64_t k = 0; k < n; ++k) {
      const bool tp = isfinite(hit[p0 + k]);
      ctp += tp; cfp += !tp;
      const float fctp = (float)ctp, fcfp = (float)cfp;
      rec[k] = fctp / denom_gt;
      prec[k] = fctp / fmaxf(fctp + fcfp, eps);
      for (int d = 0; d < n_det; ++d)
        if (k + 1 == (n < det_n[d] ? n : det_n[d])) out[3 + d] = (double)ctp;
    }
    if (n == 0) for (int d = 0; d < n_det; ++d) out[3 + d] = 0.0;
    out[1] = voc_ap(rec, prec);
    // tagging (:37-58): triplets deduplicated in rank order, hit iff the triplet occurs in the video's GT
    std::unordered_set<TripletKey, TripletHash> gt_set, seen;
    for (int64_t g = g_off[v]; g < g_off[v + 1]; ++g) gt_set.insert(TripletKey{gt_trip[3 * g], gt_trip[3 * g + 1], gt_trip[3 * g + 2]});
    tprec.clear();
    int64_t ttp = 0, tfp = 0;
    for (int64_t k = 0; k < n; ++k) {
      const int64_t idx = p0 + order[p0 + k];
      const TripletKey key{pred_trip[3 * idx], pred_trip[3 * idx + 1], pred_trip[3 * idx + 2]};
      if (!seen.insert(key).second) continue;
      const bool tp = gt_set.count(key) != 0;
      ttp += tp; tfp += !tp;
      tprec.push_back((float)ttp / fmaxf((float)ttp + (float)tfp, eps));
    }
    for (int t = 0; t < n_tag; ++t) {
      const int64_t c = (int64_t)tprec.size() < tag_n[t] ? (int64_t)tprec.size() : tag_n[t];
      out[3 + n_det + t] = c > 0 ? (double)tprec[c - 1] : 0.0;
    }
    ++n_rec;
  }
  return n_rec;
}
