// Host side of the evaluation tail: the greedy prediction <-> ground-truth matching of the detection
// mAP, i.e. the `for idx, this_pred in prediction.iterrows()` loop of compute_average_precision_detection
// (MQ/libs/utils/metrics.py:302-331) with segment_iou (:348-372).  The reference runs it on CPU through pandas + joblib
// (one DataFrame row at a time); this is the same arithmetic in double precision over plain arrays.  Host code only: no
// kernel is launched, so it is callable on a box without a GPU.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

extern "C" int vilco_ap_match(const double* pred_seg, const int64_t* pred_vid, int64_t n_pred, const double* gt_seg,
                              const int64_t* gt_start, int64_t n_vid, const double* tiou_thr, int n_thr, uint8_t* tp) {
  VILCO_CHECK_ARG(n_pred >= 0 && n_vid >= 0 && n_thr >= 0, "vilco_ap_match: negative size");
  if (n_pred == 0 || n_thr == 0) return VILCO_OK;
  VILCO_CHECK_ARG(pred_seg && pred_vid && tiou_thr && tp, "vilco_ap_match: null pointer");
  VILCO_CHECK_ARG(n_vid == 0 || (gt_seg && gt_start), "vilco_ap_match: null ground truth");
  const int64_t n_gt = n_vid ? gt_start[n_vid] : 0;
  for (int64_t v = 0; v < n_vid; ++v)
    VILCO_CHECK_ARG(gt_start[v] <= gt_start[v + 1] && gt_start[v] >= 0, "vilco_ap_match: gt_start not ascending at %lld",
                    (long long)v);
  memset(tp, 0, (size_t)n_thr * (size_t)n_pred);
  if (n_gt == 0) return VILCO_OK;
  // lock_gt of metrics.py:295 — one flag per (threshold, ground-truth row)
  uint8_t* locked = (uint8_t*)calloc((size_t)n_thr * (size_t)n_gt, 1);
  if (!locked) {
    vilco::set_error("vilco_ap_match: out of host memory");
    return VILCO_E_ARG;
  }
  int64_t max_g = 0;
  for (int64_t v = 0; v < n_vid; ++v)
    if (gt_start[v + 1] - gt_start[v] > max_g) max_g = gt_start[v + 1] - gt_start[v];
  double* iou = (double*)malloc(sizeof(double) * (size_t)max_g);
  if (!iou) {
    free(locked);
    vilco::set_error("vilco_ap_match: out of host memory");
    return VILCO_E_ARG;
  }
  for (int64_t i = 0; i < n_pred; ++i) {
    const int64_t v = pred_vid[i];
    if (v < 0) continue;  // no ground truth of this label in the video: false positive at every threshold (:308-312)
    if (v >= n_vid) {
      free(iou);
      free(locked);
      vilco::set_error("vilco_ap_match: pred_vid[%lld] = %lld out of range", (long long)i, (long long)v);
      return VILCO_E_ARG;
    }
    const int64_t g0 = gt_start[v], g1 = gt_start[v + 1];
    const double t0 = pred_seg[2 * i], t1 = pred_seg[2 * i + 1];
    for (int64_t g = g0; g < g1; ++g) {  // segment_iou, metrics.py:362-372 (same operation order)
      const double c0 = gt_seg[2 * g], c1 = gt_seg[2 * g + 1];
      const double tt1 = t0 > c0 ? t0 : c0;  // np.maximum
      const double tt2 = t1 < c1 ? t1 : c1;  // np.minimum
      double inter = tt2 - tt1;
      if (inter < 0.0) inter = 0.0;
      const double uni = (c1 - c0) + (t1 - t0) - inter;
      iou[g - g0] = inter / uni;  // 0/0 = NaN kept: argsort()[::-1] visits NaN first and `nan < thr` is False (:318-321)
    }
    for (int t = 0; t < n_thr; ++t) {
      // first not-yet-matched ground truth in descending-tIoU order whose tIoU is not below the threshold; equal tIoUs
      // are visited from the higher row index down (argsort()[::-1] of numpy's insertion sort for <= 16 rows)
      const double thr = tiou_thr[t];
      uint8_t* lk = locked + (size_t)t * (size_t)n_gt + g0;
      int64_t best = -1;
      bool best_nan = false;
      for (int64_t g = 0; g < g1 - g0; ++g) {
        const double x = iou[g];
        const bool is_nan = x != x;
        if (lk[g] || (!is_nan && x < thr)) continue;
        if (best < 0 || is_nan || (!best_nan && x >= iou[best])) {
          best = g;
          best_nan = is_nan;
        }
      }
      if (best >= 0) {
        lk[best] = 1;
        tp[(size_t)t * (size_t)n_pred + i] = 1;
      }
    }
  }
  free(iou);
  free(locked);
  return VILCO_OK;
}
