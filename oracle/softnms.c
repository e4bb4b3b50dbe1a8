/* TEST INFRASTRUCTURE ONLY — plain-C restatement of the reference's native 1-D NMS
 * (ViLCo MQ/libs/utils/csrc/nms_cpu.cpp).  Pinned bit-for-bit against the reference extension by
 * tests/test_oracle_golden.py (tests/golden/nms.npz) and tests/test_oracle_vs_reference.py.
 *
 * oracle_softnms_1d follows softnms_1d_cpu (nms_cpu.cpp:67-160): selection-sort style soft-NMS over
 * parallel arrays x1, x2, score, area(= x2 - x1 + 1e-6f), index; first maximum wins; suppressed
 * entries are replaced by the last entry (swap order is load-bearing for tie-breaking).
 * oracle_nms_1d follows nms_1d_cpu (nms_cpu.cpp:19-57) given the score-descending order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* segs: n x 2, scores: n, dets: n x 3 (out), inds: n (out).  returns number kept. */
int64_t oracle_softnms_1d(const float* segs, const float* scores, int64_t n, float iou_threshold, float sigma,
                          float min_score, int method, float* dets, int64_t* inds) {
  if (n == 0) return 0;
  float* x1 = (float*)malloc(sizeof(float) * n);
  float* x2 = (float*)malloc(sizeof(float) * n);
  float* sc = (float*)malloc(sizeof(float) * n);
  float* ar = (float*)malloc(sizeof(float) * n);
  for (int64_t i = 0; i < n; ++i) {
    x1[i] = segs[2 * i];
    x2[i] = segs[2 * i + 1];
    sc[i] = scores[i];
    ar[i] = (x2[i] - x1[i]) + 1e-6f;
    inds[i] = i;
  }
  int64_t nsegs = n;
  for (int64_t i = 0; i < nsegs; ++i) {
    float max_score = sc[i];
    int64_t max_pos = i;
    for (int64_t pos = i + 1; pos < nsegs; ++pos) {
      if (max_score < sc[pos]) {
        max_score = sc[pos];
        max_pos = pos;
      }
    }
    float ix1 = x1[max_pos], ix2 = x2[max_pos], isc = sc[max_pos], iar = ar[max_pos];
    int64_t iind = inds[max_pos];
    dets[3 * i] = ix1; dets[3 * i + 1] = ix2; dets[3 * i + 2] = isc;
    x1[max_pos] = x1[i]; x2[max_pos] = x2[i]; sc[max_pos] = sc[i]; ar[max_pos] = ar[i]; inds[max_pos] = inds[i];
    x1[i] = ix1; x2[i] = ix2; sc[i] = isc; ar[i] = iar; inds[i] = iind;
    int64_t pos = i + 1;
    while (pos < nsegs) {
      float xx1 = ix1 > x1[pos] ? ix1 : x1[pos];
      float xx2 = ix2 < x2[pos] ? ix2 : x2[pos];
      float inter = xx2 - xx1;
      if (inter < 0.f) inter = 0.f;
      float ovr = inter / (iar + ar[pos] - inter);
      float weight = 1.f;
      if (method == 0) {
        if (ovr >= iou_threshold) weight = 0.f;
      } else if (method == 1) {
        if (ovr >= iou_threshold) weight = 1.f - ovr;
      } else if (method == 2) {
        weight = expf(-(ovr * ovr) / sigma);
      }
      sc[pos] *= weight;
      if (sc[pos] < min_score) {
        x1[pos] = x1[nsegs - 1]; x2[pos] = x2[nsegs - 1]; sc[pos] = sc[nsegs - 1];
        ar[pos] = ar[nsegs - 1]; inds[pos] = inds[nsegs - 1];
        nsegs -= 1;
        pos -= 1;
      }
      pos += 1;
    }
  }
  free(x1); free(x2); free(sc); free(ar);
  return nsegs;
}

/* order: indices sorted by score descending (computed by the caller exactly like the reference: a torch sort).
 * keep: n (out, 0/1 per sorted position). */
void oracle_nms_1d(const float* segs, const int64_t* order, int64_t n, float iou_threshold, uint8_t* keep) {
  for (int64_t i = 0; i < n; ++i) keep[i] = 1;
  for (int64_t _i = 0; _i < n; ++_i) {
    if (!keep[_i]) continue;
    int64_t i = order[_i];
    float ix1 = segs[2 * i], ix2 = segs[2 * i + 1];
    float iar = (ix2 - ix1) + 1e-6f;
    for (int64_t _j = _i + 1; _j < n; ++_j) {
      if (!keep[_j]) continue;
      int64_t j = order[_j];
      float xx1 = ix1 > segs[2 * j] ? ix1 : segs[2 * j];
      float xx2 = ix2 < segs[2 * j + 1] ? ix2 : segs[2 * j + 1];
      float inter = xx2 - xx1;
      if (inter < 0.f) inter = 0.f;
      float ovr = inter / (iar + ((segs[2 * j + 1] - segs[2 * j]) + 1e-6f) - inter);
      if (ovr >= iou_threshold) keep[_j] = 0;
    }
  }
}
