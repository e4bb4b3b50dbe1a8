// Fused Moment-Query losses (forward) — PtTransformer.losses, MQ/libs/modeling/meta_archs.py:1374-1480 with
// sigmoid_focal_loss (MQ/libs/modeling/losses.py:5-51) and ctr_diou_loss_1d (losses.py:109-168).
// One pass over the concatenated pyramid rows: a warp per (video, point) row reduces the focal loss over the K classes
// with shuffles; positives add their DIoU term; the per-(video, class) maxima of softmax_K(logits) needed by the
// label-involved ("al") loss are collected with atomicMax on the (non-negative) float bit patterns.
#include "common.cuh"

namespace vilco {

struct LossParams {
  const float* logits;   // (B, P, K)
  const float* offsets;  // (B, P, 2)
  const float* pmask;    // (B, P)   1 = valid point
  const uint8_t* gap;    // (P,) 1 = pyramid gap row (does not exist in the reference), may be NULL
  const float* gt_cls;   // (B, P, K) 0/1 targets
  const float* gt_off;   // (B, P, 2)
  const float* w_cls;    // (B, P) gaussian weight of the matched segment (classification)
  const float* w_l;      // (B, P)
  const float* w_r;      // (B, P)
  int B, P, K;
  float alpha, gamma;
  float* sums;           // [0] cls_sum, [1] reg_sum, [2] num_pos   (zero-initialised by the caller)
  unsigned int* smax;    // (B, K) running max of softmax probabilities as uint bits (zero-initialised)
};

__device__ __forceinline__ float focal_term(float x, float t, float alpha, float gamma) {
  const float p = 1.0f / (1.0f + expf(-x));
  // binary_cross_entropy_with_logits: max(x,0) - x*t + log(1 + exp(-|x|))
  const float ce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
  const float p_t = p * t + (1.f - p) * (1.f - t);
  const float mod = gamma == 2.0f ? (1.f - p_t) * (1.f - p_t) : powf(1.f - p_t, gamma);
  float loss = ce * mod;
  if (alpha >= 0.f) loss *= alpha * t + (1.f - alpha) * (1.f - t);
  return loss;
}

__global__ void __launch_bounds__(256) mq_loss_kernel(const LossParams p) {
  __shared__ float s_part[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rows = (long long)p.B * p.P;
  float cls_acc = 0.f, reg_acc = 0.f, pos_acc = 0.f;
  for (long long row = blockIdx.x * 8LL + warp; row < rows; row += gridDim.x * 8LL) {
    const int b = static_cast<int>(row / p.P);
    if (p.pmask[row] == 0.f) {  // warp-uniform
      // a padded point has all logits filled with -1e7, i.e. a uniform softmax 1/K that still enters the max over
      // points (meta_archs.py:1437-1439); pyramid gap rows are an artefact of our layout and are skipped
      if (!(p.gap && p.gap[row - (long long)b * p.P]))
        for (int k = lane; k < p.K; k += 32) atomicMax(p.smax + (long long)b * p.K + k, __float_as_uint(1.0f / p.K));
      continue;
    }
    const float* lg = p.logits + row * p.K;
    const float* gt = p.gt_cls + row * p.K;
    float fl = 0.f, tsum = 0.f, mx = -INFINITY;
    for (int k = lane; k < p.K; k += 32) {
      const float x = lg[k], t = gt[k];
      fl += focal_term(x, t, p.alpha, p.gamma);
      tsum += t;
      mx = fmaxf(mx, x);
    }
    fl = warp_sum(fl); tsum = warp_sum(tsum); mx = warp_max(mx);
    float se = 0.f;
    for (int k = lane; k < p.K; k += 32) se += expf(lg[k] - mx);
    se = warp_sum(se);
    for (int k = lane; k < p.K; k += 32) {
      const float pr = expf(lg[k] - mx) / se;  // softmax over classes, >= 0
      atomicMax(p.smax + (long long)b * p.K + k, __float_as_uint(pr));
    }
    const bool pos = tsum > 0.f;
    const float wc = pos ? p.w_cls[row] : 1.0f;  // negatives weigh 1 (meta_archs.py:1430)
    if (lane == 0) {
      cls_acc += fl * wc;
      if (pos) {
        const float lp = p.offsets[2 * row], rp = p.offsets[2 * row + 1];
        const float lgt = p.gt_off[2 * row], rgt = p.gt_off[2 * row + 1];
        const float inter = fminf(lp, lgt) + fminf(rp, rgt);
        const float uni = (lp + rp) + (lgt + rgt) - inter;
        const float iou = inter / fmaxf(uni, 1e-8f);
        const float len_c = fmaxf(lp, lgt) + fmaxf(rp, rgt);
        const float rho = 0.5f * (rp - lp - rgt + lgt);
        const float q = rho / fmaxf(len_c, 1e-8f);
        const float diou = 1.0f - iou + q * q;
        reg_acc += diou * ((p.w_l[row] + p.w_r[row]) * 0.5f) * wc;
        pos_acc += 1.f;
      }
    }
  }
  if (lane == 0) { s_part[0][warp] = cls_acc; s_part[1][warp] = reg_acc; s_part[2][warp] = pos_acc; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_part[threadIdx.x][w];
    if (t != 0.f) atomicAdd(p.sums + threadIdx.x, t);
  }
}

// al loss: sum_{b,k} -inv*log(s) - (1-inv)*log(1-s)   (meta_archs.py:1436-1446); inv (B,K) 0/1 marks present labels
__global__ void al_loss_kernel(const unsigned int* smax, const float* inv, int n, float* out) {
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float s = __uint_as_float(smax[i]);
    acc += -inv[i] * logf(s) - (1.f - inv[i]) * logf(1.f - s);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0 && acc != 0.f) atomicAdd(out, acc);
}

// ---------------------------------------------------------------------------------------------
// backward of the fused losses: gradients w.r.t. logits, offsets and the three gaussian weights.
// final = cls + w_reg * reg + w_al * al, every term already divided by the loss normaliser `norm`.
// ---------------------------------------------------------------------------------------------
struct LossBwdParams {
  LossParams f;            // forward inputs (smax holds the forward maxima as uint bits)
  const float* present;    // (B, K)
  float norm, w_reg, w_al;
  float* dlogits;          // (B, P, K)
  float* doffsets;         // (B, P, 2)
  float* dw_cls; float* dw_l; float* dw_r;  // (B, P)
};

__global__ void __launch_bounds__(256) mq_loss_bwd_kernel(const LossBwdParams q) {
  const LossParams& p = q.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long rows = (long long)p.B * p.P;
  const float inv_n = 1.0f / q.norm;
  for (long long row = blockIdx.x * 8LL + warp; row < rows; row += gridDim.x * 8LL) {
    const int b = static_cast<int>(row / p.P);
    float* dl = q.dlogits + row * p.K;
    if (p.pmask[row] == 0.f) {
      for (int k = lane; k < p.K; k += 32) dl[k] = 0.f;
      if (lane == 0) { q.doffsets[2 * row] = 0.f; q.doffsets[2 * row + 1] = 0.f; q.dw_cls[row] = 0.f; q.dw_l[row] = 0.f; q.dw_r[row] = 0.f; }
      continue;
    }
    const float* lg = p.logits + row * p.K;
    const float* gt = p.gt_cls + row * p.K;
    float fl = 0.f, tsum = 0.f, mx = -INFINITY;
    for (int k = lane; k < p.K; k += 32) {
      fl += focal_term(lg[k], gt[k], p.alpha, p.gamma);
      tsum += gt[k];
      mx = fmaxf(mx, lg[k]);
    }
    fl = warp_sum(fl); tsum = warp_sum(tsum); mx = warp_max(mx);
    float se = 0.f;
    for (int k = lane; k < p.K; k += 32) se += expf(lg[k] - mx);
    se = warp_sum(se);
    const bool pos = tsum > 0.f;
    const float wc = pos ? p.w_cls[row] : 1.0f;
    // al loss: this row carries the gradient of class k when its softmax probability is the maximum over the points
    float acc_gs = 0.f;   // sum_k g_k * s_k  over the classes this row is the arg-max of
    for (int k = lane; k < p.K; k += 32) {
      const float sk = expf(lg[k] - mx) / se;
      const float sm = __uint_as_float(p.smax[(long long)b * p.K + k]);
      if (sk == sm) {
        const float inv = q.present[(long long)b * p.K + k];
        acc_gs += (-inv / sm + (1.f - inv) / (1.f - sm)) * sk;
      }
    }
    acc_gs = warp_sum(acc_gs);
    for (int k = lane; k < p.K; k += 32) {
      const float x = lg[k], t = gt[k];
      const float pr = 1.0f / (1.0f + expf(-x));
      const float ce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
      const float p_t = pr * t + (1.f - pr) * (1.f - t);
      const float om = 1.f - p_t;
      const float a_t = p.alpha >= 0.f ? p.alpha * t + (1.f - p.alpha) * (1.f - t) : 1.f;
      const float dpt = pr * (1.f - pr) * (2.f * t - 1.f);
      const float mod = p.gamma == 2.0f ? om * om : powf(om, p.gamma);
      const float dmod = p.gamma == 2.0f ? -2.f * om * dpt : -p.gamma * powf(om, p.gamma - 1.f) * dpt;
      float g = a_t * ((pr - t) * mod + ce * dmod) * wc * inv_n;
      // softmax Jacobian of the al term: sum_k' g_k' s_k' (delta_kk' - s_k)
      const float sk = expf(lg[k] - mx) / se;
      const float sm = __uint_as_float(p.smax[(long long)b * p.K + k]);
      float gal = -acc_gs * sk;
      if (sk == sm) {
        const float inv = q.present[(long long)b * p.K + k];
        gal += (-inv / sm + (1.f - inv) / (1.f - sm)) * sk;
      }
      dl[k] = g + q.w_al * inv_n * gal;
    }
    if (lane == 0) {
      float dlp = 0.f, drp = 0.f, dwc = 0.f, dwl = 0.f, dwr = 0.f;
      if (pos) {
        const float lp = p.offsets[2 * row], rp = p.offsets[2 * row + 1];
        const float lgt = p.gt_off[2 * row], rgt = p.gt_off[2 * row + 1];
        const float inter = fminf(lp, lgt) + fminf(rp, rgt);
        const float uni = (lp + rp) + (lgt + rgt) - inter;
        const float unic = fmaxf(uni, 1e-8f);
        const float iou = inter / unic;
        const float len_c = fmaxf(lp, lgt) + fmaxf(rp, rgt);
        const float lenc = fmaxf(len_c, 1e-8f);
        const float rho = 0.5f * (rp - lp - rgt + lgt);
        const float qq = rho / lenc;
        const float diou = 1.0f - iou + qq * qq;
        const float wreg = (p.w_l[row] + p.w_r[row]) * 0.5f;
        // d/d lp and d/d rp
        const float dil = lp < lgt ? 1.f : 0.f, dir = rp < rgt ? 1.f : 0.f;        // d inter
        const float dul = uni > 1e-8f ? 1.f - dil : 0.f, dur = uni > 1e-8f ? 1.f - dir : 0.f;
        const float diou_l = (dil * unic - inter * dul) / (unic * unic);
        const float diou_r = (dir * unic - inter * dur) / (unic * unic);
        const float dcl = (len_c > 1e-8f && lp > lgt) ? 1.f : 0.f, dcr = (len_c > 1e-8f && rp > rgt) ? 1.f : 0.f;
        const float dq_l = (-0.5f * lenc - rho * dcl) / (lenc * lenc);
        const float dq_r = (0.5f * lenc - rho * dcr) / (lenc * lenc);
        const float gscale = q.w_reg * wreg * wc * inv_n;
        dlp = gscale * (-diou_l + 2.f * qq * dq_l);
        drp = gscale * (-diou_r + 2.f * qq * dq_r);
        dwc = fl * inv_n + q.w_reg * diou * wreg * inv_n;
        dwl = q.w_reg * diou * 0.5f * wc * inv_n;
        dwr = dwl;
      }
      q.doffsets[2 * row] = dlp; q.doffsets[2 * row + 1] = drp;
      q.dw_cls[row] = dwc; q.dw_l[row] = dwl; q.dw_r[row] = dwr;
    }
  }
}

}  // namespace vilco

using namespace vilco;

extern "C" int vilco_mq_losses(const float* logits, const float* offsets, const float* pmask, const uint8_t* gap, const float* gt_cls,
                               const float* gt_off, const float* w_cls, const float* w_l, const float* w_r,
                               const float* present, int B, int P, int K, float alpha, float gamma, float* sums4,
                               unsigned int* smax_scratch, void* stream) {
  VILCO_CHECK_ARG(logits && offsets && pmask && gt_cls && gt_off && w_cls && w_l && w_r && present && sums4 && smax_scratch,
                  "vilco_mq_losses: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VILCO_CUDA(cudaMemsetAsync(sums4, 0, 4 * sizeof(float), st));
  VILCO_CUDA(cudaMemsetAsync(smax_scratch, 0, sizeof(unsigned int) * (size_t)B * K, st));
  LossParams p{};
  p.logits = logits; p.offsets = offsets; p.pmask = pmask; p.gap = gap; p.gt_cls = gt_cls; p.gt_off = gt_off;
  p.w_cls = w_cls; p.w_l = w_l; p.w_r = w_r; p.B = B; p.P = P; p.K = K; p.alpha = alpha; p.gamma = gamma;
  p.sums = sums4; p.smax = smax_scratch;
  const long long rows = (long long)B * P;
  int grid = static_cast<int>((rows + 7) / 8);
  if (grid > 148 * 8) grid = 148 * 8;
  mq_loss_kernel<<<grid, 256, 0, st>>>(p);
  VILCO_LAUNCH_CHECK();
  al_loss_kernel<<<1, 256, 0, st>>>(smax_scratch, present, B * K, sums4 + 3);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_mq_losses_bwd(const float* logits, const float* offsets, const float* pmask, const uint8_t* gap,
                                   const float* gt_cls, const float* gt_off, const float* w_cls, const float* w_l,
                                   const float* w_r, const float* present, const unsigned int* smax, int B, int P, int K,
                                   float alpha, float gamma, float norm, float w_reg, float w_al, float* dlogits,
                                   float* doffsets, float* dw_cls, float* dw_l, float* dw_r, void* stream) {
  VILCO_CHECK_ARG(logits && offsets && pmask && gt_cls && gt_off && w_cls && w_l && w_r && present && smax && dlogits &&
                  doffsets && dw_cls && dw_l && dw_r, "vilco_mq_losses_bwd: null pointer");
  LossBwdParams q{};
  q.f.logits = logits; q.f.offsets = offsets; q.f.pmask = pmask; q.f.gap = gap; q.f.gt_cls = gt_cls; q.f.gt_off = gt_off;
  q.f.w_cls = w_cls; q.f.w_l = w_l; q.f.w_r = w_r; q.f.B = B; q.f.P = P; q.f.K = K; q.f.alpha = alpha; q.f.gamma = gamma;
  q.f.smax = const_cast<unsigned int*>(smax);
  q.present = present; q.norm = norm; q.w_reg = w_reg; q.w_al = w_al;
  q.dlogits = dlogits; q.doffsets = doffsets; q.dw_cls = dw_cls; q.dw_l = dw_l; q.dw_r = dw_r;
  const long long rows = (long long)B * P;
  int grid = static_cast<int>((rows + 7) / 8);
  if (grid > 148 * 8) grid = 148 * 8;
  mq_loss_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(q);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}
