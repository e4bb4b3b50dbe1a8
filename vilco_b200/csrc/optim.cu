// Flat-buffer optimizer kernels for the training step: global gradient norm, clip coefficient and a fused AdamW update that
// also refreshes the bf16 (hi, lo) operand planes the GEMMs read, so the weights never need a separate re-pack pass.
// Reference semantics: torch.optim.AdamW as built by make_optimizer (MQ/libs/utils/train_utils.py:68-143) preceded by
// torch.nn.utils.clip_grad_norm_ (train_utils.py:345-349).
#include "common.cuh"

namespace vilco {

// deterministic two-stage sum of squares (data-parallel replicas must compute bit-identical clip coefficients, so no
// floating-point atomics): stage 1 writes one partial per block, stage 2 reduces the partials in a fixed order.
__global__ void __launch_bounds__(256) sumsq_kernel(const float4* __restrict__ x, long long n4, const float* __restrict__ tail,
                                                    int ntail, float* __restrict__ partial) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < ntail) acc += tail[threadIdx.x] * tail[threadIdx.x];
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    partial[blockIdx.x] = t;
  }
}

// coef = min(1, max_norm / (sqrt(sumsq) + 1e-6))  (clip_grad_norm_); norm_out = sqrt(sumsq).  One block of 256 threads.
__global__ void __launch_bounds__(256) clip_coef_kernel(const float* __restrict__ partial, int nparts, float max_norm,
                                                        float* __restrict__ coef, float* __restrict__ norm_out) {
  __shared__ float red[256];
  float t = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 256) t += partial[i];
  red[threadIdx.x] = t;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float nrm = sqrtf(red[0]);
    if (norm_out) *norm_out = nrm;
    *coef = max_norm > 0.f ? fminf(1.f, max_norm / (nrm + 1e-6f)) : 1.f;
  }
}

__device__ __forceinline__ float adamw_one(float& p, float g, float& m, float& v, float lr, float b1, float b2, float eps,
                                           float wd, float step, float bc2_sqrt) {
  float pi = p * (1.f - lr * wd);
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  pi -= step * m / (sqrtf(v) / bc2_sqrt + eps);
  p = pi;
  return pi;
}

// n is a multiple of 4 and all buffers are 16-byte aligned (FlatAdamW pads every tensor to 8 elements)
__global__ void __launch_bounds__(256) adamw_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                    float4* __restrict__ v, long long n4, float lr, float b1, float b2, float eps,
                                                    float wd, float bc1, float bc2_sqrt, const float* __restrict__ gscale,
                                                    __nv_bfloat16* __restrict__ planes, long long planes_lo, int fmt) {
  const float gs = gscale ? *gscale : 1.f;
  const float step = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pi = p[i], mi = m[i], vi = v[i];
    const float4 gi = g[i];
    adamw_one(pi.x, gi.x * gs, mi.x, vi.x, lr, b1, b2, eps, wd, step, bc2_sqrt);
    adamw_one(pi.y, gi.y * gs, mi.y, vi.y, lr, b1, b2, eps, wd, step, bc2_sqrt);
    adamw_one(pi.z, gi.z * gs, mi.z, vi.z, lr, b1, b2, eps, wd, step, bc2_sqrt);
    adamw_one(pi.w, gi.w * gs, mi.w, vi.w, lr, b1, b2, eps, wd, step, bc2_sqrt);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (planes) {
      uint32_t h0, h1, l0, l1;
      split16x2(pi.x, pi.y, fmt, h0, l0);
      split16x2(pi.z, pi.w, fmt, h1, l1);
      *reinterpret_cast<uint2*>(planes + 4 * i) = make_uint2(h0, h1);
      if (planes_lo) *reinterpret_cast<uint2*>(planes + planes_lo + 4 * i) = make_uint2(l0, l1);
    }
  }
}

}  // namespace vilco

using namespace vilco;

static inline int ogrid(long long n, int block) {
  long long b = (n + block - 1) / block;
  const long long cap = 148LL * 8;   // also bounds the number of sum-of-squares partials (VILCO_CLIP_SCRATCH)
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int vilco_grad_clip_coef(const float* g, int64_t n, float max_norm, float* scratch, float* coef, float* norm_out,
                                    void* stream) {
  VILCO_CHECK_ARG(g && scratch && coef && n > 0, "vilco_grad_clip_coef: bad arguments");
  VILCO_CHECK_ARG(reinterpret_cast<uintptr_t>(g) % 16 == 0, "vilco_grad_clip_coef: gradient buffer must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n4 = n / 4;
  const int grid = ogrid(n4, 256);   // <= 148 * 8 partials; scratch must hold VILCO_CLIP_SCRATCH floats
  sumsq_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(g), n4, g + n4 * 4, static_cast<int>(n - n4 * 4), scratch);
  VILCO_LAUNCH_CHECK();
  clip_coef_kernel<<<1, 256, 0, st>>>(scratch, grid, max_norm, coef, norm_out);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                           float weight_decay, int step, const float* grad_scale, void* planes, int64_t planes_lo, void* stream) {
  VILCO_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "vilco_adamw: bad arguments");
  VILCO_CHECK_ARG(n % 4 == 0 && reinterpret_cast<uintptr_t>(p) % 16 == 0 && reinterpret_cast<uintptr_t>(g) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(m) % 16 == 0 && reinterpret_cast<uintptr_t>(v) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(planes) % 8 == 0 && planes_lo % 4 == 0,
                  "vilco_adamw: buffers must be 16-byte aligned and n a multiple of 4");
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;   // seven 16-byte streams per thread: keep many warps in flight
  adamw_kernel<<<static_cast<unsigned>(blocks < 1 ? 1 : blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n / 4, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale, static_cast<__nv_bfloat16*>(planes), planes_lo,
      act_fmt());
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}
