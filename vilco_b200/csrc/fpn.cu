// Kernels only FPN1D needs (reference MQ/libs/modeling/necks.py:13-106 with ACConv / DenseAPP, modeling/utils.py:671-751):
//   * GroupNorm over a token-major (B, T, C) tensor: statistics per (clip, group) over all T rows and C / G channels of the
//     group — nn.GroupNorm on (B, C, T) — two-pass fp32 (mean, then centred variance, like torch), affine, optional ReLU,
//     fp32 and / or 16-bit plane outputs.  The DenseAPP stack runs on the LAST pyramid level (T = 2 at the MQ configuration),
//     so one CTA per (clip, group) is plenty.
//   * the top-down path: y[b, t, :] += x[b, t / 2, :]   (F.interpolate(scale_factor=2, mode="nearest") + add).
#include "common.cuh"

namespace vilco {

__global__ void __launch_bounds__(256) groupnorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, float* __restrict__ y32, uint16_t* __restrict__ y16,
                                                        long long y16_lo, int T, int C, int G, float eps, int relu, int fmt) {
  __shared__ float red[32];
  __shared__ float s_stat[2];
  const int g = blockIdx.x, bi = blockIdx.y;
  const int cpg = C / G;
  const long long base = (long long)bi * T * C + (long long)g * cpg;
  const int n = T * cpg;
  auto block_sum = [&](float v) -> float {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) t = warp_sum(t);
    return t;   // valid in warp 0
  };
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[base + (long long)(i / cpg) * C + (i % cpg)];
  s = block_sum(s);
  if (threadIdx.x == 0) s_stat[0] = s / n;
  __syncthreads();
  const float mean = s_stat[0];
  float q = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = x[base + (long long)(i / cpg) * C + (i % cpg)] - mean;
    q += d * d;
  }
  q = block_sum(q);
  if (threadIdx.x == 0) s_stat[1] = rsqrtf(q / n + eps);
  __syncthreads();
  const float rstd = s_stat[1];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = g * cpg + (i % cpg);
    const long long o = base + (long long)(i / cpg) * C + (i % cpg);
    float v = (x[o] - mean) * rstd * w[c] + b[c];
    if (relu) v = fmaxf(v, 0.f);
    if (y32) y32[o] = v;
    if (y16) store16_split(y16, o, y16_lo, v, fmt);
  }
}

__global__ void upsample2_add_kernel(const float* __restrict__ x, float* __restrict__ y, long long n4, int T2, int C4) {
  // y (B, T2, C) += x (B, T2 / 2, C); one float4 per thread
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / C4;
    const int c = (int)(i - row * C4);
    const long long bi = row / T2;
    const int t = (int)(row - bi * T2);
    const float4 a = reinterpret_cast<const float4*>(x)[(bi * (T2 / 2) + t / 2) * C4 + c];
    float4 v = reinterpret_cast<float4*>(y)[i];
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    reinterpret_cast<float4*>(y)[i] = v;
  }
}

}  // namespace vilco

using namespace vilco;

extern "C" int vilco_groupnorm(const float* x, const float* w, const float* b, float* y32, void* y16, int64_t y16_lo, int B, int T,
                               int C, int G, float eps, int relu, void* stream) {
  VILCO_CHECK_ARG(x && w && b && (y32 || y16), "vilco_groupnorm: null pointer");
  VILCO_CHECK_ARG(B > 0 && T > 0 && G > 0 && C % G == 0, "vilco_groupnorm: C=%d must be divisible by G=%d", C, G);
  groupnorm_kernel<<<dim3(G, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, w, b, y32, static_cast<uint16_t*>(y16), y16_lo, T, C, G,
                                                                             eps, relu, act_fmt());
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_upsample2_add(const float* x, float* y, int B, int T2, int C, void* stream) {
  VILCO_CHECK_ARG(x && y && B > 0 && T2 > 0 && T2 % 2 == 0 && C % 4 == 0, "vilco_upsample2_add: bad arguments");
  const long long n4 = (long long)B * T2 * (C / 4);
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  upsample2_add_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n4, T2, C / 4);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}
