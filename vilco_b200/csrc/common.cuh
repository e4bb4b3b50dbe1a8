// Shared device/host helpers for the vilco_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/vilco_b200.h"

namespace vilco {

// ---- host-side error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// element format of every 16-bit operand plane (VILCO_BF16 or VILCO_F16; vilco_set_plane_format).  tcgen05 kind::f16 cannot
// mix fp16 and bf16 operands in one MMA, so the gradient planes the backward kernels emit use the same format; with fp16
// they are stored pre-multiplied by grad_scale() (a power of two that centres gradient magnitudes in the fp16 range) and the
// caller folds 1 / grad_scale into the alpha of every GEMM that consumes them.
int act_fmt();
float grad_scale();

#define VILCO_CHECK_ARG(cond, ...)                         \
  do {                                                     \
    if (!(cond)) {                                         \
      ::vilco::set_error(__VA_ARGS__);                     \
      return VILCO_E_ARG;                                  \
    }                                                      \
  } while (0)

#define VILCO_CUDA(expr)                                                          \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      ::vilco::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                         __FILE__, __LINE__);                                     \
      return VILCO_E_CUDA;                                                        \
    }                                                                             \
  } while (0)

#define VILCO_LAUNCH_CHECK()                                                      \
  do {                                                                            \
    ::vilco::count_launch();                                                      \
    cudaError_t _e = cudaGetLastError();                                          \
    if (_e != cudaSuccess) {                                                      \
      ::vilco::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                     \
      return VILCO_E_CUDA;                                                        \
    }                                                                             \
  } while (0)

// ---- device helpers -----------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// erf-GELU with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7): one MUFU.RCP, one MUFU.EX2 and 9 FMAs, branch-free,
// against ~35 divergent instructions of erff().  Measured over [-12, 12] in fp32: max |gelu_fast - exact| = 4.7e-7, the same
// as the fp32 rounding of 0.5 x (1 + erff(x / sqrt 2)) itself (4.5e-7).  Used by the GEMM epilogue (FFN up-projection).
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = x * 0.70710678118654752440f;
  const float a = fabsf(z);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, a, 1.0f)));
  float q = fmaf(1.061405429f, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  q *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-a * a * 1.4426950408889634f));
  const float er = copysignf(fmaf(-q, e, 1.0f), z);
  const float hx = 0.5f * x;
  return fmaf(hx, er, hx);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == VILCO_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == VILCO_ACT_GELU) return gelu_fast(v);
  if (act == VILCO_ACT_EXP2) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v)); return e; }
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// ---- 16-bit operand planes in either format (fmt is warp-uniform: VILCO_BF16 or VILCO_F16) -------------------------------
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi, int fmt) {
  if (fmt == VILCO_F16) {   // saturate instead of producing inf (fp16 max = 65504): one F2FP.SATFINITE, NaN stays NaN
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  return pack_bf16x2(lo, hi);
}
__device__ __forceinline__ float2 unpack16x2(uint32_t u, int fmt) {
  if (fmt == VILCO_F16) return __half22float2(*reinterpret_cast<__half2*>(&u));
  return make_float2(bf16_lo(u), bf16_hi(u));
}
__device__ __forceinline__ uint16_t pack16(float v, int fmt) {
  if (fmt == VILCO_F16) return __half_as_ushort(__float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)));
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float unpack16(uint16_t u, int fmt) {
  if (fmt == VILCO_F16) return __half2float(__ushort_as_half(u));
  return __uint_as_float(static_cast<uint32_t>(u) << 16);
}
// hi plane word + the residual (lo) plane word of two values
__device__ __forceinline__ void split16x2(float a, float b, int fmt, uint32_t& hi, uint32_t& lo) {
  hi = pack16x2(a, b, fmt);
  const float2 h = unpack16x2(hi, fmt);
  lo = pack16x2(a - h.x, b - h.y, fmt);
}
// scalar store of one value into the hi (and, when lo != 0, the lo) plane
__device__ __forceinline__ void store16_split(uint16_t* D, long long off, long long lo_off, float o, int fmt) {
  const uint16_t h = pack16(o, fmt);
  D[off] = h;
  if (lo_off) D[lo_off + off] = pack16(o - unpack16(h, fmt), fmt);
}
__device__ __forceinline__ float load16_split(const uint16_t* p, long long lo, int fmt) {
  float v = unpack16(p[0], fmt);
  if (lo) v += unpack16(p[lo], fmt);
  return v;
}

}  // namespace vilco
