// Shared device/host helpers for the vilco_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/vilco_b200.h"

namespace vilco {

// ---- host-side error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

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

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == VILCO_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == VILCO_ACT_GELU) return gelu_erf(v);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

}  // namespace vilco
