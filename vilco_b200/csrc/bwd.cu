// Backward-pass building blocks (token-major fp32 gradients):
//   rows -> bf16 operand planes with optional row mask / transposition (GEMM dgrad / wgrad operands)
//   column sums (bias, AffineDropPath-scale and LayerNorm affine gradients)
//   LayerNorm backward (channel LN of blocks.py:160-175 and nn.LayerNorm), optionally through the fused ReLU
//   depthwise k=3 conv + mask + LayerNorm backward (q/k/v front of MaskedMHCA)
//   GELU / max-pool / masked-softmax backward
#include <cstdlib>
#include "common.cuh"

namespace vilco {

__device__ __forceinline__ float4 ld4f(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4f(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st_planes(__nv_bfloat16* p, long long lo, float v, int fmt) {
  store16_split(reinterpret_cast<uint16_t*>(p), 0, lo, v, fmt);
}

// four consecutive values -> hi plane (and lo plane) with one 8-byte store each; p must be 8-byte aligned, lo % 4 == 0
__device__ __forceinline__ void st_planes4(__nv_bfloat16* p, long long lo, float4 v, int fmt) {
  if (lo) {
    uint32_t h0, h1, l0, l1;
    split16x2(v.x, v.y, fmt, h0, l0);
    split16x2(v.z, v.w, fmt, h1, l1);
    *reinterpret_cast<uint2*>(p) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(p + lo) = make_uint2(l0, l1);
  } else {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack16x2(v.x, v.y, fmt), pack16x2(v.z, v.w, fmt));
  }
}
__device__ __forceinline__ float4 ld_planes4(const __nv_bfloat16* p, long long lo, int fmt) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack16x2(u.x, fmt), b = unpack16x2(u.y, fmt);
  float4 v = make_float4(a.x, a.y, b.x, b.y);
  if (lo) {
    const uint2 w = *reinterpret_cast<const uint2*>(p + lo);
    const float2 c = unpack16x2(w.x, fmt), d = unpack16x2(w.y, fmt);
    v.x += c.x; v.y += c.y; v.z += d.x; v.w += d.y;
  }
  return v;
}

// plain (no transpose) fp32 -> planes, 4 elements per thread; C % 4 == 0
__global__ void __launch_bounds__(256) to_planes_vec_kernel(const float* __restrict__ x, const float* __restrict__ rowmul,
                                                            const float* __restrict__ colmul, __nv_bfloat16* __restrict__ y,
                                                            long long y_lo, long long rows, int C, int fmt, float mul) {
  const int C4 = C >> 2;
  const long long n4 = rows * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = ld4f(x + 4 * i);
    if (rowmul) { const float m = rowmul[i / C4]; v.x *= m; v.y *= m; v.z *= m; v.w *= m; }
    if (colmul) { const float4 m = ld4f(colmul + 4 * (i % C4)); v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w; }
    v.x *= mul; v.y *= mul; v.z *= mul; v.w *= mul;
    st_planes4(y + 4 * i, y_lo, v, fmt);
  }
}

// y16[r, c] = x[r, c] * rowmul[r] * colmul[c]  as bf16 planes; optional transposed copy yT16[c, r]
__global__ void to_planes_kernel(const float* __restrict__ x, const float* __restrict__ rowmul, const float* __restrict__ colmul,
                                 __nv_bfloat16* __restrict__ y, long long y_lo, __nv_bfloat16* __restrict__ yT, long long yT_lo,
                                 int R, int C, int ldT, int fmt, float mul) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  x += (long long)blockIdx.z * R * C;
  if (y) y += (long long)blockIdx.z * R * C;
  if (yT) yT += (long long)blockIdx.z * C * ldT;
  if (rowmul) rowmul += (long long)blockIdx.z * R;
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      v = x[(long long)r * C + c];
      if (rowmul) v *= rowmul[r];
      if (colmul) v *= colmul[c];
      v *= mul;
      if (y) st_planes(y + (long long)r * C + c, y_lo, v, fmt);
    }
    tile[j][tx] = v;
  }
  if (!yT) return;
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (c < C && r < R) st_planes(yT + (long long)c * ldT + r, yT_lo, tile[tx][j], fmt);
  }
}

// y16[b, t, :] = x16[b, t + shift, :] (zero outside [0, T)) on bf16 planes: the row-shifted copies the k=3 conv wgrad needs
__global__ void shift_planes_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long lo16, int planes, int B, int T,
                                    int C8, int shift) {
  const long long total = (long long)planes * B * T * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % C8);
    long long r = i / C8;
    const int t = static_cast<int>(r % T); r /= T;
    const int b = static_cast<int>(r % B);
    const int pl = static_cast<int>(r / B);
    const int ts = t + shift;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ts >= 0 && ts < T) v = x[pl * lo16 + ((long long)b * T + ts) * C8 + c];
    y[pl * lo16 + ((long long)b * T + t) * C8 + c] = v;
  }
}

// out[c] += sum_r x[r, c] * (y ? y[r, c] : 1) * (rowmul ? rowmul[r] : 1)
// block = 32 column groups (float4 when C % 4 == 0) x 8 row lanes over `rows_per_block` rows; smem reduction over the row
// lanes, one atomicAdd per column and block.
template <int VEC>
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                     const float* __restrict__ rowmul, float* __restrict__ out, int R, int C,
                                                     int rows_per_block) {
  __shared__ float red[8][32 * VEC + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + tx) * VEC;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(R, r0 + rows_per_block);
  float acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
  if (c < C) {
#pragma unroll 4
    for (int r = r0 + ty; r < r1; r += 8) {
      const float m = rowmul ? rowmul[r] : 1.f;
      if (VEC == 4) {
        const float4 v = *reinterpret_cast<const float4*>(x + (long long)r * C + c);
        float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
        if (y) w = *reinterpret_cast<const float4*>(y + (long long)r * C + c);
        acc[0] += v.x * w.x * m; acc[1] += v.y * w.y * m; acc[2] += v.z * w.z * m; acc[3] += v.w * w.w * m;
      } else {
        float v = x[(long long)r * C + c];
        if (y) v *= y[(long long)r * C + c];
        acc[0] += v * m;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) red[ty][tx * VEC + i] = acc[i];
  __syncthreads();
  for (int j = threadIdx.x; j < 32 * VEC; j += 256) {
    const int cc = blockIdx.x * 32 * VEC + j;
    if (cc < C) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][j];
      atomicAdd(out + cc, t);
    }
  }
}

// ---- LayerNorm backward: y = act(LN(x [+ add]) * w + b); one warp per row; dw/db via per-block smem + atomics ----
static constexpr int BW_MAXCH = 8;
struct LnBwdParams {
  const float* x; const float* add; const float* w; const float* dy;
  const float* y_relu;   // output of the fused ReLU (fp32) or null: gradient passes where y_relu > 0
  float eps; float* dx; float* dw; float* db; int rows, C;
};

__global__ void __launch_bounds__(256, 2) layernorm_bwd_kernel(const LnBwdParams p) {
  extern __shared__ float s_acc[];  // [8 warps][2][C]: warp-private partial dw / db (each lane owns its columns: plain RMW)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nch = p.C >> 7;
  for (int i = threadIdx.x; i < 16 * p.C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const bool want_wb = p.dw != nullptr;
  float* my = s_acc + warp * 2 * p.C;
  for (int row = blockIdx.x * 8 + warp; row < p.rows; row += gridDim.x * 8) {
    const float* x = p.x + (long long)row * p.C;
    float4 v[BW_MAXCH], g[BW_MAXCH];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < BW_MAXCH; ++i)
      if (i < nch) {
        const int c = (i * 32 + lane) * 4;
        v[i] = ld4f(x + c);
        if (p.add) { const float4 a = ld4f(p.add + (long long)row * p.C + c); v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w; }
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    // issue the gradient loads before the reductions so that they overlap the shuffle chains
#pragma unroll
    for (int i = 0; i < BW_MAXCH; ++i)
      if (i < nch) {
        const int c = (i * 32 + lane) * 4;
        g[i] = ld4f(p.dy + (long long)row * p.C + c);
        if (p.y_relu) {
          const float4 yr = ld4f(p.y_relu + (long long)row * p.C + c);
          if (!(yr.x > 0.f)) g[i].x = 0.f; if (!(yr.y > 0.f)) g[i].y = 0.f; if (!(yr.z > 0.f)) g[i].z = 0.f; if (!(yr.w > 0.f)) g[i].w = 0.f;
        }
      }
    const float mean = warp_sum(s) / p.C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < BW_MAXCH; ++i)
      if (i < nch) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
      }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / p.C + p.eps);
    float s1 = 0.f, s2 = 0.f;  // sum(dy*w), sum(dy*w*xhat)
#pragma unroll
    for (int i = 0; i < BW_MAXCH; ++i)
      if (i < nch) {
        const int c = (i * 32 + lane) * 4;
        const float4 d = g[i];
        const float4 w = ld4f(p.w + c);
        v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;   // xhat
        if (want_wb) {
          float4 aw = ld4f(my + c), ab = ld4f(my + p.C + c);
          aw.x += d.x * v[i].x; aw.y += d.y * v[i].y; aw.z += d.z * v[i].z; aw.w += d.w * v[i].w;
          ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w;
          st4f(my + c, aw); st4f(my + p.C + c, ab);
        }
        g[i] = make_float4(d.x * w.x, d.y * w.y, d.z * w.z, d.w * w.w);
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
      }
    s1 = warp_sum(s1) / p.C; s2 = warp_sum(s2) / p.C;
#pragma unroll
    for (int i = 0; i < BW_MAXCH; ++i)
      if (i < nch) {
        const int c = (i * 32 + lane) * 4;
        st4f(p.dx + (long long)row * p.C + c,
             make_float4(rstd * (g[i].x - s1 - v[i].x * s2), rstd * (g[i].y - s1 - v[i].y * s2),
                         rstd * (g[i].z - s1 - v[i].z * s2), rstd * (g[i].w - s1 - v[i].w * s2)));
      }
  }
  __syncthreads();
  if (want_wb)
    for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += s_acc[w * 2 * p.C + i];
      float* dst = i < p.C ? p.dw + i : (p.db ? p.db + (i - p.C) : nullptr);
      if (dst) atomicAdd(dst, t);
    }
}

// ---- depthwise conv k=3 (stride s) * mask backward: given dconv (B, T/s, C) (gradient w.r.t. the masked conv output)
//      dx[b, t, c] = sum_tap w[tap, c] * dconv[b, to, c] * mask[b, to*s]  over (to, tap) with to*s + tap - 1 == t
//      dw[tap, c]  = sum_{b,to} dconv * mask * x[b, to*s + tap - 1, c]
__global__ void dwconv_bwd_kernel(const float* __restrict__ x, const float* __restrict__ mask, const float* __restrict__ w,
                                  const float* __restrict__ dconv, float* __restrict__ dx, float* __restrict__ dw, int B, int T,
                                  int C, int stride, int accumulate_dx) {
  // one thread per (b, t, c4); dw reduced per block over its rows then atomics
  extern __shared__ float s_dw[];  // [3][C]
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_dw[i] = 0.f;
  __syncthreads();
  const int To = T / stride;
  const int c4n = C / 4;
  const long long total = (long long)B * T * c4n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(idx % c4n) * 4;
    const long long bt = idx / c4n;
    const int t = static_cast<int>(bt % T), b = static_cast<int>(bt / T);
    float4 acc = make_float4(0, 0, 0, 0);
    // outputs `to` that read input position t: to*s + tap - 1 == t
#pragma unroll
    for (int tap = 0; tap < 3; ++tap) {
      const int num = t - tap + 1;
      if (num < 0 || num % stride != 0) continue;
      const int to = num / stride;
      if (to >= To) continue;
      const float m = mask[(long long)b * T + to * stride];
      if (m == 0.f) continue;
      const float4 d = ld4f(dconv + ((long long)b * To + to) * C + c);
      const float4 ww = ld4f(w + tap * C + c);
      acc.x += ww.x * d.x * m; acc.y += ww.y * d.y * m; acc.z += ww.z * d.z * m; acc.w += ww.w * d.w * m;
      if (dw) {
        const float4 xv = ld4f(x + ((long long)b * T + t) * C + c);
        atomicAdd(&s_dw[tap * C + c], d.x * m * xv.x); atomicAdd(&s_dw[tap * C + c + 1], d.y * m * xv.y);
        atomicAdd(&s_dw[tap * C + c + 2], d.z * m * xv.z); atomicAdd(&s_dw[tap * C + c + 3], d.w * m * xv.w);
      }
    }
    float* o = dx + ((long long)b * T + t) * C + c;
    if (accumulate_dx) { const float4 p0 = ld4f(o); acc.x += p0.x; acc.y += p0.y; acc.z += p0.z; acc.w += p0.w; }
    st4f(o, acc);
  }
  __syncthreads();
  if (dw)
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x)
      if (s_dw[i] != 0.f) atomicAdd(dw + i, s_dw[i]);
}

// fp32 depthwise k=3 conv * out-mask (recomputation of the LN input in the backward pass)
__global__ void dwconv_fwd32_kernel(const float* __restrict__ x, const float* __restrict__ mask, const float* __restrict__ w,
                                    float* __restrict__ out, int B, int T, int C, int stride) {
  const int To = T / stride, c4n = C / 4;
  const long long total = (long long)B * To * c4n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(idx % c4n) * 4;
    const long long r = idx / c4n;
    const int to = static_cast<int>(r % To), b = static_cast<int>(r / To);
    const int tc = to * stride;
    const float m = mask[(long long)b * T + tc];
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int tap = 0; tap < 3; ++tap) {
      const int t = tc + tap - 1;
      if (t < 0 || t >= T) continue;
      const float4 xv = ld4f(x + ((long long)b * T + t) * C + c);
      const float4 ww = ld4f(w + tap * C + c);
      acc.x = fmaf(ww.x, xv.x, acc.x); acc.y = fmaf(ww.y, xv.y, acc.y); acc.z = fmaf(ww.z, xv.z, acc.z); acc.w = fmaf(ww.w, xv.w, acc.w);
    }
    st4f(out + r * C + c, make_float4(acc.x * m, acc.y * m, acc.z * m, acc.w * m));
  }
}

// XLNet rel-shift backward: dBD[z, i, p] = dS[z, i, p - T + i] where that index is inside [0, T), else 0 (gather form: every
// output element is written once, fp32 and / or operand planes)
__global__ void __launch_bounds__(256) relshift_bwd_kernel(const float* __restrict__ dS, float* __restrict__ dBD,
                                                           __nv_bfloat16* __restrict__ dBD16, long long dbd_lo, long long Z, int T,
                                                           int fmt, float gs) {
  const int W4 = (2 * T) >> 2;   // T % 2 == 0 is checked by the host
  const long long total4 = Z * T * W4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total4; idx += (long long)gridDim.x * blockDim.x) {
    const int p0 = static_cast<int>(idx % W4) * 4;
    const long long zi = idx / W4;
    const int i = static_cast<int>(zi % T);
    const float* src = dS + zi * T;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = p0 + k - T + i;
      v[k] = (j >= 0 && j < T) ? src[j] : 0.f;
    }
    const float4 o = make_float4(v[0], v[1], v[2], v[3]);
    if (dBD) st4f(dBD + 4 * idx, o);
    if (dBD16) st_planes4(dBD16 + 4 * idx, dbd_lo, make_float4(o.x * gs, o.y * gs, o.z * gs, o.w * gs), fmt);   // gradient planes: scaled
  }
}

// generic fp32 elementwise helper: op 0: x*rowmul*colmul, 1: gelu(x), 2: relu(x), 3: x * (y > 0)  (ReLU backward);
// optionally also writes the result as bf16 (hi, lo) operand planes for the next GEMM
__device__ __forceinline__ float ew_apply(int op, float v, float y) {
  if (op == 1) return gelu_erf(v);
  if (op == 2) return fmaxf(v, 0.f);
  if (op == 3) return y > 0.f ? v : 0.f;
  return v;
}
template <bool VEC>
__global__ void __launch_bounds__(256) ew_kernel(int op, const float* __restrict__ x, const float* __restrict__ y,
                                                 const float* __restrict__ rowmul, const float* __restrict__ colmul,
                                                 float* __restrict__ out, __nv_bfloat16* __restrict__ out16, long long out16_lo,
                                                 long long rows, int C, int fmt) {
  if (VEC) {
    const int C4 = C >> 2;
    const long long n4 = rows * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      float4 v = ld4f(x + 4 * i);
      float4 yy = make_float4(0, 0, 0, 0);
      if (op == 3) yy = ld4f(y + 4 * i);
      if (op == 0) {
        if (rowmul) { const float m = rowmul[i / C4]; v.x *= m; v.y *= m; v.z *= m; v.w *= m; }
        if (colmul) { const float4 m = ld4f(colmul + 4 * (i % C4)); v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w; }
      } else {
        v.x = ew_apply(op, v.x, yy.x); v.y = ew_apply(op, v.y, yy.y); v.z = ew_apply(op, v.z, yy.z); v.w = ew_apply(op, v.w, yy.w);
      }
      if (out) st4f(out + 4 * i, v);
      if (out16) st_planes4(out16 + 4 * i, out16_lo, v, fmt);
    }
  } else {
    const long long n = rows * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      float v = x[i];
      if (op == 0) {
        if (rowmul) v *= rowmul[i / C];
        if (colmul) v *= colmul[i % C];
      } else v = ew_apply(op, v, op == 3 ? y[i] : 0.f);
      if (out) out[i] = v;
      if (out16) st_planes(out16 + i, out16_lo, v, fmt);
    }
  }
}

// inverted dropout with a counter-based generator: element i of call `seed` is kept iff mix(seed, i) >= p * 2^32.  The backward
// pass applies the same kernel to the gradient with the same seed, so no mask is stored.
__device__ __forceinline__ unsigned int mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return static_cast<unsigned int>((z ^ (z >> 31)) >> 32);
}
__device__ __forceinline__ float keep_factor(unsigned long long seed, unsigned long long i, unsigned int thr, float inv_keep) {
  return mix64(seed * 0xD1342543DE82EF95ull + i) >= thr ? inv_keep : 0.f;
}
__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                      __nv_bfloat16* __restrict__ out16, long long out16_lo, long long n,
                                                      unsigned int thr, float inv_keep, unsigned long long seed, int vec, int fmt) {
  if (vec) {
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      float4 v = ld4f(x + 4 * i);
      v.x *= keep_factor(seed, 4 * i, thr, inv_keep); v.y *= keep_factor(seed, 4 * i + 1, thr, inv_keep);
      v.z *= keep_factor(seed, 4 * i + 2, thr, inv_keep); v.w *= keep_factor(seed, 4 * i + 3, thr, inv_keep);
      if (out) st4f(out + 4 * i, v);
      if (out16) st_planes4(out16 + 4 * i, out16_lo, v, fmt);
    }
    return;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i] * keep_factor(seed, i, thr, inv_keep);
    if (out) out[i] = v;
    if (out16) st_planes(out16 + i, out16_lo, v, fmt);
  }
}

// ---- residual branch of a TransformerBlock in training mode, fused (blocks.py:404-405, 567-585, 640-670):
//   out[r,c] = resid[r,c] * rm[r] + scale[c] * (y[r,c] + bias[c]) * ymul[r] * keep(seed, r*C+c)
// y = raw GEMM output of the projection, bias its bias (added here so that dropout sees proj(x)+b like the reference),
// ymul[r] = out-mask[r] * stochastic-depth factor of the row's sample, keep = inverted dropout factor (1 when p == 0).
__global__ void __launch_bounds__(256) resid_branch_fwd_kernel(const float* __restrict__ resid, const float* __restrict__ rm,
                                                               const float* __restrict__ y, const float* __restrict__ bias,
                                                               const float* __restrict__ scale, const float* __restrict__ ymul,
                                                               float* __restrict__ out, long long rows, int C, unsigned int thr,
                                                               float inv_keep, unsigned long long seed) {
  const int C4 = C >> 2;   // C % 4 == 0 (host check)
  const long long n4 = rows * C4;
  for (long long i4 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i4 < n4; i4 += (long long)gridDim.x * blockDim.x) {
    const long long r = i4 / C4;
    const int c = static_cast<int>(i4 - r * C4) * 4;
    float4 v = ld4f(y + 4 * i4);
    if (bias) { const float4 b = ld4f(bias + c); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
    if (thr) {
      const unsigned long long i = 4ull * static_cast<unsigned long long>(i4);
      v.x *= keep_factor(seed, i, thr, inv_keep); v.y *= keep_factor(seed, i + 1, thr, inv_keep);
      v.z *= keep_factor(seed, i + 2, thr, inv_keep); v.w *= keep_factor(seed, i + 3, thr, inv_keep);
    }
    const float ym = ymul ? ymul[r] : 1.f;
    float4 sc = make_float4(ym, ym, ym, ym);
    if (scale) { const float4 t = ld4f(scale + c); sc.x *= t.x; sc.y *= t.y; sc.z *= t.z; sc.w *= t.w; }
    const float4 a = ld4f(resid + 4 * i4);
    const float m = rm ? rm[r] : 1.f;
    st4f(out + 4 * i4, make_float4(a.x * m + v.x * sc.x, a.y * m + v.y * sc.y, a.z * m + v.z * sc.z, a.w * m + v.w * sc.w));
  }
}

// backward of the above for one (128-column, row-slab) block: d resid (optional), dy as operand planes (the dZ of the
// projection's weight / data gradient GEMMs), and the column sums dbias[c] += sum_r dy, dscale[c] += sum_r t * (y + bias).
__global__ void __launch_bounds__(256) resid_branch_bwd_kernel(const float* __restrict__ g, const float* __restrict__ rm,
                                                               const float* __restrict__ y, const float* __restrict__ bias,
                                                               const float* __restrict__ scale, const float* __restrict__ ymul,
                                                               float* __restrict__ dresid, __nv_bfloat16* __restrict__ dy16,
                                                               long long dy_lo, float* __restrict__ dbias, float* __restrict__ dscale,
                                                               int R, int C, int rows_per_block, unsigned int thr, float inv_keep,
                                                               unsigned long long seed, int fmt, float gs) {
  __shared__ float red[2][8][129];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + tx) * 4;   // C % 4 == 0 (host check)
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(R, r0 + rows_per_block);
  float4 sb = make_float4(0, 0, 0, 0), ss = make_float4(0, 0, 0, 0);
  if (c < C) {
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), bc = make_float4(0, 0, 0, 0);
    if (scale) sc = ld4f(scale + c);
    if (bias) bc = ld4f(bias + c);
    for (int r = r0 + ty; r < r1; r += 8) {
      const long long i = (long long)r * C + c;
      const float4 gi = ld4f(g + i);
      if (dresid) { const float m = rm ? rm[r] : 1.f; st4f(dresid + i, make_float4(gi.x * m, gi.y * m, gi.z * m, gi.w * m)); }
      const float ym = ymul ? ymul[r] : 1.f;
      float4 t = make_float4(gi.x * ym, gi.y * ym, gi.z * ym, gi.w * ym);
      if (thr) {
        t.x *= keep_factor(seed, i, thr, inv_keep); t.y *= keep_factor(seed, i + 1, thr, inv_keep);
        t.z *= keep_factor(seed, i + 2, thr, inv_keep); t.w *= keep_factor(seed, i + 3, thr, inv_keep);
      }
      const float4 dyi = make_float4(t.x * sc.x, t.y * sc.y, t.z * sc.z, t.w * sc.w);
      st_planes4(dy16 + i, dy_lo, make_float4(dyi.x * gs, dyi.y * gs, dyi.z * gs, dyi.w * gs), fmt);   // gradient planes: scaled
      sb.x += dyi.x; sb.y += dyi.y; sb.z += dyi.z; sb.w += dyi.w;
      const float4 yi = ld4f(y + i);
      ss.x += t.x * (yi.x + bc.x); ss.y += t.y * (yi.y + bc.y); ss.z += t.z * (yi.z + bc.z); ss.w += t.w * (yi.w + bc.w);
    }
  }
  red[0][ty][tx * 4] = sb.x; red[0][ty][tx * 4 + 1] = sb.y; red[0][ty][tx * 4 + 2] = sb.z; red[0][ty][tx * 4 + 3] = sb.w;
  red[1][ty][tx * 4] = ss.x; red[1][ty][tx * 4 + 1] = ss.y; red[1][ty][tx * 4 + 2] = ss.z; red[1][ty][tx * 4 + 3] = ss.w;
  __syncthreads();
  {
    const int which = threadIdx.x >> 7, j = threadIdx.x & 127;   // 128 threads per sum
    const int cc = blockIdx.x * 128 + j;
    float* dst = which == 0 ? dbias : dscale;
    if (cc < C && dst) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[which][k][j];
      atomicAdd(dst + cc, t);
    }
  }
}

// dx = dy * gelu'(x)
__device__ __forceinline__ float gelu_grad(float v) {
  const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
  return cdf + v * pdf;
}
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                       float* __restrict__ dx, long long n) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ld4f(x + 4 * i), d = ld4f(dy + 4 * i);
    st4f(dx + 4 * i, make_float4(d.x * gelu_grad(v.x), d.y * gelu_grad(v.y), d.z * gelu_grad(v.z), d.w * gelu_grad(v.w)));
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = dy[i] * gelu_grad(x[i]);
}

// MaxPool1d(3, 2, 1) backward on token-major fp32: the gradient goes to the first maximum of the window (ATen semantics)
__global__ void maxpool3s2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int B,
                                      int T, int C) {
  const int To = T / 2;
  const long long total = (long long)B * To * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long r = i / C;
    const int to = static_cast<int>(r % To), b = static_cast<int>(r / To);
    const float* xb = x + ((long long)b * T) * C + c;
    int best = -1;
    float bv = -INFINITY;
    for (int tap = 0; tap < 3; ++tap) {
      const int t = 2 * to + tap - 1;
      if (t < 0 || t >= T) continue;
      const float v = xb[(long long)t * C];
      if (best < 0 || v > bv) { bv = v; best = t; }
    }
    atomicAdd(dx + ((long long)b * T + best) * C + c, dy[i]);
  }
}

// masked softmax backward over rows: dS[r, j] = scale * P[r, j] * (dP[r, j] - sum_k dP[r, k] P[r, k]).  P comes from the fp32
// copy when given, else from the (hi, lo) operand planes the forward softmax wrote (row stride p_ld); dS goes out as fp32
// and / or as operand planes (row stride ds_ld) for the dQ / dK GEMMs.
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const float* __restrict__ P32, const __nv_bfloat16* __restrict__ P16,
                                                          long long p_lo, long long p_ld, const float* __restrict__ dP,
                                                          float* __restrict__ dS, __nv_bfloat16* __restrict__ dS16, long long ds_lo,
                                                          long long ds_ld, long long rows, int Tk, float scale_in, int pfmt, float gs) {
  // P16 (forward probabilities) and the gradient planes dS16 share the plane format pfmt; dS16 is stored times gs
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* d = dP + row * Tk;
  const float* p32 = P32 ? P32 + row * Tk : nullptr;
  const __nv_bfloat16* p16 = P16 ? P16 + row * p_ld : nullptr;
  auto prob = [&](int j) -> float {
    if (p32) return p32[j];
    return load16_split(reinterpret_cast<const uint16_t*>(p16) + j, p_lo, pfmt);
  };
  const bool vec = !p32 && (Tk % 4 == 0) && (p_ld % 4 == 0) && (p_lo % 4 == 0) && (ds_ld % 4 == 0) && (ds_lo % 4 == 0) && dS16 && !dS;
  float s = 0.f;
  if (vec) {   // planes in, planes out: 4 elements per lane and iteration
    for (int j = lane * 4; j < Tk; j += 128) {
      const float4 pv = ld_planes4(p16 + j, p_lo, pfmt), dv = ld4f(d + j);
      s += (pv.x * dv.x + pv.y * dv.y) + (pv.z * dv.z + pv.w * dv.w);
    }
    s = warp_sum(s);
    for (int j = lane * 4; j < Tk; j += 128) {
      const float4 pv = ld_planes4(p16 + j, p_lo, pfmt), dv = ld4f(d + j);
      const float scale = scale_in * gs;
      st_planes4(dS16 + row * ds_ld + j, ds_lo, make_float4(scale * pv.x * (dv.x - s), scale * pv.y * (dv.y - s),
                                                            scale * pv.z * (dv.z - s), scale * pv.w * (dv.w - s)),
                 pfmt);
    }
    return;
  }
  for (int j = lane; j < Tk; j += 32) s += prob(j) * d[j];
  s = warp_sum(s);
  for (int j = lane; j < Tk; j += 32) {
    const float v = scale_in * prob(j) * (d[j] - s);
    if (dS) dS[row * Tk + j] = v;
    if (dS16) st_planes(dS16 + row * ds_ld + j, ds_lo, v * gs, pfmt);
  }
}

}  // namespace vilco

using namespace vilco;

static inline int bgrid(long long n, int block, int cap = 148 * 8) {
  long long g = (n + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

extern "C" int vilco_to_planes(const float* x, const float* rowmul, const float* colmul, void* y, int64_t y_lo, void* yT,
                               int64_t yT_lo, int R, int C, int ldT, int Z, int grad, void* stream) {
  VILCO_CHECK_ARG(x && (y || yT) && R > 0 && C > 0 && Z > 0, "vilco_to_planes: bad arguments");
  const int fmt = act_fmt();
  const float mul = grad ? grad_scale() : 1.0f;
  if (!yT && C % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 8 == 0 && y_lo % 4 == 0 &&
      (!colmul || reinterpret_cast<uintptr_t>(colmul) % 16 == 0)) {
    // Z independent matrices are contiguous: one flat pass (rowmul, when given, indexes z * R + r as well)
    const long long rows = (long long)Z * R;
    long long g = (rows * (C / 4) + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    to_planes_vec_kernel<<<static_cast<unsigned>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, rowmul, colmul, static_cast<__nv_bfloat16*>(y), y_lo, rows, C, fmt, mul);
    VILCO_LAUNCH_CHECK();
    return VILCO_OK;
  }
  dim3 grid((C + 31) / 32, (R + 31) / 32, Z);
  to_planes_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      x, rowmul, colmul, static_cast<__nv_bfloat16*>(y), y_lo, static_cast<__nv_bfloat16*>(yT), yT_lo, R, C, ldT, fmt, mul);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_colsum(const float* x, const float* y, const float* rowmul, float* out, int R, int C, void* stream) {
  VILCO_CHECK_ARG(x && out && R > 0 && C > 0, "vilco_colsum: bad arguments");
  const bool vec = (C % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (!y || reinterpret_cast<uintptr_t>(y) % 16 == 0);
  const int cols_per_block = vec ? 128 : 32;
  const int cblocks = (C + cols_per_block - 1) / cols_per_block;
  // enough row slabs to fill the GPU a few times over, at least 64 rows per slab
  int slabs = (148 * 4 + cblocks - 1) / cblocks;
  int rpb = (R + slabs - 1) / slabs;
  if (rpb < 64) rpb = 64;
  rpb = (rpb + 7) / 8 * 8;
  dim3 grid(cblocks, (R + rpb - 1) / rpb);
  if (vec) colsum_kernel<4><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, rowmul, out, R, C, rpb);
  else colsum_kernel<1><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, rowmul, out, R, C, rpb);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_layernorm_bwd(const float* x, const float* add, const float* w, const float* dy, const float* y_relu,
                                   float eps, float* dx, float* dw, float* db, int rows, int C, void* stream) {
  VILCO_CHECK_ARG(x && w && dy && dx, "vilco_layernorm_bwd: null pointer");
  VILCO_CHECK_ARG(C % 128 == 0 && C <= 128 * BW_MAXCH, "vilco_layernorm_bwd: C=%d unsupported", C);
  LnBwdParams p{x, add, w, dy, y_relu, eps, dx, dw, db, rows, C};
  int grid = (rows + 7) / 8;
  static int cap = [] { const char* e = getenv("VILCO_LNB_GRID"); return e ? atoi(e) : 148 * 2; }();
  if (grid > cap) grid = cap;   // 2 resident CTAs per SM (64 KB of warp-private accumulators each)
  const int smem = 16 * C * static_cast<int>(sizeof(float));
  static int configured = 0;
  if (smem > configured) {
    VILCO_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  layernorm_bwd_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(p);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_dwconv_bwd(const float* x, const float* mask, const float* w, const float* dconv, float* dx, float* dw,
                                int B, int T, int C, int stride, int accumulate_dx, void* stream) {
  VILCO_CHECK_ARG(x && mask && w && dconv && dx, "vilco_dwconv_bwd: null pointer");
  VILCO_CHECK_ARG(C % 4 == 0 && (stride == 1 || stride == 2) && T % stride == 0, "vilco_dwconv_bwd: bad shape");
  dwconv_bwd_kernel<<<bgrid((long long)B * T * (C / 4), 256, 148 * 4), 256, 3 * C * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      x, mask, w, dconv, dx, dw, B, T, C, stride, accumulate_dx);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_gelu_bwd(const float* x, const float* dy, float* dx, int64_t n, void* stream) {
  VILCO_CHECK_ARG(x && dy && dx && n > 0, "vilco_gelu_bwd: bad arguments");
  VILCO_CHECK_ARG(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(dy) % 16 == 0 &&
                      reinterpret_cast<uintptr_t>(dx) % 16 == 0, "vilco_gelu_bwd: buffers must be 16-byte aligned");
  gelu_bwd_kernel<<<bgrid((n + 3) / 4, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, dy, dx, n);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_maxpool3s2_bwd(const float* x, const float* dy, float* dx, int B, int T, int C, void* stream) {
  VILCO_CHECK_ARG(x && dy && dx && T % 2 == 0, "vilco_maxpool3s2_bwd: bad arguments");
  maxpool3s2_bwd_kernel<<<bgrid((long long)B * (T / 2) * C, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, dy, dx, B, T, C);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_softmax_bwd(const float* P32, const void* P16, int64_t p_lo, int64_t p_ld, const float* dP, float* dS,
                                 void* dS16, int64_t ds_lo, int64_t ds_ld, int64_t rows, int Tk, float scale, void* stream) {
  VILCO_CHECK_ARG((P32 || P16) && dP && (dS || dS16) && rows > 0 && Tk > 0, "vilco_softmax_bwd: bad arguments");
  softmax_bwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      P32, static_cast<const __nv_bfloat16*>(P16), p_lo, p_ld, dP, dS, static_cast<__nv_bfloat16*>(dS16), ds_lo, ds_ld, rows, Tk, scale,
      act_fmt(), grad_scale());
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_shift_planes(const void* x, void* y, int64_t lo, int B, int T, int C, int shift, void* stream) {
  VILCO_CHECK_ARG(x && y && C % 8 == 0 && lo % 8 == 0, "vilco_shift_planes: bad arguments");
  const int planes = lo ? 2 : 1;
  shift_planes_kernel<<<bgrid((long long)planes * B * T * (C / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), lo / 8, planes, B, T, C / 8, shift);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_dwconv_fwd32(const float* x, const float* mask, const float* w, float* out, int B, int T, int C, int stride,
                                  void* stream) {
  VILCO_CHECK_ARG(x && mask && w && out && C % 4 == 0 && (stride == 1 || stride == 2) && T % stride == 0, "vilco_dwconv_fwd32: bad arguments");
  dwconv_fwd32_kernel<<<bgrid((long long)B * (T / stride) * (C / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, mask, w, out, B, T, C, stride);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_relshift_bwd(const float* dS, float* dBD, void* dBD16, int64_t dbd_lo, int64_t Z, int T, void* stream) {
  VILCO_CHECK_ARG(dS && (dBD || dBD16) && Z > 0 && T > 0 && T % 2 == 0 && dbd_lo % 4 == 0, "vilco_relshift_bwd: bad arguments");
  relshift_bwd_kernel<<<bgrid(Z * T * (T / 2LL), 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dS, dBD, static_cast<__nv_bfloat16*>(dBD16), dbd_lo, Z, T, act_fmt(), grad_scale());
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_dropout(const float* x, float* out, void* out16, int64_t out16_lo, int64_t n, float p, uint64_t seed, void* stream) {
  VILCO_CHECK_ARG(x && (out || out16) && n > 0 && p >= 0.f && p < 1.f, "vilco_dropout: bad arguments");
  const unsigned int thr = static_cast<unsigned int>(static_cast<double>(p) * 4294967296.0);
  auto al = [](const void* q, int a) { return !q || reinterpret_cast<uintptr_t>(q) % a == 0; };
  const int vec = (n % 4 == 0 && al(x, 16) && al(out, 16) && al(out16, 8) && out16_lo % 4 == 0) ? 1 : 0;
  dropout_kernel<<<bgrid(vec ? n / 4 : n, 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, out, static_cast<__nv_bfloat16*>(out16), out16_lo, n, thr, 1.0f / (1.0f - p), seed, vec, act_fmt());
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

int vilco_ew(int op, const float* x, const float* y, const float* rowmul, const float* colmul, float* out, void* out16,
             int64_t out16_lo, int64_t rows, int C, void* stream) {
  VILCO_CHECK_ARG(x && (out || out16) && rows > 0 && C > 0 && op >= 0 && op <= 3 && (op != 3 || y), "vilco_ew: bad arguments");
  auto al = [](const void* q, int a) { return !q || reinterpret_cast<uintptr_t>(q) % a == 0; };
  const bool vec = C % 4 == 0 && al(x, 16) && al(y, 16) && al(colmul, 16) && al(out, 16) && al(out16, 8) && out16_lo % 4 == 0;
  if (vec)
    ew_kernel<true><<<bgrid(rows * (C / 4), 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        op, x, y, rowmul, colmul, out, static_cast<__nv_bfloat16*>(out16), out16_lo, rows, C, act_fmt());
  else
    ew_kernel<false><<<bgrid(rows * C, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        op, x, y, rowmul, colmul, out, static_cast<__nv_bfloat16*>(out16), out16_lo, rows, C, act_fmt());
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_resid_branch_fwd(const float* resid, const float* rm, const float* y, const float* bias, const float* scale,
                                      const float* ymul, float* out, int64_t rows, int C, float p, uint64_t seed, void* stream) {
  VILCO_CHECK_ARG(resid && y && out && rows > 0 && C > 0 && C % 4 == 0 && p >= 0.f && p < 1.f, "vilco_resid_branch_fwd: bad arguments");
  const unsigned int thr = static_cast<unsigned int>(static_cast<double>(p) * 4294967296.0);
  resid_branch_fwd_kernel<<<bgrid(rows * (C / 4), 256, 148 * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(resid, rm, y, bias, scale, ymul, out, rows,
                                                                                        C, thr, 1.0f / (1.0f - p), seed);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_resid_branch_bwd(const float* g, const float* rm, const float* y, const float* bias, const float* scale,
                                      const float* ymul, float* dresid, void* dy16, int64_t dy_lo, float* dbias, float* dscale,
                                      int R, int C, float p, uint64_t seed, void* stream) {
  VILCO_CHECK_ARG(g && y && dy16 && R > 0 && C > 0 && C % 4 == 0 && dy_lo % 4 == 0 && p >= 0.f && p < 1.f,
                  "vilco_resid_branch_bwd: bad arguments");
  const unsigned int thr = static_cast<unsigned int>(static_cast<double>(p) * 4294967296.0);
  const int cblocks = (C + 127) / 128;
  int slabs = (148 * 8 + cblocks - 1) / cblocks;
  int rpb = (R + slabs - 1) / slabs;
  if (rpb < 32) rpb = 32;
  rpb = (rpb + 7) / 8 * 8;
  dim3 grid(cblocks, (R + rpb - 1) / rpb);
  resid_branch_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(g, rm, y, bias, scale, ymul, dresid,
                                                                             static_cast<__nv_bfloat16*>(dy16), dy_lo, dbias, dscale, R,
                                                                             C, rpb, thr, 1.0f / (1.0f - p), seed, act_fmt(),
                                                                             grad_scale());
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}
