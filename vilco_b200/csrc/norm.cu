// Bandwidth-bound token-major kernels: channel LayerNorm (+ fused add / ReLU / positional encoding),
// depthwise k=3 conv + mask + LayerNorm (the q/k/v front of MaskedMHCA), max-pool skip, layout packing.
// One warp owns one token row (C contiguous): 128-bit coalesced loads, warp-shuffle statistics.
#include "common.cuh"

namespace vilco {

static constexpr int MAXCH = 8;  // float4 chunks per lane: C <= 8*128 = 1024

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {   // bf16 INPUT rows (x_dtype = VILCO_BF16), not operand planes
  uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}
// operand-plane store (format fmt) with optional lo plane (split precision: x ~= hi + lo)
__device__ __forceinline__ void st4s(__nv_bfloat16* p, long long lo, float4 v, int fmt) {
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

// two-pass statistics exactly like the reference (mean, then mean of squared residuals)
__device__ __forceinline__ void row_stats(const float4 (&v)[MAXCH], int nch, int C, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i)
    if (i < nch) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  mean = warp_sum(s) / static_cast<float>(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i)
    if (i < nch) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  rstd = 1.0f / sqrtf(warp_sum(q) / static_cast<float>(C) + eps);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the channel dim.  y = act(LN(x [+ add]) * w + b) [+ pe[t] * rowmul[row]], zeroed where zero_rows.
// ---------------------------------------------------------------------------------------------
struct LnParams {
  const void* x; int x_bf16;
  const float* add;
  const float* w; const float* b;
  float eps; int relu;
  const float* pe; int pe_T;          // (pe_T, C) table indexed by row % rows_per_batch
  const float* rowmul;                // per-row multiplier for the pe term
  const uint8_t* zero_rows;
  float* y32; __nv_bfloat16* y16; long long y16_lo;
  long long y_ld, y_bs;               // output row stride / batch stride (elements)
  int rows, rows_per_batch, C;
  int fmt;                            // element format of y16
};

template <typename TIn>
__global__ void __launch_bounds__(256) layernorm_kernel(const LnParams p) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const int nch = p.C >> 7;
  const TIn* x = static_cast<const TIn*>(p.x) + (long long)row * p.C;
  float4 v[MAXCH];
#pragma unroll
  for (int i = 0; i < MAXCH; ++i)
    if (i < nch) {
      const int c = (i * 32 + lane) * 4;
      v[i] = ld4(x + c);
      if (p.add) {
        const float4 a = ld4(p.add + (long long)row * p.C + c);
        v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w;
      }
    }
  float mean, rstd;
  row_stats(v, nch, p.C, p.eps, mean, rstd);
  const int bi = row / p.rows_per_batch, t = row - bi * p.rows_per_batch;
  const long long yoff = (long long)bi * p.y_bs + (long long)t * p.y_ld;
  const bool zero = p.zero_rows && p.zero_rows[row];
  const float rm = (p.pe && p.rowmul) ? p.rowmul[row] : 1.0f;
#pragma unroll
  for (int i = 0; i < MAXCH; ++i)
    if (i < nch) {
      const int c = (i * 32 + lane) * 4;
      const float4 w = ld4(p.w + c), b = ld4(p.b + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * w.x + b.x;
      o.y = (v[i].y - mean) * rstd * w.y + b.y;
      o.z = (v[i].z - mean) * rstd * w.z + b.z;
      o.w = (v[i].w - mean) * rstd * w.w + b.w;
      if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      if (p.pe) {
        const float4 e = ld4(p.pe + (long long)t * p.C + c);
        o.x += e.x * rm; o.y += e.y * rm; o.z += e.z * rm; o.w += e.w * rm;
      }
      if (zero) o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.y32) st4(p.y32 + yoff + c, o);
      if (p.y16) st4s(p.y16 + yoff + c, p.y16_lo, o, p.fmt);
    }
}

// ---------------------------------------------------------------------------------------------
// depthwise conv (k=3, stride s, zero pad 1) * out_mask -> LayerNorm, for q/k/v at once (blockIdx.y picks one)
// ---------------------------------------------------------------------------------------------
struct DwParams {
  const void* x; int x_bf16;          // (B, T, C)
  const float* mask;                  // (B, T) 1/0, input resolution
  const int* tlen;                    // optional (B,): rows t >= tlen[b] are treated as outside the sequence (zero)
  const float* wconv[3];              // (3 taps, C) each: tap-major
  const float* lnw[3]; const float* lnb[3];
  __nv_bfloat16* out[3];              // (B, T/s, C)
  long long out_lo;
  int B, T, C, stride; float eps;
  int fmt;
};

template <typename TIn>
__global__ void __launch_bounds__(256) dwconv_ln_kernel(const DwParams p) {
  const int lane = threadIdx.x & 31;
  const int To = p.T / p.stride;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.B * To) return;
  const int which = blockIdx.y;
  const int bi = row / To, to = row - bi * To;
  const int tc = to * p.stride;
  const int nch = p.C >> 7;
  const TIn* xb = static_cast<const TIn*>(p.x) + (long long)bi * p.T * p.C;
  const float m = p.mask[(long long)bi * p.T + tc];
  const float* wc = p.wconv[which];
  const int Tb = p.tlen ? min(p.T, p.tlen[bi]) : p.T;  // effective sequence end for the zero padding
  float4 v[MAXCH];
#pragma unroll
  for (int i = 0; i < MAXCH; ++i)
    if (i < nch) {
      const int c = (i * 32 + lane) * 4;
      const float4 w1 = ld4(wc + p.C + c);
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tc < Tb) {
        const float4 x1 = ld4(xb + (long long)tc * p.C + c);
        a = make_float4(w1.x * x1.x, w1.y * x1.y, w1.z * x1.z, w1.w * x1.w);
      }
      if (tc > 0 && tc - 1 < Tb) {
        const float4 w0 = ld4(wc + c);
        const float4 x0 = ld4(xb + (long long)(tc - 1) * p.C + c);
        a.x = fmaf(w0.x, x0.x, a.x); a.y = fmaf(w0.y, x0.y, a.y); a.z = fmaf(w0.z, x0.z, a.z); a.w = fmaf(w0.w, x0.w, a.w);
      }
      if (tc + 1 < Tb) {
        const float4 w2 = ld4(wc + 2 * p.C + c);
        const float4 x2 = ld4(xb + (long long)(tc + 1) * p.C + c);
        a.x = fmaf(w2.x, x2.x, a.x); a.y = fmaf(w2.y, x2.y, a.y); a.z = fmaf(w2.z, x2.z, a.z); a.w = fmaf(w2.w, x2.w, a.w);
      }
      v[i] = make_float4(a.x * m, a.y * m, a.z * m, a.w * m);
    }
  float mean, rstd;
  row_stats(v, nch, p.C, p.eps, mean, rstd);
  __nv_bfloat16* o = p.out[which] + (long long)row * p.C;
  const float* lw = p.lnw[which];
  const float* lb = p.lnb[which];
#pragma unroll
  for (int i = 0; i < MAXCH; ++i)
    if (i < nch) {
      const int c = (i * 32 + lane) * 4;
      const float4 w = ld4(lw + c), b = ld4(lb + c);
      st4s(o + c, p.out_lo, make_float4((v[i].x - mean) * rstd * w.x + b.x, (v[i].y - mean) * rstd * w.y + b.y,
                                        (v[i].z - mean) * rstd * w.z + b.z, (v[i].w - mean) * rstd * w.w + b.w), p.fmt);
    }
}

// ---------------------------------------------------------------------------------------------
// MaxPool1d(kernel 3, stride 2, pad 1) over time on token-major fp32 (implicit -inf padding)
// ---------------------------------------------------------------------------------------------
__global__ void maxpool3s2_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int T, int C) {
  const int To = T / 2;
  const long long n4 = (long long)B * To * (C / 4);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % (C / 4)) * 4;
    const long long r = i / (C / 4);
    const int to = static_cast<int>(r % To);
    const int bi = static_cast<int>(r / To);
    const float* xb = x + ((long long)bi * T) * C + c;
    const int tc = 2 * to;
    float4 m = ld4(xb + (long long)tc * C);
    if (tc > 0) {
      const float4 a = ld4(xb + (long long)(tc - 1) * C);
      m.x = fmaxf(m.x, a.x); m.y = fmaxf(m.y, a.y); m.z = fmaxf(m.z, a.z); m.w = fmaxf(m.w, a.w);
    }
    if (tc + 1 < T) {
      const float4 a = ld4(xb + (long long)(tc + 1) * C);
      m.x = fmaxf(m.x, a.x); m.y = fmaxf(m.y, a.y); m.z = fmaxf(m.z, a.z); m.w = fmaxf(m.w, a.w);
    }
    st4(y + r * C + c, m);
  }
}

// out = a*x + b*y (fp32), optional bf16 copy
__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float a, float b,
                             float* __restrict__ o32, __nv_bfloat16* __restrict__ o16, long long o16_lo, long long n4, int fmt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 u = ld4(x + 4 * i);
    float4 o = make_float4(a * u.x, a * u.y, a * u.z, a * u.w);
    if (y) {
      const float4 w = ld4(y + 4 * i);
      o.x = fmaf(b, w.x, o.x); o.y = fmaf(b, w.y, o.y); o.z = fmaf(b, w.z, o.z); o.w = fmaf(b, w.w, o.w);
    }
    if (o32) st4(o32 + 4 * i, o);
    if (o16) st4s(o16 + 4 * i, o16_lo, o, fmt);
  }
}

// out[r, c] = x[r, c] * rowmul[r] + scale[c] * y[r, c]   (skip * mask + AffineDropPath-scale * adapter branch)
__global__ void scale_add_kernel(const float* __restrict__ x, const float* __restrict__ rowmul, const float* __restrict__ y,
                                 const float* __restrict__ scale, float* __restrict__ out, long long rows, int C) {
  const long long n4 = rows * (C / 4);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (C / 4);
    const int c = static_cast<int>(i % (C / 4)) * 4;
    const float rm = rowmul ? rowmul[r] : 1.0f;
    const float4 a = ld4(x + 4 * i), b = ld4(y + 4 * i);
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
    if (scale) sc = ld4(scale + c);
    st4(out + 4 * i, make_float4(a.x * rm + sc.x * b.x, a.y * rm + sc.y * b.y, a.z * rm + sc.z * b.z, a.w * rm + sc.w * b.w));
  }
}

// (B, C, T) fp32 channel-major (the reference layout) -> (B, T, C) bf16 token-major, zero padded to T_out rows
__global__ void pack_feats_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long y_lo, int C, int T,
                                  int T_out, int fmt) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, t = t0 + tx;
    tile[j][tx] = (c < C && t < T) ? x[((long long)b * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int t = t0 + j, c = c0 + tx;
    if (t < T_out && c < C) {
      const long long o = ((long long)b * T_out + t) * C + c;
      store16_split(reinterpret_cast<uint16_t*>(y), o, y_lo, tile[tx][j], fmt);
    }
  }
}

// Data path in front of the model: raw clip features as stored on disk — (T_in, C) fp32 rows, clip b owns
// rows row_start[b] .. row_start[b+1]-1 of x — resized along time to T_out rows like
// F.interpolate(feats.permute(1,0)[None], size=max_seq_len, mode='linear', align_corners=False) (MQ/libs/datasets/ego4d.py:644-651;
// ATen's area_pixel_compute_source_index / guard_index_and_lambda in float), written token-major as fp32 and/or bf16 operand
// planes.  One thread per 4 channels of an output row: both source rows are read with 16-byte loads that are contiguous
// across the warp, every input byte is needed by at most ceil(T_out/T_in)+1 neighbouring output rows (L2 hits).
__global__ void resize_feats_kernel(const float* __restrict__ x, const long long* __restrict__ row_start, float* __restrict__ o32,
                                    __nv_bfloat16* __restrict__ o16, long long o16_lo, int C4, int T_out, int fmt) {
  const int b = blockIdx.y;
  const long long r0 = row_start[b];
  const int T_in = static_cast<int>(row_start[b + 1] - r0);
  if (T_in <= 0) return;
  const float scale = static_cast<float>(T_in) / static_cast<float>(T_out);
  const long long n = static_cast<long long>(T_out) * C4;
  const long long C = 4ll * C4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / C4);
    const long long c = (i - static_cast<long long>(t) * C4) * 4;
    float src = __fadd_rn(__fmul_rn(scale, static_cast<float>(t) + 0.5f), -0.5f);
    if (src < 0.f) src = 0.f;
    int i0 = static_cast<int>(src);
    if (i0 > T_in - 1) i0 = T_in - 1;
    const int i1 = i0 + (i0 < T_in - 1 ? 1 : 0);
    float l1 = src - static_cast<float>(i0);
    l1 = fminf(fmaxf(l1, 0.f), 1.f);
    const float l0 = 1.f - l1;
    const float4 a = ld4(x + (r0 + i0) * C + c), d = ld4(x + (r0 + i1) * C + c);
    // separate multiplies and add (no FMA contraction), like the scalar expression w0*x0 + w1*x1 of the CPU kernel
    const float4 v = make_float4(__fadd_rn(__fmul_rn(l0, a.x), __fmul_rn(l1, d.x)), __fadd_rn(__fmul_rn(l0, a.y), __fmul_rn(l1, d.y)),
                                 __fadd_rn(__fmul_rn(l0, a.z), __fmul_rn(l1, d.z)), __fadd_rn(__fmul_rn(l0, a.w), __fmul_rn(l1, d.w)));
    const long long o = (static_cast<long long>(b) * T_out + t) * C + c;
    if (o32) st4(o32 + o, v);
    if (o16) st4s(o16 + o, o16_lo, v, fmt);
  }
}

// (B, T, C) fp32 token-major -> (B, C, T) fp32 channel-major (outputs handed back in the reference layout)
__global__ void unpack_kernel(const float* __restrict__ x, float* __restrict__ y, int T, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int j = ty; j < 32; j += 8) {
    const int t = t0 + j, c = c0 + tx;
    tile[j][tx] = (t < T && c < C) ? x[((long long)b * T + t) * C + c] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, t = t0 + tx;
    if (c < C && t < T) y[((long long)b * C + c) * T + t] = tile[tx][j];
  }
}

}  // namespace vilco

using namespace vilco;

static inline int grid_for(long long n, int block, int cap = 148 * 8) {
  long long g = (n + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

extern "C" int vilco_layernorm(const void* x, int x_dtype, const float* add, const float* w, const float* b, float eps,
                               int relu, const float* pe, int pe_T, const float* rowmul, const uint8_t* zero_rows,
                               float* y32, void* y16, int64_t y16_lo, int64_t y_ld, int64_t y_bs, int rows,
                               int rows_per_batch, int C, void* stream) {
  VILCO_CHECK_ARG(x && w && b && (y32 || y16), "vilco_layernorm: null pointer");
  VILCO_CHECK_ARG(C % 128 == 0 && C <= 128 * MAXCH, "vilco_layernorm: C=%d must be a multiple of 128 and <= %d", C, 128 * MAXCH);
  VILCO_CHECK_ARG(rows > 0 && rows_per_batch > 0 && y_ld % 4 == 0 && y_bs % 4 == 0, "vilco_layernorm: bad shape");
  LnParams p{};
  p.x = x; p.x_bf16 = x_dtype == VILCO_BF16; p.add = add; p.w = w; p.b = b; p.eps = eps; p.relu = relu;
  p.pe = pe; p.pe_T = pe_T; p.rowmul = rowmul; p.zero_rows = zero_rows;
  p.y32 = y32; p.y16 = static_cast<__nv_bfloat16*>(y16); p.y16_lo = y16_lo; p.y_ld = y_ld; p.y_bs = y_bs;
  p.rows = rows; p.rows_per_batch = rows_per_batch; p.C = C;
  p.fmt = act_fmt();
  const int grid = (rows + 7) / 8;
  if (p.x_bf16) layernorm_kernel<__nv_bfloat16><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else layernorm_kernel<float><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_dwconv_ln(const void* x, int x_dtype, const float* mask, const int* tlen, const float* const* wconv,
                               const float* const* lnw, const float* const* lnb, void* const* out, int64_t out_lo,
                               int n_out, int B, int T, int C, int stride, float eps, void* stream) {
  VILCO_CHECK_ARG(x && mask && wconv && lnw && lnb && out, "vilco_dwconv_ln: null pointer");
  VILCO_CHECK_ARG(n_out >= 1 && n_out <= 3, "vilco_dwconv_ln: n_out must be 1..3");
  VILCO_CHECK_ARG(C % 128 == 0 && C <= 128 * MAXCH, "vilco_dwconv_ln: C=%d unsupported", C);
  VILCO_CHECK_ARG((stride == 1 || stride == 2) && T % stride == 0, "vilco_dwconv_ln: stride %d / T %d", stride, T);
  DwParams p{};
  p.x = x; p.x_bf16 = x_dtype == VILCO_BF16; p.mask = mask; p.tlen = tlen;
  for (int i = 0; i < n_out; ++i) {
    p.wconv[i] = wconv[i]; p.lnw[i] = lnw[i]; p.lnb[i] = lnb[i]; p.out[i] = static_cast<__nv_bfloat16*>(out[i]);
  }
  p.out_lo = out_lo;
  p.B = B; p.T = T; p.C = C; p.stride = stride; p.eps = eps;
  p.fmt = act_fmt();
  const int rows = B * (T / stride);
  dim3 grid((rows + 7) / 8, n_out);
  if (p.x_bf16) dwconv_ln_kernel<__nv_bfloat16><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else dwconv_ln_kernel<float><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_maxpool3s2(const float* x, float* y, int B, int T, int C, void* stream) {
  VILCO_CHECK_ARG(x && y && T % 2 == 0 && C % 4 == 0, "vilco_maxpool3s2: bad arguments");
  const long long n4 = (long long)B * (T / 2) * (C / 4);
  maxpool3s2_kernel<<<grid_for(n4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, B, T, C);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_axpby(const float* x, const float* y, float a, float b, float* o32, void* o16, int64_t o16_lo,
                           int64_t n, void* stream) {
  VILCO_CHECK_ARG(x && (o32 || o16) && n % 4 == 0, "vilco_axpby: bad arguments");
  axpby_kernel<<<grid_for(n / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, y, a, b, o32, static_cast<__nv_bfloat16*>(o16), o16_lo, n / 4, act_fmt());
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_pack_feats(const float* x, void* y, int64_t y_lo, int B, int C, int T, int T_out, void* stream) {
  VILCO_CHECK_ARG(x && y && B > 0 && C > 0 && T > 0 && T_out >= T, "vilco_pack_feats: bad arguments");
  dim3 grid((T_out + 31) / 32, (C + 31) / 32, B);
  pack_feats_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(y), y_lo, C, T, T_out,
                                                                            act_fmt());
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_resize_feats(const float* x, const int64_t* row_start, int B, int C, int T_out, float* out32, void* out16,
                                  int64_t out16_lo, void* stream) {
  VILCO_CHECK_ARG(x && row_start && (out32 || out16) && B > 0 && T_out > 0, "vilco_resize_feats: bad arguments");
  VILCO_CHECK_ARG(C > 0 && C % 4 == 0, "vilco_resize_feats: C = %d must be a multiple of 4", C);
  static_assert(sizeof(long long) == sizeof(int64_t), "int64_t row offsets");
  const long long n = static_cast<long long>(T_out) * (C / 4);
  dim3 grid(grid_for(n, 256, 148 * 8 / (B < 8 ? B : 8)), B);
  resize_feats_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<const long long*>(row_start), out32, static_cast<__nv_bfloat16*>(out16), out16_lo, C / 4, T_out,
      act_fmt());
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_unpack(const float* x, float* y, int B, int T, int C, void* stream) {
  VILCO_CHECK_ARG(x && y && B > 0 && C > 0 && T > 0, "vilco_unpack: bad arguments");
  dim3 grid((C + 31) / 32, (T + 31) / 32, B);
  unpack_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(x, y, T, C);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_scale_add(const float* x, const float* rowmul, const float* y, const float* scale, float* out,
                               int64_t rows, int C, void* stream) {
  VILCO_CHECK_ARG(x && y && out && rows > 0 && C % 4 == 0, "vilco_scale_add: bad arguments");
  scale_add_kernel<<<grid_for(rows * (C / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rowmul, y, scale, out, rows, C);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}
