// Attention-side kernels that are not GEMMs:
//  * masked row softmax over materialised scores (global MaskedMHCA / cross MaskedMHA)        blocks.py:228-269, 351-410
//  * XLNet relative-attention softmax: (ac + rel_shift(bd)) * scale - 1e30 * mask             modeling_xlnet_x.py:256-320
//  * LocalMaskedMHCA core: sliding-window attention with the window staged in shared memory   blocks.py:1038-1138
//  * ChannelAttention core: softmax_d((k*s)^T v) then q A^T                                     blocks.py:423-436
#include "common.cuh"

namespace vilco {

// ---------------------------------------------------------------------------------------------
// row softmax.  S: (Z2, Z1, Tq, Tk) fp32 -> P: same shape bf16 with row stride p_ld.
// mode 0: keys with kmask == 0 get -inf (probability exactly 0).
// mode 1 (XLNet): score = (S[i,j] + BD[i, Tk + j - i]) * scale; minus 1e30 where key j is padding and i != j.
// ---------------------------------------------------------------------------------------------
struct SmParams {
  const float* S; const float* BD; const float* kmask;  // kmask (Z2, Tk)
  __nv_bfloat16* P; long long p_lo;
  float* P32;   // optional fp32 copy of the probabilities (row stride Tk), kept for the backward pass
  int Z1, Tq, Tk; long long p_ld; float scale; int mode;
  long long rows;
  int fmt;      // element format of P
};

__global__ void __launch_bounds__(256) softmax_rows_kernel(const SmParams p) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const int i = static_cast<int>(row % p.Tq);
  const long long zh = row / p.Tq;
  const int b = static_cast<int>(zh / p.Z1);
  const float* s = p.S + row * p.Tk;
  const float* km = p.kmask ? p.kmask + (long long)b * p.Tk : nullptr;
  const float* bd = p.mode == 1 ? p.BD + row * (2LL * p.Tk) + (p.Tk - i) : nullptr;
  __nv_bfloat16* o = p.P + row * p.p_ld;

  auto score = [&](int j) -> float {
    float v = s[j];
    if (p.mode == 0) {
      if (km && km[j] == 0.f) v = -INFINITY;
    } else {
      v = (v + bd[j]) * p.scale;
      if (km && km[j] == 0.f && j != i) v -= 1e30f;
    }
    return v;
  };
  float mx = -INFINITY;
  for (int j = lane; j < p.Tk; j += 32) mx = fmaxf(mx, score(j));
  mx = warp_max(mx);
  if (mx == -INFINITY) {  // fully masked row (empty clip): the reference would produce NaN; emit zeros
    for (int j = lane; j < p.p_ld; j += 32) {
      o[j] = __float2bfloat16_rn(0.f);              // 0x0000 in both formats
      if (p.p_lo) o[p.p_lo + j] = __float2bfloat16_rn(0.f);
    }
    return;
  }
  float sum = 0.f;
  for (int j = lane; j < p.Tk; j += 32) sum += __expf(score(j) - mx);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j = lane; j < p.p_ld; j += 32) {
    const float v = j < p.Tk ? __expf(score(j) - mx) * inv : 0.f;
    store16_split(reinterpret_cast<uint16_t*>(o), j, p.p_lo, v, p.fmt);
    if (p.P32 && j < p.Tk) p.P32[row * p.Tk + j] = v;
  }
}

// single-pass variant: the whole row lives in registers (Tk even, Tk <= 64*KMAX); lane l holds columns 2l+64k, 2l+64k+1
// so loads are 8-byte and bf16 stores 4-byte per lane, fully coalesced; S is read from HBM exactly once.
template <int KMAX>
__global__ void __launch_bounds__(256) softmax_rows_reg_kernel(const SmParams p) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const int i = static_cast<int>(row % p.Tq);
  const int b = static_cast<int>((row / p.Tq) / p.Z1);
  const float* s = p.S + row * p.Tk;
  const float* km = p.kmask ? p.kmask + (long long)b * p.Tk : nullptr;
  const float* bd = p.mode == 1 ? p.BD + row * (2LL * p.Tk) + (p.Tk - i) : nullptr;
  __nv_bfloat16* o = p.P + row * p.p_ld;
  float v[KMAX][2];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int j = 2 * lane + 64 * k;
    v[k][0] = -INFINITY; v[k][1] = -INFINITY;
    if (j < p.Tk) {
      const float2 x = *reinterpret_cast<const float2*>(s + j);
      float a0 = x.x, a1 = x.y;
      float m0 = 1.f, m1 = 1.f;
      if (km) { const float2 mm = *reinterpret_cast<const float2*>(km + j); m0 = mm.x; m1 = mm.y; }
      if (p.mode == 0) {
        if (m0 == 0.f) a0 = -INFINITY;
        if (m1 == 0.f) a1 = -INFINITY;
      } else {
        a0 = (a0 + bd[j]) * p.scale;
        a1 = (a1 + bd[j + 1]) * p.scale;
        if (m0 == 0.f && j != i) a0 -= 1e30f;
        if (m1 == 0.f && j + 1 != i) a1 -= 1e30f;
      }
      v[k][0] = a0; v[k][1] = a1;
      mx = fmaxf(mx, fmaxf(a0, a1));
    }
  }
  mx = warp_max(mx);
  const bool dead = mx == -INFINITY;  // fully masked row: emit zeros (the reference would produce NaN)
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    v[k][0] = dead ? 0.f : __expf(v[k][0] - mx);
    v[k][1] = dead ? 0.f : __expf(v[k][1] - mx);
    sum += v[k][0] + v[k][1];
  }
  sum = warp_sum(sum);
  const float inv = dead ? 0.f : 1.0f / sum;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int j = 2 * lane + 64 * k;
    if (j < p.p_ld) {
      const float a0 = v[k][0] * inv, a1 = v[k][1] * inv;
      if (p.p_lo) {
        uint32_t h, l;
        split16x2(a0, a1, p.fmt, h, l);
        *reinterpret_cast<uint32_t*>(o + j) = h;
        *reinterpret_cast<uint32_t*>(o + p.p_lo + j) = l;
      } else {
        *reinterpret_cast<uint32_t*>(o + j) = pack16x2(a0, a1, p.fmt);
      }
      if (p.P32 && j < p.Tk) *reinterpret_cast<float2*>(p.P32 + row * p.Tk + j) = make_float2(a0, a1);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// local (windowed) attention.  q,k,v: (B, T, C) bf16 token-major, head h = channels [h*d, (h+1)*d).
// One CTA = (token tile, head, batch); K/V rows [t0-w, t0+TI+w) staged in shared memory once and reused by all
// queries of the tile; one warp per query, lanes over the head dim, warp-shuffle dot products.
// ---------------------------------------------------------------------------------------------
static constexpr int LW_TI = 32;      // queries per CTA
static constexpr int LW_MAXW = 65;    // max window size
static constexpr int LW_MAXD = 128;   // max head dim

int local_attn_tc(const void* q, const void* k, const void* v, const float* mask, const float* rel_pe, void* out, long long lo,
                  int B, int T, int C, int H, int W, float scale, int fmt, void* stream);   // local_attn_tc.cu

struct LwParams {
  const __nv_bfloat16 *q, *k, *v; const float* mask;  // mask (B, T)
  const float* rel_pe;                                // (H, W) or null
  __nv_bfloat16* out;
  long long lo;                                       // lo-plane offset of q/k/v/out (0 = single plane)
  int B, T, C, H, d, W; float scale;
  int fmt;
};

__device__ __forceinline__ float2 ld2_split(const __nv_bfloat16* p, long long lo, int fmt) {
  float2 r = unpack16x2(*reinterpret_cast<const uint32_t*>(p), fmt);
  if (lo) {
    const float2 y = unpack16x2(*reinterpret_cast<const uint32_t*>(p + lo), fmt);
    r.x += y.x; r.y += y.y;
  }
  return r;
}
__device__ __forceinline__ void st2_split(__nv_bfloat16* p, long long lo, float a, float b, int fmt) {
  uint32_t h, l;
  split16x2(a, b, fmt, h, l);
  *reinterpret_cast<uint32_t*>(p) = h;
  if (lo) *reinterpret_cast<uint32_t*>(p + lo) = l;
}

__global__ void __launch_bounds__(256) local_attn_kernel(const LwParams p) {
  extern __shared__ float lw_smem[];
  const int w = p.W / 2;
  const int t0 = blockIdx.x * LW_TI, h = blockIdx.y, b = blockIdx.z;
  const int nrows = LW_TI + 2 * w;
  float* sk = lw_smem;
  float* sv = lw_smem + (size_t)nrows * p.d;
  const long long base = (long long)b * p.T * p.C + (long long)h * p.d;
  // stage the K/V window as fp32 (hi + lo planes summed); rows outside the sequence are never read
  const int half = p.d / 2;
  for (int idx = threadIdx.x; idx < nrows * half; idx += blockDim.x) {
    const int r = idx / half, c = (idx - r * half) * 2;
    const int t = t0 - w + r;
    if (t >= 0 && t < p.T) {
      const float2 kk = ld2_split(p.k + base + (long long)t * p.C + c, p.lo, p.fmt);
      const float2 vv = ld2_split(p.v + base + (long long)t * p.C + c, p.lo, p.fmt);
      sk[r * p.d + c] = kk.x; sk[r * p.d + c + 1] = kk.y;
      sv[r * p.d + c] = vv.x; sv[r * p.d + c + 1] = vv.y;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const float* mk = p.mask + (long long)b * p.T;
  for (int qi = warp; qi < LW_TI; qi += nwarp) {
    const int t = t0 + qi;
    if (t >= p.T) break;
    __nv_bfloat16* o = p.out + base + (long long)t * p.C;
    if (mk[t] == 0.f) {  // padded query: probabilities are zeroed (blocks.py:1192-1194)
      for (int c = lane * 2; c < p.d; c += 64) st2_split(o + c, p.lo, 0.f, 0.f, p.fmt);
      continue;
    }
    float qv[LW_MAXD / 32];
#pragma unroll
    for (int u = 0; u < LW_MAXD / 64; ++u) {
      const int c = lane * 2 + u * 64;
      if (c < p.d) {
        const float2 x = ld2_split(p.q + base + (long long)t * p.C + c, p.lo, p.fmt);
        qv[2 * u] = x.x * p.scale; qv[2 * u + 1] = x.y * p.scale;
      } else { qv[2 * u] = 0.f; qv[2 * u + 1] = 0.f; }
    }
    // scores for j = t-w .. t+w   (lane jj%32 keeps the score of window slot jj)
    float sc0 = -INFINITY, sc1 = -INFINITY, sc2 = -INFINITY;
    for (int jj = 0; jj < p.W; ++jj) {
      const int j = t - w + jj;
      float part = 0.f;
      if (j >= 0 && j < p.T) {
        const float* kr = sk + (qi + jj) * p.d;
#pragma unroll
        for (int u = 0; u < LW_MAXD / 64; ++u) {
          const int c = lane * 2 + u * 64;
          if (c < p.d) {
            part = fmaf(qv[2 * u], kr[c], part);
            part = fmaf(qv[2 * u + 1], kr[c + 1], part);
          }
        }
      }
      float s = warp_sum(part);
      if (j < 0 || j >= p.T) s = -INFINITY;
      else {
        if (p.rel_pe) s += p.rel_pe[h * p.W + jj];
        if (mk[j] == 0.f) s += -1e4f;
      }
      if ((jj & 31) == lane) { if (jj < 32) sc0 = s; else if (jj < 64) sc1 = s; else sc2 = s; }
    }
    float mx = warp_max(fmaxf(fmaxf(sc0, sc1), sc2));
    float e0 = sc0 == -INFINITY ? 0.f : __expf(sc0 - mx);
    float e1 = sc1 == -INFINITY ? 0.f : __expf(sc1 - mx);
    float e2 = sc2 == -INFINITY ? 0.f : __expf(sc2 - mx);
    const float inv = 1.0f / warp_sum(e0 + e1 + e2);
    float acc[LW_MAXD / 32];
#pragma unroll
    for (int u = 0; u < LW_MAXD / 32; ++u) acc[u] = 0.f;
    for (int jj = 0; jj < p.W; ++jj) {
      const int j = t - w + jj;
      const float pe = __shfl_sync(0xffffffffu, jj < 32 ? e0 : (jj < 64 ? e1 : e2), jj & 31) * inv;
      if (j < 0 || j >= p.T) continue;
      const float* vr = sv + (qi + jj) * p.d;
#pragma unroll
      for (int u = 0; u < LW_MAXD / 64; ++u) {
        const int c = lane * 2 + u * 64;
        if (c < p.d) {
          acc[2 * u] = fmaf(pe, vr[c], acc[2 * u]);
          acc[2 * u + 1] = fmaf(pe, vr[c + 1], acc[2 * u + 1]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < LW_MAXD / 64; ++u) {
      const int c = lane * 2 + u * 64;
      if (c < p.d) st2_split(o + c, p.lo, acc[2 * u], acc[2 * u + 1], p.fmt);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// local (windowed) attention, backward.  Two passes with the staging scheme of the forward kernel:
//   pass Q (one warp per query t): recompute the W probabilities, dP[jj] = dO[t] . v[j], dS = P * (dP - sum P dP),
//          dq[t] = scale * sum_jj dS[jj] k[j]; P and dS rows go to a (B, H, T, W) fp32 scratch
//   pass KV (one warp per key j): dk[j] = scale * sum_jj dS[t, jj] q[t], dv[j] = sum_jj P[t, jj] dO[t] with t = j + w - jj
//          (the window is symmetric, so the queries that see key j are t in [j - w, j + w]); no atomics.
// ---------------------------------------------------------------------------------------------
struct LwBwdParams {
  const __nv_bfloat16 *q, *k, *v; const float* mask; const float* rel_pe; const float* dO;
  float *sP, *sdS;               // scratch (B, H, T, W)
  float *dq, *dk, *dv;           // (B, T, C) fp32
  long long lo;
  int B, T, C, H, d, W; float scale;
  int fmt;
};

__global__ void __launch_bounds__(256) local_attn_bwd_q_kernel(const LwBwdParams p) {
  extern __shared__ float lw_smem[];
  const int w = p.W / 2;
  const int t0 = blockIdx.x * LW_TI, h = blockIdx.y, b = blockIdx.z;
  const int nrows = LW_TI + 2 * w;
  float* sk = lw_smem;
  float* sv = lw_smem + (size_t)nrows * p.d;
  const long long base = (long long)b * p.T * p.C + (long long)h * p.d;
  const int half = p.d / 2;
  for (int idx = threadIdx.x; idx < nrows * half; idx += blockDim.x) {
    const int r = idx / half, c = (idx - r * half) * 2;
    const int t = t0 - w + r;
    if (t >= 0 && t < p.T) {
      const float2 kk = ld2_split(p.k + base + (long long)t * p.C + c, p.lo, p.fmt);
      const float2 vv = ld2_split(p.v + base + (long long)t * p.C + c, p.lo, p.fmt);
      sk[r * p.d + c] = kk.x; sk[r * p.d + c + 1] = kk.y;
      sv[r * p.d + c] = vv.x; sv[r * p.d + c + 1] = vv.y;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const float* mk = p.mask + (long long)b * p.T;
  for (int qi = warp; qi < LW_TI; qi += nwarp) {
    const int t = t0 + qi;
    if (t >= p.T) break;
    float* rowP = p.sP + (((long long)b * p.H + h) * p.T + t) * p.W;
    float* rowS = p.sdS + (((long long)b * p.H + h) * p.T + t) * p.W;
    float* dqo = p.dq + base + (long long)t * p.C;
    if (mk[t] == 0.f) {      // padded query: its probabilities were zeroed in the forward pass
      for (int jj = lane; jj < p.W; jj += 32) { rowP[jj] = 0.f; rowS[jj] = 0.f; }
      for (int c = lane * 2; c < p.d; c += 64) { dqo[c] = 0.f; dqo[c + 1] = 0.f; }
      continue;
    }
    float qv[LW_MAXD / 32], gv[LW_MAXD / 32];
#pragma unroll
    for (int u = 0; u < LW_MAXD / 64; ++u) {
      const int c = lane * 2 + u * 64;
      if (c < p.d) {
        const float2 x = ld2_split(p.q + base + (long long)t * p.C + c, p.lo, p.fmt);
        qv[2 * u] = x.x * p.scale; qv[2 * u + 1] = x.y * p.scale;
        gv[2 * u] = p.dO[base + (long long)t * p.C + c]; gv[2 * u + 1] = p.dO[base + (long long)t * p.C + c + 1];
      } else { qv[2 * u] = qv[2 * u + 1] = gv[2 * u] = gv[2 * u + 1] = 0.f; }
    }
    float sc0 = -INFINITY, sc1 = -INFINITY, sc2 = -INFINITY, dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
    for (int jj = 0; jj < p.W; ++jj) {
      const int j = t - w + jj;
      float part = 0.f, dpart = 0.f;
      if (j >= 0 && j < p.T) {
        const float* kr = sk + (qi + jj) * p.d;
        const float* vr = sv + (qi + jj) * p.d;
#pragma unroll
        for (int u = 0; u < LW_MAXD / 64; ++u) {
          const int c = lane * 2 + u * 64;
          if (c < p.d) {
            part = fmaf(qv[2 * u], kr[c], part); part = fmaf(qv[2 * u + 1], kr[c + 1], part);
            dpart = fmaf(gv[2 * u], vr[c], dpart); dpart = fmaf(gv[2 * u + 1], vr[c + 1], dpart);
          }
        }
      }
      float s = warp_sum(part);
      const float dpv = warp_sum(dpart);
      if (j < 0 || j >= p.T) s = -INFINITY;
      else {
        if (p.rel_pe) s += p.rel_pe[h * p.W + jj];
        if (mk[j] == 0.f) s += -1e4f;
      }
      if ((jj & 31) == lane) {
        if (jj < 32) { sc0 = s; dp0 = dpv; } else if (jj < 64) { sc1 = s; dp1 = dpv; } else { sc2 = s; dp2 = dpv; }
      }
    }
    const float mx = warp_max(fmaxf(fmaxf(sc0, sc1), sc2));
    float e0 = sc0 == -INFINITY ? 0.f : __expf(sc0 - mx);
    float e1 = sc1 == -INFINITY ? 0.f : __expf(sc1 - mx);
    float e2 = sc2 == -INFINITY ? 0.f : __expf(sc2 - mx);
    const float inv = 1.0f / warp_sum(e0 + e1 + e2);
    e0 *= inv; e1 *= inv; e2 *= inv;
    const float delta = warp_sum(e0 * dp0 + e1 * dp1 + e2 * dp2);
    const float ds0 = e0 * (dp0 - delta), ds1 = e1 * (dp1 - delta), ds2 = e2 * (dp2 - delta);
    if (lane < p.W) { rowP[lane] = e0; rowS[lane] = ds0; }
    if (lane + 32 < p.W) { rowP[lane + 32] = e1; rowS[lane + 32] = ds1; }
    if (lane + 64 < p.W) { rowP[lane + 64] = e2; rowS[lane + 64] = ds2; }
    float acc[LW_MAXD / 32];
#pragma unroll
    for (int u = 0; u < LW_MAXD / 32; ++u) acc[u] = 0.f;
    for (int jj = 0; jj < p.W; ++jj) {
      const int j = t - w + jj;
      const float ds = __shfl_sync(0xffffffffu, jj < 32 ? ds0 : (jj < 64 ? ds1 : ds2), jj & 31);
      if (j < 0 || j >= p.T) continue;
      const float* kr = sk + (qi + jj) * p.d;
#pragma unroll
      for (int u = 0; u < LW_MAXD / 64; ++u) {
        const int c = lane * 2 + u * 64;
        if (c < p.d) { acc[2 * u] = fmaf(ds, kr[c], acc[2 * u]); acc[2 * u + 1] = fmaf(ds, kr[c + 1], acc[2 * u + 1]); }
      }
    }
#pragma unroll
    for (int u = 0; u < LW_MAXD / 64; ++u) {
      const int c = lane * 2 + u * 64;
      if (c < p.d) { dqo[c] = acc[2 * u] * p.scale; dqo[c + 1] = acc[2 * u + 1] * p.scale; }
    }
  }
}

__global__ void __launch_bounds__(256) local_attn_bwd_kv_kernel(const LwBwdParams p) {
  extern __shared__ float lw_smem[];
  const int w = p.W / 2;
  const int t0 = blockIdx.x * LW_TI, h = blockIdx.y, b = blockIdx.z;
  const int nrows = LW_TI + 2 * w;
  float* sq = lw_smem;                              // queries  t0 - w .. t0 + TI + w
  float* sg = lw_smem + (size_t)nrows * p.d;        // dO rows of the same queries
  const long long base = (long long)b * p.T * p.C + (long long)h * p.d;
  const int half = p.d / 2;
  for (int idx = threadIdx.x; idx < nrows * half; idx += blockDim.x) {
    const int r = idx / half, c = (idx - r * half) * 2;
    const int t = t0 - w + r;
    if (t >= 0 && t < p.T) {
      const float2 qq = ld2_split(p.q + base + (long long)t * p.C + c, p.lo, p.fmt);
      sq[r * p.d + c] = qq.x; sq[r * p.d + c + 1] = qq.y;
      sg[r * p.d + c] = p.dO[base + (long long)t * p.C + c]; sg[r * p.d + c + 1] = p.dO[base + (long long)t * p.C + c + 1];
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int ki = warp; ki < LW_TI; ki += nwarp) {
    const int j = t0 + ki;
    if (j >= p.T) break;
    float ak[LW_MAXD / 32], av[LW_MAXD / 32];
#pragma unroll
    for (int u = 0; u < LW_MAXD / 32; ++u) { ak[u] = 0.f; av[u] = 0.f; }
    for (int jj = 0; jj < p.W; ++jj) {
      const int t = j + w - jj;                     // the query for which key j sits in window slot jj
      if (t < 0 || t >= p.T) continue;
      const long long so = (((long long)b * p.H + h) * p.T + t) * p.W + jj;
      const float ds = p.sdS[so], pp = p.sP[so];
      const float* qr = sq + (t - (t0 - w)) * p.d;
      const float* gr = sg + (t - (t0 - w)) * p.d;
#pragma unroll
      for (int u = 0; u < LW_MAXD / 64; ++u) {
        const int c = lane * 2 + u * 64;
        if (c < p.d) {
          ak[2 * u] = fmaf(ds, qr[c], ak[2 * u]); ak[2 * u + 1] = fmaf(ds, qr[c + 1], ak[2 * u + 1]);
          av[2 * u] = fmaf(pp, gr[c], av[2 * u]); av[2 * u + 1] = fmaf(pp, gr[c + 1], av[2 * u + 1]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < LW_MAXD / 64; ++u) {
      const int c = lane * 2 + u * 64;
      if (c < p.d) {
        p.dk[base + (long long)j * p.C + c] = ak[2 * u] * p.scale; p.dk[base + (long long)j * p.C + c + 1] = ak[2 * u + 1] * p.scale;
        p.dv[base + (long long)j * p.C + c] = av[2 * u]; p.dv[base + (long long)j * p.C + c + 1] = av[2 * u + 1];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// channel attention.  qkv: (B, T, 3C) bf16 with column = which*C + h*64 + d  (d = 64 fixed).
// phase 1: G[b,h] (64x64 fp32, zero-initialised by the caller) += sum over a T chunk of (k*scale)^T v
// phase 2: A = softmax_rows(G[b,h]);  y[t, h*64 + i] = sum_j A[i,j] q[t, j]
// ---------------------------------------------------------------------------------------------
static constexpr int CA_D = 64;
static constexpr int CA_TCHUNK = 64;

// One CTA per (head, clip) walks all T chunks in order: the sum over tokens is formed in a FIXED order (no atomics), so the
// result — which feeds a d x d softmax that amplifies its absolute error — is bit-reproducible from run to run.
__global__ void __launch_bounds__(256) chan_attn_kv_kernel(const __nv_bfloat16* __restrict__ qkv, long long lo,
                                                           float* __restrict__ G, const int* __restrict__ tlen, int T,
                                                           int C, int H, float scale, int fmt) {
  __shared__ float sk[CA_TCHUNK][CA_D + 1];
  __shared__ float sv[CA_TCHUNK][CA_D + 1];
  const int h = blockIdx.y, b = blockIdx.z;
  const int Tb = tlen ? min(T, tlen[b]) : T;  // tokens that take part in the k^T v sum
  const long long ld = 3LL * C;
  // thread (i, j4): 4x4 sub-block of the 64x64 output
  const int ti = (threadIdx.x / 16) * 4, tj = (threadIdx.x % 16) * 4;
  float acc[4][4] = {};
  for (int t0 = 0; t0 < Tb; t0 += CA_TCHUNK) {
    const int nt = min(CA_TCHUNK, Tb - t0);
    const __nv_bfloat16* kb = qkv + ((long long)b * T + t0) * ld + C + h * CA_D;
    const __nv_bfloat16* vb = kb + C;
    __syncthreads();   // the previous chunk has been consumed
    for (int idx = threadIdx.x; idx < CA_TCHUNK * (CA_D / 2); idx += blockDim.x) {
      const int r = idx / (CA_D / 2), c = (idx % (CA_D / 2)) * 2;
      float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
      if (r < nt) {
        const float2 xk = ld2_split(kb + r * ld + c, lo, fmt);
        const float2 xv = ld2_split(vb + r * ld + c, lo, fmt);
        k0 = xk.x * scale; k1 = xk.y * scale; v0 = xv.x; v1 = xv.y;
      }
      sk[r][c] = k0; sk[r][c + 1] = k1; sv[r][c] = v0; sv[r][c + 1] = v1;
    }
    __syncthreads();
    for (int r = 0; r < CA_TCHUNK; ++r) {
      float a[4], c[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = sk[r][ti + u]; c[u] = sv[r][tj + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int w = 0; w < 4; ++w) acc[u][w] = fmaf(a[u], c[w], acc[u][w]);
    }
  }
  float* g = G + ((long long)b * H + h) * CA_D * CA_D;
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int w = 0; w < 4; ++w) g[(ti + u) * CA_D + tj + w] = acc[u][w];
}

__global__ void __launch_bounds__(256) chan_attn_apply_kernel(const __nv_bfloat16* __restrict__ qkv, long long lo,
                                                              const float* __restrict__ G, __nv_bfloat16* __restrict__ y,
                                                              long long y_lo, int T, int C, int H, int fmt) {
  __shared__ float sa[CA_D][CA_D + 1];
  __shared__ float sq[CA_TCHUNK][CA_D + 1];
  const int h = blockIdx.y, b = blockIdx.z;
  const int t0 = blockIdx.x * CA_TCHUNK;
  const int nt = min(CA_TCHUNK, T - t0);
  const float* g = G + ((long long)b * H + h) * CA_D * CA_D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // softmax of each row i over j (8 warps x 8 rows)
  for (int i = warp; i < CA_D; i += 8) {
    const float a0 = g[i * CA_D + lane], a1 = g[i * CA_D + lane + 32];
    const float mx = warp_max(fmaxf(a0, a1));
    const float e0 = __expf(a0 - mx), e1 = __expf(a1 - mx);
    const float inv = 1.0f / warp_sum(e0 + e1);
    sa[i][lane] = e0 * inv; sa[i][lane + 32] = e1 * inv;
  }
  const long long ld = 3LL * C;
  const __nv_bfloat16* qb = qkv + ((long long)b * T + t0) * ld + h * CA_D;
  for (int idx = threadIdx.x; idx < CA_TCHUNK * (CA_D / 2); idx += blockDim.x) {
    const int r = idx / (CA_D / 2), c = (idx % (CA_D / 2)) * 2;
    float q0 = 0.f, q1 = 0.f;
    if (r < nt) {
      const float2 x = ld2_split(qb + r * ld + c, lo, fmt);
      q0 = x.x; q1 = x.y;
    }
    sq[r][c] = q0; sq[r][c + 1] = q1;
  }
  __syncthreads();
  // thread: token r = tid / 4 (64 tokens), outputs i in [ (tid%4)*16, +16 )
  const int r = threadIdx.x / 4, i0 = (threadIdx.x % 4) * 16;
  if (r < nt) {
    float acc[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) acc[u] = 0.f;
    for (int j = 0; j < CA_D; ++j) {
      const float qj = sq[r][j];
#pragma unroll
      for (int u = 0; u < 16; ++u) acc[u] = fmaf(sa[i0 + u][j], qj, acc[u]);
    }
    __nv_bfloat16* o = y + ((long long)b * T + t0 + r) * C + h * CA_D + i0;
#pragma unroll
    for (int u = 0; u < 16; u += 2) st2_split(o + u, y_lo, acc[u], acc[u + 1], fmt);
  }
}

// ---------------------------------------------------------------------------------------------
// channel attention backward.  y[t,i] = sum_j A[i,j] q[t,j],  A = softmax_j(G),  G[i,j] = sum_t (s k[t,i]) v[t,j]
//   dA[i,j] = sum_t dy[t,i] q[t,j]           (kernel 1, atomics over T chunks, dA pre-zeroed)
//   dG = A o (dA - rowsum(dA o A)); dq[t,j] = sum_i dy[t,i] A[i,j]; dk[t,i] = s sum_j dG[i,j] v[t,j];
//   dv[t,j] = s sum_i dG[i,j] k[t,i]         (kernel 2, per T chunk)
// ---------------------------------------------------------------------------------------------
static constexpr int CB_T = 32;

__global__ void __launch_bounds__(256) chan_attn_dA_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ qkv,
                                                           long long lo, float* __restrict__ dA, int T, int C, int H, int fmt) {
  __shared__ float sd[CA_TCHUNK][CA_D + 1];
  __shared__ float sq[CA_TCHUNK][CA_D + 1];
  const int h = blockIdx.y, b = blockIdx.z;
  const int t0 = blockIdx.x * CA_TCHUNK;
  const int nt = min(CA_TCHUNK, T - t0);
  const long long ld = 3LL * C;
  const __nv_bfloat16* qb = qkv + ((long long)b * T + t0) * ld + h * CA_D;
  const float* db = dy + ((long long)b * T + t0) * C + h * CA_D;
  for (int idx = threadIdx.x; idx < CA_TCHUNK * (CA_D / 2); idx += blockDim.x) {
    const int r = idx / (CA_D / 2), c = (idx % (CA_D / 2)) * 2;
    float q0 = 0.f, q1 = 0.f, d0 = 0.f, d1 = 0.f;
    if (r < nt) {
      const float2 x = ld2_split(qb + r * ld + c, lo, fmt);
      q0 = x.x; q1 = x.y;
      d0 = db[(long long)r * C + c]; d1 = db[(long long)r * C + c + 1];
    }
    sq[r][c] = q0; sq[r][c + 1] = q1; sd[r][c] = d0; sd[r][c + 1] = d1;
  }
  __syncthreads();
  const int ti = (threadIdx.x / 16) * 4, tj = (threadIdx.x % 16) * 4;
  float acc[4][4] = {};
  for (int r = 0; r < CA_TCHUNK; ++r) {
    float a[4], c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { a[u] = sd[r][ti + u]; c[u] = sq[r][tj + u]; }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int w = 0; w < 4; ++w) acc[u][w] = fmaf(a[u], c[w], acc[u][w]);
  }
  float* g = dA + ((long long)b * H + h) * CA_D * CA_D;
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int w = 0; w < 4; ++w) atomicAdd(g + (ti + u) * CA_D + tj + w, acc[u][w]);
}

__global__ void __launch_bounds__(256) chan_attn_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ qkv,
                                                            long long lo, const float* __restrict__ G, const float* __restrict__ dA,
                                                            float* __restrict__ dqkv, int T, int C, int H, float scale, int fmt) {
  __shared__ float sA[CA_D][CA_D + 1];
  __shared__ float sG[CA_D][CA_D + 1];   // dG
  __shared__ float sx[CB_T][CA_D + 1];
  const int h = blockIdx.y, b = blockIdx.z;
  const int t0 = blockIdx.x * CB_T;
  const int nt = min(CB_T, T - t0);
  const float* g = G + ((long long)b * H + h) * CA_D * CA_D;
  const float* da = dA + ((long long)b * H + h) * CA_D * CA_D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < CA_D; i += 8) {
    const float a0 = g[i * CA_D + lane], a1 = g[i * CA_D + lane + 32];
    const float mx = warp_max(fmaxf(a0, a1));
    float e0 = __expf(a0 - mx), e1 = __expf(a1 - mx);
    const float inv = 1.0f / warp_sum(e0 + e1);
    e0 *= inv; e1 *= inv;
    const float d0 = da[i * CA_D + lane], d1 = da[i * CA_D + lane + 32];
    const float dot = warp_sum(d0 * e0 + d1 * e1);
    sA[i][lane] = e0; sA[i][lane + 32] = e1;
    sG[i][lane] = e0 * (d0 - dot); sG[i][lane + 32] = e1 * (d1 - dot);
  }
  const long long ld = 3LL * C;
  const int r = threadIdx.x / 8, c0 = (threadIdx.x % 8) * 8;   // token r of the chunk, 8 outputs
  float* out = dqkv + ((long long)b * T + t0 + r) * ld + h * CA_D + c0;
  auto load = [&](int which) {  // 0: dy, 1: v, 2: k  -> sx
    __syncthreads();
    for (int idx = threadIdx.x; idx < CB_T * CA_D; idx += blockDim.x) {
      const int rr = idx / CA_D, cc = idx % CA_D;
      float v = 0.f;
      if (rr < nt) {
        if (which == 0) v = dy[((long long)b * T + t0 + rr) * C + h * CA_D + cc];
        else {
          const __nv_bfloat16* p = qkv + ((long long)b * T + t0 + rr) * ld + (which == 1 ? 2 : 1) * C + h * CA_D + cc;
          v = load16_split(reinterpret_cast<const uint16_t*>(p), lo, fmt);
        }
      }
      sx[rr][cc] = v;
    }
    __syncthreads();
  };
  float acc[8];
  // dq[t, j] = sum_i dy[t, i] A[i, j]
  load(0);
#pragma unroll
  for (int u = 0; u < 8; ++u) acc[u] = 0.f;
  for (int i = 0; i < CA_D; ++i) {
    const float d = sx[r][i];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = fmaf(d, sA[i][c0 + u], acc[u]);
  }
  if (r < nt)
#pragma unroll
    for (int u = 0; u < 8; ++u) out[u] = acc[u];
  // dk[t, i] = s * sum_j dG[i, j] v[t, j]
  load(1);
#pragma unroll
  for (int u = 0; u < 8; ++u) acc[u] = 0.f;
  for (int j = 0; j < CA_D; ++j) {
    const float v = sx[r][j];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = fmaf(sG[c0 + u][j], v, acc[u]);
  }
  if (r < nt)
#pragma unroll
    for (int u = 0; u < 8; ++u) out[C + u] = acc[u] * scale;
  // dv[t, j] = s * sum_i dG[i, j] k[t, i]
  load(2);
#pragma unroll
  for (int u = 0; u < 8; ++u) acc[u] = 0.f;
  for (int i = 0; i < CA_D; ++i) {
    const float k = sx[r][i];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = fmaf(sG[i][c0 + u], k, acc[u]);
  }
  if (r < nt)
#pragma unroll
    for (int u = 0; u < 8; ++u) out[2 * C + u] = acc[u] * scale;
}

}  // namespace vilco

using namespace vilco;

extern "C" int vilco_softmax_rows(const float* S, const float* BD, const float* kmask, void* P, int64_t p_lo, float* P32, int Z2,
                                  int Z1, int Tq, int Tk, int64_t p_ld, float scale, int mode, void* stream) {
  VILCO_CHECK_ARG(S && P && Z1 > 0 && Z2 > 0 && Tq > 0 && Tk > 0 && p_ld >= Tk, "vilco_softmax_rows: bad arguments");
  VILCO_CHECK_ARG(mode == 0 || (mode == 1 && BD && Tq == Tk), "vilco_softmax_rows: mode 1 needs BD and Tq == Tk");
  SmParams p{};
  p.S = S; p.BD = BD; p.kmask = kmask; p.P = static_cast<__nv_bfloat16*>(P); p.p_lo = p_lo; p.P32 = P32;
  p.Z1 = Z1; p.Tq = Tq; p.Tk = Tk; p.p_ld = p_ld; p.scale = scale; p.mode = mode;
  p.rows = (long long)Z2 * Z1 * Tq;
  p.fmt = act_fmt();
  const long long grid = (p.rows + 7) / 8;
  const bool reg_ok = (Tk % 2 == 0) && (p_ld % 2 == 0) && Tk <= 2048 && p_ld <= ((Tk + 63) / 64) * 64 &&
                      (reinterpret_cast<uintptr_t>(S) % 8 == 0) && (!kmask || reinterpret_cast<uintptr_t>(kmask) % 8 == 0);
  if (reg_ok && Tk <= 512)
    softmax_rows_reg_kernel<8><<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else if (reg_ok && Tk <= 1024)
    softmax_rows_reg_kernel<16><<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else if (reg_ok)
    softmax_rows_reg_kernel<32><<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else
    softmax_rows_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_local_attention(const void* q, const void* k, const void* v, const float* mask, const float* rel_pe,
                                     void* out, int64_t lo, int B, int T, int C, int H, int W, void* stream) {
  VILCO_CHECK_ARG(q && k && v && mask && out, "vilco_local_attention: null pointer");
  VILCO_CHECK_ARG(H > 0 && C % H == 0, "vilco_local_attention: C %% H != 0");
  const int d = C / H;
  VILCO_CHECK_ARG(d % 8 == 0 && d <= LW_MAXD, "vilco_local_attention: head dim %d unsupported", d);
  VILCO_CHECK_ARG(W >= 3 && (W & 1) && W <= LW_MAXW, "vilco_local_attention: window %d unsupported (odd, 3..%d)", W, LW_MAXW);
  LwParams p{};
  p.q = static_cast<const __nv_bfloat16*>(q); p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v); p.mask = mask; p.rel_pe = rel_pe;
  p.out = static_cast<__nv_bfloat16*>(out); p.lo = lo;
  p.B = B; p.T = T; p.C = C; p.H = H; p.d = d; p.W = W; p.scale = 1.0f / sqrtf(static_cast<float>(d));
  p.fmt = act_fmt();
  // tensor-core kernel (local_attn_tc.cu) for the template set of head dims / windows; VILCO_LOCAL_ATTN=simt keeps the scalar one
  static const bool force_simt = [] { const char* e = getenv("VILCO_LOCAL_ATTN"); return e && !strcmp(e, "simt"); }();
  if (!force_simt && local_attn_tc(q, k, v, mask, rel_pe, out, lo, B, T, C, H, W, p.scale, p.fmt, stream) == VILCO_OK) return VILCO_OK;
  const size_t smem = (size_t)2 * (LW_TI + 2 * (W / 2)) * d * sizeof(float);
  static bool configured = false;
  if (!configured) {
    VILCO_CUDA(cudaFuncSetAttribute(local_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    2 * (LW_TI + LW_MAXW) * LW_MAXD * (int)sizeof(float)));
    configured = true;
  }
  dim3 grid((T + LW_TI - 1) / LW_TI, H, B);
  local_attn_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(p);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_local_attention_bwd(const float* dO, const void* q, const void* k, const void* v, int64_t lo,
                                         const float* mask, const float* rel_pe, float* scratch_p, float* scratch_ds, float* dq,
                                         float* dk, float* dv, int B, int T, int C, int H, int W, void* stream) {
  VILCO_CHECK_ARG(dO && q && k && v && mask && scratch_p && scratch_ds && dq && dk && dv, "vilco_local_attention_bwd: null pointer");
  VILCO_CHECK_ARG(H > 0 && C % H == 0, "vilco_local_attention_bwd: C %% H != 0");
  const int d = C / H;
  VILCO_CHECK_ARG(d % 8 == 0 && d <= LW_MAXD, "vilco_local_attention_bwd: head dim %d unsupported", d);
  VILCO_CHECK_ARG(W >= 3 && (W & 1) && W <= LW_MAXW, "vilco_local_attention_bwd: window %d unsupported (odd, 3..%d)", W, LW_MAXW);
  LwBwdParams p{};
  p.q = static_cast<const __nv_bfloat16*>(q); p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v); p.mask = mask; p.rel_pe = rel_pe; p.dO = dO;
  p.sP = scratch_p; p.sdS = scratch_ds; p.dq = dq; p.dk = dk; p.dv = dv; p.lo = lo;
  p.B = B; p.T = T; p.C = C; p.H = H; p.d = d; p.W = W; p.scale = 1.0f / sqrtf(static_cast<float>(d));
  p.fmt = act_fmt();
  const size_t smem = (size_t)2 * (LW_TI + 2 * (W / 2)) * d * sizeof(float);
  static bool configured = false;
  if (!configured) {
    const int mx = 2 * (LW_TI + LW_MAXW) * LW_MAXD * (int)sizeof(float);
    VILCO_CUDA(cudaFuncSetAttribute(local_attn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    VILCO_CUDA(cudaFuncSetAttribute(local_attn_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    configured = true;
  }
  dim3 grid((T + LW_TI - 1) / LW_TI, H, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  local_attn_bwd_q_kernel<<<grid, 256, smem, st>>>(p);
  VILCO_LAUNCH_CHECK();
  local_attn_bwd_kv_kernel<<<grid, 256, smem, st>>>(p);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_channel_attention(const void* qkv, int64_t qkv_lo, float* G, void* y, int64_t y_lo, const int* tlen,
                                       int B, int T, int C, int H, void* stream) {
  VILCO_CHECK_ARG(qkv && G, "vilco_channel_attention: null pointer");   // y == NULL: only G = (k / 8)^T v is wanted
  VILCO_CHECK_ARG(H > 0 && C == H * CA_D, "vilco_channel_attention: head dim must be 64 (C=%d H=%d)", C, H);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((T + CA_TCHUNK - 1) / CA_TCHUNK, H, B);
  chan_attn_kv_kernel<<<dim3(1, H, B), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(qkv), qkv_lo, G, tlen, T, C, H, 1.0f / sqrtf((float)CA_D),
                                            act_fmt());
  VILCO_LAUNCH_CHECK();
  if (!y) return VILCO_OK;
  chan_attn_apply_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(qkv), qkv_lo, G, static_cast<__nv_bfloat16*>(y), y_lo, T, C, H,
                                               act_fmt());
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

extern "C" int vilco_channel_attention_bwd(const float* dy, const void* qkv, int64_t qkv_lo, const float* G, float* dA_scratch,
                                           float* dqkv, int B, int T, int C, int H, void* stream) {
  VILCO_CHECK_ARG(dy && qkv && G && dA_scratch && dqkv, "vilco_channel_attention_bwd: null pointer");
  VILCO_CHECK_ARG(H > 0 && C == H * CA_D, "vilco_channel_attention_bwd: head dim must be 64");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VILCO_CUDA(cudaMemsetAsync(dA_scratch, 0, sizeof(float) * (size_t)B * H * CA_D * CA_D, st));
  chan_attn_dA_kernel<<<dim3((T + CA_TCHUNK - 1) / CA_TCHUNK, H, B), 256, 0, st>>>(
      dy, static_cast<const __nv_bfloat16*>(qkv), qkv_lo, dA_scratch, T, C, H, act_fmt());
  VILCO_LAUNCH_CHECK();
  chan_attn_bwd_kernel<<<dim3((T + CB_T - 1) / CB_T, H, B), 256, 0, st>>>(
      dy, static_cast<const __nv_bfloat16*>(qkv), qkv_lo, G, dA_scratch, dqkv, T, C, H, 1.0f / sqrtf((float)CA_D), act_fmt());
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}
