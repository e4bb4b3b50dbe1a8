// Sliding-window attention (LocalMaskedMHCA core, reference MQ/libs/modeling/blocks.py:1129-1211 and the NLQ copy of it) on the
// warp-level tensor-core path.  The op is HBM-bound (q, k, v read once, out written once: 4 T C e bytes per clip), so the job of
// the kernel is to keep the arithmetic out of the way of the memory stream:
//
//   * one CTA = TQ queries of one (clip, head); the K / V rows [t0 - w, t0 + TQ + w) and the Q tile are staged with 16-byte
//     cp.async copies straight from the 16-bit planes (no fp32 expansion) into rows padded by 16 bytes, so that every ldmatrix
//     phase (8 rows x 16 bytes) touches 8 distinct bank groups;
//   * each warp owns 16 queries: S = Q K^T over its 16 + 2w key rows (2 KC n-tiles of 8 keys) with mma.sync m16n8k16, the
//     band / sequence / key-mask terms and the relative bias applied on the accumulator fragments, a quad-shuffle softmax, and
//     O = P V with the probabilities re-used from the accumulator registers as A fragments (no shared-memory round trip);
//   * two-plane operands (exact modes) add the hi*lo and lo*hi products, and split P the same way;
//   * the output tile goes back through the warp's own Q rows in shared memory and is written with 16-byte coalesced stores.
//
// The scalar kernel in attention.cu remains for head dims / windows outside the template set.
#include "common.cuh"
#include <cuda_fp16.h>

namespace vilco {

struct LtParams {
  const uint16_t *q, *k, *v; const float* mask; const float* rel_pe; uint16_t* out;
  long long lo;
  int B, T, C, H, W; float scale; int fmt;
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <int FMT>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (FMT == VILCO_F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// packs without the fp16 saturation of common.cuh's pack16x2: probabilities are <= 1 and the outputs are convex combinations
// of fp16-representable values
template <int FMT>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (FMT == VILCO_F16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
  return pack_bf16x2(a, b);
}
template <int FMT>
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack2<FMT>(a, b);
  const float2 h = unpack16x2(hi, FMT);
  lo = pack2<FMT>(a - h.x, b - h.y);
}
__device__ __forceinline__ float lt_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// D = head dim, KC = 16-key chunks per warp (16 + 2w <= 16 KC), PLANES = 16-bit planes per operand, FMT = plane format
template <int D, int KC, int PLANES, int FMT>
__global__ void __launch_bounds__(256) local_attn_tc_kernel(const LtParams p) {
  extern __shared__ __align__(16) unsigned char lt_smem[];
  constexpr int PITCH = D + 8;                      // elements; row pitch 2 D + 16 bytes = odd multiple of 16
  constexpr int CH = D / 8;                         // 16-byte chunks per row
  constexpr int RW = 32 / CH;                       // rows one warp copies per pass (CH = 12: lanes 24..31 idle)
  constexpr float LOG2E = 1.4426950408889634f;
  const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int TQ = nwarp * 16, NKR = TQ - 16 + 16 * KC;   // staged key / value rows (>= TQ + 2w)
  const int w = p.W / 2;
  const int t0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  uint16_t* sq = reinterpret_cast<uint16_t*>(lt_smem);            // [PLANES][TQ][PITCH]
  uint16_t* sk = sq + (size_t)PLANES * TQ * PITCH;                 // [PLANES][NKR][PITCH]
  uint16_t* sv = sk + (size_t)PLANES * NKR * PITCH;
  float* smadd = reinterpret_cast<float*>(sv + (size_t)PLANES * NKR * PITCH);   // [NKR] additive key term * log2 e (-inf outside the clip)
  float* srel = smadd + NKR;                                                    // [W] relative bias * log2 e
  const long long base = (long long)b * p.T * p.C + (long long)h * D;
  const float* mk = p.mask + (long long)b * p.T;
  const uint32_t sq_u = (uint32_t)__cvta_generic_to_shared(sq), sk_u = (uint32_t)__cvta_generic_to_shared(sk),
                 sv_u = (uint32_t)__cvta_generic_to_shared(sv);
  const uint32_t q_plane = 2u * TQ * PITCH, k_plane = 2u * NKR * PITCH;
  // ---- stage Q, K, V: each warp copies RW rows per pass, one 16-byte cp.async per lane; rows outside the clip are zero-filled
  //      (0 * garbage must not reach the accumulators).  Addresses advance by increments: the kernel is issue-bound. ----
  {
    const int lr = lane / CH, lc = lane - lr * CH;
    const bool act = lr < RW;
    const int rstep = nwarp * RW;
    const long long sstep = (long long)rstep * p.C;
    const uint32_t dstep = 2u * rstep * PITCH;
    int r = warp * RW + lr;
    const long long off0 = base + (long long)(t0 + r) * p.C + lc * 8;
    const uint32_t doff = 2u * (r * PITCH + lc * 8);
    {
      const uint16_t* src = p.q + off0;
      uint32_t dst = sq_u + doff;
#pragma unroll 1
      for (int rr = r; rr < TQ; rr += rstep, src += sstep, dst += dstep) {
        if (!act) continue;
        if (t0 + rr < p.T) {
          cp_async16(dst, src);
          if (PLANES == 2) cp_async16(dst + q_plane, src + p.lo);
        } else {
          asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0) : "memory");
          if (PLANES == 2) asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dst + q_plane), "r"(0) : "memory");
        }
      }
    }
    {
      const long long off = off0 - (long long)w * p.C;
      const uint16_t *srck = p.k + off, *srcv = p.v + off;
      uint32_t dk = sk_u + doff, dv = sv_u + doff;
      const int rlim = TQ + 2 * w;
#pragma unroll 1
      for (int rr = r; rr < NKR; rr += rstep, srck += sstep, srcv += sstep, dk += dstep, dv += dstep) {
        if (!act) continue;
        const int t = t0 - w + rr;
        if (t >= 0 && t < p.T && rr < rlim) {
          cp_async16(dk, srck);
          cp_async16(dv, srcv);
          if (PLANES == 2) { cp_async16(dk + k_plane, srck + p.lo); cp_async16(dv + k_plane, srcv + p.lo); }
        } else {
          asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dk), "r"(0) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dv), "r"(0) : "memory");
          if (PLANES == 2) {
            asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dk + k_plane), "r"(0) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dv + k_plane), "r"(0) : "memory");
          }
        }
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int r = threadIdx.x; r < NKR; r += blockDim.x) {
    const int t = t0 - w + r;
    smadd[r] = (t >= 0 && t < p.T) ? (mk[t] == 0.f ? -1e4f * LOG2E : 0.f) : -INFINITY;
  }
  const bool has_rel = p.rel_pe != nullptr;
  if (has_rel)
    for (int j = threadIdx.x; j < p.W; j += blockDim.x) srel[j] = p.rel_pe[h * p.W + j] * LOG2E;
  const int g = lane >> 2, tq = lane & 3;
  const int kr0 = warp * 16;                         // first key row of this warp's window (clip time t0 - w + kr0)
  float inv[2];                                      // 0 for padded queries (blocks.py:1192-1194) and rows past the clip
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int t = t0 + kr0 + g + 8 * r;
    inv[r] = (t < p.T && mk[t] != 0.f) ? 1.f : 0.f;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- S = Q K^T ----
  float S[2 * KC][4];
#pragma unroll
  for (int i = 0; i < 2 * KC; ++i) { S[i][0] = S[i][1] = S[i][2] = S[i][3] = 0.f; }
  const uint32_t sq_a = sq_u + 2u * ((kr0 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 8);
  const uint32_t sk_a = sk_u + 2u * ((kr0 + (lane & 7) + (lane >> 4) * 8) * PITCH + ((lane >> 3) & 1) * 8);
  const uint32_t sv_a = sv_u + 2u * ((kr0 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 8);
#pragma unroll
  for (int kc = 0; kc < D / 16; ++kc) {
    uint32_t a[PLANES][4];
#pragma unroll
    for (int pl = 0; pl < PLANES; ++pl) ldsm_x4(sq_a + pl * q_plane + kc * 32, a[pl]);
#pragma unroll
    for (int np = 0; np < KC; ++np) {
      uint32_t bk[PLANES][4];
#pragma unroll
      for (int pl = 0; pl < PLANES; ++pl) ldsm_x4(sk_a + pl * k_plane + 2u * (np * 16 * PITCH) + kc * 32, bk[pl]);
      mma16816<FMT>(S[2 * np], a[0], bk[0][0], bk[0][1]);
      mma16816<FMT>(S[2 * np + 1], a[0], bk[0][2], bk[0][3]);
      if (PLANES == 2) {
        mma16816<FMT>(S[2 * np], a[0], bk[PLANES - 1][0], bk[PLANES - 1][1]);
        mma16816<FMT>(S[2 * np + 1], a[0], bk[PLANES - 1][2], bk[PLANES - 1][3]);
        mma16816<FMT>(S[2 * np], a[PLANES - 1], bk[0][0], bk[0][1]);
        mma16816<FMT>(S[2 * np + 1], a[PLANES - 1], bk[0][2], bk[0][3]);
      }
    }
  }
  // ---- band / clip / key-mask terms, relative bias, softmax (base 2) over the W window slots of each query row ----
  // element e of S[nt]: query row i = g + 8 (e >> 1), key row jn = 8 nt + 2 tq + (e & 1); window slot jj = jn - i
  const float sc2 = p.scale * LOG2E;
  const int dj = 2 * tq - g;
  const unsigned Wu = (unsigned)p.W;
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int nt = 0; nt < 2 * KC; ++nt) {
    const float2 ma = *reinterpret_cast<const float2*>(smadd + kr0 + 8 * nt + 2 * tq);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int jj = dj + (8 * nt + (e & 1) - 8 * (e >> 1));
      const bool valid = (unsigned)jj < Wu;
      float s = fmaf(S[nt][e], sc2, (e & 1) ? ma.y : ma.x);
      if (has_rel && valid) s += srel[jj];
      s = valid ? s : -INFINITY;
      S[nt][e] = s;
      mx[e >> 1] = fmaxf(mx[e >> 1], s);
    }
  }
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    if (mx[r] == -INFINITY) mx[r] = 0.f;             // rows past the end of the clip (never stored)
  }
#pragma unroll
  for (int nt = 0; nt < 2 * KC; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float pe = lt_ex2(S[nt][e] - mx[e >> 1]);          // ex2(-inf) = 0
      S[nt][e] = pe;
      sum[e >> 1] += pe;
    }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
    sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
    inv[r] = inv[r] / sum[r];                        // sum >= 1 for every stored row (its own key is in the window)
  }
  // ---- O = P V (normalised P from the accumulator registers) ----
  float O[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { O[i][0] = O[i][1] = O[i][2] = O[i][3] = 0.f; }
#pragma unroll
  for (int c = 0; c < KC; ++c) {
    uint32_t ph[4], pl_[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float x = S[2 * c + (u >> 1)][2 * (u & 1)] * inv[u & 1], y = S[2 * c + (u >> 1)][2 * (u & 1) + 1] * inv[u & 1];
      if (PLANES == 2) split2<FMT>(x, y, ph[u], pl_[u]);
      else ph[u] = pack2<FMT>(x, y);
    }
#pragma unroll
    for (int nd = 0; nd < D / 16; ++nd) {
      uint32_t bv[PLANES][4];
#pragma unroll
      for (int pl = 0; pl < PLANES; ++pl) ldsm_x4_t(sv_a + pl * k_plane + 2u * (c * 16 * PITCH) + nd * 32, bv[pl]);
      mma16816<FMT>(O[2 * nd], ph, bv[0][0], bv[0][1]);
      mma16816<FMT>(O[2 * nd + 1], ph, bv[0][2], bv[0][3]);
      if (PLANES == 2) {
        mma16816<FMT>(O[2 * nd], pl_, bv[0][0], bv[0][1]);
        mma16816<FMT>(O[2 * nd + 1], pl_, bv[0][2], bv[0][3]);
        mma16816<FMT>(O[2 * nd], ph, bv[PLANES - 1][0], bv[PLANES - 1][1]);
        mma16816<FMT>(O[2 * nd + 1], ph, bv[PLANES - 1][2], bv[PLANES - 1][3]);
      }
    }
  }
  // ---- stage the output tile through this warp's own Q rows, then 16-byte coalesced stores ----
  __syncwarp();                                       // every lane is done reading the Q rows
  {
    const uint32_t d0 = sq_u + 2u * ((kr0 + g) * PITCH + 2 * tq);
#pragma unroll
    for (int nd = 0; nd < D / 8; ++nd)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const uint32_t dst = d0 + 2u * (8 * r * PITCH + nd * 8);
        uint32_t hi, lo;
        if (PLANES == 2) {
          split2<FMT>(O[nd][2 * r], O[nd][2 * r + 1], hi, lo);
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + q_plane), "r"(lo) : "memory");
        } else {
          hi = pack2<FMT>(O[nd][2 * r], O[nd][2 * r + 1]);
        }
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(hi) : "memory");
      }
  }
  __syncwarp();
  {
    const int lr = lane / CH, lc = lane - lr * CH;
    if (lr < RW) {
      uint32_t src = sq_u + 2u * ((kr0 + lr) * PITCH + lc * 8);
      uint16_t* dst = p.out + base + (long long)(t0 + kr0 + lr) * p.C + lc * 8;
#pragma unroll 1
      for (int r = lr; r < 16; r += RW, src += 2u * RW * PITCH, dst += (long long)RW * p.C) {
        if (t0 + kr0 + r >= p.T) break;
        uint32_t x0, x1, x2, x3;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(src));
        *reinterpret_cast<uint4*>(dst) = make_uint4(x0, x1, x2, x3);
        if (PLANES == 2) {
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(src + q_plane));
          *reinterpret_cast<uint4*>(dst + p.lo) = make_uint4(x0, x1, x2, x3);
        }
      }
    }
  }
}

template <int D, int KC, int PLANES, int FMT>
static int launch_lt(const LtParams& p, cudaStream_t st) {
  // queries per CTA: 64 keeps 6+ CTAs resident per SM (the loads of one overlap the math of the others)
  static const int tq_env = [] { const char* e = getenv("VILCO_LT_TQ"); return e ? atoi(e) : 64; }();   // tuning knob: 16..128
  const int tq_max = tq_env >= 16 && tq_env <= 128 ? tq_env / 16 * 16 : 64;
  const int TQ = p.T >= tq_max ? tq_max : (p.T + 15) / 16 * 16;
  const int NKR = TQ - 16 + 16 * KC;
  const size_t smem = (size_t)PLANES * (TQ + 2 * NKR) * (D + 8) * 2 + (size_t)(NKR + p.W) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    VILCO_CUDA(cudaFuncSetAttribute(local_attn_tc_kernel<D, KC, PLANES, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((p.T + TQ - 1) / TQ, p.H, p.B);
  local_attn_tc_kernel<D, KC, PLANES, FMT><<<grid, TQ * 2, smem, st>>>(p);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

template <int D, int KC>
static int launch_lt_pf(const LtParams& p, cudaStream_t st) {
  if (p.fmt == VILCO_F16) return p.lo ? launch_lt<D, KC, 2, VILCO_F16>(p, st) : launch_lt<D, KC, 1, VILCO_F16>(p, st);
  return p.lo ? launch_lt<D, KC, 2, VILCO_BF16>(p, st) : launch_lt<D, KC, 1, VILCO_BF16>(p, st);
}

// returns VILCO_OK when launched, -1 when the shape is outside the template set (the caller falls back to the scalar kernel)
int local_attn_tc(const void* q, const void* k, const void* v, const float* mask, const float* rel_pe, void* out, long long lo,
                  int B, int T, int C, int H, int W, float scale, int fmt, void* stream) {
  const int d = C / H, w = W / 2;
  if (C % 8 || (lo & 7) || w > 16) return -1;
  if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) return -1;
  LtParams p{static_cast<const uint16_t*>(q), static_cast<const uint16_t*>(k), static_cast<const uint16_t*>(v), mask, rel_pe,
             static_cast<uint16_t*>(out), lo, B, T, C, H, W, scale, fmt};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool big = w > 8;      // 16 + 2w keys per warp: <= 32 -> KC 2, <= 48 -> KC 3
#define LT_CASE(DD) case DD: return big ? launch_lt_pf<DD, 3>(p, st) : launch_lt_pf<DD, 2>(p, st);
  switch (d) {
    LT_CASE(16) LT_CASE(32) LT_CASE(64) LT_CASE(96) LT_CASE(128)
    default: return -1;
  }
#undef LT_CASE
}

}  // namespace vilco
