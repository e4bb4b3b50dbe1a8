// Fused XLNet relative attention (XLNetRelativeAttention.rel_attn_core + rel_shift_bnij, MQ/libs/modeling/modeling_xlnet_x.py:256-320):
//
//   score[i, j] = ((q_i + r_w) . k_j  +  (q_i + r_r) . k_r[T + j - i]) / sqrt(d)   - 1e30 where key j is padding and i != j
//   O[i]        = softmax_j(score[i, :]) @ V
//
// One CTA = 128 query rows of one (clip, head); keys are walked in tiles of 128.  Neither the (T x T) content scores, nor the
// (T x 2T) position scores, nor the probabilities ever exist in HBM (the materialised chain wrote and re-read ~10 GB per
// 32-clip step).  Per key tile the tensor core forms, into TMEM,
//     S  = Qw K^T          (128 x 128)
//     BD = Qr Kr[p0:p0+256]^T (128 x 256),  p0 = T + j0 - i0 - 127, so that  bd[i, j] = BD[i - i0, (j - j0) - (i - i0) + 127]
// — the relative shift is a per-ROW column offset, which no tcgen05.ld shape expresses (a load reads the same columns for all
// lanes).  Each softmax thread owns one row: it fetches the 64-column window that covers its 32 wanted columns for every lane
// of its warp, parks it in a private row of shared memory (128-bit stores, conflict-free at a pitch of 68 floats) and reads it
// back at its own offset (31 - lane) with conflict-free scalar loads — a shared-memory barrel shifter, 48 instructions per 32
// scores instead of ~190 register selects.  Online softmax in the log2 domain with a LAZY reference maximum: the running
// output in TMEM is only rescaled when a row's maximum grows by more than 2^8 (tcgen05.ld / st of O, rare after the first
// tiles); P goes to the P V MMA as fp16 / bf16 in 128B-swizzled shared memory.
//
// warp 0: TMA producer, warp 1: TMEM allocator + single-thread tcgen05.mma issuer, warps 2..9: softmax (TMEM lane quadrant =
// warp % 4, two warps per quadrant split the 128 key columns).  Single-plane operands only (the split-operand modes keep the
// materialised chain).
#include "tc_common.cuh"

namespace vilco {

static constexpr int XL_THREADS = 320;
static constexpr int XL_BQ = 128, XL_BKV = 128, XL_D = 64;
static constexpr int XL_MAX_T = 2048;
static constexpr int XL_PITCH = 68;                      // floats per lane row of the skew buffer
static constexpr int XL_TILE = XL_BQ * XL_D * 2;         // 16 KB: 128 rows x 128 bytes
static constexpr int XL_SKEW_BYTES = 8 * 32 * XL_PITCH * 4;
static constexpr float XL_LAZY = 8.0f;                   // rescale O only when the row maximum grows by more than 2^8

struct XlDev {
  int q_slot_row, q_slot_z1, q_slot_z2;      // qw and qr share strides
  int k_slot_row, k_slot_z1, k_slot_z2;      // k and v share strides
  int r_slot_row, r_slot_z1, r_slot_z2;      // k_r (2T, C): no batch dim
  int T, H;
  float scale;
  const float* kmask;                        // (B, T) 1 = valid key, or null
  uint16_t* O; long long o_ld, o_sh, o_sb;   // element strides: row, head, batch
  float* lse2;                               // (B, H, T) row log-sum-exp in base 2, or null
  int fmt;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// REL = true: XLNet relative attention as described above.  REL = false: plain masked self-attention (MaskedMHCA core,
// blocks.py:351-410) with the same single-pass structure — no position branch, padded keys invisible to every query, 256 TMEM
// columns and ~83 KB of shared memory, so two CTAs share an SM and one CTA's softmax overlaps the other's MMAs.
template <bool REL>
__global__ void __launch_bounds__(XL_THREADS, 1)
xl_attn_kernel(const __grid_constant__ CUtensorMap tmQw, const __grid_constant__ CUtensorMap tmQr,
               const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               const __grid_constant__ CUtensorMap tmR, const __grid_constant__ XlDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQw = smem;
  uint8_t* sQr = sQw + XL_TILE;                    // (REL only)
  uint8_t* sK = sQr + (REL ? XL_TILE : 0);
  uint8_t* sR = sK + XL_TILE;                      // 256 k_r rows: 32 KB (REL only)
  uint8_t* sV = sR + (REL ? 2 * XL_TILE : 0);
  uint8_t* sP = sV + XL_TILE;                      // 128 rows x 128 keys = two 64-key K-major blocks
  float* s_skew = reinterpret_cast<float*>(sP + 2 * XL_TILE);
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(s_skew) + (REL ? XL_SKEW_BYTES : 0));   // key validity bits
  float* s_x = reinterpret_cast<float*>(s_bits + XL_MAX_T / 32);     // [2 tile parities][2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_x + 2 * 2 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t qfull = smem_u32(bars + 0);
  const uint32_t kfull = smem_u32(bars + 1), kempty = smem_u32(bars + 2);
  const uint32_t vfull = smem_u32(bars + 3), vempty = smem_u32(bars + 4);
  const uint32_t sfull = smem_u32(bars + 5), sempty = smem_u32(bars + 6);
  const uint32_t pfull = smem_u32(bars + 7), pempty = smem_u32(bars + 8), ofull = smem_u32(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * XL_BQ, h = blockIdx.y, b = blockIdx.z;
  const int nkv = p.T / XL_BKV;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQw) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQr) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
    mbar_init(qfull, 1);
    mbar_init(kfull, 1); mbar_init(kempty, 1);
    mbar_init(vfull, 1); mbar_init(vempty, 1);
    mbar_init(sfull, 1); mbar_init(sempty, 8);
    mbar_init(pfull, 8); mbar_init(pempty, 1); mbar_init(ofull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(REL ? 512 : 256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {  // key validity bits of this clip
    for (int j0 = (warp - 2) * 32; j0 < p.T; j0 += 8 * 32) {
      const int j = j0 + lane;
      const bool ok = !p.kmask || p.kmask[(long long)b * p.T + j] != 0.f;
      const uint32_t w = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) s_bits[j0 >> 5] = w;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tBD = tmem_base + 128, tO = tmem_base + (REL ? 384 : 128);

  if (warp == 0) {
    if (lane == 0) {
      int c[4];
      mbar_expect_tx(qfull, (REL ? 2 : 1) * XL_TILE);
      c[0] = 0; c[p.q_slot_row] = q0; c[p.q_slot_z1] = h; c[p.q_slot_z2] = b;
      tma_load_5d(smem_u32(sQw), &tmQw, qfull, c[0], c[1], c[2], c[3], 0);
      if (REL) tma_load_5d(smem_u32(sQr), &tmQr, qfull, c[0], c[1], c[2], c[3], 0);
      for (int j = 0; j < nkv; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait(kempty, ph ^ 1);
        mbar_expect_tx(kfull, (REL ? 3 : 1) * XL_TILE);
        c[0] = 0; c[p.k_slot_row] = j * XL_BKV; c[p.k_slot_z1] = h; c[p.k_slot_z2] = b;
        tma_load_5d(smem_u32(sK), &tmK, kfull, c[0], c[1], c[2], c[3], 0);
        if (REL) {
          int r[4];
          r[0] = 0; r[p.r_slot_row] = p.T + j * XL_BKV - q0 - (XL_BQ - 1); r[p.r_slot_z1] = h; r[p.r_slot_z2] = 0;
          tma_load_5d(smem_u32(sR), &tmR, kfull, r[0], r[1], r[2], r[3], 0);     // one 256-row box (row 2T reads as zero)
        }
        mbar_wait(vempty, ph ^ 1);
        mbar_expect_tx(vfull, XL_TILE);
        tma_load_5d(smem_u32(sV), &tmV, vfull, c[0], c[1], c[2], c[3], 0);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = make_idesc(XL_BKV, 0, 0, p.fmt, p.fmt);     // N = 128 keys, B K-major
    const uint32_t idesc_r = make_idesc(256, 0, 0, p.fmt, p.fmt);        // N = 256 relative positions
    const uint32_t idesc_o = make_idesc(XL_D, 1, 0, p.fmt, p.fmt);       // N = 64, B (V) MN-major
    mbar_wait(qfull, 0);
    auto issue_pv = [&](int j) {
      mbar_wait(pfull, j & 1);
      mbar_wait(vfull, j & 1);
      tcgen05_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int ks = 0; ks < XL_BKV / UMMA_K; ++ks) {
          const uint32_t aoff = (ks >> 2) * XL_TILE + (ks & 3) * 32;  // 64-key block, then 16-key step inside the atom
          const uint32_t boff = ks * UMMA_K * 128;                    // 16 key rows of 128 bytes
          tcgen05_mma_f16(tO, make_smem_desc(smem_u32(sP) + aoff, 16, 1024), make_smem_desc(smem_u32(sV) + boff, 16, 1024),
                          idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        }
        tcgen05_commit(pempty);
        tcgen05_commit(vempty);
      }
      __syncwarp();
    };
    for (int j = 0; j < nkv; ++j) {
      const uint32_t ph = j & 1;
      mbar_wait(kfull, ph);
      mbar_wait(sempty, ph ^ 1);        // the softmax warps hold tile j-1's scores in registers
      tcgen05_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < XL_D / UMMA_K; ++k)
          tcgen05_mma_f16(tS, make_smem_desc(smem_u32(sQw) + k * 32, 16, 1024), make_smem_desc(smem_u32(sK) + k * 32, 16, 1024),
                          idesc_s, k > 0 ? 1u : 0u);
        if (REL) {
#pragma unroll
          for (int k = 0; k < XL_D / UMMA_K; ++k)
            tcgen05_mma_f16(tBD, make_smem_desc(smem_u32(sQr) + k * 32, 16, 1024), make_smem_desc(smem_u32(sR) + k * 32, 16, 1024),
                            idesc_r, k > 0 ? 1u : 0u);
        }
        tcgen05_commit(kempty);
        tcgen05_commit(sfull);
      }
      __syncwarp();
      if (j > 0) issue_pv(j - 1);
    }
    issue_pv(nkv - 1);
    if (lane == 0) tcgen05_commit(ofull);
    __syncwarp();
  } else {
    // ===== 8 softmax warps: TMEM lane quadrant q = warp % 4 (one thread per query row), column half = (warp - 2) / 4 =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int gi = q0 + row;                           // global query index
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const float sc2 = p.scale * 1.4426950408889634f;
    float* my_skew = s_skew + ((warp - 2) * 32 + lane) * XL_PITCH;
    const int sh = 31 - lane;
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const uint32_t ph = j & 1;
      mbar_wait(sfull, ph);
      tcgen05_fence_after();
      float x[64];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = half * 64 + cc * 32;
        uint32_t r[32];
        // position scores: the 64-column window [base, base + 64) covers columns c + 127 - row for every lane of the warp
        const int base = c0 + 96 - 32 * q;
        __syncwarp();
        if (REL) {
          tmem_ld32(tBD + lane_addr + base, r);
#pragma unroll
          for (int t = 0; t < 8; ++t)
            *reinterpret_cast<uint4*>(my_skew + 4 * t) = make_uint4(r[4 * t], r[4 * t + 1], r[4 * t + 2], r[4 * t + 3]);
          tmem_ld32(tBD + lane_addr + base + 32, r);
#pragma unroll
          for (int t = 0; t < 8; ++t)
            *reinterpret_cast<uint4*>(my_skew + 32 + 4 * t) = make_uint4(r[4 * t], r[4 * t + 1], r[4 * t + 2], r[4 * t + 3]);
        }
        tmem_ld32(tS + lane_addr + c0, r);             // content scores
        const uint32_t bits = s_bits[(j * XL_BKV + c0) >> 5];
        const int jj0 = j * XL_BKV + c0;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (REL) {
            const float v = (__uint_as_float(r[c]) + my_skew[sh + c]) * sc2;
            const bool ok = ((bits >> c) & 1u) || (jj0 + c == gi);   // a padded key is only visible to itself
            x[cc * 32 + c] = ok ? v : -INFINITY;
          } else {
            x[cc * 32 + c] = ((bits >> c) & 1u) ? __uint_as_float(r[c]) * sc2 : -INFINITY;   // padded keys: masked_fill(-inf)
          }
        }
      }
      // the scores of this tile live in registers: the tensor core may overwrite S / BD with the next tile
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sempty);
      float cmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; ++c) cmax = fmaxf(cmax, x[c]);
      // both column halves of a row must use the same reference maximum
      float* xs = s_x + (ph * 2) * 128;
      xs[half * 128 + row] = cmax;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      const float tmax = fmaxf(cmax, xs[(half ^ 1) * 128 + row]);
      float f = 1.0f;
      const bool grow = tmax > m_ref + XL_LAZY;          // also true for the first tile (m_ref = -inf, tmax finite)
      if (grow) {
        f = ex2f(m_ref - tmax);                          // 0 on the first tile
        l *= f;
        m_ref = tmax;
      }
      mbar_wait(pempty, ph ^ 1);                         // P V of tile j-1 has retired: P is free and O is up to date
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        tcgen05_fence_after();
        uint32_t o[32];
        tmem_ld32(tO + lane_addr + half * 32, o);
#pragma unroll
        for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * f);
        tmem_st32(tO + lane_addr + half * 32, o);
      }
      const float mr = m_ref == -INFINITY ? 0.f : m_ref;  // (a row that has not seen a visible key yet: every x is -inf -> P = 0)
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {                   // 8 keys = one 16-byte chunk of the row's 128-byte line
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float a = ex2f(x[ch * 8 + 2 * u] - mr), bq = ex2f(x[ch * 8 + 2 * u + 1] - mr);
          sum0 += a; sum1 += bq;
          w[u] = pack16x2(a, bq, p.fmt);
        }
        uint8_t* dst = sP + half * XL_TILE + row * 128 + ((ch ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      l += sum0 + sum1;
      tcgen05_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> tensor core reads
      __syncwarp();
      if (lane == 0) mbar_arrive(pfull);
    }
    // ===== output: O / l as one 16-bit plane; 32 columns per warp =====
    mbar_wait(ofull, 0);
    tcgen05_fence_after();
    float* xs = s_x;                                     // (all tile exchanges are behind the bar.sync of the last tile)
    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
    xs[half * 128 + row] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
    const float ltot = l + xs[(half ^ 1) * 128 + row];
    const float inv = 1.0f / ltot;
    if (p.lse2 && half == 0) p.lse2[((long long)b * p.H + h) * p.T + gi] = m_ref + log2f(ltot);
    uint32_t o[32];
    __syncwarp();
    tmem_ld32(tO + lane_addr + half * 32, o);
    uint16_t* orow = p.O + (long long)b * p.o_sb + (long long)h * p.o_sh + (long long)gi * p.o_ld + half * 32;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        w[u] = pack16x2(__uint_as_float(o[ch * 8 + 2 * u]) * inv, __uint_as_float(o[ch * 8 + 2 * u + 1]) * inv, p.fmt);
      *reinterpret_cast<uint4*>(orow + ch * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(REL ? 512 : 256) : "memory");
  }
}

}  // namespace vilco

using namespace vilco;

extern "C" int vilco_xl_attention(const void* qw, const void* qr, const void* k, const void* v, const void* kr, const float* kmask,
                                  void* out, int B, int H, int T, int C, float scale, void* stream) {
  VILCO_CHECK_ARG(qw && qr && k && v && kr && out, "vilco_xl_attention: null pointer");
  VILCO_CHECK_ARG(H > 0 && C == H * XL_D, "vilco_xl_attention: head dim must be 64 (C=%d, H=%d)", C, H);
  VILCO_CHECK_ARG(T >= XL_BQ && T % XL_BQ == 0 && T <= XL_MAX_T, "vilco_xl_attention: T=%d must be a multiple of %d, <= %d", T,
                  XL_BQ, XL_MAX_T);
  VILCO_CHECK_ARG(reinterpret_cast<uintptr_t>(out) % 16 == 0, "vilco_xl_attention: out alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tmQw, tmQr, tmK, tmV, tmR;
  int sq[3], sk[3], sr[3], tmp[3];
  int rc = encode_map(&tmQw, qw, XL_D, (uint64_t)T, C, (uint64_t)H, XL_D, (uint64_t)B, (int64_t)T * C, 0, XL_D, XL_BQ, sq);
  if (rc) return rc;
  rc = encode_map(&tmQr, qr, XL_D, (uint64_t)T, C, (uint64_t)H, XL_D, (uint64_t)B, (int64_t)T * C, 0, XL_D, XL_BQ, tmp);
  if (rc) return rc;
  rc = encode_map(&tmK, k, XL_D, (uint64_t)T, C, (uint64_t)H, XL_D, (uint64_t)B, (int64_t)T * C, 0, XL_D, XL_BKV, sk);
  if (rc) return rc;
  rc = encode_map(&tmV, v, XL_D, (uint64_t)T, C, (uint64_t)H, XL_D, (uint64_t)B, (int64_t)T * C, 0, XL_D, XL_BKV, tmp);
  if (rc) return rc;
  rc = encode_map(&tmR, kr, XL_D, (uint64_t)2 * T, C, (uint64_t)H, XL_D, 1, 0, 0, XL_D, 256, sr);
  if (rc) return rc;
  XlDev p{};
  p.q_slot_row = sq[0]; p.q_slot_z1 = sq[1]; p.q_slot_z2 = sq[2];
  p.k_slot_row = sk[0]; p.k_slot_z1 = sk[1]; p.k_slot_z2 = sk[2];
  p.r_slot_row = sr[0]; p.r_slot_z1 = sr[1]; p.r_slot_z2 = sr[2];
  p.T = T; p.H = H; p.scale = scale; p.kmask = kmask;
  p.O = static_cast<uint16_t*>(out); p.o_ld = C; p.o_sh = XL_D; p.o_sb = (long long)T * C;
  p.fmt = act_fmt();
  const int smem = 8 * XL_TILE + XL_SKEW_BYTES + XL_MAX_T / 8 + 2 * 2 * 128 * 4 + 16 * 8 + 16 + 1024;
  static bool cfg = false;
  if (!cfg) { VILCO_CUDA(cudaFuncSetAttribute(xl_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); cfg = true; }
  dim3 grid(T / XL_BQ, H, B);
  xl_attn_kernel<true><<<grid, XL_THREADS, smem, st>>>(tmQw, tmQr, tmK, tmV, tmR, p);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

// Single-pass masked self-attention for single-plane operands (see include/vilco_b200.h)
extern "C" int vilco_self_attention_lse(const void* q, const void* k, const void* v, const float* kmask, void* out, float* lse2, int B,
                                        int H, int T, int C, float scale, void* stream);
extern "C" int vilco_self_attention(const void* q, const void* k, const void* v, const float* kmask, void* out, int B, int H, int T,
                                    int C, float scale, void* stream) {
  return vilco_self_attention_lse(q, k, v, kmask, out, nullptr, B, H, T, C, scale, stream);
}

extern "C" int vilco_self_attention_lse(const void* q, const void* k, const void* v, const float* kmask, void* out, float* lse2, int B,
                                        int H, int T, int C, float scale, void* stream) {
  VILCO_CHECK_ARG(q && k && v && out, "vilco_self_attention: null pointer");
  VILCO_CHECK_ARG(H > 0 && C == H * XL_D, "vilco_self_attention: head dim must be 64 (C=%d, H=%d)", C, H);
  VILCO_CHECK_ARG(T >= XL_BQ && T % XL_BQ == 0 && T <= XL_MAX_T, "vilco_self_attention: T=%d must be a multiple of %d, <= %d", T,
                  XL_BQ, XL_MAX_T);
  VILCO_CHECK_ARG(reinterpret_cast<uintptr_t>(out) % 16 == 0, "vilco_self_attention: out alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tmQ, tmK, tmV;
  int sq[3], sk[3], tmp[3];
  int rc = encode_map(&tmQ, q, XL_D, (uint64_t)T, C, (uint64_t)H, XL_D, (uint64_t)B, (int64_t)T * C, 0, XL_D, XL_BQ, sq);
  if (rc) return rc;
  rc = encode_map(&tmK, k, XL_D, (uint64_t)T, C, (uint64_t)H, XL_D, (uint64_t)B, (int64_t)T * C, 0, XL_D, XL_BKV, sk);
  if (rc) return rc;
  rc = encode_map(&tmV, v, XL_D, (uint64_t)T, C, (uint64_t)H, XL_D, (uint64_t)B, (int64_t)T * C, 0, XL_D, XL_BKV, tmp);
  if (rc) return rc;
  XlDev p{};
  p.q_slot_row = sq[0]; p.q_slot_z1 = sq[1]; p.q_slot_z2 = sq[2];
  p.k_slot_row = sk[0]; p.k_slot_z1 = sk[1]; p.k_slot_z2 = sk[2];
  p.T = T; p.H = H; p.scale = scale; p.kmask = kmask;
  p.O = static_cast<uint16_t*>(out); p.o_ld = C; p.o_sh = XL_D; p.o_sb = (long long)T * C;
  p.lse2 = lse2;
  p.fmt = act_fmt();
  const int smem = 5 * XL_TILE + XL_MAX_T / 8 + 2 * 2 * 128 * 4 + 16 * 8 + 16 + 1024;
  static bool cfg = false;
  if (!cfg) { VILCO_CUDA(cudaFuncSetAttribute(xl_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); cfg = true; }
  dim3 grid(T / XL_BQ, H, B);
  xl_attn_kernel<false><<<grid, XL_THREADS, smem, st>>>(tmQ, tmQ, tmK, tmV, tmK, p);
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}
