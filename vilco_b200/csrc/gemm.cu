// Dense contraction kernel for every GEMM-shaped op on the Moment-Query path (see include/vilco_b200.h).
//
// tcgen05 path, persistent and warp-specialised.  One output tile is (128 * CG) x BN, CG = 1 or 2 CTAs (cta_group::2: the two
// CTAs of a cluster pair share one 256-row MMA; each loads its own 128 A rows and HALF of the B tile, so the operand bytes
// every SM pulls from L2 per MMA are halved).  Warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle), warp 1 = TMEM
// allocator + single-thread tcgen05.mma issuer (kind::f16: fp16 / bf16 operands, mixed per operand, fp32 accumulate in TMEM;
// leader CTA only when CG = 2), warps 2..9 = epilogue (tcgen05.ld 32x32b -> swizzled smem transpose -> fused bias / row-mask /
// activation / channel-scale / residual -> 16-byte coalesced stores).  A k=3 convolution is three row-shifted TMA loads of the
// same activation tile accumulating into the same TMEM tile; TMA out-of-bounds zero fill implements the conv zero padding and
// every M/N/K tail.  Operands come as 1 or 2 planes each (x ~= hi + lo): the MMAs issued per k-step are hi*hi, plus hi*lo when
// B has a lo plane, plus lo*hi when A has one.
#include <mutex>
#include "tc_common.cuh"
#include <type_traits>

namespace vilco {

static constexpr int NUM_THREADS = 320;  // 10 warps: TMA, MMA, 8 x epilogue
static constexpr int EPI_WARPS = 8;
static constexpr int EPI_SMEM = EPI_WARPS * 32 * 32 * 4;  // one swizzled 32x32 fp32 transpose buffer per epilogue warp
static constexpr int MAX_STAGES = 8;
static constexpr int BAR_BYTES = 1024;   // barriers + TMEM slot live in the first KB of the (1024-aligned) dynamic smem
static constexpr int SMEM_LIMIT = 232448;

struct GemmDev {
  // coordinate slots: the three outer tensor-map dims are sorted by stride on the host
  int a_slot_row, a_slot_z1, a_slot_z2;
  int b_slot_row, b_slot_z1, b_slot_z2;
  int M, N, K, taps, Z1, Ztot;
  int a_major;
  int band_lo, band_hi;  // when band_hi > band_lo: only elements with band_lo <= m + n < band_hi are needed; tiles outside are skipped
  int b_major, b_batched;
  int pa, pb, stages;    // operand planes (1 or 2) and smem ring depth
  int a_fmt, b_fmt;      // VILCO_BF16 / VILCO_F16 per operand
  void* D; int d_dtype; long long d_ld, d_s1, d_s2, d_lo;
  float alpha;
  const float* bias;
  const float* rowmul; long long rowmul_zs;
  int act;
  const float* colscale;
  const float* resid; int resid_masked;
  int vec_ok;  // 16-byte aligned vector stores allowed
  const float* rowsub; long long rowsub_s1, rowsub_s2;   // fused softmax-recompute epilogues (see VilcoGemm)
  long long colscale_zs;
  const uint16_t* emul;
};

// ---------------------------------------------------------------------------------------------
// epilogue math shared by both implementations
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float epi_value(const GemmDev& p, float acc, int n, float rm, float res) {
  float v = acc * p.alpha;
  if (p.bias) v += __ldg(p.bias + n);
  v *= rm;
  v = apply_act(v, p.act);
  if (p.colscale) v *= __ldg(p.colscale + n);
  return v + res;
}

// ---------------------------------------------------------------------------------------------
// tcgen05 kernel
// ---------------------------------------------------------------------------------------------
// Persistent: each CTA pair (CG = 2) or CTA (CG = 1) loops over output tiles t = id, id + n_workers, ...
//   warp 0      : TMA producer (ring of p.stages smem stages, full/empty mbarriers).  CG = 2: both CTAs load their own A rows
//                 and their half of B; all transaction bytes of a stage are signalled on the LEADER's full barrier.
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (leader CTA only); two accumulator stages in TMEM so the
//                 epilogue of tile i overlaps the main loop of tile i+1 (tmem_full / tmem_empty mbarriers; tmem_empty lives in
//                 the leader and counts the epilogue warps of both CTAs)
//   warps 2..9  : epilogue of this CTA's 128 rows
// GX = true: the instantiation that carries the fused attention-gradient epilogue terms (rowsub / emul / batched colscale);
// they are compiled out of the common kernels, whose epilogue sits at the register cap
template <int BN, int CG, bool GX>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmDev p) {
  constexpr int BNH = BN / CG;                      // B rows (output columns) this CTA stages
  constexpr int ACC_COLS = BN < 32 ? 32 : BN;       // TMEM columns of one accumulator stage
  constexpr int TMEM_COLS = 2 * ACC_COLS;           // power of two >= 64
  constexpr uint32_t A_BYTES = BM * BK * 2;         // 16 KB per plane
  constexpr uint32_t B_BYTES = BNH * BK * 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  float* stage_buf = reinterpret_cast<float*>(smem + BAR_BYTES);
  uint8_t* ring = smem + BAR_BYTES + EPI_SMEM;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int PA = p.pa, PB = p.pb, STAGES = p.stages;
  const uint32_t stage_bytes = PA * A_BYTES + PB * B_BYTES;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;

  const uint32_t full0 = smem_u32(bars);
  const uint32_t empty0 = smem_u32(bars + MAX_STAGES);
  const uint32_t tfull0 = smem_u32(bars + 2 * MAX_STAGES);       // [2]
  const uint32_t tempty0 = smem_u32(bars + 2 * MAX_STAGES + 2);  // [2]

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, EPI_WARPS * CG);  // one arrival per epilogue warp (of both CTAs of a pair)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kblocks = (p.K + BK - 1) / BK;
  const int iters = p.taps * kblocks;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int m_tiles = (p.M + BM * CG - 1) / (BM * CG);
  const int tiles_per_z = n_tiles * m_tiles;
  const int total_tiles = tiles_per_z * p.Ztot;
  const int worker = blockIdx.x / CG, n_workers = gridDim.x / CG;
  // band filter (XLNet relative scores: only bd_raw[i, p] with T <= i + p < 2T is ever read)
  auto tile_needed = [&](int m0, int n0) -> bool {
    if (p.band_hi <= p.band_lo) return true;
    return (m0 + n0 + (BM * CG - 1) + (BN - 1) >= p.band_lo) && (m0 + n0 < p.band_hi);
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int ca[4], cb[4];
      uint32_t it_g = 0;  // global k-iteration counter (continues across tiles)
      uint32_t s = 0, ph = 0;   // ring position of it_g
      for (int t = worker; t < total_tiles; t += n_workers) {
        const int z = t / tiles_per_z, r = t - z * tiles_per_z;
        const int mt0 = (r / n_tiles) * BM * CG, n0 = (r % n_tiles) * BN;
        const int z1 = z % p.Z1, z2 = z / p.Z1;
        if (!tile_needed(mt0, n0)) continue;
        const int m0 = mt0 + static_cast<int>(rank) * BM;     // this CTA's A rows
        const int nb0 = n0 + static_cast<int>(rank) * BNH;    // this CTA's share of the B tile
        ca[p.a_slot_z1] = z1; ca[p.a_slot_z2] = z2;
        cb[p.b_slot_z2] = p.b_batched ? z2 : 0;
        int tap = 0, kb = 0;
        for (int it = 0; it < iters; ++it, ++it_g) {
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t sa = smem_u32(ring + s * stage_bytes);
          const uint32_t sb = sa + A_BYTES * PA;
          uint32_t fb = full0 + 8 * s;
          if (CG == 2) {
            if (rank == 0) mbar_expect_tx(fb, stage_bytes * 2);   // the bytes of both CTAs land on the leader's barrier
            fb &= PEER_BIT_MASK;
          } else {
            mbar_expect_tx(fb, stage_bytes);
          }
          if (p.a_major == 0) {
            ca[0] = kb * BK;
            ca[p.a_slot_row] = m0 + tap - (p.taps >> 1);
          } else {
            ca[0] = m0;
            ca[p.a_slot_row] = kb * BK;
          }
          cb[p.b_slot_z1] = p.b_batched ? z1 : tap;
          if (p.b_major == 0) {
            cb[0] = kb * BK;
            cb[p.b_slot_row] = nb0;
          } else {
            cb[0] = nb0;
            cb[p.b_slot_row] = kb * BK;
          }
          auto load = [&](uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4) {
            if (CG == 2) tma_load_5d_2sm(dst, tm, fb, c0, c1, c2, c3, c4);
            else tma_load_5d(dst, tm, fb, c0, c1, c2, c3, c4);
          };
          for (int pl = 0; pl < PA; ++pl) {
            load(sa + pl * A_BYTES, &tmA, ca[0], ca[1], ca[2], ca[3], pl);
            if (p.a_major != 0)   // MN-major A: two (64 m x BK k) boxes, BK*128 bytes apart
              load(sa + pl * A_BYTES + BK * 128, &tmA, ca[0] + 64, ca[1], ca[2], ca[3], pl);
          }
          for (int pl = 0; pl < PB; ++pl) {
            if (BNH <= 64 || p.b_major == 0) {
              load(sb + pl * B_BYTES, &tmB, cb[0], cb[1], cb[2], cb[3], pl);
            } else {
              // MN-major B wider than one 128-byte swizzle atom: one (64 n x BK k) box per 64 columns, BK*128 bytes apart
#pragma unroll
              for (int h = 0; h < BNH / 64; ++h)
                load(sb + pl * B_BYTES + h * (BK * 128), &tmB, cb[0] + h * 64, cb[1], cb[2], cb[3], pl);
            }
          }
          if (++s == static_cast<uint32_t>(STAGES)) { s = 0; ph ^= 1; }
          if (++kb == kblocks) { kb = 0; ++tap; }
        }
      }
      // tail: wait until the MMAs have released every stage this CTA filled (no arrival may target an exited CTA)
      for (int k = 0; k < STAGES; ++k, ++it_g) {
        if (it_g < static_cast<uint32_t>(STAGES)) continue;   // stage never used
        mbar_wait(empty0 + 8 * (it_g % STAGES), ((it_g / STAGES) & 1) ^ 1);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane of the leader CTA) =====
    // The loop is latency-critical: it is ONE warp's instruction stream that has to stay ahead of the tensor pipe (a 256 x 256 x 16
    // MMA retires in ~128 cycles).  Everything below is warp-uniform (descriptors are built by adding a stage offset to a
    // constant, the ring position is a counter, the plane combinations are separate straight-line paths), so the compiler
    // keeps it in uniform registers; only the tcgen05 instructions themselves sit under the elected-lane predicate.
    if (rank == 0) {
      const uint32_t idesc = make_idesc(BN, p.b_major, p.a_major, p.a_fmt, p.b_fmt, BM * CG);
      // K-major: advance 16 elements = 32 bytes inside the swizzle atom; MN-major: 16 k-rows of 128 bytes (descriptor units of 16 B)
      const uint32_t a_k16 = (p.a_major == 0 ? UMMA_K * 2 : UMMA_K * 128) >> 4;
      const uint32_t b_k16 = (p.b_major == 0 ? UMMA_K * 2 : UMMA_K * 128) >> 4;
      const uint64_t a_desc0 = make_smem_desc(0, p.a_major == 0 ? 16 : BK * 128, 1024);
      const uint64_t b_desc0 = make_smem_desc(0, p.b_major == 0 ? 16 : BK * 128, 1024);   // MN-major: distance between 64-column swizzle atoms
      const uint32_t ring16 = (smem_u32(ring) & 0x3FFFFu) >> 4, stage16 = stage_bytes >> 4;
      const uint32_t a_pl16 = A_BYTES >> 4, b_pl16 = B_BYTES >> 4, b_off16 = (A_BYTES * PA) >> 4;
      const bool leader_lane = elect_one();
      uint32_t s = 0, ph = 0, tc = 0;
      for (int t = worker; t < total_tiles; t += n_workers) {
        {
          const int r = t % tiles_per_z;
          if (!tile_needed((r / n_tiles) * BM * CG, (r % n_tiles) * BN)) continue;
        }
        const uint32_t acc = tc & 1;
        const uint32_t tc_cur = tc++;
        mbar_wait(tempty0 + 8 * acc, ((tc_cur >> 1) & 1) ^ 1);  // epilogue has drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * ACC_COLS;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(full0 + 8 * s, ph);
          tcgen05_fence_after();
          const uint64_t ad = a_desc0 + (ring16 + s * stage16);
          const uint64_t bd = ad - a_desc0 + b_desc0 + b_off16;
          const uint32_t first = it > 0 ? 1u : 0u;
          if (leader_lane) {
            auto mma = [&](uint64_t x, uint64_t y, uint32_t accum) {
              if (CG == 2) tcgen05_mma_f16_2sm(tmem_d, x, y, idesc, accum);
              else tcgen05_mma_f16(tmem_d, x, y, idesc, accum);
            };
            if (PA == 1 && PB == 1) {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) mma(ad + k * a_k16, bd + k * b_k16, k ? 1u : first);
            } else {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                const uint64_t a_hi = ad + k * a_k16, b_hi = bd + k * b_k16;
                mma(a_hi, b_hi, k ? 1u : first);
                if (PB == 2) mma(a_hi, b_hi + b_pl16, 1u);
                if (PA == 2) mma(a_hi + a_pl16, b_hi, 1u);
              }
            }
            if (CG == 2) {
              tcgen05_commit_2sm(empty0 + 8 * s);                        // frees the stage in both CTAs when these MMAs retire
              if (it == iters - 1) tcgen05_commit_2sm(tfull0 + 8 * acc);  // accumulator complete (both CTAs' epilogues)
            } else {
              tcgen05_commit(empty0 + 8 * s);
              if (it == iters - 1) tcgen05_commit(tfull0 + 8 * acc);
            }
          }
          __syncwarp();
          if (++s == static_cast<uint32_t>(STAGES)) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===== 8 epilogue warps: TMEM lane quadrant = warp % 4; the two warps of a quadrant split the 32-column chunks =====
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;
    float* sbuf = stage_buf + (warp - 2) * (32 * 32);
    const float alpha = p.alpha;
    const float* __restrict__ bias = p.bias;
    const float* __restrict__ resid = p.resid;
    const int act = p.act;
    const bool resid_masked = p.resid_masked != 0;
    const long long d_ld = p.d_ld, d_lo = p.d_lo;
    const bool is_f32 = p.d_dtype == VILCO_F32;
    const int d_fmt = p.d_dtype;
    const bool has_rm = p.rowmul != nullptr;
    // 16-byte vector loads of the per-column terms need aligned pointers; a 16-bit output with a residual takes the general path
    const bool fast_ok = p.vec_ok && (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0) &&
                         (!p.colscale || ((reinterpret_cast<uintptr_t>(p.colscale) & 15) == 0 && (p.colscale_zs & 3) == 0)) && (is_f32 || resid == nullptr);
    uint32_t tc = 0;
    for (int t = worker; t < total_tiles; t += n_workers) {
      const int z = t / tiles_per_z, r = t - z * tiles_per_z;
      const int mt0 = (r / n_tiles) * BM * CG, n0 = (r % n_tiles) * BN;
      if (!tile_needed(mt0, n0)) continue;
      const int m0 = mt0 + static_cast<int>(rank) * BM;
      const int z1 = z % p.Z1, z2 = z / p.Z1;
      const uint32_t acc = tc & 1;
      const uint32_t tc_cur = tc++;
      mbar_wait(tfull0 + 8 * acc, (tc_cur >> 1) & 1);
      tcgen05_fence_after();
      const int mrow0 = m0 + q * 32;            // first row of this warp's 32-row slab
      float rm = 1.0f;                          // row multiplier of the accumulator row this lane owns
      if (p.rowmul && mrow0 + lane < p.M) rm = __ldg(p.rowmul + z2 * p.rowmul_zs + mrow0 + lane);
      float row_sub = 0.0f;                     // per-row subtrahend (row log-sum-exp / delta of the attention gradients)
      if (GX && p.rowsub && mrow0 + lane < p.M) row_sub = __ldg(p.rowsub + z1 * p.rowsub_s1 + z2 * p.rowsub_s2 + mrow0 + lane);
      const float* __restrict__ colscale = p.colscale ? p.colscale + (GX ? z2 * p.colscale_zs : 0) : nullptr;
      const long long zoff = z1 * p.d_s1 + z2 * p.d_s2;
      auto run_chunks = [&](auto act_c) {
        constexpr int ACT = decltype(act_c)::value;   // -1: runtime `act` (slow path only)
      for (int c = chalf; c < BN / 32; c += 2) {
          const int nb = n0 + c * 32;
          if (nb >= p.N || mrow0 >= p.M) break;  // warp-uniform
          // ---- fast path: whole 32 x 32 chunk in range, 16-byte accesses allowed.  The epilogue is one of the two pipelines
          // that bound a K = 1024 tile (8 warps x 4 chunks of it per 128 x 256 sub-tile), so this path has no bounds checks, no
          // per-element branches, shared-space ld / st and vector loads of the per-column terms issued ahead of the TMEM read ----
          if (fast_ok && (nb + 32 <= p.N) && (mrow0 + 32 <= p.M)) {
            const uint32_t sb_u = smem_u32(sbuf);
            if (is_f32) {
              const int rsub = lane >> 3, g = lane & 7;
              const int n = nb + 4 * g;
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = make_float4(1.f, 1.f, 1.f, 1.f);
              if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
              if (colscale) s4 = __ldg(reinterpret_cast<const float4*>(colscale + n));
              float* dptr = static_cast<float*>(p.D) + zoff + (long long)(mrow0 + rsub) * d_ld + n;
              float4 rs[8];
              if (resid) {
                const float* rptr = resid + zoff + (long long)(mrow0 + rsub) * d_ld + n;
#pragma unroll
                for (int i = 0; i < 8; ++i) rs[i] = __ldg(reinterpret_cast<const float4*>(rptr + (long long)(4 * i) * d_ld));
              }
              uint32_t rr[32];
              __syncwarp();
              tmem_ld32(tmem_base + acc * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16) + c * 32, rr);
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4)
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sb_u + 4u * (lane * 32 + ((j4 ^ (lane & 7)) << 2))),
                             "r"(rr[4 * j4]), "r"(rr[4 * j4 + 1]), "r"(rr[4 * j4 + 2]), "r"(rr[4 * j4 + 3]) : "memory");
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rloc = 4 * i + rsub;
                const float rmr = has_rm ? __shfl_sync(0xffffffffu, rm, rloc) : 1.0f;
                float4 a;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                             : "r"(sb_u + 4u * (rloc * 32 + ((g ^ (rloc & 7)) << 2))));
                a.x = apply_act((a.x * alpha + b4.x) * rmr, ACT) * s4.x;
                a.y = apply_act((a.y * alpha + b4.y) * rmr, ACT) * s4.y;
                a.z = apply_act((a.z * alpha + b4.z) * rmr, ACT) * s4.z;
                a.w = apply_act((a.w * alpha + b4.w) * rmr, ACT) * s4.w;
                if (resid) {
                  const float f = resid_masked ? rmr : 1.0f;
                  a.x += rs[i].x * f; a.y += rs[i].y * f; a.z += rs[i].z * f; a.w += rs[i].w * f;
                }
                *reinterpret_cast<float4*>(dptr + (long long)(4 * i) * d_ld) = a;
              }
            } else {
              const int rsub = lane >> 2, g2 = (lane & 3) * 2;
              const int n = nb + 4 * g2;
              float4 b8[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
              float4 s8[2] = {make_float4(1.f, 1.f, 1.f, 1.f), make_float4(1.f, 1.f, 1.f, 1.f)};
              if (bias) { b8[0] = __ldg(reinterpret_cast<const float4*>(bias + n)); b8[1] = __ldg(reinterpret_cast<const float4*>(bias + n + 4)); }
              if (colscale) { s8[0] = __ldg(reinterpret_cast<const float4*>(colscale + n)); s8[1] = __ldg(reinterpret_cast<const float4*>(colscale + n + 4)); }
              uint16_t* dptr = static_cast<uint16_t*>(p.D) + zoff + (long long)(mrow0 + rsub) * d_ld + n;
              uint32_t rr[32];
              __syncwarp();
              tmem_ld32(tmem_base + acc * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16) + c * 32, rr);
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4)
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sb_u + 4u * (lane * 32 + ((j4 ^ (lane & 7)) << 2))),
                             "r"(rr[4 * j4]), "r"(rr[4 * j4 + 1]), "r"(rr[4 * j4 + 2]), "r"(rr[4 * j4 + 3]) : "memory");
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rloc = 8 * i + rsub;
                const float rmr = has_rm ? __shfl_sync(0xffffffffu, rm, rloc) : 1.0f;
                const float rsr = (GX && p.rowsub) ? __shfl_sync(0xffffffffu, row_sub, rloc) : 0.0f;
                float v[8];
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
                             : "r"(sb_u + 4u * (rloc * 32 + ((g2 ^ (rloc & 7)) << 2))));
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                             : "r"(sb_u + 4u * (rloc * 32 + (((g2 + 1) ^ (rloc & 7)) << 2))));
                const float bb[8] = {b8[0].x, b8[0].y, b8[0].z, b8[0].w, b8[1].x, b8[1].y, b8[1].z, b8[1].w};
                const float ss[8] = {s8[0].x, s8[0].y, s8[0].z, s8[0].w, s8[1].x, s8[1].y, s8[1].z, s8[1].w};
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = apply_act((v[u] * alpha + bb[u] - rsr) * rmr, ACT) * ss[u];
                uint32_t h[4], l[4];
                uint16_t* o = dptr + (long long)(8 * i) * d_ld;
                if (GX && p.emul) {      // elementwise factor laid out like D (one 16-bit plane): dS = P * (dP - delta)
                  const uint4 e = __ldg(reinterpret_cast<const uint4*>(p.emul + (o - static_cast<uint16_t*>(p.D))));
                  const uint32_t ew[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    const float2 f = unpack16x2(ew[u], d_fmt);
                    v[2 * u] *= f.x; v[2 * u + 1] *= f.y;
                  }
                }
                if (d_lo) {
#pragma unroll
                  for (int u = 0; u < 4; ++u) split16x2(v[2 * u], v[2 * u + 1], d_fmt, h[u], l[u]);
                  *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
                  *reinterpret_cast<uint4*>(o + d_lo) = make_uint4(l[0], l[1], l[2], l[3]);
                } else {
#pragma unroll
                  for (int u = 0; u < 4; ++u) h[u] = pack16x2(v[2 * u], v[2 * u + 1], d_fmt);
                  *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
                }
              }
            }
            continue;
          }
          // general path (tails, unaligned pointers, 16-bit output with a residual)
          uint32_t rr[32];
          __syncwarp();
          tmem_ld32(tmem_base + acc * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16) + c * 32, rr);
          // stage 1: raw accumulator row -> XOR-swizzled 32x32 transpose buffer (conflict-free 128-bit accesses)
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            *reinterpret_cast<uint4*>(sbuf + lane * 32 + ((j4 ^ (lane & 7)) << 2)) =
                make_uint4(rr[4 * j4], rr[4 * j4 + 1], rr[4 * j4 + 2], rr[4 * j4 + 3]);
          __syncwarp();
          // stage 2: lanes own columns -> per-column vectors loaded once, wide fully coalesced stores
          if (is_f32) {
            // 8 lanes cover one 128-byte row segment, 4 rows per instruction
            const int rsub = lane >> 3, g = lane & 7;
            const int n = nb + 4 * g;
            const bool vec = p.vec_ok && (n + 3 < p.N);
            float b4[4] = {0.f, 0.f, 0.f, 0.f}, s4[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (n + u < p.N) {
                if (bias) b4[u] = __ldg(bias + n + u);
                if (colscale) s4[u] = __ldg(colscale + n + u);
              }
            float* Df = static_cast<float*>(p.D);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rloc = 4 * i + rsub;
              const int mm = mrow0 + rloc;
              const float rmr = __shfl_sync(0xffffffffu, rm, rloc);
              if (mm < p.M && n < p.N) {
                const float4 a = *reinterpret_cast<const float4*>(sbuf + rloc * 32 + ((g ^ (rloc & 7)) << 2));
                float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = apply_act((v[u] * alpha + b4[u]) * rmr, act) * s4[u];
                const long long o = zoff + (long long)mm * d_ld + n;
                const float f = resid_masked ? rmr : 1.0f;
                if (vec) {
                  if (resid) {
                    const float4 rv = *reinterpret_cast<const float4*>(resid + o);
                    v[0] += rv.x * f; v[1] += rv.y * f; v[2] += rv.z * f; v[3] += rv.w * f;
                  }
                  *reinterpret_cast<float4*>(Df + o) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                  for (int u = 0; u < 4; ++u)
                    if (n + u < p.N) Df[o + u] = v[u] + (resid ? resid[o + u] * f : 0.0f);
                }
              }
            }
          } else {
            // 4 lanes cover one 64-byte row segment (8 columns each), 8 rows per instruction
            const int rsub = lane >> 2, g2 = (lane & 3) * 2;
            const int n = nb + 4 * g2;
            const bool vec = p.vec_ok && (n + 7 < p.N);
            float b8[8], s8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              b8[u] = (bias && n + u < p.N) ? __ldg(bias + n + u) : 0.f;
              s8[u] = (colscale && n + u < p.N) ? __ldg(colscale + n + u) : 1.f;
            }
            uint16_t* Db = static_cast<uint16_t*>(p.D);
#pragma unroll 2
            for (int r0 = 0; r0 < 32; r0 += 8) {
              const int rloc = r0 + rsub;
              const int mm = mrow0 + rloc;
              const float rmr = __shfl_sync(0xffffffffu, rm, rloc);
              if (mm < p.M && n < p.N) {
                const float4 a = *reinterpret_cast<const float4*>(sbuf + rloc * 32 + ((g2 ^ (rloc & 7)) << 2));
                const float4 b = *reinterpret_cast<const float4*>(sbuf + rloc * 32 + (((g2 + 1) ^ (rloc & 7)) << 2));
                float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = apply_act((v[u] * alpha + b8[u]) * rmr, act) * s8[u];
                const long long o = zoff + (long long)mm * d_ld + n;
                if (resid) {
                  const float f = resid_masked ? rmr : 1.0f;
#pragma unroll
                  for (int u = 0; u < 8; ++u)
                    if (n + u < p.N) v[u] += resid[o + u] * f;
                }
                if (vec) {
                  uint32_t h[4], l[4];
                  if (d_lo) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) split16x2(v[2 * u], v[2 * u + 1], d_fmt, h[u], l[u]);
                    *reinterpret_cast<uint4*>(Db + o) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4*>(Db + d_lo + o) = make_uint4(l[0], l[1], l[2], l[3]);
                  } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) h[u] = pack16x2(v[2 * u], v[2 * u + 1], d_fmt);
                    *reinterpret_cast<uint4*>(Db + o) = make_uint4(h[0], h[1], h[2], h[3]);
                  }
                } else {
#pragma unroll
                  for (int u = 0; u < 8; ++u)
                    if (n + u < p.N) store16_split(Db, o + u, d_lo, v[u], d_fmt);
                }
              }
            }
          }
        }
      };
      if (act == VILCO_ACT_NONE) run_chunks(std::integral_constant<int, VILCO_ACT_NONE>{});
      else if (act == VILCO_ACT_RELU) run_chunks(std::integral_constant<int, VILCO_ACT_RELU>{});
      else if (GX && act == VILCO_ACT_EXP2) run_chunks(std::integral_constant<int, VILCO_ACT_EXP2>{});
      else if (!GX) run_chunks(std::integral_constant<int, VILCO_ACT_GELU>{});
      // this warp is done reading the accumulator stage
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster((tempty0 + 8 * acc) & PEER_BIT_MASK);   // the leader's barrier
        else mbar_arrive(tempty0 + 8 * acc);
      }
    }
  }

  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// SIMT cross-check kernel: same contract, one thread per output element, reads global memory directly.
// ---------------------------------------------------------------------------------------------
struct SimtAddr {
  const __nv_bfloat16* A; long long a_ld, a_s1, a_s2, a_lo; int a_rows;
  const __nv_bfloat16* B; long long b_ld, b_s1, b_s2, b_lo;
  int Z2;
};
__device__ __forceinline__ float ld_split(const __nv_bfloat16* p, long long lo, int fmt) {
  return load16_split(reinterpret_cast<const uint16_t*>(p), lo, fmt);
}

__global__ void gemm_simt_kernel(const GemmDev p, const SimtAddr q) {
  const long long total = (long long)p.M * p.N;
  const int z = blockIdx.y;
  const int z1 = z % p.Z1, z2 = z / p.Z1;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int m = static_cast<int>(idx / p.N);
    const int n = static_cast<int>(idx % p.N);
    float acc = 0.0f;
    for (int tap = 0; tap < p.taps; ++tap) {
      const int row = m + tap - (p.taps >> 1);
      if (row < 0 || row >= q.a_rows) continue;
      if (p.a_major == 1) {   // A stored (K, M): element (m, k) at k * a_ld + m; B MN-major or K-major
        const __nv_bfloat16* a = q.A + z1 * q.a_s1 + z2 * q.a_s2 + m;
        const __nv_bfloat16* b = q.B + (p.b_batched ? z1 * q.b_s1 + z2 * q.b_s2 : 0) + (p.b_major == 0 ? (long long)n * q.b_ld : n);
        const long long bs = p.b_major == 0 ? 1 : q.b_ld;
        for (int k = 0; k < p.K; ++k) acc = fmaf(ld_split(a + (long long)k * q.a_ld, q.a_lo, p.a_fmt), ld_split(b + k * bs, q.b_lo, p.b_fmt), acc);
        continue;
      }
      const __nv_bfloat16* a = q.A + z1 * q.a_s1 + z2 * q.a_s2 + (long long)row * q.a_ld;
      if (p.b_major == 0) {
        const __nv_bfloat16* b = q.B + (p.b_batched ? z1 * q.b_s1 + z2 * q.b_s2 : tap * q.b_s1) + (long long)n * q.b_ld;
        for (int k = 0; k < p.K; ++k) acc = fmaf(ld_split(a + k, q.a_lo, p.a_fmt), ld_split(b + k, q.b_lo, p.b_fmt), acc);
      } else {
        const __nv_bfloat16* b = q.B + (p.b_batched ? z1 * q.b_s1 + z2 * q.b_s2 : tap * q.b_s1) + n;
        for (int k = 0; k < p.K; ++k) acc = fmaf(ld_split(a + k, q.a_lo, p.a_fmt), ld_split(b + (long long)k * q.b_ld, q.b_lo, p.b_fmt), acc);
      }
    }
    float rm = 1.0f;
    if (p.rowmul) rm = p.rowmul[z2 * p.rowmul_zs + m];
    const long long doff = z1 * p.d_s1 + z2 * p.d_s2 + (long long)m * p.d_ld + n;
    const float res = p.resid ? p.resid[doff] * (p.resid_masked ? rm : 1.0f) : 0.0f;
    const float o = epi_value(p, acc, n, rm, res);
    if (p.d_dtype == VILCO_F32) static_cast<float*>(p.D)[doff] = o;
    else store16_split(static_cast<uint16_t*>(p.D), doff, p.d_lo, o, p.d_dtype);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

// Build a 4-D bf16 tensor map (inner, row, z1, z2); the three outer dims are sorted by stride so the
// descriptor always sees non-decreasing strides.  slots[] receives the coordinate slot of (row, z1, z2).
int encode_map(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t rows, int64_t ld, uint64_t n1,
               int64_t s1, uint64_t n2, int64_t s2, int64_t lo_off, uint32_t box_inner, uint32_t box_rows,
               int slots[3]) {
  auto fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return VILCO_E_CUDA; }
  struct Dim { uint64_t n; int64_t stride; uint32_t box; int id; };
  Dim d[3] = {{rows, ld, box_rows, 0}, {n1, s1, 1, 1}, {n2, s2, 1, 2}};
  // extent-1 dims may carry any stride: give them a harmless large one
  int64_t big = 16;
  for (auto& x : d) if (x.n > 1 && x.stride * (int64_t)x.n > big) big = x.stride * (int64_t)x.n;
  for (auto& x : d) if (x.n <= 1) { x.n = 1; x.stride = (big + 7) / 8 * 8; }
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (d[j].stride < d[i].stride) { Dim t = d[i]; d[i] = d[j]; d[j] = t; }
  // 5th dim = operand plane (hi, lo); extent 1 when the operand has no lo plane
  const uint64_t planes = lo_off ? 2 : 1;
  const int64_t pstride = lo_off ? lo_off : (big + 7) / 8 * 8 * 2;
  cuuint64_t gdim[5] = {inner, d[0].n, d[1].n, d[2].n, planes};
  cuuint64_t gstr[4] = {(cuuint64_t)d[0].stride * 2, (cuuint64_t)d[1].stride * 2, (cuuint64_t)d[2].stride * 2,
                        (cuuint64_t)pstride * 2};
  cuuint32_t box[5] = {box_inner, d[0].box, d[1].box, d[2].box, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 3; ++i) slots[d[i].id] = i + 1;
  for (int i = 0; i < 4; ++i)
    if (gstr[i] % 16 != 0) { set_error("tensor map stride %llu bytes is not a multiple of 16", (unsigned long long)gstr[i]); return VILCO_E_ARG; }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) { set_error("tensor map base not 16-byte aligned"); return VILCO_E_ARG; }
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu %llu strides %llu %llu %llu box %u %u", (int)r,
              (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2],
              (unsigned long long)gdim[3], (unsigned long long)gstr[0], (unsigned long long)gstr[1],
              (unsigned long long)gstr[2], box[0], box[1]);
    return VILCO_E_CUDA;
  }
  return VILCO_OK;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BN, int CG, bool GX = false>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmDev& p, int Z, cudaStream_t st) {
  const int stage_bytes = p.pa * (BM * BK * 2) + p.pb * ((BN / CG) * BK * 2);
  int stages = (SMEM_LIMIT - 1024 - BAR_BYTES - EPI_SMEM) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) { set_error("vilco_gemm: tile does not fit shared memory"); return VILCO_E_UNSUPPORTED; }
  p.stages = stages;
  const int smem = 1024 + BAR_BYTES + EPI_SMEM + stages * stage_bytes;
  static bool configured = false;
  if (!configured) {
    VILCO_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, CG, GX>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    configured = true;
  }
  const long long tiles = (long long)((p.N + BN - 1) / BN) * ((p.M + BM * CG - 1) / (BM * CG)) * Z;
  const int workers = num_sms() / CG;
  const int grid = static_cast<int>(tiles < workers ? tiles : workers) * CG;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CG == 2 ? 1 : 0;
  VILCO_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CG, GX>, tmA, tmB, p));
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}

// =============================================================================================
// Fused masked attention (global MaskedMHCA core and the MaskedMHA cross attention), head dim 64:
//   O[b, i, h*64:(h+1)*64] = softmax_j( scale * q_i . k_j  |  keys with kmask == 0 excluded ) @ v
// One CTA = 128 query rows of one (batch, head).  S = Q K^T lives only in TMEM (two 128-column buffers), never in HBM.
// Two passes over the key tiles: pass 1 accumulates the row max / sum (online, one thread per row), pass 2 recomputes
// S, writes P = exp(S - m) / l as bf16 (hi, lo) into 128B-swizzled shared memory and accumulates O += P V in TMEM —
// recomputing QK^T (K = 64) is cheaper than rescaling O.  warp 0: TMA, warp 1: MMA issuer, warps 2..5: softmax.
// =============================================================================================
static constexpr int AT_THREADS = 320;  // TMA warp, MMA warp, 8 softmax warps
static constexpr int AT_BQ = 128, AT_BKV = 128, AT_D = 64;
static constexpr int AT_MAX_TK = 2048;

struct AttnDev {
  int q_slot_row, q_slot_z1, q_slot_z2;
  int k_slot_row, k_slot_z1, k_slot_z2;
  int v_slot_row, v_slot_z1, v_slot_z2;
  int Tq, Tk, H;
  float scale;
  const float* kmask;  // (B, Tk) or null
  __nv_bfloat16* O; long long o_lo; long long o_ld, o_sh, o_sb;  // element strides: row, head, batch
  int fmt;             // element format of q / k / v / P / O (activation planes)
};

template <bool SPLIT>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnDev p) {
  constexpr int PL = SPLIT ? 2 : 1;
  constexpr int TILE = AT_BQ * AT_D * 2;            // 16 KB: 128 rows x 128 bytes
  constexpr int Q_BYTES = PL * TILE, K_BYTES = PL * TILE, V_BYTES = PL * TILE;
  constexpr int P_BYTES = PL * 2 * TILE;            // 128 rows x 128 keys = two 64-key K-major blocks per plane
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;                       // 2 stages
  uint8_t* sV = sK + 2 * K_BYTES;                   // 1 stage
  uint8_t* sP = sV + V_BYTES;
  uint32_t* s_bits = reinterpret_cast<uint32_t*>(sP + P_BYTES);      // key validity, one bit per key (AT_MAX_TK / 32 words)
  float* s_part = reinterpret_cast<float*>(s_bits + AT_MAX_TK / 32);  // (m, l) of 2 halves x 128 rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 2 * 128 * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t qfull = smem_u32(bars + 0);
  const uint32_t kfull0 = smem_u32(bars + 1), kempty0 = smem_u32(bars + 3);
  const uint32_t vfull = smem_u32(bars + 5), vempty = smem_u32(bars + 6);
  const uint32_t sfull0 = smem_u32(bars + 7), sempty0 = smem_u32(bars + 9);
  const uint32_t pfull = smem_u32(bars + 11), pempty = smem_u32(bars + 12), ofull = smem_u32(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.Tk + AT_BKV - 1) / AT_BKV;
  const int G = 2 * nkv;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    mbar_init(qfull, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(kfull0 + 8 * i, 1); mbar_init(kempty0 + 8 * i, 1);
      mbar_init(sfull0 + 8 * i, 1); mbar_init(sempty0 + 8 * i, 8);
    }
    mbar_init(vfull, 1); mbar_init(vempty, 1);
    mbar_init(pfull, 8); mbar_init(pempty, 1); mbar_init(ofull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {  // key validity bits of this batch element (tail keys are invalid too)
    for (int j0 = (warp - 2) * 32; j0 < nkv * AT_BKV; j0 += 8 * 32) {
      const int j = j0 + lane;
      const bool ok = j < p.Tk && (!p.kmask || p.kmask[(long long)b * p.Tk + j] != 0.f);
      const uint32_t w = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) s_bits[j0 >> 5] = w;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS0 = tmem_base, tO = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {
      int c[4];
      mbar_expect_tx(qfull, Q_BYTES);
      c[0] = 0; c[p.q_slot_row] = q0; c[p.q_slot_z1] = h; c[p.q_slot_z2] = b;
      for (int pl = 0; pl < PL; ++pl) tma_load_5d(smem_u32(sQ) + pl * TILE, &tmQ, qfull, c[0], c[1], c[2], c[3], pl);
      for (int g = 0; g < G; ++g) {
        const int st = g & 1, j = g % nkv;
        mbar_wait(kempty0 + 8 * st, ((g >> 1) & 1) ^ 1);
        mbar_expect_tx(kfull0 + 8 * st, K_BYTES);
        c[0] = 0; c[p.k_slot_row] = j * AT_BKV; c[p.k_slot_z1] = h; c[p.k_slot_z2] = b;
        for (int pl = 0; pl < PL; ++pl)
          tma_load_5d(smem_u32(sK) + st * K_BYTES + pl * TILE, &tmK, kfull0 + 8 * st, c[0], c[1], c[2], c[3], pl);
        if (g >= nkv) {
          mbar_wait(vempty, (j & 1) ^ 1);
          mbar_expect_tx(vfull, V_BYTES);
          c[0] = 0; c[p.v_slot_row] = j * AT_BKV; c[p.v_slot_z1] = h; c[p.v_slot_z2] = b;
          for (int pl = 0; pl < PL; ++pl) tma_load_5d(smem_u32(sV) + pl * TILE, &tmV, vfull, c[0], c[1], c[2], c[3], pl);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = make_idesc(AT_BKV, 0, 0, p.fmt, p.fmt);   // N = 128 keys, B K-major
    const uint32_t idesc_o = make_idesc(AT_D, 1, 0, p.fmt, p.fmt);     // N = 64, B (V) MN-major
    mbar_wait(qfull, 0);
    auto issue_pv = [&](int j) {
      mbar_wait(pfull, j & 1);
      mbar_wait(vfull, j & 1);
      tcgen05_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int ks = 0; ks < AT_BKV / UMMA_K; ++ks) {
          const uint32_t aoff = (ks >> 2) * TILE + (ks & 3) * 32;  // 64-key block, then 16-key step inside the atom
          const uint32_t boff = ks * UMMA_K * 128;                 // 16 key rows of 128 bytes
          const uint64_t a_hi = make_smem_desc(smem_u32(sP) + aoff, 16, 1024);
          const uint64_t b_hi = make_smem_desc(smem_u32(sV) + boff, 16, 1024);
          tcgen05_mma_f16(tO, a_hi, b_hi, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
          if (SPLIT) {
            const uint64_t a_lo = make_smem_desc(smem_u32(sP) + 2 * TILE + aoff, 16, 1024);
            const uint64_t b_lo = make_smem_desc(smem_u32(sV) + TILE + boff, 16, 1024);
            tcgen05_mma_f16(tO, a_hi, b_lo, idesc_o, 1u);
            tcgen05_mma_f16(tO, a_lo, b_hi, idesc_o, 1u);
          }
        }
        tcgen05_commit(pempty);
        tcgen05_commit(vempty);
      }
      __syncwarp();
    };
    for (int g = 0; g < G; ++g) {
      const int st = g & 1;
      const uint32_t ph = (g >> 1) & 1;
      mbar_wait(kfull0 + 8 * st, ph);
      mbar_wait(sempty0 + 8 * st, ph ^ 1);
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t kb = smem_u32(sK) + st * K_BYTES;
#pragma unroll
        for (int k = 0; k < AT_D / UMMA_K; ++k) {
          const uint64_t a_hi = make_smem_desc(smem_u32(sQ) + k * 32, 16, 1024);
          const uint64_t b_hi = make_smem_desc(kb + k * 32, 16, 1024);
          tcgen05_mma_f16(tS0 + st * 128, a_hi, b_hi, idesc_s, k > 0 ? 1u : 0u);
          if (SPLIT) {
            const uint64_t a_lo = make_smem_desc(smem_u32(sQ) + TILE + k * 32, 16, 1024);
            const uint64_t b_lo = make_smem_desc(kb + TILE + k * 32, 16, 1024);
            tcgen05_mma_f16(tS0 + st * 128, a_hi, b_lo, idesc_s, 1u);
            tcgen05_mma_f16(tS0 + st * 128, a_lo, b_hi, idesc_s, 1u);
          }
        }
        tcgen05_commit(kempty0 + 8 * st);
        tcgen05_commit(sfull0 + 8 * st);
      }
      __syncwarp();
      if (g > nkv) issue_pv(g - nkv - 1);  // P V of the previous key tile, after the next S is already in flight
    }
    issue_pv(nkv - 1);
    if (lane == 0) tcgen05_commit(ofull);
    __syncwarp();
  } else {
    // ===== 8 softmax warps: TMEM lane quadrant q = warp % 4 (one thread per query row), column half = (warp-2)/4 =====
    // everything is kept in the log2 domain: x2 = s * scale * log2(e), p = 2^(x2 - m2l)
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;                 // row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const float sc2 = p.scale * 1.4426950408889634f;
    float m = -INFINITY, l = 0.f, m2l = 0.f;
    bool dead = false;
    for (int g = 0; g < G; ++g) {
      const int st = g & 1, j = g % nkv;
      mbar_wait(sfull0 + 8 * st, (g >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t tS = tS0 + st * 128 + lane_addr + half * 64;
      if (g < nkv) {
        // pass 1: online row max / sum over this warp's 64 columns of the tile
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const uint32_t bits = s_bits[j * 4 + half * 2 + c];   // validity of the 32 keys (warp-uniform)
          uint32_t r[32];
          __syncwarp();
          tmem_ld32(tS + c * 32, r);
          if (bits == 0u) continue;
          float x[32];
          float cmax = -INFINITY;
          if (bits == 0xFFFFFFFFu) {
#pragma unroll
            for (int u = 0; u < 32; ++u) { x[u] = __uint_as_float(r[u]) * sc2; cmax = fmaxf(cmax, x[u]); }
          } else {
#pragma unroll
            for (int u = 0; u < 32; ++u) {
              x[u] = ((bits >> u) & 1u) ? __uint_as_float(r[u]) * sc2 : -INFINITY;
              cmax = fmaxf(cmax, x[u]);
            }
          }
          const float m_new = fmaxf(m, cmax);      // finite: at least one valid key in this chunk
          float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
          for (int u = 0; u < 32; u += 2) { sum0 += ex2f(x[u] - m_new); sum1 += ex2f(x[u + 1] - m_new); }
          l = l * ex2f(m - m_new) + (sum0 + sum1);  // ex2(-inf) = 0 covers the first chunk
          m = m_new;
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sempty0 + 8 * st);
      } else {
        if (g == nkv) {
          // combine the two column halves of every row (only cross-warp exchange of the kernel)
          s_part[(half * 128 + row) * 2] = m;
          s_part[(half * 128 + row) * 2 + 1] = l;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const float mo = s_part[((half ^ 1) * 128 + row) * 2], lo_ = s_part[((half ^ 1) * 128 + row) * 2 + 1];
          const float M = fmaxf(m, mo);
          dead = M == -INFINITY;                    // fully masked row -> zeros
          const float Lsum = dead ? 1.f : l * ex2f(m - M) + lo_ * ex2f(mo - M);
          m2l = dead ? 0.f : M + __log2f(Lsum);
        }
        mbar_wait(pempty, (j & 1) ^ 1);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const uint32_t bits = dead ? 0u : s_bits[j * 4 + half * 2 + c];
          uint32_t r[32];
          __syncwarp();
          tmem_ld32(tS + c * 32, r);
          // 32 keys = four 16-byte chunks per plane; 64-key block = half, chunk index inside its 128-byte row
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int k0 = ch * 8 + 2 * u;
              float a = ex2f(__uint_as_float(r[k0]) * sc2 - m2l);
              float bq = ex2f(__uint_as_float(r[k0 + 1]) * sc2 - m2l);
              if (bits != 0xFFFFFFFFu) {
                if (!((bits >> k0) & 1u)) a = 0.f;
                if (!((bits >> (k0 + 1)) & 1u)) bq = 0.f;
              }
              if (SPLIT) split16x2(a, bq, p.fmt, hi[u], lo[u]);
              else hi[u] = pack16x2(a, bq, p.fmt);
            }
            const int chunk = c * 4 + ch;
            uint8_t* dst = sP + half * TILE + row * 128 + ((chunk ^ (row & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (SPLIT) *reinterpret_cast<uint4*>(dst + 2 * TILE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        tcgen05_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> tensor core reads
        __syncwarp();
        if (lane == 0) { mbar_arrive(sempty0 + 8 * st); mbar_arrive(pfull); }
      }
    }
    // ===== output: O (128 x 64 fp32 in TMEM) -> bf16 (hi, lo) rows in global memory; 32 columns per warp =====
    mbar_wait(ofull, 0);
    tcgen05_fence_after();
    const int grow = q0 + row;
    __nv_bfloat16* orow = p.O + (long long)b * p.o_sb + (long long)h * p.o_sh + (long long)grow * p.o_ld + half * 32;
    {
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(tO + lane_addr + half * 32, r);
      if (grow < p.Tq) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float a = __uint_as_float(r[ch * 8 + 2 * u]), bq = __uint_as_float(r[ch * 8 + 2 * u + 1]);
            split16x2(a, bq, p.fmt, hi[u], lo[u]);
          }
          *reinterpret_cast<uint4*>(orow + ch * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (p.o_lo) *reinterpret_cast<uint4*>(orow + p.o_lo + ch * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace vilco

using namespace vilco;

extern "C" int vilco_gemm(const VilcoGemm* g, void* stream) {
  VILCO_CHECK_ARG(g && g->A && g->B && g->D, "vilco_gemm: null pointer");
  VILCO_CHECK_ARG(g->M > 0 && g->N > 0 && g->K > 0 && g->Z1 > 0 && g->Z2 > 0, "vilco_gemm: empty problem");
  VILCO_CHECK_ARG(g->taps == 1 || g->taps == 3, "vilco_gemm: taps must be 1 or 3");
  VILCO_CHECK_ARG(!(g->b_batched && g->taps != 1), "vilco_gemm: batched B cannot have taps");
  VILCO_CHECK_ARG(g->d_dtype == VILCO_F32 || g->d_dtype == VILCO_BF16 || g->d_dtype == VILCO_F16, "vilco_gemm: bad d_dtype");
  VILCO_CHECK_ARG(!(g->resid_masked && !g->rowmul), "vilco_gemm: resid_masked needs rowmul");
  VILCO_CHECK_ARG(g->a_major == 0 || (g->a_major == 1 && g->taps == 1), "vilco_gemm: MN-major A needs taps == 1");
  const int a_fmt = g->a_fmt ? g->a_fmt : VILCO_BF16, b_fmt = g->b_fmt ? g->b_fmt : VILCO_BF16;   // 0 = legacy callers: bf16
  VILCO_CHECK_ARG((a_fmt == VILCO_BF16 || a_fmt == VILCO_F16) && (b_fmt == VILCO_BF16 || b_fmt == VILCO_F16),
                  "vilco_gemm: operand formats must be VILCO_BF16 or VILCO_F16");
  // measured on B200: a kind::f16 MMA whose A and B formats differ raises "illegal instruction"
  VILCO_CHECK_ARG(a_fmt == b_fmt || g->impl == 1, "vilco_gemm: A and B must have the same 16-bit format (tcgen05 kind::f16)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  GemmDev p{};
  p.M = g->M; p.N = g->N; p.K = g->K; p.taps = g->taps; p.Z1 = g->Z1; p.Ztot = g->Z1 * g->Z2;
  p.band_lo = g->band_lo; p.band_hi = g->band_hi;
  p.a_major = g->a_major;
  p.b_major = g->b_major; p.b_batched = g->b_batched;
  p.pa = g->a_lo ? 2 : 1; p.pb = g->b_lo ? 2 : 1;
  p.a_fmt = a_fmt; p.b_fmt = b_fmt;
  p.D = g->D; p.d_dtype = g->d_dtype; p.d_ld = g->d_ld; p.d_s1 = g->d_s1; p.d_s2 = g->d_s2;
  p.d_lo = g->d_dtype != VILCO_F32 ? g->d_lo : 0;
  p.alpha = g->alpha; p.bias = g->bias; p.rowmul = g->rowmul; p.rowmul_zs = g->rowmul_zs;
  p.act = g->act; p.colscale = g->colscale; p.resid = g->resid; p.resid_masked = g->resid_masked;
  p.rowsub = g->rowsub; p.rowsub_s1 = g->rowsub_s1; p.rowsub_s2 = g->rowsub_s2; p.colscale_zs = g->colscale_zs;
  p.emul = static_cast<const uint16_t*>(g->emul);
  const int esz = g->d_dtype == VILCO_F32 ? 4 : 2;
  // 16-byte vector stores / residual loads: every (row, column % vec == 0) element must be 16-byte aligned
  p.vec_ok = (reinterpret_cast<uintptr_t>(g->D) % 16 == 0) && ((g->d_ld * esz) % 16 == 0) &&
             ((g->d_s1 * esz) % 16 == 0) && ((g->d_s2 * esz) % 16 == 0) && ((p.d_lo * esz) % 16 == 0) &&
             (!g->resid || (reinterpret_cast<uintptr_t>(g->resid) % 16 == 0 && (g->d_ld * 4) % 16 == 0 &&
                            (g->d_s1 * 4) % 16 == 0 && (g->d_s2 * 4) % 16 == 0));
  const int Z = g->Z1 * g->Z2;
  if (g->rowsub || g->emul || g->colscale_zs || g->act == VILCO_ACT_EXP2) {
    // the fused gradient epilogues live in the bounds-free 16-bit path of the tensor-core kernel only
    VILCO_CHECK_ARG(g->impl != 1 && g->d_dtype != VILCO_F32 && p.vec_ok && !g->resid && g->M % 32 == 0 && g->N % 32 == 0 &&
                        (!g->bias || reinterpret_cast<uintptr_t>(g->bias) % 16 == 0) &&
                        (!g->colscale || (reinterpret_cast<uintptr_t>(g->colscale) % 16 == 0 && g->colscale_zs % 4 == 0)) &&
                        (!g->emul || reinterpret_cast<uintptr_t>(g->emul) % 16 == 0),
                    "vilco_gemm: rowsub / emul / colscale_zs / ACT_EXP2 need a 16-bit output, M %% 32 == N %% 32 == 0 and 16-byte alignment");
  }

  if (g->impl == 1) {
    SimtAddr q{};
    q.A = static_cast<const __nv_bfloat16*>(g->A); q.a_ld = g->a_ld; q.a_s1 = g->a_s1; q.a_s2 = g->a_s2; q.a_rows = g->a_rows;
    q.a_lo = g->a_lo; q.b_lo = g->b_lo;
    q.B = static_cast<const __nv_bfloat16*>(g->B); q.b_ld = g->b_ld; q.b_s1 = g->b_s1; q.b_s2 = g->b_s2; q.Z2 = g->Z2;
    const long long total = (long long)g->M * g->N;
    int blocks = static_cast<int>((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    gemm_simt_kernel<<<dim3(blocks, Z), 256, 0, st>>>(p, q);
    VILCO_LAUNCH_CHECK();
    return VILCO_OK;
  }

  // ---- tcgen05 path: tile shape ----
  int BN, CG = 1;
  const long long t128 = (long long)((g->N + 127) / 128) * ((g->M + BM - 1) / BM) * Z;   // 128 x 128 tiles
  const int force_cg = g->impl >= 2 ? g->impl - 1 : 0;    // impl 2 / 3: force cta_group 1 / 2 (probes); 0: automatic
  if (g->b_major == 1) {
    BN = (g->N <= 64 || t128 * 2 <= num_sms()) ? 64 : 128;
  } else if (g->N <= 32) BN = 32;
  else if (g->N <= 64) BN = 64;
  else {
    // small problems (text stem, deep pyramid levels): 128x64 tiles double the number of CTAs when 128x128 tiles
    // would leave most of the 148 SMs idle
    BN = (t128 * 2 <= num_sms()) ? 64 : 128;
  }
  // CTA pairs (cta_group::2, 256 x 256 or 256 x 128 tiles), any operand majors (the gradient GEMMs read their operands MN-major).
  // The main loop is bound by the L2 -> shared-memory stream, not by the MMA rate (ncu: the MMA warp waits on the full
  // barriers), and a pair tile moves half the bytes per FLOP of a 128 x 128 tile: measured 1.5 - 1.9x faster per launch as soon
  // as the 256 x 256 tiles fill half of the 74 CTA pairs (tools/gemm2_probe.py bwd), so that is the threshold.
  if (g->band_hi <= g->band_lo && force_cg != 1 && g->N >= 128 && (g->M > 128 || force_cg == 2)) {
    const long long t256 = (long long)((g->N + 255) / 256) * ((g->M + 255) / 256) * Z;
    if (g->N >= 256 && (t256 >= num_sms() / 4 || force_cg == 2)) { BN = 256; CG = 2; }
    else if (t128 / 2 >= 2 * (num_sms() / 2) || force_cg == 2) { BN = 128; CG = 2; }
  }

  CUtensorMap tmA, tmB;
  int sa[3], sb[3];
  int rc;
  if (g->a_major == 0)
    rc = encode_map(&tmA, g->A, (uint64_t)g->K, (uint64_t)g->a_rows, g->a_ld, (uint64_t)g->Z1, g->a_s1,
                    (uint64_t)g->Z2, g->a_s2, g->a_lo, BK, BM, sa);
  else   // (m inner, k rows, z1, z2), one 64-wide swizzle atom per box
    rc = encode_map(&tmA, g->A, (uint64_t)g->M, (uint64_t)g->K, g->a_ld, (uint64_t)g->Z1, g->a_s1,
                    (uint64_t)g->Z2, g->a_s2, g->a_lo, 64, BK, sa);
  if (rc) return rc;
  const uint64_t n1 = g->b_batched ? (uint64_t)g->Z1 : (uint64_t)g->taps;
  const uint64_t n2 = g->b_batched ? (uint64_t)g->Z2 : 1;
  if (g->b_major == 0)   // (k inner, n rows, z1|tap, z2)
    rc = encode_map(&tmB, g->B, (uint64_t)g->K, (uint64_t)g->N, g->b_ld, n1, g->b_s1, n2, g->b_s2, g->b_lo, BK, BN / CG, sb);
  else                   // (n inner, k rows, z1, z2)
    rc = encode_map(&tmB, g->B, (uint64_t)g->N, (uint64_t)g->K, g->b_ld, n1, g->b_s1, n2, g->b_s2, g->b_lo, 64, BK, sb);
  if (rc) return rc;
  p.a_slot_row = sa[0]; p.a_slot_z1 = sa[1]; p.a_slot_z2 = sa[2];
  p.b_slot_row = sb[0]; p.b_slot_z1 = sb[1]; p.b_slot_z2 = sb[2];

  if (g->rowsub || g->emul || g->colscale_zs || g->act == VILCO_ACT_EXP2) {      // fused attention-gradient epilogues
    VILCO_CHECK_ARG(g->act == VILCO_ACT_NONE || g->act == VILCO_ACT_RELU || g->act == VILCO_ACT_EXP2, "vilco_gemm: act unsupported with rowsub / emul");
    if (CG == 2) return BN == 256 ? launch_tc<256, 2, true>(tmA, tmB, p, Z, st) : launch_tc<128, 2, true>(tmA, tmB, p, Z, st);
    VILCO_CHECK_ARG(BN >= 64, "vilco_gemm: N too small for the fused gradient epilogues");
    return BN == 64 ? launch_tc<64, 1, true>(tmA, tmB, p, Z, st) : launch_tc<128, 1, true>(tmA, tmB, p, Z, st);
  }
  if (CG == 2) return BN == 256 ? launch_tc<256, 2>(tmA, tmB, p, Z, st) : launch_tc<128, 2>(tmA, tmB, p, Z, st);
  switch (BN) {
    case 32: return launch_tc<32, 1>(tmA, tmB, p, Z, st);
    case 64: return launch_tc<64, 1>(tmA, tmB, p, Z, st);
    default: return launch_tc<128, 1>(tmA, tmB, p, Z, st);
  }
}


extern "C" int vilco_attention(const void* q, int64_t q_lo, const void* k, const void* v, int64_t kv_lo, const float* kmask,
                               void* out, int64_t out_lo, int B, int H, int Tq, int Tk, int C, float scale, void* stream) {
  VILCO_CHECK_ARG(q && k && v && out, "vilco_attention: null pointer");
  VILCO_CHECK_ARG(H > 0 && C == H * AT_D, "vilco_attention: head dim must be 64 (C=%d, H=%d)", C, H);
  VILCO_CHECK_ARG(Tq > 0 && Tk > 0 && Tk <= AT_MAX_TK, "vilco_attention: Tk=%d unsupported (1..%d)", Tk, AT_MAX_TK);
  VILCO_CHECK_ARG(reinterpret_cast<uintptr_t>(out) % 16 == 0 && out_lo % 8 == 0, "vilco_attention: out alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool split = q_lo != 0 && kv_lo != 0;
  CUtensorMap tmQ, tmK, tmV;
  int sq[3], sk[3], sv[3];
  int rc = encode_map(&tmQ, q, AT_D, (uint64_t)Tq, C, (uint64_t)H, AT_D, (uint64_t)B, (int64_t)Tq * C, split ? q_lo : 0,
                      AT_D, AT_BQ, sq);
  if (rc) return rc;
  rc = encode_map(&tmK, k, AT_D, (uint64_t)Tk, C, (uint64_t)H, AT_D, (uint64_t)B, (int64_t)Tk * C, split ? kv_lo : 0, AT_D,
                  AT_BKV, sk);
  if (rc) return rc;
  rc = encode_map(&tmV, v, AT_D, (uint64_t)Tk, C, (uint64_t)H, AT_D, (uint64_t)B, (int64_t)Tk * C, split ? kv_lo : 0, AT_D,
                  AT_BKV, sv);
  if (rc) return rc;
  AttnDev p{};
  p.q_slot_row = sq[0]; p.q_slot_z1 = sq[1]; p.q_slot_z2 = sq[2];
  p.k_slot_row = sk[0]; p.k_slot_z1 = sk[1]; p.k_slot_z2 = sk[2];
  p.v_slot_row = sv[0]; p.v_slot_z1 = sv[1]; p.v_slot_z2 = sv[2];
  p.Tq = Tq; p.Tk = Tk; p.H = H; p.scale = scale; p.kmask = kmask;
  p.O = static_cast<__nv_bfloat16*>(out); p.o_lo = out_lo; p.o_ld = C; p.o_sh = AT_D; p.o_sb = (long long)Tq * C;
  p.fmt = act_fmt();
  const int PL = split ? 2 : 1;
  const int smem = PL * 16384 * (1 + 2 + 1 + 2) + AT_MAX_TK / 8 + 2 * 128 * 2 * 4 + 16 * 8 + 16 + 1024;
  dim3 grid((Tq + AT_BQ - 1) / AT_BQ, H, B);
  if (split) {
    static bool cfg = false;
    if (!cfg) { VILCO_CUDA(cudaFuncSetAttribute(attn_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); cfg = true; }
    attn_fused_kernel<true><<<grid, AT_THREADS, smem, st>>>(tmQ, tmK, tmV, p);
  } else {
    static bool cfg = false;
    if (!cfg) { VILCO_CUDA(cudaFuncSetAttribute(attn_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); cfg = true; }
    attn_fused_kernel<false><<<grid, AT_THREADS, smem, st>>>(tmQ, tmK, tmV, p);
  }
  VILCO_LAUNCH_CHECK();
  return VILCO_OK;
}
