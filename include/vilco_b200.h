/*
 * vilco_b200 — C ABI of the B200-native (sm_100a) Moment-Query hot path.
 *
 * This is the drop-in boundary: every entry point takes plain
 * device/host pointers, sizes and a CUDA stream (as void*), returns an int status
 * (0 = ok, otherwise a VILCO_E_* code; vilco_last_error() gives the text), never
 * allocates, never synchronises (except where stated) and never throws.
 *
 * Each function names the reference interface it replaces (paths relative to the
 * ViLCo repository root).  Layout convention inside the library: activations are
 * TOKEN-MAJOR (B, T, C) with C contiguous; fp32 for the residual stream / statistics,
 * bf16 for GEMM operands.  The reference's (B, C, T) tensors are converted once at
 * the boundary (vilco_pack_feats / the host layer).
 */
#ifndef VILCO_B200_H
#define VILCO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VILCO_OK 0
#define VILCO_E_ARG 1      /* bad argument (shape / alignment / null pointer) */
#define VILCO_E_CUDA 2     /* CUDA runtime / driver error */
#define VILCO_E_UNSUPPORTED 3

/* activation codes for fused epilogues */
#define VILCO_ACT_NONE 0
#define VILCO_ACT_RELU 1
#define VILCO_ACT_GELU 2   /* exact erf GELU (torch.nn.GELU default) */

/* dtype codes */
#define VILCO_F32 0
#define VILCO_BF16 1

const char* vilco_last_error(void);
int vilco_version(void);
/* number of kernels launched by this library since process start (bench.py "gpu_launches") */
uint64_t vilco_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Dense contraction:  D[z, m, n] = epi( alpha * sum_{tap,k} A[z, m + tap - taps/2, k] * B[tap|z, n, k] )
 *
 * Replaces every dense conv / linear / einsum on the path:
 *   MaskedConv1D.forward            MQ/libs/modeling/blocks.py:106-130 (k=1 and k=3, stride 1, groups 1)
 *   MaskedMHCA / MaskedMHA q,k,v,proj 1x1 convs and the two attention matmuls   blocks.py:228-269, 351-410
 *   ChannelAttention qkv / proj, ChannelBlock mlp                               blocks.py:423-466
 *   TransformerBlock.mlp            blocks.py:533-539
 *   XLNetRelativeAttention einsums  MQ/libs/modeling/modeling_xlnet_x.py:270-332, 440-446; XLNetFeedForward :482-490
 *   PtTransformerClsHead / RegHead convs  MQ/libs/modeling/meta_archs.py:259-275, 334-349
 *
 * A is bf16, K-major (k contiguous), addressed as a 4-D tensor (k, row, z1, z2) with element strides
 * (1, a_ld, a_s1, a_s2); rows outside [0, a_rows) read as zero (this implements the conv zero padding).
 * B is bf16; b_major = 0: K-major rows n (k contiguous), b_major = 1: MN-major rows k (n contiguous; N <= 64).
 *   b_batched = 0: B is a weight [taps][N][K] (tap stride b_s1 elements);  b_batched = 1: B indexed by (z1, z2)
 *   with strides (b_s1, b_s2) like A (taps must be 1).
 * epilogue order:  v = alpha*acc + bias[n];  v *= rowmul[z2*rowmul_zs + m];  v = act(v);
 *                  v = v*colscale[n] + resid[z,m,n] * (resid_masked ? rowmul[..] : 1);   store as d_dtype.
 * Any of bias / rowmul / colscale / resid may be NULL.  resid is fp32 with the same strides as D.
 * impl = 0: tcgen05 + TMA + TMEM kernel;  impl = 1: plain SIMT kernel (debug cross-check of the same math).
 * ------------------------------------------------------------------------------------ */
typedef struct VilcoGemm {
  const void* A; int64_t a_ld, a_s1, a_s2; int32_t a_rows;
  const void* B; int64_t b_ld, b_s1, b_s2; int32_t b_major, b_batched;
  int32_t M, N, K, taps, Z1, Z2;
  void* D; int32_t d_dtype; int64_t d_ld, d_s1, d_s2;
  float alpha;
  const float* bias;
  const float* rowmul; int64_t rowmul_zs;
  int32_t act;
  const float* colscale;
  const float* resid; int32_t resid_masked;
  int32_t impl;
} VilcoGemm;

int vilco_gemm(const VilcoGemm* g, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VILCO_B200_H */
