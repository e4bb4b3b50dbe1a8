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

#include <stddef.h>
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
#define VILCO_ACT_EXP2 3   /* 2^x (softmax recompute in the attention gradients, see VilcoGemm.rowsub) */

/* dtype codes */
#define VILCO_F32 0
#define VILCO_BF16 1
#define VILCO_F16 2

/* Element format of the 16-bit operand planes every kernel of this library reads and writes (VILCO_BF16, the initial
 * value, or VILCO_F16).  It is process-wide state read at launch time (and therefore baked into a captured CUDA graph):
 * set it once per precision mode, before packing weights.  fp16 planes carry 11 significant bits (8x tighter than bf16)
 * at the same tensor-core rate.  tcgen05 kind::f16 cannot mix fp16 and bf16 operands in one MMA, so the GRADIENT planes
 * the backward kernels emit (vilco_to_planes with grad = 1, vilco_resid_branch_bwd, vilco_softmax_bwd,
 * vilco_relshift_bwd) use the same format: they are stored multiplied by the gradient scale (vilco_set_grad_scale, a
 * power of two, 1 for bf16) and the caller folds 1 / scale into the alpha of every vilco_gemm that consumes them.
 * vilco_gemm takes the operand formats explicitly (a_fmt == b_fmt required by the hardware). */
int vilco_set_plane_format(int fmt);
int vilco_get_plane_format(void);
int vilco_set_grad_scale(float scale);

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
 * A and B are 16-bit operands (a_fmt / b_fmt: fp16 or bf16, independently) in one or two planes: a non-zero a_lo / b_lo is
 * the element offset of a second plane with x ~= hi + lo (lo = round16(x - hi)).  Per k-step the kernel issues hi*hi, plus
 * hi*lo when B has a lo plane, plus lo*hi when A has one (1, 2 or 3 tcgen05.mma into the same fp32 accumulator): single
 * planes for the contractions whose operand rounding the 1e-3 parity bar tolerates, split operands where it does not
 * (DESIGN.md section 2).  d_dtype VILCO_BF16 / VILCO_F16 with d_lo != 0 writes both planes of the output.
 * impl = 0: tcgen05 + TMA + TMEM kernel, tile shape chosen automatically (128 x {32,64,128} tiles on one CTA, 256 x {128,256}
 * tiles on a CTA pair with cta_group::2 for large K-major problems);  impl = 1: plain SIMT kernel (debug cross-check of the
 * same math);  impl = 2 / 3: tcgen05 kernel forced to one CTA / a CTA pair per tile (probes).
 * ------------------------------------------------------------------------------------ */
typedef struct VilcoGemm {
  const void* A; int64_t a_ld, a_s1, a_s2, a_lo; int32_t a_rows;
  const void* B; int64_t b_ld, b_s1, b_s2, b_lo; int32_t b_major, b_batched;
  int32_t M, N, K, taps, Z1, Z2;
  void* D; int32_t d_dtype; int64_t d_ld, d_s1, d_s2, d_lo;
  float alpha;
  const float* bias;
  const float* rowmul; int64_t rowmul_zs;
  int32_t act;
  const float* colscale;
  const float* resid; int32_t resid_masked;
  int32_t impl;
  int32_t band_lo, band_hi;   /* band_hi > band_lo: only outputs with band_lo <= m + n < band_hi are computed (others untouched) */
  int32_t a_major;            /* 0: A is (a_rows, K) K-major (row stride a_ld).  1: A is stored (K, M) MN-major — element (m, k) at
                                 A[k * a_ld + m] — the natural layout of a gradient matrix used as dZ^T in dW = dZ^T X; taps must be 1 */
  int32_t a_fmt, b_fmt;       /* element format of A / B: VILCO_BF16 or VILCO_F16 (0 = VILCO_BF16); must be equal on the tensor core */
  /* Fused softmax-recompute epilogues of the attention gradients (16-bit output, whole 32 x 32 chunks only; ABI version 3):
   *   out = act((acc * alpha + bias[n] - rowsub[z, m]) * rowmul[m]) * colscale[z2, n] * emul[z, m, n]
   * rowsub: fp32, element (z1, z2, m) at rowsub[z1 * rowsub_s1 + z2 * rowsub_s2 + m]; colscale gets the batch offset
   * z2 * colscale_zs; emul: ONE 16-bit plane laid out like D (d_ld / d_s1 / d_s2).  With act = VILCO_ACT_EXP2 and rowsub = the
   * row log-sum-exp (base 2) saved by vilco_self_attention, the QK^T GEMM writes the probabilities P directly; with emul = P and
   * rowsub = delta = rowsum(dO * O), the dO V^T GEMM writes dS = P (dP - delta) directly — no (T x T) fp32 tensor and no softmax
   * kernel in the backward pass. */
  const float* rowsub; int64_t rowsub_s1, rowsub_s2;
  int64_t colscale_zs;
  const void* emul;
} VilcoGemm;

int vilco_gemm(const VilcoGemm* g, void* stream);

/* ------------------------------------------------------------------------------------
 * Channel LayerNorm over token-major rows (one warp per token, two-pass statistics):
 *   y = act(LN(x [+ add]) * w + b) [+ pe[t] * rowmul[row]]   (rows flagged in zero_rows are written as 0)
 * Replaces LayerNorm.forward MQ/libs/modeling/blocks.py:160-175 (eps 1e-5), nn.LayerNorm of ChannelBlock
 * (blocks.py:449) and of the XLNet layer (eps 1e-12, modeling_xlnet_x.py:330, 489), the `relu(embd_norm(..))`
 * of the embedding network and the `+ pos_embd * mask` that follows it (MQ/libs/modeling/backbones.py:217-236),
 * FPNIdentity.forward (MQ/libs/modeling/necks.py:173-198).
 * Every bf16 OUTPUT of this library takes a `*_lo` element offset: when non-zero a second plane lo = bf16(v - hi) is
 * written there (split precision, see vilco_gemm); bf16 inputs of the non-GEMM kernels take the same offset.
 * x is fp32 or bf16 (x_dtype), rows x C contiguous; C % 128 == 0, C <= 1024.  y32 / y16 (either may be NULL) are
 * written at  batch * y_bs + t * y_ld  with batch = row / rows_per_batch, t = row % rows_per_batch.
 * ------------------------------------------------------------------------------------ */
int vilco_layernorm(const void* x, int x_dtype, const float* add, const float* w, const float* b, float eps, int relu,
                    const float* pe, int pe_T, const float* rowmul, const uint8_t* zero_rows, float* y32, void* y16,
                    int64_t y16_lo, int64_t y_ld, int64_t y_bs, int rows, int rows_per_batch, int C, void* stream);

/* depthwise conv (k=3, stride 1|2, zero pad 1, no bias) * out_mask -> LayerNorm, for up to three (weight, norm)
 * sets sharing one input: the q/k/v front of MaskedMHCA / LocalMaskedMHCA (blocks.py:315-345, 364-371; the mask is
 * nearest-downsampled, blocks.py:117-127).  x (B,T,C) fp32|bf16, mask (B,T) float 1/0, wconv[i] (3,C) tap-major,
 * out[i] (B,T/stride,C) bf16.  tlen (B,) int or NULL: rows t >= tlen[b] are treated as lying outside the sequence (zero
 * padding), which makes a batch of differently padded sequences equal to running each one alone. */
int vilco_dwconv_ln(const void* x, int x_dtype, const float* mask, const int* tlen, const float* const* wconv,
                    const float* const* lnw,
                    const float* const* lnb, void* const* out, int64_t out_lo, int n_out, int B, int T, int C, int stride,
                    float eps, void* stream);

/* nn.MaxPool1d(3, 2, 1) over time (TransformerBlock.pool_skip, blocks.py:519-525) on (B,T,C) fp32 -> (B,T/2,C). */
int vilco_maxpool3s2(const float* x, float* y, int B, int T, int C, void* stream);

/* o = a*x + b*y (fp32; y may be NULL), optional bf16 copy: the `t_c_alpha` mix (blocks.py:581) and dtype casts. */
int vilco_axpby(const float* x, const float* y, float a, float b, float* o32, void* o16, int64_t o16_lo, int64_t n,
                void* stream);

/* out[r,c] = x[r,c]*rowmul[r] + scale[c]*y[r,c] (fp32): `pool_skip(x)*mask + drop_path_attn(adapter branch)` of a
 * TransformerBlock that carries a parallel adapter (blocks.py:45-54, 564-567; meta_archs.py:139-148). */
int vilco_scale_add(const float* x, const float* rowmul, const float* y, const float* scale, float* out, int64_t rows, int C,
                    void* stream);

/* layout changes at the boundary: (B,C,T) fp32 reference layout -> (B,T_out,C) bf16 token-major (zero padded), and
 * (B,T,C) fp32 -> (B,C,T) fp32.  (PtTransformer.preprocessing, MQ/libs/modeling/meta_archs.py:1134-1181) */
int vilco_pack_feats(const float* x, void* y, int64_t y_lo, int B, int C, int T, int T_out, void* stream);
int vilco_unpack(const float* x, float* y, int B, int T, int C, void* stream);

/* Data path in front of the model: the `force_upsampling` resize of Ego4dCLDataset.__getitem__,
 * MQ/libs/datasets/ego4d.py:644-651 — F.interpolate(feats.permute(1,0)[None], size=max_seq_len, mode='linear',
 * align_corners=False) — on the features as stored on disk.  x: rows of C floats (token-major, the `.pt` layout of
 * ego4d.py:612); clip b owns rows row_start[b] .. row_start[b+1]-1 (device int64, B+1 entries, every clip >= 1 row).
 * Outputs, either may be NULL: out32 (B, T_out, C) fp32 and out16 (B, T_out, C) bf16 operand planes (lo plane at element
 * offset out16_lo, 0 = none) — the layout vilco_pack_feats produces, so the resized clip never exists in the reference's
 * (C, T) layout unless the caller asks for it (vilco_unpack). */
int vilco_resize_feats(const float* x, const int64_t* row_start, int B, int C, int T_out, float* out32, void* out16,
                       int64_t out16_lo, void* stream);

/* Row softmax over materialised attention scores S (Z2,Z1,Tq,Tk) fp32 -> P bf16 (row stride p_ld >= Tk, tail zeroed).
 * mode 0: keys with kmask[z2, j] == 0 get probability 0 (masked_fill(-inf) + softmax, blocks.py:258-260, 388-391).
 * mode 1: XLNet: (S[i,j] + BD[i, Tk + j - i]) * scale - 1e30 * [key j padded and i != j]
 *         (rel_shift_bnij + rel_attn_core, modeling_xlnet_x.py:256-304; mask prep :1152-1188); BD is (Z2,Z1,Tq,2Tk).
 * P32 (optional, row stride Tk): fp32 copy of the probabilities kept for the backward pass. */
int vilco_softmax_rows(const float* S, const float* BD, const float* kmask, void* P, int64_t p_lo, float* P32, int Z2, int Z1,
                       int Tq, int Tk, int64_t p_ld, float scale, int mode, void* stream);

/* LocalMaskedMHCA attention core (blocks.py:1038-1138, 1165-1200): q,k,v (B,T,C) bf16 token-major after the q/k/v
 * projections, window W (odd), out-of-range keys -inf, padded keys -1e4, padded queries -> 0; rel_pe (H,W) or NULL.
 * out (B,T,C) bf16 (heads side by side), to be followed by the proj GEMM. */
int vilco_local_attention(const void* q, const void* k, const void* v, const float* mask, const float* rel_pe, void* out,
                          int64_t lo, int B, int T, int C, int H, int W, void* stream);

/* Backward of vilco_local_attention (LocalMaskedMHCA core, blocks.py:1140-1200): dO (B,T,C) fp32 -> dq, dk, dv (B,T,C) fp32.
 * The probabilities are recomputed; scratch_p / scratch_ds are (B, H, T, W) fp32 work buffers (P and dS rows). The gradient
 * of rel_pe is not produced (use_rel_pe is off in every reference config). */
int vilco_local_attention_bwd(const float* dO, const void* q, const void* k, const void* v, int64_t lo, const float* mask,
                              const float* rel_pe, float* scratch_p, float* scratch_ds, float* dq, float* dk, float* dv,
                              int B, int T, int C, int H, int W, void* stream);

/* ChannelAttention core (blocks.py:423-436): qkv (B,T,3C) bf16 -> y (B,T,C) bf16;  G is a (B,H,64,64) fp32 scratch.
 * tlen (B,) or NULL limits the tokens summed in k^T v (the reference sums over every position of its padded batch). */
/* (y may be NULL: only G is computed) */
int vilco_channel_attention(const void* qkv, int64_t qkv_lo, float* G, void* y, int64_t y_lo, const int* tlen, int B, int T,
                            int C, int H, void* stream);

/* ------------------------------------------------------------------------------------
 * Decode of the head outputs into candidate segments — PtTransformer.inference_single_video,
 * MQ/libs/modeling/meta_archs.py:1594-1692: per pyramid level  prob = sigmoid(logit) * mask;  keep prob > pre_nms_thresh;
 * sort descending (ties: lower flat index first), keep the first `topk`;  pt = idx / K, cls = idx % K;
 * seg = (t - off_l * stride, t + off_r * stride) with t = pt * stride;  keep (right - left) > duration_thresh.
 * logits (B,P,K), offsets (B,P,2), pmask (B,P) fp32 over the concatenated pyramid rows; level l = rows
 * [lvl_off[l], lvl_off[l] + lvl_len[l]) (host arrays).  Output: per (video, level) a region of `topk` slots in
 * cand_* (B, n_levels*topk [,2]) holding cand_count[b, l] candidates in sorted order.
 * ------------------------------------------------------------------------------------ */
int vilco_decode(const float* logits, const float* offsets, const float* pmask, int B, int P, int K, int n_levels,
                 const int* lvl_off, const int* lvl_len, const float* lvl_stride, float pre_nms_thresh,
                 float duration_thresh, int topk, float* cand_segs, float* cand_scores, int* cand_labels, int* cand_count,
                 void* stream);

/* ------------------------------------------------------------------------------------
 * batched_nms — MQ/libs/utils/nms.py:103-190 with its native backend MQ/libs/utils/csrc/nms_cpu.cpp
 * (softnms_1d_cpu :67-160 for method 0/1/2 = SoftNMSop with method vanilla/linear/gaussian; method 3 = NMSop / nms_1d_cpu
 * :19-57, i.e. score pre-filter + greedy hard NMS).  Candidates of video b are the first region_count[b, r] entries of each
 * of its n_regions regions of region_cap slots (decode output layout; a flat array is one region).  Per class (ascending
 * original order) the reference's array algorithm is reproduced exactly, at most max_seg_num picks are kept per class,
 * classes are concatenated in ascending id, sorted by score (descending, ties by concatenation order) and cut to
 * max_seg_num.  multiclass = 0 runs one class-agnostic pass (follow it with vilco_seg_voting when voting_thresh > 0).
 * Outputs (B, max_seg_num [,2]) + out_count (B).  workspace: vilco_nms_workspace_bytes(...) device bytes.
 * ------------------------------------------------------------------------------------ */
size_t vilco_nms_workspace_bytes(int B, int n_regions, int region_cap, int num_classes, int det_cap);
int vilco_batched_nms(const float* segs, const float* scores, const int* labels, const int* region_count, int B,
                      int n_regions, int region_cap, int num_classes, int multiclass, int method, float iou_threshold,
                      float sigma, float min_score, int max_seg_num, void* workspace, size_t workspace_bytes,
                      float* out_segs, float* out_scores, long long* out_labels, int* out_count, void* stream);

/* Segment voting of the class-agnostic branch of batched_nms (MQ/libs/utils/nms.py:66-101, 174-181): every kept segment
 * (out_segs (B, max_seg_num, 2), first out_count[b] rows) is replaced in place by the score*IoU-weighted mean of all
 * candidates of its clip (same layout as vilco_batched_nms' inputs) with IoU >= voting_thresh. */
int vilco_seg_voting(float* out_segs, const int* out_count, const float* segs, const float* scores, const int* region_count,
                     int B, int n_regions, int region_cap, int max_seg_num, float voting_thresh, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused forward of the Moment-Query losses — PtTransformer.losses, MQ/libs/modeling/meta_archs.py:1374-1480:
 * sigmoid focal loss (alpha .25, gamma 2; MQ/libs/modeling/losses.py:5-51) summed over classes and weighted by the
 * gaussian weight (1 on negatives), DIoU (losses.py:109-168) on positives weighted by (w_l + w_r)/2 * w_cls, the number of
 * positives, and the label-involved loss  sum_{b,k} BCE(max_t softmax_K(logits)[b,k], present[b,k]).
 * All tensors over the concatenated pyramid rows (B,P,...); gap (P,) flags layout-only rows.  sums4 (device) receives
 * [cls_sum, reg_sum, num_pos, al_sum] (un-normalised; the caller divides by the EMA loss normaliser).
 * ------------------------------------------------------------------------------------ */
int vilco_mq_losses(const float* logits, const float* offsets, const float* pmask, const uint8_t* gap, const float* gt_cls,
                    const float* gt_off, const float* w_cls, const float* w_l, const float* w_r, const float* present, int B,
                    int P, int K, float alpha, float gamma, float* sums4, unsigned int* smax_scratch, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused masked attention core (head dim 64) — the `att = (q*scale) @ k^T; masked_fill(~kv_mask, -inf); softmax;
 * att @ v` of MaskedMHCA.forward (MQ/libs/modeling/blocks.py:386-396) and MaskedMHA.forward (blocks.py:255-263):
 * q (B,Tq,C), k/v (B,Tk,C) bf16 token-major (q_lo / kv_lo = lo-plane offsets, 0 for single plane), kmask (B,Tk) or NULL,
 * out (B,Tq,C) bf16 (+ lo plane).  Scores stay in TMEM; Tk <= 2048.  Fully masked rows produce zeros.
 * ------------------------------------------------------------------------------------ */
int vilco_attention(const void* q, int64_t q_lo, const void* k, const void* v, int64_t kv_lo, const float* kmask, void* out,
                    int64_t out_lo, int B, int H, int Tq, int Tk, int C, float scale, void* stream);

/* Fused XLNet relative attention (XLNetRelativeAttention.rel_attn_core + rel_shift_bnij, MQ/libs/modeling/modeling_xlnet_x.py:
 * 256-320) for single-plane operands, head dim 64, T a multiple of 128 (<= 2048):
 *   out[b, i, h*64:(h+1)*64] = softmax_j( ((qw_i . k_j) + (qr_i . kr[T + j - i])) * scale | key j visible ) @ v
 * qw = q + r_w_bias, qr = q + r_r_bias, k, v: (B, T, C) 16-bit planes; kr = pos_emb W_r: (2T, C) (no batch dim); a key j with
 * kmask[b, j] == 0 is visible only to query j itself (the reference's -1e30 non-self padding mask).  Scores, relative
 * shift and probabilities stay in TMEM / shared memory; out is one 16-bit plane (B, T, C). */
int vilco_xl_attention(const void* qw, const void* qr, const void* k, const void* v, const void* kr, const float* kmask,
                       void* out, int B, int H, int T, int C, float scale, void* stream);

/* FPN1D building blocks (MQ/libs/modeling/necks.py:13-106; ACConv / DenseAPP, MQ/libs/modeling/utils.py:671-751).
 * vilco_groupnorm: nn.GroupNorm(G, C) of a token-major fp32 (B, T, C) tensor — statistics per (clip, group) over all T rows and
 * the C / G channels of the group, eps inside the square root, affine w / b (C,), optional ReLU; writes fp32 y32 and / or the
 * 16-bit operand planes y16 (either may be NULL).
 * vilco_upsample2_add: the top-down path, y (B, T2, C) += x (B, T2 / 2, C) repeated twice along time
 * (F.interpolate(scale_factor=2, mode="nearest") + add, necks.py:88-93). */
int vilco_groupnorm(const float* x, const float* w, const float* b, float* y32, void* y16, int64_t y16_lo, int B, int T, int C,
                    int G, float eps, int relu, void* stream);
int vilco_upsample2_add(const float* x, float* y, int B, int T2, int C, void* stream);

/* Single-pass masked self-attention (MaskedMHCA core, MQ/libs/modeling/blocks.py:351-410: att = softmax(q k^T * scale with
 * masked_fill(~kv_mask, -inf)); out = att @ v) for single-plane operands, head dim 64, T a multiple of 128 (<= 2048) — the
 * kernel of vilco_xl_attention without the position branch (online softmax with lazy rescale, one pass over the keys).
 * q, k, v, out: (B, T, C) 16-bit planes; kmask (B, T) fp32, 0 = padded key, or NULL.  vilco_attention remains the general
 * entry (Tq != Tk, tails, two planes). */
int vilco_self_attention(const void* q, const void* k, const void* v, const float* kmask, void* out, int B, int H, int T, int C,
                         float scale, void* stream);
/* same, and lse2 (B, H, T) fp32 (may be NULL) receives log2(sum_j exp(score_ij)) = the row log-sum-exp in base 2 of the scaled,
 * masked scores — what the fused gradient epilogues of vilco_gemm (rowsub) need to recompute P without a softmax pass */
int vilco_self_attention_lse(const void* q, const void* k, const void* v, const float* kmask, void* out, float* lse2, int B, int H,
                             int T, int C, float scale, void* stream);

/* ------------------------------------------------------------------------------------
 * Backward-pass building blocks (training; token-major fp32 gradients).  The GEMM-shaped gradients reuse vilco_gemm:
 *   dX = dZ W      : A = dZ (K-major over the output channels), B = W as MN-major operand (b_major = 1)
 *   dW = dZ^T X    : A = dZ^T (vilco_to_planes writes the transposed planes), B = X as MN-major operand
 * ------------------------------------------------------------------------------------ */
/* y16[r,c] = x[r,c]*rowmul[r]*colmul[c] as 16-bit (hi, lo) planes and / or its transpose yT16[c,r] with row stride ldT
 * (either output may be NULL).  grad != 0: x is a gradient, the planes hold x * grad_scale (vilco_set_grad_scale). */
int vilco_to_planes(const float* x, const float* rowmul, const float* colmul, void* y, int64_t y_lo, void* yT, int64_t yT_lo,
                    int R, int C, int ldT, int Z, int grad, void* stream);   /* Z independent (R,C) matrices */
/* XLNet rel-shift backward (modeling_xlnet_x.py:256-268): dBD[z,i,p] = dS[z,i,p-T+i] inside the band, 0 elsewhere; every
 * element of the (Z,T,2T) output is written, as fp32 (dBD) and / or as bf16 operand planes (dBD16, lo plane at + dbd_lo). */
int vilco_relshift_bwd(const float* dS, float* dBD, void* dBD16, int64_t dbd_lo, int64_t Z, int T, void* stream);
/* y16[b,t,:] = x16[b,t+shift,:] (zero outside the clip) on bf16 planes: shifted operands of the k=3 conv weight gradient */
int vilco_shift_planes(const void* x, void* y, int64_t lo, int B, int T, int C, int shift, void* stream);
/* out[c] += sum_r x[r,c] * (y ? y[r,c] : 1) * (rowmul ? rowmul[r] : 1)   (bias / scale / affine gradients; out pre-zeroed) */
int vilco_colsum(const float* x, const float* y, const float* rowmul, float* out, int R, int C, void* stream);
/* LayerNorm backward of y = act(LN(x [+ add]) * w + b) (blocks.py:160-175): dx (also the gradient of `add`), dw, db
 * (accumulated; pre-zeroed by the caller).  y_relu = forward output when the ReLU was fused, else NULL. */
int vilco_layernorm_bwd(const float* x, const float* add, const float* w, const float* dy, const float* y_relu, float eps,
                        float* dx, float* dw, float* db, int rows, int C, void* stream);
/* depthwise k=3 conv * mask backward (MaskedConv1D with groups = C, blocks.py:106-130): dconv (B,T/stride,C) is the
 * gradient w.r.t. the masked conv output; dx (B,T,C) written or accumulated, dw (3,C) accumulated. */
int vilco_dwconv_bwd(const float* x, const float* mask, const float* w, const float* dconv, float* dx, float* dw, int B,
                     int T, int C, int stride, int accumulate_dx, void* stream);
/* fp32 depthwise k=3 conv * out-mask (the LayerNorm input of vilco_dwconv_ln, recomputed in the backward pass) */
int vilco_dwconv_fwd32(const float* x, const float* mask, const float* w, float* out, int B, int T, int C, int stride,
                       void* stream);
/* fp32 elementwise helper of the training path: op 0: x*rowmul[r]*colmul[c], 1: gelu(x), 2: relu(x), 3: x*(y>0).
 * out (fp32) and / or out16 (bf16 hi plane, lo plane at out16 + out16_lo when out16_lo != 0) receive the result. */
int vilco_ew(int op, const float* x, const float* y, const float* rowmul, const float* colmul, float* out, void* out16,
             int64_t out16_lo, int64_t rows, int C, void* stream);
int vilco_gelu_bwd(const float* x, const float* dy, float* dx, int64_t n, void* stream);
/* MaxPool1d(3,2,1) backward (TransformerBlock.pool_skip): dx must be pre-zeroed; gradient goes to the first maximum. */
int vilco_maxpool3s2_bwd(const float* x, const float* dy, float* dx, int B, int T, int C, void* stream);
/* dS[r,j] = scale * P[r,j] * (dP[r,j] - sum_k dP[r,k] P[r,k]).  P: fp32 rows (P32) or, when P32 is NULL, the bf16 hi/lo planes
 * of the forward softmax (P16, lo plane at + p_lo, row stride p_ld).  dS: fp32 (row stride Tk) and / or operand planes
 * (dS16, lo plane at + ds_lo, row stride ds_ld). */
int vilco_softmax_bwd(const float* P32, const void* P16, int64_t p_lo, int64_t p_ld, const float* dP, float* dS, void* dS16,
                      int64_t ds_lo, int64_t ds_ld, int64_t rows, int Tk, float scale, void* stream);

/* ChannelAttention core backward (blocks.py:423-436): dy (B,T,C) fp32, qkv planes and G (B,H,64,64) from the forward call
 * -> dqkv (B,T,3C) fp32; dA_scratch is a (B,H,64,64) fp32 work buffer. */
int vilco_channel_attention_bwd(const float* dy, const void* qkv, int64_t qkv_lo, const float* G, float* dA_scratch,
                                float* dqkv, int B, int T, int C, int H, void* stream);

/* Inverted dropout (nn.Dropout in training mode — blocks.py:222-223, 536-538): out = keep(seed, i) ? x / (1 - p) : 0 with a
 * counter-based generator; calling it on the gradient with the same seed is the backward pass.  Outputs as in vilco_ew. */
int vilco_dropout(const float* x, float* out, void* out16, int64_t out16_lo, int64_t n, float p, uint64_t seed, void* stream);

/* ------------------------------------------------------------------------------------
 * Optimizer step over flat fp32 buffers (all parameters / gradients / AdamW moments of a parameter group contiguous):
 *   vilco_grad_clip_coef : coef = min(1, max_norm / (||g||_2 + 1e-6))   — torch.nn.utils.clip_grad_norm_ as called by
 *                          train_one_epoch (MQ/libs/utils/train_utils.py:345-349); max_norm <= 0 gives coef = 1.
 *                          Deterministic (no floating-point atomics: data-parallel replicas get bit-identical
 *                          coefficients).  scratch: VILCO_CLIP_SCRATCH device floats; coef / norm_out: single device floats.
 *   vilco_adamw          : torch.optim.AdamW update (decoupled weight decay, bias correction; make_optimizer,
 *                          train_utils.py:124-140) with the gradient scaled by *grad_scale (device scalar or NULL).  When
 *                          `planes` is given the updated parameter is also written as bf16 hi (planes[i]) and lo
 *                          (planes[planes_lo + i], skipped when planes_lo == 0) operand planes for the GEMM kernels.
 * ------------------------------------------------------------------------------------ */
#define VILCO_CLIP_SCRATCH 1184
int vilco_grad_clip_coef(const float* g, int64_t n, float max_norm, float* scratch, float* coef, float* norm_out, void* stream);
int vilco_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                float weight_decay, int step, const float* grad_scale, void* planes, int64_t planes_lo, void* stream);

/* Residual branch of a TransformerBlock in training mode, fused (blocks.py:404-405, 567-585, drop_path :640-652):
 *   out[r,c] = resid[r,c]*rm[r] + scale[c] * dropout_p(y[r,c] + bias[c]) * ymul[r]
 * with y the raw projection GEMM output, ymul[r] = out-mask * stochastic-depth factor of the row's sample, dropout from the
 * counter-based generator of vilco_dropout (p = 0: none).  rm / bias / scale / ymul may be NULL (= 1 / 0 / 1 / 1).
 * The backward writes d resid = g*rm (optional), dy as bf16 operand planes (the dZ of the projection's gradient GEMMs) and
 * accumulates dbias[c] += sum_r dy, dscale[c] += sum_r (d out * ymul * keep) * (y + bias)  (either may be NULL). */
int vilco_resid_branch_fwd(const float* resid, const float* rm, const float* y, const float* bias, const float* scale,
                           const float* ymul, float* out, int64_t rows, int C, float p, uint64_t seed, void* stream);
int vilco_resid_branch_bwd(const float* g, const float* rm, const float* y, const float* bias, const float* scale,
                           const float* ymul, float* dresid, void* dy16, int64_t dy_lo, float* dbias, float* dscale, int R, int C,
                           float p, uint64_t seed, void* stream);

/* Backward of vilco_mq_losses for final = cls + w_reg*reg + w_al*al (each divided by `norm`): gradients w.r.t. the logits,
 * the offsets and the three gaussian weights (the latter feed torch autograd of the target-assignment glue, which owns
 * mu / sigma).  smax = the (B,K) scratch filled by the forward call. */
int vilco_mq_losses_bwd(const float* logits, const float* offsets, const float* pmask, const uint8_t* gap, const float* gt_cls,
                        const float* gt_off, const float* w_cls, const float* w_l, const float* w_r, const float* present,
                        const unsigned int* smax, int B, int P, int K, float alpha, float gamma, float norm, float w_reg,
                        float w_al, float* dlogits, float* doffsets, float* dw_cls, float* dw_l, float* dw_r, void* stream);

/* ------------------------------------------------------------------------------------
 * Evaluation tail (host function, no kernel launch; all pointers are HOST pointers): the greedy matching loop of
 * compute_average_precision_detection, MQ/libs/utils/metrics.py:277-346 (tIoU = segment_iou :348-372, float64) for the
 * predictions and ground truths of ONE label.  pred_seg (n_pred, 2) sorted by descending score; pred_vid[i] = dense index
 * of the prediction's video among the videos that own ground truth of this label, -1 if it has none (false positive at
 * every threshold, :308-312); ground-truth rows are grouped by video in their original order: video v owns rows
 * gt_start[v] .. gt_start[v+1]-1 of gt_seg (n_gt, 2).  Output tp (n_thr, n_pred) bytes: 1 = true positive; every other
 * prediction is a false positive (fp = 1 - tp).
 * ------------------------------------------------------------------------------------ */
int vilco_ap_match(const double* pred_seg, const int64_t* pred_vid, int64_t n_pred, const double* gt_seg,
                   const int64_t* gt_start, int64_t n_vid, const double* tiou_thr, int n_thr, uint8_t* tp);

#ifdef __cplusplus
}
#endif
#endif /* VILCO_B200_H */
