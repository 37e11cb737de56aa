/* univs_b200 -- C ABI of the B200 (sm_100a) hot-path kernels of the UniVS per-clip forward.
 *
 * Every entry point is stream-ordered (no internal synchronisation, no allocation), takes
 * plain device pointers + sizes, borrows its inputs and writes into caller-owned outputs.
 * Return value: 0 on success, a negative UNIVS_E_* code otherwise; univs_b200_last_error()
 * returns a human-readable message for the last failure on the calling thread.
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *
 * Reference interfaces replaced (paths relative to the MinghanLi/UniVS tree):
 *   mask2former/modeling/pixel_decoder/ops/src/ms_deform_attn.h:25-44      (ms_deform_attn_forward)
 *   mask2former/modeling/pixel_decoder/ops/src/cuda/ms_deform_im2col_cuda.cuh:928-959
 *   mask2former/modeling/backbone/swin.py:131-171, 247-289                  (window attention + addressing)
 *   univs/modeling/transformer_decoder/video_mask2former_transformer_decoder_univs.py:527  (mask einsum)
 *   ...:555-566 (attention-mask generation), :390 (fully-blocked-row rule)
 *   univs/modeling/transformer_decoder/transformer_layers.py:34-44, 95-115  (MHA cores)
 *   ...video_mask2former_transformer_decoder_univs.py:456-496                (ProCA)
 */
#ifndef UNIVS_B200_H_
#define UNIVS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNIVS_OK 0
#define UNIVS_E_BADARG (-1)   /* invalid size / null pointer / unsupported shape      */
#define UNIVS_E_LAUNCH (-2)   /* cudaGetLastError() != cudaSuccess after the launch   */
#define UNIVS_E_NOTIMPL (-3)  /* entry point exported for ABI parity, not implemented */

/* precision of the tensor-core contractions inside the attention kernels */
#define UNIVS_PREC_TF32X3 0   /* 3xTF32 split (fp32-equivalent products), default */
#define UNIVS_PREC_TF32 1     /* single TF32 product, round-to-nearest operands   */

const char* univs_b200_last_error(void);
int univs_b200_abi_version(void);

/* ---- MSDeformAttn forward (a7).  Mirrors ms_deformable_im2col_cuda (ms_deform_im2col_cuda.cuh:928-942):
 * value [N,S,M,D] f32, spatial_shapes [L,2] i64 (H,W), level_start_index [L] i64,
 * sampling_loc [N,Lq,M,L,P,2] f32 (x,y in [0,1]), attn_weight [N,Lq,M,L,P] f32 -> out [N,Lq,M*D] f32. */
int univs_ms_deform_attn_forward_f32(void* stream, const float* value, const int64_t* spatial_shapes,
                                     const int64_t* level_start_index, const float* sampling_loc,
                                     const float* attn_weight, int batch, int spatial_size, int num_heads,
                                     int channels, int num_levels, int num_query, int num_point, float* out);

/* Exported for ABI parity with ms_deform_attn_backward (ms_deform_attn.h:46-66); training is out of scope. */
int univs_ms_deform_attn_backward_f32(void);

/* ---- MSDeformAttn forward with the encoder front end fused (ms_deform_attn.py:98-117): queries are the
 * pixels of the level pyramid (Lq == S, reference point = pixel centre, msdeformattn.py:143-155);
 * offs_logits [N,S,M*L*P*3] = [sampling_offsets(q) (M,L,P,2) | attention_weights(q) (M,L*P)] raw linear outputs;
 * the kernel applies softmax over L*P and loc = ref + off/(W_l,H_l).  D must be 32. */
int univs_ms_deform_attn_encoder_f32(void* stream, const float* value, const int64_t* spatial_shapes,
                                     const int64_t* level_start_index, const float* offs_logits, int batch,
                                     int spatial_size, int num_heads, int num_levels, int num_point, float* out);
/* Same operation, same arithmetic per (frame, query, head) -- results are bit-identical -- with a different work
 * distribution: one CTA = a tile_width x (32 / tile_width) block of neighbouring queries of one head, so that the
 * overlapping sampling footprints are served from L1 (tile_width: power of two <= 32).
 * Optional fusions (all nullable / 0): value_bias [heads*32] = bias of value_proj, applied to the in-bounds samples as
 * the reference's zero padding requires, so that GEMM runs bias-free; offs_logits_bias [heads*L*P*3] = biases of the
 * sampling_offsets / attention_weights linears; split != 0 writes `out` as a GEMM operand (see the row-wise kernels). */
int univs_ms_deform_attn_encoder_tiled_f32(void* stream, const float* value, const int64_t* spatial_shapes,
                                           const int64_t* level_start_index, const float* offs_logits, int batch,
                                           int spatial_size, int num_heads, int num_levels, int num_point,
                                           int tile_width, const float* value_bias, const float* offs_logits_bias,
                                           int split, float* out);

/* ---- Swin (shifted-)window attention, addressing folded in (a2,a3).
 * qkv [B,H,W,3C] f32 = LN(x) * Wqkv^T on the unpadded token grid WITHOUT the bias; qkv_bias [3C] is added by the kernel
 * to every token (and is the whole value of a pad token: zero-padded after norm1, swin.py:247-255);
 * rel_bias_table [(2*window-1)^2, num_heads]; head_dim = C/num_heads must be 32; window in {7,12} (or any <=12);
 * shift = 0 or window/2.  out [B,H,W,C] f32 (pre output projection).  precision UNIVS_PREC_TF32X3 runs the fp16 hi|lo
 * split kernel (m16n8k16, fp32-equivalent products), UNIVS_PREC_TF32 the single-pass TF32 kernel. */
int univs_swin_window_attention_f32(void* stream, const float* qkv, const float* qkv_bias,
                                    const float* rel_bias_table, int batch, int height, int width, int channels,
                                    int num_heads, int window, int shift, int precision, float* out);

/* Same operator (strict precision), writing its result directly as the A operand of the fp16x3 projection GEMM:
 * out16 = __half [B,H,W,3C] = [lo*2^11 | hi*2^-11 | hi] (channels <= 1536). */
int univs_swin_window_attention_f16x3out(void* stream, const float* qkv, const float* qkv_bias,
                                         const float* rel_bias_table, int batch, int height, int width, int channels,
                                         int num_heads, int window, int shift, void* out16);

/* Same operator for 12x12 windows (Swin-B / Swin-L) with both contractions on the tcgen05 tensor cores: one persistent
 * CTA per SM, fp16 hi|lo operand tiles in shared memory, S and O accumulated in TMEM, softmax on rows read back with
 * tcgen05.ld (swin_window_attn_tc.cu).  Strict precision only (same arithmetic contract as UNIVS_PREC_TF32X3 above).
 * out (f32 [B,H,W,C]) and out16 (__half [B,H,W,3C] operand layout as above, channels <= 1536) are both optional, at
 * least one must be given.  flags bit 0: the second version of the kernel (swin_window_attn_tc2.cu: the 16 tail rows of
 * a window as a second row tile served by two warps with the main warps' code, conflict-free bias table, incremental
 * unit decoding); bit 1 (needs bit 0): out16 is the compact operand __half [B,H,W,2C] = [hi | lo*2^11] that
 * univs_gemm_f16x3_tc consumes (no hi*2^-11 block; any channel count).  debug_scores: NULL, or f32
 * [B*windows*heads, 144, 144] receiving the biased, masked scores before the softmax (diagnostics). */
int univs_swin_window_attention_tc(void* stream, const float* qkv, const float* qkv_bias, const float* rel_bias_table,
                                   int batch, int height, int width, int channels, int num_heads, int window, int shift,
                                   int flags, float* out, void* out16, float* debug_scores);

/* ---- Mask einsum "btqc,btchw->btqhw" + transpose(1,2) (a11), on the tcgen05 tensor cores
 * (TMA -> smem -> tcgen05.mma kind::tf32 -> TMEM -> tcgen05.ld -> coalesced stores).
 * mask_embed [T,Q,C] f32; mask_features channel-last [T,HW,C] f32; out [Q,T,HW] f32.  C % 32 == 0, Q <= 256,
 * operands 16-byte aligned.  The tensor core consumes the upper 19 bits of each fp32 operand (truncation) and
 * accumulates in fp32: pass operands through univs_round_tf32_f32 first for round-to-nearest behaviour. */
int univs_mask_einsum_f32(void* stream, const float* mask_embed, const float* mask_features_cl, int frames,
                          int queries, int channels, int pixels, float* out);

/* fp32-equivalent variant on the tensor cores: operands are fp16 [hi | lo] pairs (univs_split_tf32_f32 with
 * chunk = UNIVS_SPLIT_F16U): mask_embed16 [T,Q,2C], mask_features16 [T,HW,2C] (__half); three tcgen05.mma kind::f16 per
 * k-step (lo*hi + hi*lo + hi*hi, fp32 accumulate in TMEM).  Same bytes per element as the fp32 tensors.  C % 64 == 0. */
int univs_mask_einsum_f16x3(void* stream, const void* mask_embed16, const void* mask_features16, int frames, int queries,
                            int channels, int pixels, float* out);

/* Cluster variant of univs_mask_einsum_f16x3 (csrc/mask_einsum_mc.cu; same operands, same result up to the fp32
 * accumulation order, which is identical per output element): two CTAs of a cluster split the queries, each keeps its half
 * of mask_embed16 resident in shared memory for a whole frame, mask_features16 tiles reach both through TMA multicast.
 * 16 < Q <= 256, C % 32 == 0 and C <= 256.  Opt-in (never run on hardware in round 1). */
int univs_mask_einsum_f16x3_cluster(void* stream, const void* mask_embed16, const void* mask_features16, int frames,
                                    int queries, int channels, int pixels, float* out);

/* Same contraction with register operands (mma.sync): precision UNIVS_PREC_TF32 rounds operands to nearest TF32
 * in-kernel (bit-identical to the tcgen05 kernel on pre-rounded operands); UNIVS_PREC_TF32X3 uses the 3xTF32 split
 * (fp32-equivalent products) -- the strict-parity policy and the on-device cross-check of the tcgen05 kernel. */
int univs_mask_einsum_mma_f32(void* stream, const float* mask_embed, const float* mask_features_cl, int frames,
                              int queries, int channels, int pixels, int precision, float* out);

/* ---- Attention-mask bits from mask logits (a11, ..._univs.py:555-566 + :390).
 * logits [Q,T,H,W] f32; target (h,w) with H % h == 0 and W % w == 0 and even ratios (bilinear
 * align_corners=False then equals the mean of the 2x2 centre pixels of each cell);
 * bits [T,Q,words] u32, words = ceil(h*w/32), bit k of word j set <=> key 32*j+k is BLOCKED (sigmoid<0.5);
 * row_open [T,Q] i32 = 1 if at least one key of the row is allowed. */
int univs_attn_mask_bits_f32(void* stream, const float* logits, int queries, int frames, int height, int width,
                             int tgt_h, int tgt_w, uint32_t* bits, int32_t* row_open);

/* Pooled-feature variant of the same step (decoder_glue.cu): the resize is linear, so resize(E . F) = E . resize(F).
 * univs_mask_feature_pool_f32 pools the channel-last mask features [T,H,W,C] once per clip to a memory size (even integer
 * ratios: mean of the centre 2x2 block of each cell, the bilinear weights of align_corners=False) into [T, h*w, C] plain
 * f32 (split 0) or the fp16 [hi|lo] einsum operand (split UNIVS_SPLIT_F16U); the mask einsum on it yields the logits at
 * the memory resolution and univs_attn_mask_bits_direct_f32 turns those ([Q,T,S]) into the bits / row flags above. */
int univs_mask_feature_pool_f32(void* stream, const float* feats_cl, int frames, int height, int width, int channels,
                                int tgt_h, int tgt_w, void* out, int split);
int univs_attn_mask_bits_direct_f32(void* stream, const float* logits, int queries, int frames, int keys, uint32_t* bits,
                                    int32_t* row_open);

/* ---- Multi-head attention core (a12 masked cross-attention, a13 spatio-temporal self-attention).
 * q [B,Lq,C], k,v [B,Lk,C] f32, already in-projected (q unscaled), head_dim 32, heads = C/32.
 * mask_bits (nullable) [Bm,Lq,ceil(Lk/32)] u32, Bm in {1,B}, bit set = blocked;
 * row_open (nullable) [Bm,Lq] i32: rows with 0 ignore the mask (..._univs.py:390).
 * workspace: >= univs_mha_workspace_bytes(...) bytes of device scratch (split-K partials).
 * out [B,Lq,C] f32 (pre output projection). */
int64_t univs_mha_workspace_bytes(int batch, int len_q, int len_k, int channels);
int univs_mha_forward_f32(void* stream, const float* q, const float* k, const float* v, const uint32_t* mask_bits,
                          const int32_t* row_open, int mask_batch, int batch, int len_q, int len_k, int channels,
                          int precision, void* workspace, float* out);

/* Same operator, strict precision, on the tcgen05 tensor cores for the decoder's cross-attention shape: len_q <= 256
 * queries against a long key sequence (mha_tc.cu: S and the per-block O in TMEM, softmax on rows read back with
 * tcgen05.ld, running output in registers).  Same arguments, mask semantics and split-K partial format; the workspace
 * size comes from univs_mha_tc_workspace_bytes (its key-split plan differs: one CTA per SM).  flags bit 0 (diagnostic):
 * stage V transposed and read it through the K-major B descriptor instead of the MN-major one. */
int64_t univs_mha_tc_workspace_bytes(int batch, int len_q, int len_k, int channels);
int univs_mha_tc_forward_f32(void* stream, const float* q, const float* k, const float* v, const uint32_t* mask_bits,
                             const int32_t* row_open, int mask_batch, int batch, int len_q, int len_k, int channels,
                             int flags, void* workspace, float* out);

/* ---- Dense layer with fused epilogue on the tcgen05 tensor cores (gemm_tc.cu).  Replaces nn.Linear +
 * the activation / residual behind it (swin.py:35-41, 138-169, 292-293; ms_deform_attn.py:98-120; transformer_layers.py):
 *   y[t, n] = act(alpha * sum_k x[t, k] * w[n, k] + bias[n]) + addend[t, n]
 * Operands are fp16 pairs, value = hi + lo' * 2^-11 (lo' = the residual scaled by 2^11): x16 rows [tokens, ldx] halfs with the
 * hi block at column x_hi_off and the lo' block at x_lo_off (k columns each); w16 rows [channels, ldw] likewise (weights
 * pre-scaled by a power of two that `alpha` undoes).  Both the [lo' | hi*2^-11 | hi] K-chunk container of the row-wise kernels
 * (split = -Kc; offsets 2*Kc and 0) and the compact [hi | lo'] container written by this function fit.  k % 8 == 0; one
 * call accumulates at most one K-chunk (callers slice k > 1536 and pass the partial result as `addend`).
 * Outputs (each nullable, at least one): out f32 [tokens, ldo]; out16 = y as the next layer's operand, fp16 [tokens, ld16] with
 * hi at column n and lo' at column out16_lo_off + n.  addend f32 [tokens, ldadd] may alias out.
 * activation: 0 none, 1 GELU (erf), 2 ReLU.  bias f32 [channels] nullable. */
int univs_gemm_f16x3_tc(void* stream, const void* x16, int64_t ldx, int64_t x_hi_off, int64_t x_lo_off, const void* w16,
                        int64_t ldw, int64_t w_hi_off, int64_t w_lo_off, int64_t tokens, int channels, int k, float alpha,
                        const float* bias, const float* addend, int64_t ldadd, float* out, int64_t ldo, void* out16,
                        int64_t ld16, int64_t out16_lo_off, int activation);
/* The same with shifted-row taps: a k x k "same" convolution over a zero-padded channel-last activation viewed as a token
 * matrix is sum_t X[m + tap_row_offsets[t], :] * W_t^T (msdeformattn.py:345-360, the 3x3 FPN output convolutions), accumulated in
 * ONE tensor-core chain instead of k*k GEMMs that re-read and re-write the fp32 result:
 *   y[m, n] = act(alpha * sum_t sum_kk x[m + tap_row_offsets[t], kk] * w[n, t * k_tap + kk] + bias[n]) + addend[m, n]
 * x16 has x_rows rows (rows beyond read as zeros); the hi / lo' blocks of w16 are taps * k_tap columns wide, tap-major.
 * taps <= 9, k_tap % 64 == 0 when taps > 1, taps * k_tap <= 1536 per call (callers group the taps and pass `addend`);
 * tap_row_offsets is a HOST array.  taps == 1 with offset 0 is univs_gemm_f16x3_tc. */
int univs_gemm_f16x3_tc_taps(void* stream, const void* x16, int64_t ldx, int64_t x_hi_off, int64_t x_lo_off, int64_t x_rows,
                             const void* w16, int64_t ldw, int64_t w_hi_off, int64_t w_lo_off, int64_t tokens, int channels,
                             int k_tap, int taps, const int64_t* tap_row_offsets, float alpha, const float* bias,
                             const float* addend, int64_t ldadd, float* out, int64_t ldo, void* out16, int64_t ld16,
                             int64_t out16_lo_off, int activation);

/* ---- ProCA attention core (a14): every (prompt p, frame t) query attends to its own token and its L
 * prompt-memory tokens.  q,k_self,v_self [P,T,C]; k_mem,v_mem [P,Tm,L,C], Tm in {1,T}; out [P,T,C]. */
int univs_proca_forward_f32(void* stream, const float* q, const float* k_self, const float* v_self,
                            const float* k_mem, const float* v_mem, int prompts, int frames, int mem_frames,
                            int mem_len, int channels, float* out);

/* ---- fused row-wise kernels around the library GEMMs -------------------------------------------------------
 * `split` = 0 writes plain [rows, C]; `split` = Kc > 0 (Kc | C, Kc % 4 == 0) writes [rows, 2*C] in K-chunks of Kc
 * columns, chunk c = [hi_c | lo_c] with hi = the value rounded to nearest TF32 and lo = x - hi: the operand format
 * of the 3xTF32 (fp32-equivalent) GEMM policy, X*W^T ~= sum_c ([Xh|Xl]_c*[Wh|Wh]_c^T + Xh_c*Wl_c^T).  K is chunked
 * because the tensor cores accumulate with truncation: the bias grows linearly with the accumulation chain
 * (tools/accum_probe.py), so chains are kept <= Kc/8 MMA steps and chunks are summed in the fp32 epilogue.
 * `split` = UNIVS_SPLIT_F16U writes fp16 [rows, 2*C] = [hi | lo]; `split` = -Kc writes fp16 [rows, 3*C] in K-chunks
 * [lo*2^11 | hi*2^-11 | hi] (hi = fp16(x) saturating, lo = x - hi): fp16 has TF32's 11-bit significand at twice the
 * tensor-core rate; X*W^T = [Xl' | Xh_s | Xh][Wh_s | Wl' | Wh]^T is ONE GEMM (correction terms first).  `out` is
 * then a __half buffer.
 * layernorm: s = x (+ residual (+ residual_bias), nullable); sum_out (nullable) = s; out = LayerNorm(s)*gamma+beta (nn.LayerNorm,
 *   e.g. swin.py:246,292).  channels % 4 == 0, <= 4096.
 * gelu: exact erf GELU (nn.GELU default, swin.py:24-41). */
#define UNIVS_SPLIT_F16U (-2)  /* fp16 [rows,2C] = [hi | lo]  (operands of the fp16 mask einsum)                            */
#define UNIVS_SPLIT_F16C (-3)  /* fp16 [rows,2C] = [hi | lo*2^11]  (compact operand of univs_gemm_f16x3_tc: 4 bytes / element) */
/* split = -Kc (Kc >= 4, Kc | C): fp16 [rows,3C] in K-chunks [lo*2^11 | hi*2^-11 | hi]  (A operand of the fp16x3 GEMM) */
int univs_layernorm_f32(void* stream, const float* x, const float* residual, const float* residual_bias,
                        const float* gamma, const float* beta, int64_t rows, int channels, float eps, float* sum_out,
                        float* out, int split);
/* bias (nullable, [channels]) = the bias of the GEMM that produced x (resp. `residual`), deferred into these kernels */
int univs_gelu_f32(void* stream, const float* x, const float* bias, int64_t rows, int channels, float* out, int split);
int univs_relu_f32(void* stream, const float* x, const float* bias, int64_t rows, int channels, float* out, int split);
int univs_split_tf32_f32(void* stream, const float* x, int64_t rows, int channels, int chunk, float* out);

/* ---- channel-last GroupNorm fused with the FPN glue of the pixel decoder (detectron2 Conv2d = conv -> GroupNorm ->
 * [ReLU], msdeformattn.py:218-221, :249-283; top-down add of the bilinearly upsampled coarser level, :345-354).
 * x is addressed as x[n*img_stride + y*row_stride + xw*channels + c] (elements), so it may be the convolution output
 * inside a spatially padded row buffer.  stats: mean_rstd [frames, groups, 2] f32 = (mean, 1/sqrt(var + eps)) per frame
 * and group (biased variance, as nn.GroupNorm); workspace >= univs_groupnorm_workspace_bytes(...) bytes of device scratch.
 * apply: y = (x - mean) * rstd * gamma + beta  [+ bilinear resize of lowres [frames, low_height, low_width, channels] (frame n at
 * lowres + n*lowres_img_stride) to (height, width), align_corners = False]  [ReLU if relu != 0], written to out_f32 [frames, height, width, channels]
 * (nullable) and / or to out_split (nullable) in the operand format `split` (see above) at the zero-padded position
 * (y + pad, xw + pad) of a [frames, height + 2*pad, width + 2*pad, .] token matrix whose border the caller keeps zero.
 * Requirements: channels % groups == 0, (channels / groups) % 4 == 0, 256 % (channels / 4) == 0. */
int64_t univs_groupnorm_workspace_bytes(int frames, int height, int width, int groups);
int univs_groupnorm_stats_f32(void* stream, const float* x, int frames, int height, int width, int channels,
                              int64_t img_stride, int64_t row_stride, int groups, float eps, void* workspace,
                              float* mean_rstd);
int univs_groupnorm_apply_f32(void* stream, const float* x, int frames, int height, int width, int channels,
                              int64_t img_stride, int64_t row_stride, const float* mean_rstd, const float* gamma,
                              const float* beta, int groups, const float* lowres, int64_t lowres_img_stride, int low_height,
                              int low_width, int relu, float* out_f32, void* out_split, int split, int pad);

/* ---- gather-fused kernels at the edges of the Swin stages ------------------------------------------------------
 * patchify_normalize: frames [num_frames, 3, height, width] (uint8 if is_uint8 else f32, values 0..255) -> the patch
 *   matrix [num_frames * (padded_height/4) * (padded_width/4), 48] (column = c*16 + ky*4 + kx, the flattening of the
 *   PatchEmbed.proj weight [E,3,4,4], swin.py:456-495) of the normalised ((x - mean[c]) / std[c], univs_prompt.py:165-168)
 *   and zero-padded frames, plain f32 or in the operand format `split`.  mean3 / std3 are HOST arrays of 3 floats.
 * layernorm_merge2x2: PatchMerging (swin.py:298-337) without the concatenated copy: x [num_frames, height, width,
 *   channels] -> LayerNorm over [x(2i,2j) | x(2i+1,2j) | x(2i,2j+1) | x(2i+1,2j+1)] (zeros beyond odd borders),
 *   out [num_frames * ceil(height/2) * ceil(width/2), 4*channels] plain or split.  4*channels <= 4096. */
int univs_patchify_normalize(void* stream, const void* frames, int is_uint8, int num_frames, int height, int width,
                             int padded_height, int padded_width, int patch, const float* mean3, const float* std3,
                             float* out, int split);
int univs_layernorm_merge2x2_f32(void* stream, const float* x, int num_frames, int height, int width, int channels,
                                 const float* gamma, const float* beta, float eps, float* out, int split);
/* LayerNorm with several consumers (post-norm layers, msdeformattn.py:126-133): y = LN(x (+ residual (+ residual_bias)));
 * out_f32 (nullable) = y; out_split (nullable) = y as a GEMM operand (`split`); out_split_pos (nullable) = (y +
 * pos[row % pos_rows]) as a GEMM operand -- the query of the next deformable attention (src + pos, ms_deform_attn.py). */
int univs_layernorm_multi_f32(void* stream, const float* x, const float* residual, const float* residual_bias,
                              const float* gamma, const float* beta, int64_t rows, int channels, float eps, float* out_f32,
                              void* out_split, int split, const float* pos, int64_t pos_rows, void* out_split_pos);

/* ---- helpers ---- */
/* in-place/out-of-place round-to-nearest-even to TF32 (19-bit) of n floats */
int univs_round_tf32_f32(void* stream, const float* in, float* out, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* UNIVS_B200_H_ */
