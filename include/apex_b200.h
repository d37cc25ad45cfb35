/*
 * apex_b200.h -- C ABI of libapex_b200.so, the sm_100a (B200) implementation of the denoising hot path of
 * Apex Studio's generation server (reference: totokunda/apex-studio, apps/api).
 *
 * The reference has no native ABI on this path: it is Python calling torch.  Each entry point below
 * replaces one Python-level call site of the reference (cited as file:line under apps/api/src/), taking
 * what that call site holds at that moment -- device pointers, shapes, element strides, scalars -- plus
 * the CUDA stream to launch on.  No torch types, no allocation inside, no global state besides lazily
 * resolved driver entry points.  Every function returns 0 (B200_OK) or a negative error code and never
 * throws; the Python shim (apex-studio_b200/ops.py) turns codes into the exceptions the reference raises
 * (ValueError for shape/dtype/alignment problems, RuntimeError otherwise).
 *
 * All tensors are bf16 unless stated otherwise; "stride" arguments are in ELEMENTS.
 * `stream` is a cudaStream_t passed as void*.
 */
#ifndef APEX_B200_H_
#define APEX_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_SHAPE (-1)  /* unsupported shape (e.g. head_dim != 128)            -> ValueError   */
#define B200_ERR_ALIGN (-2)  /* pointer/stride not 16-byte aligned                   -> ValueError   */
#define B200_ERR_DRIVER (-3) /* CUDA driver entry point unavailable (no GPU/driver)  -> RuntimeError */
#define B200_ERR_TMAP (-4)   /* cuTensorMapEncodeTiled rejected the layout           -> RuntimeError */
#define B200_ERR_LAUNCH (-5) /* kernel launch failed                                 -> RuntimeError */
#define B200_ERR_ARG (-6)    /* null pointer / bad enum                              -> ValueError   */

/* Library / ABI version (major*100+minor). */
int b200_version(void);

/* Human-readable message for an error code (static storage). */
const char* b200_strerror(int code);

/*
 * Attention core: o = softmax(q k^T * scale) v, non-causal, no mask, no dropout.
 * Replaces attention_register.call(q, k, v, ...) -- attention/functions.py:84,338-377 (`sdpa` is the gold
 * backend), called from transformer/wan/base/attention.py:397 (self and cross attention) and from every
 * other DiT family (flux/base/attention.py:89, qwenimage/base/attention.py:138, hunyuanvideo15/base/model.py:150).
 *   q: [B,H,Sq,D]  k,v: [B,H,Sk,D]  o: [B,H,Sq,D]   D must be 128; last dim contiguous; every other stride
 *   arbitrary but a multiple of 8 elements (16 bytes).  fp32 softmax / accumulation, P rounded to bf16
 *   before P*V (as flash kernels do).
 */
int b200_attn_fwd(const void* q, const void* k, const void* v, void* o, int B, int H, int Sq, int Sk, int D,
                  int64_t q_sb, int64_t q_sh, int64_t q_ss, int64_t k_sb, int64_t k_sh, int64_t k_ss,
                  int64_t v_sb, int64_t v_sh, int64_t v_ss, int64_t o_sb, int64_t o_sh, int64_t o_ss,
                  float scale, void* stream);

/*
 * b200_attn_fwd for a query whose RMS-norm over all H * 128 channels of a token (norm_q of the Wan cross-attention,
 * attention.py:345-370 / efficiency/mod.py:24-35) is folded in: the logits of query row r are multiplied by
 * rstd[r] = bf16(rsqrt(sum_i q_rowsumsq[(b * Sq + r) * q_parts + i] / q_norm_dim + q_eps)); q itself is the projection output
 * already multiplied by the norm weight (b200_linear_normw, which also writes q_rowsumsq).
 */
int b200_attn_fwd_qnorm(const void* q, const void* k, const void* v, void* o, int B, int H, int Sq, int Sk, int D,
                        int64_t q_sb, int64_t q_sh, int64_t q_ss, int64_t k_sb, int64_t k_sh, int64_t k_ss,
                        int64_t v_sb, int64_t v_sh, int64_t v_ss, int64_t o_sb, int64_t o_sh, int64_t o_ss,
                        float scale, const float* q_rowsumsq, int q_parts, int q_norm_dim, float q_eps, void* stream);

/*
 * Linear layer with fused epilogue: C = epilogue(A @ W^T + bias).
 *   A: [M,K] row stride lda;  W: [N,K] row stride ldw (nn.Linear weight layout);  bias: [N] or NULL.
 * epilogue:
 *   B200_EPI_BIAS      C[M,N] (ldc) = acc + bias                         attention.py:345-347,407 (to_q/k/v, to_out)
 *   B200_EPI_GELU_TANH C = gelu_tanh(acc + bias)                         diffusers FeedForward("gelu-approximate"), model.py:1062,1270
 *   B200_EPI_GATE_RES  C (read-modify-write, the residual stream) += gate[n] * (acc + bias); gate NULL => 1
 *                                                                        model.py:1212-1213, 1245-1251, 1278-1279
 *   B200_EPI_BIAS_F32  C is float32 [M,N] (ldc in float elements) = acc + bias (output head / embedders)
 * K must be a multiple of 8; lda, ldw multiples of 8; ldc multiple of 8 (bf16) / 4 (f32); pointers 16 B aligned.
 */
#define B200_EPI_BIAS 0
#define B200_EPI_GELU_TANH 1
#define B200_EPI_GATE_RES 2
#define B200_EPI_BIAS_F32 3
#define B200_EPI_SILU 4     /* C = silu(acc + bias): FeedForward("linear-silu"), hunyuanvideo15/base/model.py:545-550 */
#define B200_EPI_GELU_ERF 5 /* C = gelu(acc + bias), exact erf form: nn.GELU(), hunyuanvideo15/base/model.py:571,589 */
#define B200_EPI_NORMW 6    /* b200_linear_normw only: C = bf16(acc + bias) * w[n], row sums of squares on the side */
/* OR-ed into `epilogue`: bias is indexed by ROW ([M]) instead of by column -- used for transposed projections
 * (V^T = W_v x^T + b_v of the VAE mid-block attention, vae/wan/model.py:470-478). */
#define B200_EPI_ROW_BIAS 16
int b200_linear(const void* A, const void* W, const void* bias, void* C, const void* gate, int M, int N, int K,
                int64_t lda, int64_t ldw, int64_t ldc, int epilogue, void* stream);

/*
 * The cross-attention query projection with its RMS-norm folded in (attention.py:345-370: q = norm_q(to_q(x)) with the norm
 * taken over ALL heads' channels): C = bf16(q * norm_w[n]) with q = bf16(A W^T + bias) as the reference stores it, and
 * row_sumsq[row * n_parts + part] = sum of q^2 over the columns of column tile `part` (*n_parts tiles per row, decided by the
 * kernel form; row_sumsq must hold M * row_sumsq_capacity floats with row_sumsq_capacity >= ceil(N / 64)).  The row's
 * rsqrt(mean(q^2) + eps) is applied to the attention logits by b200_attn_fwd_qnorm: one kernel launch and one read + write
 * pass over q less than projection -> b200_rmsnorm_rope, and one rounding of q less (q * rstd is never rounded to bf16).
 */
int b200_linear_normw(const void* A, const void* W, const void* bias, const void* norm_w, void* C, float* row_sumsq,
                      int row_sumsq_capacity, int* n_parts, int M, int N, int K, int64_t lda, int64_t ldw, int64_t ldc,
                      void* stream);

/*
 * y = LayerNorm_fp32(x, eps, no affine) * (1 + scale) + shift      (adaLN modulate)
 * Replaces _chunked_modulated_norm -- transformer/wan/base/model.py:56-116 with apply_scale_shift_inplace
 * (transformer/efficiency/ops.py:37-56).  x,y: [rows, dim] (row strides ldx, ldy); scale, shift: [dim] bf16
 * shared by all rows (mod_stride = 0) or per-row [rows, dim] with row stride mod_stride (Wan 2.2 5B ti2v).
 * With scale == NULL it is the plain FP32LayerNorm with optional affine weight/bias (norm2, model.py:1219;
 * ln_w, ln_b: [dim] fp32-or-bf16 given as bf16 here) or no affine at all.
 */
int b200_layernorm_modulate(const void* x, void* y, const void* scale, const void* shift, const void* ln_w,
                            const void* ln_b, int rows, int dim, int64_t ldx, int64_t ldy, int64_t mod_stride,
                            float eps, void* stream);

/*
 * In-place q/k RMS-norm over the full channel dim (across heads) followed by Wan 3-axis RoPE on
 * (even, odd) pairs.  Replaces InplaceRMSNorm.forward (transformer/efficiency/mod.py:24-35) +
 * apply_wan_rope_inplace (transformer/efficiency/ops.py:101-160) at attention.py:349-370.
 *   x: [rows, heads*head_dim] row stride ldx, modified in place; w: [heads*head_dim] bf16 norm weight or NULL;
 *   rope: bf16 [rows, head_dim] = (cos0, sin0, cos1, sin1, ...) already cast to bf16 as the reference casts them,
 *   shared by all heads, or NULL for no RoPE (cross-attn q/k).  eps < 0 skips the norm (RoPE only).
 * The reference's bf16 rounding points are reproduced (rsqrt factor, weight multiply, cos/sin cast to bf16,
 * mul_ then addcmul_).
 */
int b200_rmsnorm_rope(void* x, const void* w, const void* rope, int rows, int heads, int head_dim, int64_t ldx,
                      float eps, void* stream);

/*
 * h += y * gate   (apply_gate_inplace + add_, transformer/efficiency/ops.py:19-34, model.py:1212-1213).
 * gate: [dim] bf16 or NULL (plain residual add, model.py:1245-1251).  Stand-alone form of B200_EPI_GATE_RES
 * for callers whose y comes from somewhere else.
 */
int b200_gate_residual(void* h, const void* y, const void* gate, int rows, int dim, int64_t ldh, int64_t ldy,
                       void* stream);

/*
 * Classifier-free-guidance combine: out = u + g * (c - u), computed as the reference's bf16 tensor expression
 * does (three bf16 roundings; engine/wan/shared/__init__.py:565).  out is bf16 like the reference's noise_pred,
 * which the scheduler then consumes (scheduler/unipc.py:317).
 */
int b200_cfg_combine(const void* cond, const void* uncond, void* out, float guidance, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Sequence-parallel exchange fused into the kernels (no reference counterpart: the reference runs one job on one
 * GPU).  `peers[i]` are device pointers to rank i's receive buffer, mapped into this process
 * over NVLink (CUDA peer / symmetric memory); stores go straight to the owning GPU, so the all-to-all of the Ulysses
 * layout overlaps the arithmetic tile by tile and no pack / NCCL step remains.
 * --------------------------------------------------------------------------------------------------------- */

/*
 * b200_rmsnorm_rope over `n_batch` column blocks of the same rows in ONE launch: block b is x + b * x_batch_stride (elements)
 * with weight w + b * w_batch_stride; the RoPE table is shared.  Call sites: q and k of the self-attention processor
 * (attention.py:349-370: norm_q, norm_k, RoPE on both) with n_batch = 2, x_batch_stride = dim inside the fused q|k|v buffer;
 * norm_k of every layer's cross-attention (attention.py:349-352 on the text context) with n_batch = num_layers over the
 * [L_text, num_layers * 2 * dim] output of the batched to_k|to_v projection.
 */
int b200_rmsnorm_rope_batched(void* x, const void* w, const void* rope, int rows, int heads, int head_dim, int64_t ldx,
                              float eps, int n_batch, int64_t x_batch_stride, int64_t w_batch_stride, void* stream);

/*
 * b200_rmsnorm_rope with the "tokens -> heads" scatter: channel c of local token r is written to
 *   peers[c / width] + dst_elem_offset + (row0 + r) * width + c % width,   width = (heads / n_peers) * head_dim,
 * i.e. into the [S_total, width] plane (q, k or v -- selected by dst_elem_offset) of the rank that owns that head
 * group.  x is not modified.  eps < 0 and rope == NULL make it a pure scatter copy (used for v).
 */
int b200_rmsnorm_rope_scatter(const void* x, const void* w, const void* rope, int rows, int heads, int head_dim,
                              int64_t ldx, float eps, void* const* peers, int n_peers, int64_t dst_elem_offset,
                              int row0, void* stream);

/*
 * b200_attn_fwd (batch 1) with the "heads -> tokens" scatter: this rank computes heads [head_off, head_off + H) for
 * ALL Sq query rows; output row r is written to o_peers[r / rows_per_rank] at row r % rows_per_rank, head
 * head_off + h of that rank's [rows_per_rank, H_total * 128] buffer (row stride o_ss, head stride o_sh).
 */
int b200_attn_fwd_scatter(const void* q, const void* k, const void* v, int H, int Sq, int Sk, int D, int64_t q_sh,
                          int64_t q_ss, int64_t k_sh, int64_t k_ss, int64_t v_sh, int64_t v_ss, void* const* o_peers,
                          int n_peers, int rows_per_rank, int head_off, int64_t o_sh, int64_t o_ss, float scale,
                          void* stream);

/*
 * The same for the JOINT sequences of the dual-stream families (hunyuanvideo15/base/model.py:617-694 concatenates latent and
 * text tokens before the attention call; flux/base/attention.py and qwenimage/base/model.py put the text first): Sq - rep_rows
 * rows are token-sharded as above, `rep_rows` rows -- the replicated text stream, first (rep_first != 0) or last -- are written
 * to EVERY peer.  Each peer's buffer is [rep_rows + rows_per_rank, H_total * 128] in its local joint order.
 */
int b200_attn_fwd_scatter_joint(const void* q, const void* k, const void* v, int H, int Sq, int Sk, int D, int64_t q_sh,
                                int64_t q_ss, int64_t k_sh, int64_t k_ss, int64_t v_sh, int64_t v_ss, void* const* o_peers,
                                int n_peers, int rows_per_rank, int head_off, int64_t o_sh, int64_t o_ss, int rep_rows,
                                int rep_first, float scale, void* stream);

/*
 * Diagnostics (not a reference call site): b200_attn_fwd with SM-clock timestamps of the softmax / MMA hand-offs of CTA
 * (0,0,0) written to prof[prof_steps][32] (int64, device memory): columns 0-4 softmax of query tile 0 (S seen ready,
 * S in registers, row max done, P stores issued, "P ready" signalled), 5-9 the same for tile 1, 10-13 the MMA thread
 * (V tile landed, PV0 issued, S0(j+1) issued, PV1 issued).  Used by scripts/attn_timeline.py; profiles/r02_attn_timeline*.
 */
int b200_attn_fwd_prof(const void* q, const void* k, const void* v, void* o, int B, int H, int Sq, int Sk, int D,
                       int64_t q_sb, int64_t q_sh, int64_t q_ss, int64_t k_sb, int64_t k_sh, int64_t k_ss, int64_t v_sb,
                       int64_t v_sh, int64_t v_ss, int64_t o_sb, int64_t o_sh, int64_t o_ss, float scale,
                       long long* prof, int prof_steps, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Wan 3D-VAE decode (AutoencoderKLWan.decode, vae/wan/model.py:1378; BaseEngine.vae_decode, engine/base_engine.py:2030).
 * Activations are channels-last bf16 [T, H, W, C] (one spatial tile of one video at a time).
 * --------------------------------------------------------------------------------------------------------- */

/*
 * Causal 3-D convolution (WanCausalConv3d, vae/wan/model.py:136-185; also the per-frame Conv2d of WanResample
 * :291-353 with KT = 1) as an implicit GEMM on the tensor cores:
 *   y[t,h,w,:] = bias + sum_{kt,kh,kw} W[kt,kh,kw] x[t+kt-(KT-1), h+kh-KH/2, w+kw-KW/2, :]   (zero outside the volume)
 * x: [T,H,W,Cin]; w: [KT*KH*KW, Cout, Cin] (tap-major, host re-layout of the reference's [Cout,Cin,KT,KH,KW]);
 * bias: [Cout] or NULL; residual: [T,H,W,Cout] added to the result or NULL (WanResidualBlock `x + h`, :441).
 * out_mode 0: channels-last bf16; output channel n goes to frame t*out_t_mul + out_t_off + n / c_split, channel
 *             n % c_split (c_split = Cout, mul = 1, off = 0 for a plain conv; c_split = Cout/2, mul = 2, off = 1 for
 *             the 2x temporal interleave of upsample3d's time_conv, :332-334).
 * out_mode 1: planar bf16 [c_valid, T, H, W] (conv_out; channels >= c_valid are padding).
 * Cin % 32 == 0, Cout % 16 == 0, KH and KW odd.
 */
int b200_conv3d_cl(const void* x, const void* w, const void* bias, const void* residual, void* out, int T, int H, int W,
                   int Cin, int Cout, int KT, int KH, int KW, int out_mode, int out_t_mul, int out_t_off, int c_split,
                   int c_valid, void* stream);

/*
 * b200_conv3d_cl followed by WanRMS_norm + SiLU of its output in the SAME kernel: conv1 -> norm2 -> nonlinearity of
 * WanResidualBlock (vae/wan/model.py:404-413; the conv output has no other consumer).  The epilogue rounds the conv result to
 * bf16 as the reference stores it, normalises over the channels of the pixel, applies gamma and SiLU, and stores only that.
 * Needs Cout in one N tile (Cout % 16 == 0, Cout <= 256 and a divisor rule that keeps it whole: 96, 192; 384 -> B200_ERR_SHAPE,
 * the caller then issues conv + b200_rmsnorm_silu_cl).
 */
int b200_conv3d_cl_norm_silu(const void* x, const void* w, const void* bias, const void* gamma, void* out, int T, int H,
                             int W, int Cin, int Cout, int KT, int KH, int KW, void* stream);

/* y = x / max(||x||_2 over channels, 1e-12) * sqrt(C) * gamma, then SiLU if `silu` (WanRMS_norm :216-222 + the
 * nonlinearity at :404-405, :1005-1006); x, y: [pixels, C]; gamma: [C]. */
int b200_rmsnorm_silu_cl(const void* x, void* y, const void* gamma, int64_t pixels, int C, int silu, void* stream);

/* Nearest-exact 2x spatial upsample (WanUpsample :226-237): [T,H,W,C] -> [T,2H,2W,C]. */
int b200_upsample2x_cl(const void* in, void* out, int T, int H, int W, int C, void* stream);

/* P = softmax(S * scale) row-wise; S fp32 [rows, cols] (row stride lds), P bf16 (row stride ldp).  Mid-block
 * single-head attention (WanAttentionBlock :461-490), whose head dim (= channels) is not 128. */
int b200_softmax_rows(const float* s, void* p, int rows, int cols, int64_t lds, int64_t ldp, float scale, void* stream);

/* Blend a decoded tile (planar bf16 [planes, th, tw], planes = 3*T) with its upper / left neighbours, in place, in
 * the reference's order and bf16 arithmetic (blend_v then blend_h, :1404-1422), then write the cropped, clamped
 * [-1,1] tile into the frame buffer [planes, OH, OW] at (y0, x0) (:1600-1619). */
int b200_blend_tile(void* tile, const void* up, const void* left, void* frame, int planes, int th, int tw, int up_h,
                    int up_w, int left_h, int left_w, int blend, int crop_h, int crop_w, int y0, int x0, int OH, int OW,
                    void* stream);

/* Frame hand-off after decode: planar bf16 video [3, T, H, W] in [-1, 1] -> uint8 [T, H, W, 3] (the layout the
 * encoder / PIL consumes), with the arithmetic of BaseEngine._tensor_to_frames (engine/base_engine.py:2945-2949) ->
 * diffusers VideoProcessor.postprocess_video on a bf16 tensor: (x * 0.5 + 0.5) in bf16 (two roundings), clamp(0, 1),
 * float32 * 255, round-half-even, uint8.  Bit-exact. */
int b200_frames_to_uint8(const void* video, void* out, int T, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Dual-stream (MMDiT) block families -- Flux, HunyuanVideo-1.5, QwenImage.  They reach the
 * attention core and the GEMM above unchanged; these are their row kernels.
 * --------------------------------------------------------------------------------------------------------- */

/* diffusers AdaLayerNormZero / AdaLayerNormZeroSingle / AdaLayerNormContinuous and the explicit
 * `norm2(x) * (1 + scale) + shift` of the dual-stream blocks: y = bf16(bf16(bf16(LN(x)) * bf16(1 + scale)) + shift),
 * LN without affine, fp32 statistics; scale/shift [dim] shared by all rows.
 * Replaces flux/base/model.py:266-272,297-300,320-324,215,649; hunyuanvideo15/base/model.py:634-639,664-673;
 * qwenimage/base/model.py:679-750 (_modulate). */
int b200_adaln_zero_modulate(const void* x, void* y, const void* scale, const void* shift, int rows, int dim,
                             int64_t ldx, int64_t ldy, float eps, void* stream);

/* Per-head q/k RMS-norm (over head_dim = 128) + rotary embedding on interleaved channel pairs, in place on the q and k
 * column blocks of a fused QKV buffer (row stride ldx), both tensors in one launch (k, wq, wk, rope may be NULL).
 *   norm_mode 0: no norm; 1: torch.nn.RMSNorm rounding (flux/base/attention.py:70-71,78-79);
 *             2: InplaceRMSNorm rounding (efficiency/mod.py:24-35 as used hunyuanvideo15/base/model.py:115-116,142-145);
 *             3: diffusers RMSNorm rounding (qwenimage/base/attention.py:115-122).
 *   rope: fp32 [rows, 64, 2] = (cos, sin) per channel pair; re' = re*c - im*s, im' = im*c + re*s in fp32, one rounding
 *         (flux/base/attention.py:86-88; efficiency/ops.py:163-235; qwenimage/base/attention.py:124-130). */
int b200_headnorm_rope(void* q, void* k, const void* wq, const void* wk, const float* rope, int rows, int heads,
                       int head_dim, int64_t ldx, float eps, int norm_mode, void* stream);

/* Row RMS-norm with gain over all `dim` channels, out of place: the text-stream input norm of QwenImage (diffusers
 * RMSNorm(joint_attention_dim), qwenimage/base/model.py:826, called :920).  norm_mode 1..3 as in b200_headnorm_rope. */
int b200_rmsnorm_rows(const void* x, void* y, const void* w, int rows, int dim, int64_t ldx, int64_t ldy, float eps,
                      int norm_mode, void* stream);

/* SwiGLU of a fused projection: y[:, j] = bf16(bf16(silu(x[:, j])) * x[:, inner + j]), x [rows, 2*inner] (row stride
 * ldx), y [rows, inner] (row stride ldy).  Replaces Flux2SwiGLU.forward, flux2/base/model.py:91-105. */
int b200_swiglu(const void* x, void* y, int rows, int inner, int64_t ldx, int64_t ldy, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * HunyuanVideo-1.5 3D-VAE decode (vae/hunyuanvideo15/model.py).  Its causal convs pad with
 * mode="replicate" (:72-90), which TMA zero fill cannot express: the producer kernel materialises the padded tensor
 * (fused with the RMS-norm + SiLU that precedes the conv) and the conv runs on it without padding.
 * --------------------------------------------------------------------------------------------------------- */

/* y [T+pad_t, H+2*pad_h, W+2*pad_w, C] = replicate_pad(f(x [T,H,W,C])), f = channel RMS-norm * gamma (+ SiLU) when gamma
 * is not NULL (HunyuanVideo15RMS_norm :93-127 + nonlinearity :366-376), identity otherwise (conv_in :708, the upsample
 * conv :251).  Time is padded in FRONT only (causal). */
int b200_pad_norm_silu_cl(const void* x, void* y, const void* gamma, int T, int H, int W, int C, int pad_t, int pad_h,
                          int pad_w, int silu, void* stream);

/* b200_conv3d_cl on an input that already carries its padding: x_padded [T+KT-1, H+KH-1, W+KW-1, Cin] -> out [T,H,W,Cout]
 * (out_mode 0) or planar [c_valid, T, H, W] (out_mode 1).  Replaces HunyuanVideo15CausalConv3d.forward :86-90. */
int b200_conv3d_cl_padded(const void* x_padded, const void* w, const void* bias, const void* residual, void* out, int T,
                          int H, int W, int Cin, int Cout, int KT, int KH, int KW, int out_mode, int c_valid, void* stream);

/* HunyuanVideo15Upsample.forward after its conv (:249-274): DCAE channel-to-space rearrangement of h [T,H,W,F*Cout]
 * (F = 8 with temporal upsampling, first frame not doubled; else 4) plus the channel-repeated, rearranged shortcut of the
 * conv input x [T,H,W,Cin] -> out [T',2H,2W,Cout], T' = 2T-1 or T. */
int b200_dcae_upsample_cl(const void* h, const void* x, void* out, int T, int H, int W, int Cout, int Cin, int temporal,
                          void* stream);

/* b200_softmax_rows with the frame-causal mask of HunyuanVideo15AttnBlock (:143-165): row r sees columns
 * [0, (r / block + 1) * block), block = H*W; masked probabilities are written as 0. */
int b200_softmax_rows_block_causal(const float* s, void* p, int rows, int cols, int64_t lds, int64_t ldp, float scale,
                                   int block, void* stream);

/* b200_blend_tile without the final clamp (AutoencoderKLHunyuanVideo15.tiled_decode :1060-1119 does not clamp). */
int b200_blend_tile_noclamp(void* tile, const void* up, const void* left, void* frame, int planes, int th, int tw,
                            int up_h, int up_w, int left_h, int left_w, int blend, int crop_h, int crop_w, int y0, int x0,
                            int OH, int OW, void* stream);

/*
 * The whole Wan 3D-VAE decoder of ONE latent tile in one call (SURVEY.md section 8b minimum set: b200_wan_vae_decode <-
 * AutoencoderKLWan.decode, vae/wan/model.py:1378 -> _decode :1333-1375 -> post_quant_conv + WanDecoder3d.forward :972-1021
 * over all latent frames; the reference streams frame by frame through feat_cache, which equals one causal pass over the
 * whole time axis -- tests/test_oracle_vae.py).  Non-residual decoder (Wan 2.1 / 2.2-A14B): conv_in, mid block (res, single-head
 * attention, res), four up blocks of three residual blocks, three upsamplers, norm_out + conv_out.  All pointers are bf16 device
 * memory in the layouts AutoencoderKLWan.load_state_dict of this package produces: conv weights tap-major [taps * Cout, Cin],
 * 1x1 convs as [Cout, Cin], gammas [C].
 *   z_cl       [T, h, w, 64]   the latent tile channels-last, z_dim real channels + zero padding to 64
 *   out        [3, 1 + 4 (T - 1), 8 h, 8 w]   planar bf16, BEFORE the clamp (the blend kernel clamps)
 *   workspace  >= b200_wan_vae_decode_workspace(...) bytes, 256-byte aligned; nothing is allocated inside
 * ~197 launches for a 32 x 32 x 21 tile, all on `stream`.
 */
typedef struct {
  const void *norm1_gamma, *conv1_w, *conv1_b, *norm2_gamma, *conv2_w, *conv2_b;
  const void *shortcut_w, *shortcut_b; /* 1x1 conv [cout, cin] when cin != cout, else NULL */
  int cin, cout;
} B200WanResBlock;
typedef struct {
  const void *norm_gamma, *to_qkv_w, *to_qkv_b, *proj_w, *proj_b; /* to_qkv [3C, C], proj [C, C] */
  int channels;
} B200WanAttn;
typedef struct {
  const void *resample_w, *resample_b;   /* (1,3,3) conv C -> C/2, tap-major */
  const void *time_conv_w, *time_conv_b; /* (3,1,1) conv C -> 2C (temporal upsample) or NULL */
  int channels, temporal;
} B200WanUpsample;
typedef struct {
  const void *post_quant_w, *post_quant_b; /* [z_pad, 64], [z_pad] */
  const void *conv_in_w, *conv_in_b;       /* [27 * dims[0], z_pad] */
  B200WanResBlock mid_res[2];
  B200WanAttn mid_attn;
  B200WanResBlock up_res[4][3];
  B200WanUpsample up_samp[3];
  const void *norm_out_gamma, *conv_out_w, *conv_out_b; /* conv_out padded to 16 output channels */
  int dims[5];                             /* decoder widths, e.g. 384 384 384 192 96 */
  int z_pad;                               /* channels conv_in consumes (32) */
} B200WanVaeWeights;
int64_t b200_wan_vae_decode_workspace(const B200WanVaeWeights* w, int T, int h, int wd);
int b200_wan_vae_decode(const void* z_cl, const B200WanVaeWeights* w, void* out, void* workspace, int64_t workspace_bytes,
                        int T, int h, int wd, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Composite entry points at the granularity of the reference's call sites.  Each enqueues the
 * kernels above back to back on `stream`; tensors are contiguous ([rows, dim] unless a stride is given).
 * --------------------------------------------------------------------------------------------------------- */

/* adaLN modulate: y = LN_fp32(x) * (1 + scale) + shift with scale/shift [dim] shared by all rows
 * (_chunked_modulated_norm, transformer/wan/base/model.py:56-116). */
int b200_ln_modulate(const void* x, const void* scale, const void* shift, void* y, int rows, int dim, float eps,
                     void* stream);

/* Self-attention front end of WanAttnProcessor2_0.__call__ (transformer/wan/base/attention.py:345-370): the fused
 * to_q|to_k|to_v projection (w_qkv [3*dim, dim], b_qkv [3*dim]) into qkv [rows, 3*dim], then RMS-norm across heads +
 * RoPE in place on the q and k column blocks (3 launches).  q, k, v for b200_attn_fwd are the column blocks of `qkv`
 * (head stride head_dim, token stride 3*dim).  x: [rows, dim] with row stride ldx. */
int b200_qkv_rmsnorm_rope(const void* x, const void* w_qkv, const void* b_qkv, const void* wq_norm, const void* wk_norm,
                          const void* rope, void* qkv, int rows, int dim, int heads, int64_t ldx, float eps,
                          void* stream);

/* Feed-forward + gated residual (transformer/wan/base/model.py:1265-1279, diffusers FeedForward "gelu-approximate"):
 * h += gate * (W2 gelu_tanh(W1 x + b1) + b2); workspace: [rows, ffn_dim] bf16 (2 launches, the activation and the
 * gate/residual are GEMM epilogues). */
int b200_mlp_gelu(const void* x, const void* w1, const void* b1, const void* w2, const void* b2, const void* gate,
                  void* h, void* workspace, int rows, int dim, int ffn_dim, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* APEX_B200_H_ */
