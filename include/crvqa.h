/* crvqa.h -- C ABI of libcrvqa.so: the sm_100a kernels behind the stage-2 mask-training hot path of
 * Compress-Robust-VQA.
 *
 * The reference (PhoebusSi/Compress-Robust-VQA) is pure Python on PyTorch and has NO FFI layer of its
 * own; the boundary it offers is the Python object protocol of
 * masking/maskers*.py and hg_transformers/mask_trainer_*VQA.py.  Each entry point below therefore
 * cites the reference *call site* (file:line, relative to the reference root) whose PyTorch library
 * calls it replaces.  The Python mirror of the reference interface lives in
 * compress-robust-vqa_b200/{masking,hg_transformers}/ and binds these symbols with ctypes
 * (compress-robust-vqa_b200/crvqa/_lib.py); INTEGRATION.md shows the stub a reference maintainer
 * would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *   - the library never allocates or frees caller memory; scratch comes from a caller workspace whose
 *     size is returned by the matching *_workspace_bytes() query
 *   - return value: 0 = success, < 0 = argument error (CRV_E_*), > 0 = cudaError_t / CUresult
 *   - row-major tensors; "bf16" = __nv_bfloat16 bit pattern in uint16_t; thresholds are passed as a
 *     device pointer to one float so that no host synchronisation is ever needed
 */
#ifndef CRVQA_H_
#define CRVQA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRV_OK 0
#define CRV_E_BADARG (-1)   /* null pointer / non-positive size */
#define CRV_E_ALIGN (-2)    /* pointer or leading dimension violates the stated alignment */
#define CRV_E_SHAPE (-3)    /* shape not supported by this entry point */
#define CRV_E_WORKSPACE (-4) /* workspace too small */
#define CRV_E_DRIVER (-5)   /* CUDA driver entry point (tensor-map encode) unavailable */

#define CRV_DTYPE_F32 0
#define CRV_DTYPE_BF16 1

/* Library / device info -------------------------------------------------------------------- */
int crv_version(void);                       /* ABI version, currently 1 */
const char* crv_error_string(int code);      /* static string for CRV_E_* codes; "cuda error" otherwise */
int crv_last_cuda_error(void);               /* last cudaError_t / CUresult seen by this thread */
unsigned long long crv_launch_count(void);   /* kernels this library has launched so far (process-wide) */

/* Elementwise helpers ------------------------------------------------------------------------ */
/* fp32 -> bf16 (round-to-nearest-even) operand conversion in front of the bf16 MMAs. */
int crv_cast_f32_to_bf16(const float* src, uint16_t* dst, int64_t n, void* stream);

/* Stage-3 frozen-mask fine-tune (run_vqa_stage3.py:227-300: torch.nn.utils.prune.CustomFromMask, whose forward
 * pre-hook sets weight = weight_orig * weight_mask): dst = bf16(w * mask), product in fp32, one pass.  The
 * masked GEMMs then take dst as their plain operand and crv_masked_linear_bwd_ds with w_f32 = the fp32 0/1 mask
 * yields dW = (dY^T X) (.) mask, the gradient autograd gives weight_orig. */
int crv_mul_cast_bf16(const float* w, const float* mask, uint16_t* dst, int64_t n, void* stream);

/* binarizer_fn1 (masking/maskers.py:325-329): mask[i] = scores[i] > *thr ? 1 : 0 (strict >),
 * written as fp32 0/1 (mask_f32) and/or bytes (mask_u8) -- either may be NULL.  Also the body of
 * Trainer.save_model_mask / Trainer.binarizer_fn1 (hg_transformers/mask_trainer_VQA.py:930-968).
 * If kept_count != NULL, *kept_count (device, int64) receives the number of ones. */
int crv_binarize(const float* scores, const float* thr, float* mask_f32, uint8_t* mask_u8,
                 long long* kept_count, int64_t n, void* stream);

/* Materialised masked weight  Wm = W (.) (S > *thr)  in bf16 -- the `self.weight * M_w` of
 * MaskedLinear1.forward (masking/maskers.py:363-366) for callers that reuse one mask across calls. */
int crv_apply_mask_bf16(const uint16_t* w_bf16, const float* scores, const float* thr, uint16_t* wm_bf16,
                        int64_t n, void* stream);

/* The same over a whole score arena in ONE launch (mask cache of the training engine: scores only change
 * at the optimiser step, thresholds only at reset_threshold, so W (.) M can be refreshed once per step
 * instead of once per module call).  w / scores / wm are flat buffers with identical element offsets;
 * chunks = nchunks x int4 {start / 8, length, segment index, 0} (device), no chunk straddling a module;
 * thr_vec[segment] is that module's threshold. */
int crv_apply_mask_segmented(const uint16_t* w_bf16, const float* scores, const float* thr_vec,
                             const int* chunks, int nchunks, uint16_t* wm_bf16, void* stream);

/* Masked linear (tcgen05 GEMMs) --------------------------------------------------------------- */
/* Forward of MaskedLinear1 (masking/maskers.py:359-366 = _Binarizer1 + `weight * M_w` + F.linear):
 *     Y[M,N] = X[M,K] . (W[N,K] (.) (S[N,K] > *thr))^T + bias[N]
 * X, W bf16; S fp32; accumulation fp32; Y fp32 or bf16 (y_dtype).  The mask is applied to the W tile
 * in shared memory between the TMA load and the MMA; no masked weight is materialised in HBM.
 * scores == NULL means "W is already masked / dense" (plain bf16 GEMM).  bias may be NULL.
 * Requirements: K % 8 == 0 (use crv_masked_linear_small_k otherwise), X/W/S 16-byte aligned. */
int crv_masked_linear_fwd(const uint16_t* x_bf16, const uint16_t* w_bf16, const float* scores,
                          const float* thr, const float* bias, void* y, int y_dtype, int M, int N, int K,
                          void* stream);

/* Debug aid (not part of the drop-in path): when buf != NULL every masked-GEMM CTA writes 8 int64
 * globaltimer stamps at buf[8 * blockIdx.x]; pass NULL to switch it off. */
int crv_gemm_debug_timestamps(long long* buf);

/* dX of the above (autograd of masking/maskers.py:365-366 with frozen weights, :564-569):
 *     dX[M,K] = dY[M,N] . (W (.) (S > *thr))
 * dY bf16; dX fp32 or bf16.  Requirements: N % 8 == 0, K % 8 == 0. */
int crv_masked_linear_bwd_dx(const uint16_t* dy_bf16, const uint16_t* w_bf16, const float* scores,
                             const float* thr, void* dx, int dx_dtype, int M, int N, int K, void* stream);

/* Straight-through score gradient (autograd of masking/maskers.py:337-339,365-366):
 *     dS[N,K] (+)= (dY[M,N]^T . X[M,K]) (.) W[N,K]
 * The (.)W happens in the GEMM epilogue with W in FP32, as the reference multiplies (dM * self.weight, an fp32
 * Parameter): only the MMA operands dY and X are bf16.  accumulate != 0 adds into dS (second invocation of the
 * shared cross-attention modules, hg_transformers/modeling_lxmert.py:947-958), otherwise dS is overwritten.
 * The reduction over M may be split across CTAs (TMA reduce-add, fp32): the summation order of the splits is
 * not fixed, so dS is reproducible to fp32 rounding of a handful of partial sums, not bit for bit.
 * Requirements: N % 8 == 0, K % 8 == 0; W 16-byte aligned. */
int crv_masked_linear_bwd_ds(const uint16_t* dy_bf16, const uint16_t* x_bf16, const float* w_f32,
                             float* dscores, int accumulate, int M, int N, int K, void* stream);

/* Grouped launch of the three GEMMs above (mask-cache form: `b` is the materialised W (.) M, no scores): up to four
 * independent problems of one layer phase run as ONE persistent 2-CTA kernel -- dX and dS of a module (both read the
 * same dY), the two modalities of a cross layer (hg_transformers/modeling_lxmert.py:947-1009), the language and the
 * vision stack in lockstep -- so that pipeline fill / drain and wave quantisation are paid per group, not per GEMM.
 * Longer lists are cut into groups of four in order; problems whose shape the 2-CTA kernel does not take (output
 * width not a multiple of 256, fewer than 256 output rows) fall back to the single entry points.
 *   kind  CRV_GEMM_FWD  out[M,N] = a[M,K] . b[N,K]^T + bias          a = X,  b = W (.) M
 *         CRV_GEMM_DX   out[M,K] = a[M,N] . b[N,K]                   a = dY, b = W (.) M
 *         CRV_GEMM_DS   out[N,K] (+)= (a[M,N]^T . b[M,K]) (.) w_f32   a = dY, b = X
 *   act   CRV_ACT_GELU  fuses the erf-GELU of LxmertIntermediate (hg_transformers/modeling_lxmert.py:876-886) into
 *         the GEMMs around it (bf16 outputs only): FWD writes aux = y (the pre-activation, bf16) and out = gelu(y);
 *         DX reads aux = u [M,K] (bf16) and writes out = dX (.) gelu'(u).
 * Two DS problems of one call may name the same `out` (a shared module applied to both modalities): they reduce-add
 * into it.  problems_host is a HOST array. */
#define CRV_GEMM_FWD 0
#define CRV_GEMM_DX 1
#define CRV_GEMM_DS 2
#define CRV_ACT_NONE 0
#define CRV_ACT_GELU 1
/* `accumulate` of a DS problem / crv_masked_linear_bwd_ds: 0 = overwrite dS, 1 = add to dS, CRV_DS_ZEROED = dS is known
 * to hold zeros (cleared by crv_adamw_segmented): plain store when the reduction is not split, reduce-add without a
 * clearing memset when it is. */
#define CRV_DS_ZEROED 3
typedef struct crv_gemm_problem {
  int kind, act, out_dtype, accumulate;
  int M, N, K, reserved;
  const uint16_t* a;
  const uint16_t* b;
  const float* bias;   /* FWD, may be NULL */
  const float* w_f32;  /* DS */
  void* out;
  void* aux;           /* CRV_ACT_GELU only */
} crv_gemm_problem;
int crv_masked_gemm_grouped(const crv_gemm_problem* problems_host, int count, void* stream);

/* Same three operations for inner dimensions the TMA path cannot take (box_fc has K = 4,
 * hg_transformers/modeling_lxmert.py:1025): fp32 in / fp32 out SIMT kernels, exact fp32 math. */
int crv_masked_linear_small_k_fwd(const float* x, const float* w, const float* scores, const float* thr,
                                  const float* bias, float* y, int M, int N, int K, void* stream);
int crv_masked_linear_small_k_bwd(const float* dy, const float* x, const float* w, const float* scores,
                                  const float* thr, float* dx /* may be NULL */, float* dscores,
                                  int accumulate, int M, int N, int K, void* stream);

/* Masked embedding ---------------------------------------------------------------------------- */
/* `F.embedding(x, self.weight * M_w, padding_idx)` branch of MaskedLinear1.forward
 * (masking/maskers.py:362-363): out[t,:] = W[ids[t],:] (.) (S[ids[t],:] > *thr), fp32, without masking
 * the whole table. */
int crv_masked_embedding_fwd(const long long* ids, const float* w, const float* scores, const float* thr,
                             float* out, int64_t n_tokens, int64_t vocab, int dim, void* stream);
/* backward: dS[ids[t],:] += dOut[t,:] (.) W[ids[t],:] for ids[t] != padding_idx (pass -1 for none);
 * rows never looked up stay untouched, so the caller zeroes dS (or accumulates). */
int crv_masked_embedding_bwd(const long long* ids, const float* dout, const float* w, float* dscores,
                             int64_t n_tokens, int64_t vocab, int dim, long long padding_idx, void* stream);

/* Thresholds: exact k-th smallest value --------------------------------------------------------- */
/* Trainer.reset_threshold (hg_transformers/mask_trainer_Robust_VQA.py:467-482,
 * mask_trainer_VQA.py:470-477): for every segment i, thr_out[i] = k[i]-th smallest (1-based, as
 * torch.kthvalue) of the n[i] floats at ptrs[i]; use_abs != 0 selects on |x| instead (magnitude
 * init).  ptrs/n/k are HOST arrays of length count; thr_out is a device array of `count` floats.
 * NaNs order above +inf (torch semantics).
 * Segments larger than 65536 elements take a sample -> filter -> select path that reads the data ONCE
 * (pivots from a 16384-element stratified sample, one streaming pass that counts against the pivots and
 * compacts the ~2.5 % of keys between them, exact radix select of the survivors); it needs candidate space,
 * so size the workspace with crv_kth_value_workspace_bytes_for(n_host, count).  With the smaller
 * crv_kth_value_workspace_bytes(count) every segment takes the 3-pass radix core.  Both are exact. */
size_t crv_kth_value_workspace_bytes(int count);
size_t crv_kth_value_workspace_bytes_for(const long long* n_host, int count);
int crv_kth_value_batched(const float* const* ptrs_host, const long long* n_host, const long long* k_host,
                          int count, int use_abs, float* thr_out, void* workspace, size_t workspace_bytes,
                          void* stream);

/* MaskedLinearX.controlled_init._magnitude (masking/maskers.py:204-215):
 *     S[i] = |W[i]| > *w_thr ? hi : lo      (hi = 2*threshold, lo = 0*threshold in the reference) */
int crv_magnitude_init(const float* w, const float* w_thr, float hi, float lo, float* scores, int64_t n,
                       void* stream);

/* Losses over the answer head ------------------------------------------------------------------- */
/* All three: logits/labels/bias are [B, A] fp32 row-major; outputs: loss_out[0] = scalar loss,
 * loss_out[1] = batch VQA score sum_b labels[b, argmax_a logits[b,a]]
 * (compute_score_with_logits, hg_transformers/data/metrics/__init__.py:90-104); dlogits[B, A] =
 * d loss / d logits.  workspace: crv_vqa_loss_workspace_bytes(B). */
size_t crv_vqa_loss_workspace_bytes(int B);
/* instance_bce_with_logits (hg_transformers/modeling_lxmert.py:248-253): mean BCE * A. */
int crv_vqa_loss_bce(const float* logits, const float* labels, float* loss_out, float* dlogits, int B, int A,
                     void* workspace, void* stream);
/* LPF_loss (hg_transformers/mask_trainer_VQA.py:111-129): mean_b (1-q[b,y])^gamma * -log p[b,y]. */
int crv_vqa_loss_lpf(const float* logits, const float* bias, const long long* max_label, float gamma,
                     float* loss_out, const float* labels /* for the score; may be NULL */, float* dlogits,
                     int B, int A, void* workspace, void* stream);
/* RUBI_loss (hg_transformers/mask_trainer_VQA.py:131-135): mean_b CE(logits * sigmoid(bias), max_label). */
int crv_vqa_loss_rubi(const float* logits, const float* bias, const long long* max_label, float* loss_out,
                      const float* labels /* for the score; may be NULL */, float* dlogits, int B, int A,
                      void* workspace, void* stream);
/* LearnedMixin.forward (hg_transformers/vqa_debias_loss_functions.py:148-196) with entropy weight w:
 * factor_pre[B] = bias_lin(pooled) (pre-softplus), smooth = sigmoid(smooth_param) + constant_smooth.
 * Outputs dlogits[B,A] and dfactor_pre[B] (gradient w.r.t. the pre-softplus factor).
 * BiasProduct.forward (:83-122) is the same expression with factor = 1 and w = 0: pass factor_pre = log(e - 1). */
int crv_vqa_loss_lmh(const float* logits, const float* bias, const float* labels, const float* factor_pre,
                     float smooth, float w, float* loss_out, float* dlogits, float* dfactor_pre, int B, int A,
                     void* workspace, void* stream);

/* Fused elementwise ops between the GEMMs ("next" row f3; hg_transformers/modeling_lxmert.py:830-903) -- */
/* LxmertAttentionOutput / LxmertOutput tail: z = dropout(g) + res; y = LayerNorm(z).  g is the GEMM output
 * (bias already added) in fp32 or bf16; res fp32 or NULL; y is written as fp32 and/or bf16 (either may be
 * NULL); mean / rstd [M] are saved for the backward.  Dropout is counter based on (rng_state[0] = seed,
 * rng_state[1] = step counter, site, element index); rng_state == NULL or p_drop == 0 disables it.
 * H % 128 == 0, H <= 1024. */
int crv_ln_fwd(const void* g, int g_dtype, const float* res, const float* gamma, const float* beta, float eps,
               float p_drop, const unsigned long long* rng_state, int site, float* y_f32, uint16_t* y_bf16,
               float* mean, float* rstd, int M, int H, void* stream);
/* backward: dz = LayerNorm'(dy_f32 + dy_bf16) (either may be NULL); d_res = dz (fp32, may be NULL);
 * d_g = dropout'(dz) in fp32 or bf16 (may be NULL).  The forward mask is regenerated from the same (rng_state, site).
 * gamma / beta are frozen in stage 2 (param_partials = NULL).  Stage 3 trains them (run_vqa_stage3.py:577-598): pass
 * crv_ln_bwd_partials_bytes(M, H) bytes; the kernel writes ceil(M / 8) rows of {dgamma[H] | dbeta[H]} partial sums
 * and crv_partial_reduce adds them in index order (deterministic). */
size_t crv_ln_bwd_partials_bytes(int M, int H);
int crv_ln_bwd(const float* dy_f32, const uint16_t* dy_bf16, const void* g, int g_dtype, const float* res,
               const float* gamma, const float* mean, const float* rstd, float p_drop,
               const unsigned long long* rng_state, int site, void* dg, int dg_dtype, float* dres,
               float* param_partials, int M, int H, void* stream);
/* out[j] (+)= sum_i part[i][j], i = 0 .. nparts-1 in order (n columns). */
int crv_partial_reduce(const float* part, int nparts, int n, float* out, int accumulate, void* stream);
/* Bias gradient of a linear layer (stage 3): out[n] (+)= sum_m x[m, n], x bf16 [M, N] (N % 8 == 0), deterministic;
 * workspace: crv_colsum_workspace_bytes(N). */
size_t crv_colsum_workspace_bytes(int N);
int crv_colsum_bf16(const uint16_t* x, int M, int N, float* out, int accumulate, void* workspace, void* stream);
/* Small-sequence multi-head attention, head dim 64, Sq, Sk <= 64 (LxmertAttention.forward,
 * hg_transformers/modeling_lxmert.py:798-827): out = dropout(softmax(Q K^T * scale + mask)) V.
 * q / k / v are bf16 and addressed as base + b * bs + s * ss + head * 64 + d (element strides), so they can
 * be slices of a fused QKV projection; mask is additive fp32 [B, Sk] or NULL; out is [B, Sq, heads * 64].
 * One warp per (batch, head); warp-level tensor-core MMAs; the backward recomputes the probabilities and
 * regenerates the dropout mask from (rng_state, site). */
int crv_attention_fwd(const uint16_t* q, long long q_bs, long long q_ss, const uint16_t* k, long long k_bs,
                      long long k_ss, const uint16_t* v, long long v_bs, long long v_ss, const float* mask,
                      uint16_t* out, int B, int heads, int Sq, int Sk, float scale, float p_drop,
                      const unsigned long long* rng_state, int site, void* stream);
int crv_attention_bwd(const uint16_t* q, long long q_bs, long long q_ss, const uint16_t* k, long long k_bs,
                      long long k_ss, const uint16_t* v, long long v_bs, long long v_ss, const float* mask,
                      const uint16_t* dout, uint16_t* dq, long long dq_bs, long long dq_ss, uint16_t* dk,
                      long long dk_bs, long long dk_ss, uint16_t* dv, long long dv_bs, long long dv_ss, int B,
                      int heads, int Sq, int Sk, float scale, float p_drop, const unsigned long long* rng_state,
                      int site, void* stream);

/* Training pair of the same attention (hg_transformers/modeling_lxmert.py:798-827 and its autograd): the forward also
 * writes probs [B * heads][Sq][skp] bf16, skp = crv_attention_probs_pitch(Sk) = Sk rounded up to 8: +P_ij where the
 * dropout kept (i, j), -P_ij where it dropped it (P = softmax probability before dropout).  The backward reads probs
 * instead of recomputing Q K^T, the softmax and the dropout decisions: dQ, dK, dV from five small tensor-core GEMMs per
 * (batch, head).  No mask / rng arguments in the backward: both are folded into probs. */
int crv_attention_probs_pitch(int Sk);
int crv_attention_fwd_p(const uint16_t* q, long long q_bs, long long q_ss, const uint16_t* k, long long k_bs,
                        long long k_ss, const uint16_t* v, long long v_bs, long long v_ss, const float* mask,
                        uint16_t* out, uint16_t* probs, int B, int heads, int Sq, int Sk, float scale, float p_drop,
                        const unsigned long long* rng_state, int site, void* stream);
int crv_attention_bwd_p(const uint16_t* q, long long q_bs, long long q_ss, const uint16_t* k, long long k_bs,
                        long long k_ss, const uint16_t* v, long long v_bs, long long v_ss, const uint16_t* probs,
                        const uint16_t* dout, uint16_t* dq, long long dq_bs, long long dq_ss, uint16_t* dk,
                        long long dk_bs, long long dk_ss, uint16_t* dv, long long dv_bs, long long dv_ss, int B,
                        int heads, int Sq, int Sk, float scale, float p_drop, void* stream);

/* LayerNorm -> (average of two) -> dropout, the entry blocks of LXMERT (hg_transformers/modeling_lxmert.py:576-592
 * LxmertVisualFeatureEncoder: y = dropout((LN_a(a) + LN_b(b)) / 2); :744-770 LxmertEmbeddings: y = dropout(LN_a(a)),
 * b = NULL).  a, b fp32 [M, H]; y as fp32 and bf16; stats [M][4] = {mean_a, rstd_a, mean_b, rstd_b}.  The backward
 * regenerates the dropout mask (same rng_state / site) and writes da (and db); either may be NULL when its input
 * needs no gradient.  No gamma / beta gradients (frozen in stage 2).  H % 128 == 0, H <= 1024. */
int crv_ln_avg_drop_fwd(const float* a, const float* b, const float* gamma_a, const float* beta_a, const float* gamma_b,
                        const float* beta_b, float eps, float p_drop, const unsigned long long* rng_state, int site,
                        float* y_f32, uint16_t* y_bf16, float* stats, int M, int H, void* stream);
int crv_ln_avg_drop_bwd(const float* dy_f32, const uint16_t* dy_bf16, const float* a, const float* b,
                        const float* gamma_a, const float* gamma_b, const float* stats, float p_drop,
                        const unsigned long long* rng_state, int site, float* da, float* db, int M, int H, void* stream);

/* erf GELU on bf16 (LxmertIntermediate): y = gelu(u);  du = dy * gelu'(u).  n % 8 == 0. */
int crv_gelu_fwd(const uint16_t* u, uint16_t* y, int64_t n, void* stream);
int crv_gelu_bwd(const uint16_t* u, const uint16_t* dy, uint16_t* du, int64_t n, void* stream);
/* QuickGELU of mPLUG's CLIP vision tower (mPLUG/models/clip/model.py:25-27) on bf16: y = u sigmoid(1.702 u);
 * du = dy * s (1 + 1.702 u (1 - s)), s = sigmoid(1.702 u).  n % 8 == 0. */
int crv_quick_gelu_fwd(const uint16_t* u, uint16_t* y, int64_t n, void* stream);
int crv_quick_gelu_bwd(const uint16_t* u, const uint16_t* dy, uint16_t* du, int64_t n, void* stream);
/* Momentum update of mPLUG's distillation twins (mPLUG/models/model_vqa_mplug.py:152-156) over many separately
 * allocated fp32 tensors in ONE launch: twins[t][e] = twins[t][e] * m + online[t][e] * one_minus_m (two rounded
 * products, one rounded sum: bit-identical to the PyTorch expression).  online_dev / twins_dev: device arrays of
 * tensor pointers (16-byte aligned tensors); rows_dev: device int4 rows {tensor, first element / 4, float4 count,
 * scalar tail count} cutting the tensors into chunks. */
int crv_momentum_update(const float* const* online_dev, float* const* twins_dev, const int* rows_dev, int nrows,
                        float m, float one_minus_m, void* stream);
/* Global-norm clip + optimiser step of mPLUG's engine over many separately allocated fp32 tensors (the step DeepSpeed
 * runs for the reference: mPLUG/vqa_mplug.py:171-204 with mPLUG/configs/ds_config.json gradient_clipping 1.0 and the
 * torch AdamW of mPLUG/optim/optim_factory.py:60-89).  rows_dev: device int4 rows {tensor, first element / 8, elements
 * in the row, flags}; every *_dev argument is a device array of tensor pointers (16-byte aligned tensors).
 * crv_sumsq_multi: *out += sum of squares over all rows (deterministic, same workspace contract as crv_sumsq).
 * crv_adamw_multi: g' = g * min(1, max_norm / (sqrt(*total_sumsq) + 1e-6)) (total_sumsq NULL: no clip);
 * p *= 1 - lr * weight_decay; m += (g' - m)(1 - b1); v = b2 v + (1 - b2) g'^2;
 * p -= lr / (1 - b1^step) * m / (sqrt(v) / sqrt(1 - b2^step) + eps)   (torch.optim.AdamW, step 1-based).
 * Rows with flags bit 0 also refresh the tensor's masked bf16 operand from the new scores in registers:
 * wm[t] = w16[t] (.) (p_new > *thr[t])  (masking/maskers.py:342-358; thr[t] points at one device float).
 * w16_dev / wm_dev / thr_dev: all NULL or all given (entries of tensors without the flag are not read). */
int crv_sumsq_multi(const float* const* xs_dev, const int* rows_dev, int nrows, float* out, void* workspace,
                    void* stream);
int crv_adamw_multi(float* const* p_dev, const float* const* g_dev, float* const* m_dev, float* const* v_dev,
                    const uint16_t* const* w16_dev, uint16_t* const* wm_dev, const float* const* thr_dev,
                    const int* rows_dev, int nrows, double lr, int step, double beta1, double beta2, double eps,
                    double weight_decay, const float* total_sumsq, float max_norm, void* stream);
/* Few-query attention (mPLUG text side: BertSelfAttention of the text / fusion / decoder stacks,
 * mPLUG/models/modeling_mplug.py:205-300): out = dropout(softmax(Q K^T * scale + mask)) V for Lq <= 16 queries against
 * Lk <= 1024 keys per (batch, head), head dim 64.  q / out / dq: [B, Lq, heads * 64] bf16 contiguous; k, v, dk, dv:
 * [B, Lk, heads * 64].  mask: additive fp32, element (b, i, j) at mask[b * mask_batch_stride + i * mask_query_stride +
 * j] (strides 0 broadcast), or NULL.  probs: [B, heads, Lq, Lk] bf16, written by the forward: the softmax probability
 * with the dropout decision in its sign (+P kept, -P dropped; dropout hashed from rng_state / site as crv_ln_fwd does;
 * rng_state NULL or p_drop 0: none) -- the backward reads it instead of recomputing scores, softmax and hash, and
 * must be given the same p_drop. */
int crv_fq_attention_fwd(const uint16_t* q, const uint16_t* k, const uint16_t* v, const float* mask,
                         long long mask_batch_stride, long long mask_query_stride, uint16_t* out, uint16_t* probs,
                         int B, int heads, int Lq, int Lk, float scale, float p_drop,
                         const unsigned long long* rng_state, int site, void* stream);
int crv_fq_attention_bwd(const uint16_t* dout, const uint16_t* q, const uint16_t* k, const uint16_t* v,
                         const uint16_t* probs, uint16_t* dq, uint16_t* dk, uint16_t* dv, int B, int heads, int Lq,
                         int Lk, float scale, float p_drop, void* stream);
/* rng_state[1] += 1 on the device (once per training step, inside the captured graph). */
int crv_rng_advance(unsigned long long* rng_state, void* stream);

/* Optimiser ("next" row f1; hg_transformers/mask_trainer_VQA.py:646-659 + optimization.py:66-129) */
/* sum of squares of n floats accumulated into *out (device, fp32; caller zeroes it).  DETERMINISTIC: block partials
 * are added in a fixed order (the last-arriving block sums them by index), so every data-parallel rank derives the
 * same clip coefficient bit for bit.  workspace: crv_sumsq_workspace_bytes() bytes, ZEROED ONCE by the caller before
 * its first use and then owned by these calls (launches sharing a workspace must be stream-ordered). */
size_t crv_sumsq_workspace_bytes(void);
int crv_sumsq(const float* x, int64_t n, float* out, void* workspace, void* stream);
/* the same over the chunks {start / 8, length, -, -} of a table: the gradient shard one rank owns when the optimiser
 * state is sharded across data-parallel ranks */
int crv_sumsq_segmented(const float* x, const int* chunks, int nchunks, float* out, void* workspace, void* stream);
/* One AdamW step of the reference optimiser on a flat fp32 segment, with the clip coefficient of
 * torch.nn.utils.clip_grad_norm_ folded in:  g' = g * min(1, max_norm / (sqrt(*total_sumsq) + 1e-6));
 * sum += |g'|; m = b1 m + (1-b1) g'; v = b2 v + (1-b2) g'^2; p -= step_size * m / (sqrt(v) + eps);
 * p -= lr * weight_decay * p.  step_size = lr * sqrt(1-b2^t)/(1-b1^t) is computed by the caller.
 * total_sumsq may be NULL (no clipping).  `sum` may be NULL.  If hyper_dev != NULL, {lr, step_size} are
 * read from that device array instead of the by-value arguments (so a captured CUDA graph of the step
 * follows the learning-rate schedule). */
int crv_adamw_step(float* p, const float* g, float* m, float* v, float* sum, int64_t n, float lr,
                   float step_size, double beta1, double beta2, float eps, float weight_decay,
                   const float* total_sumsq, float max_norm, const float* hyper_dev, int mode, float inv_bc2_sqrt,
                   void* stream);
/* mode: 0 = the rule above (the reference's root optimization.AdamW, stage 2).  1 = torch.optim.Adam as the stage-3
 * driver builds it (run_vqa_stage3.py:577-598): g' += weight_decay * p; m, v as above; step_size = lr / (1 - b1^t)
 * (caller); p -= step_size * m / (sqrt(v) * inv_bc2_sqrt + eps) with inv_bc2_sqrt = 1 / sqrt(1 - b2^t); `sum` unused.
 * hyper_dev, when given, holds {lr, step_size, inv_bc2_sqrt}. */

/* The same step over a whole score arena in ONE launch that also (a) refreshes the masked bf16 operand of every
 * segment, wm = w_bf16 (.) (p_new > thr_vec[segment]) -- the scores are in registers, so the separate
 * crv_apply_mask_segmented pass and its re-read of p disappear -- and (b) with zero_grad != 0 clears g after
 * consuming it (the reference loop's model.zero_grad(), hg_transformers/mask_trainer_VQA.py:659), which lets the next
 * step's split score-gradient GEMMs reduce-add without a memset per module (accumulate = CRV_DS_ZEROED).
 * chunks = nchunks x int4 {start / 8, length, segment, flags} (device; no chunk straddles a segment; flags bit 0:
 * the segment has a bf16 operand; bit 1: clear-only chunk -- under the sharded data-parallel optimiser a slice owned
 * by another rank, whose gradient buffer holds reduce-scatter leftovers); w_bf16 / wm_bf16 may both be NULL. */
int crv_adamw_segmented(float* p, float* g, float* m, float* v, float* sum, const int* chunks, int nchunks,
                        const float* thr_vec, const uint16_t* w_bf16, uint16_t* wm_bf16, float lr, float step_size,
                        double beta1, double beta2, float eps, float weight_decay, const float* total_sumsq,
                        float max_norm, const float* hyper_dev, int zero_grad, int mode, float inv_bc2_sqrt,
                        void* stream);
/* mode 1 (stage 3): p are the trained weights themselves, w_bf16 holds the frozen 0/1 mask as bf16 and the refreshed
 * operand is wm = bf16(p_new) where the mask is set, 0 elsewhere; thr_vec is not read and may be NULL. */

#ifdef __cplusplus
}
#endif
#endif /* CRVQA_H_ */
