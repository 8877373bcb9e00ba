/* libtapclip — C ABI of the B200-native TAP-CLIP hot path (attribution-instrumented CLIP forward
 * + backward to the prompt context vectors).
 *
 * The reference (3300786/TAP-CLIP) is pure Python and has no FFI of its own; the boundary it exposes
 * for this path is the Python surface of `CLIPWrapper` (models/clip_wrapper.py:9-65) and `FullModel`
 * (models/model_wrapper.py:12-100).  `tapclip_b200/` keeps that surface and binds the functions below
 * through ctypes; each export cites the reference lines whose arithmetic it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; `tapclip_last_error()` then returns a
 *     thread-local message.  Unsupported shapes / dtypes / devices are errors: there is NO CPU fallback.
 *   - all pointers are DEVICE pointers (borrowed for the duration of the call, never freed by the
 *     library) unless stated; tensors are dense row-major fp32 unless stated; `stream` is a cudaStream_t
 *     passed as void* (0 = default stream).  All work is enqueued on `stream`; no hidden host syncs.
 *   - a handle owns its converted weights and workspaces; it is not thread-safe; one handle per device.
 *     The library also keeps process-wide launch caches (tensor maps, kernel attributes, per-device scratch): drive it
 *     from ONE host thread per process (the deployment model is one process per GPU; several devices from one thread
 *     are supported).
 */
#ifndef TAPCLIP_H_
#define TAPCLIP_H_

#include <stdint.h>

#if defined(__GNUC__)
#define TAPCLIP_API __attribute__((visibility("default")))
#else
#define TAPCLIP_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tapclip_engine* tapclip_handle;

enum { TAPCLIP_ACT_GELU_ERF = 0, TAPCLIP_ACT_QUICK_GELU = 1 };
/* Engine precision (tapclip_config.dtype).  Residual streams, LayerNorm, softmax, L2-norm, logits and the loss are
 * always fp32; the setting picks the tensor-core OPERAND type:
 *   FP32  : fp32 SIMT kernels everywhere (parity mode, logits within 1e-4 of the reference);
 *   BF16  : bf16 operands everywhere;
 *   MIXED : bf16 operands in the image tower and in the backward pass, fp16 operands in the text-tower forward
 *           (same tcgen05 kind::f16 rate and bytes; needed for the 1e-2 logit bar — DESIGN.md "Precision").
 * For the single-kernel tapclip_op_* entry points the same integers name the element type: 0 fp32, 1 bf16, 2 fp16. */
enum { TAPCLIP_DTYPE_FP32 = 0, TAPCLIP_DTYPE_BF16 = 1, TAPCLIP_DTYPE_MIXED = 2, TAPCLIP_DTYPE_FP16 = 2 };
enum { TAPCLIP_ATTR_LITERAL = 0, TAPCLIP_ATTR_INTENDED = 1 };  /* SURVEY.md 8a "mode definitions" */

typedef struct {
    int32_t image_size, patch_size;
    int32_t vision_width, vision_layers, vision_heads;
    int32_t text_width, text_layers, text_heads;
    int32_t embed_dim;
    int32_t context_length;   /* tokens per class prompt in the token bank (77) */
    int32_t act;              /* TAPCLIP_ACT_* : open_clip `quick_gelu` flag */
    int32_t dtype;            /* TAPCLIP_DTYPE_* */
} tapclip_config;

/* ---- lifetime ------------------------------------------------------------------------------------ */
/* Replaces open_clip.create_model_and_transforms + .to(device).eval() + freeze (clip_wrapper.py:13-20).
 * Uses the calling thread's current CUDA device; fails if it is not sm_100. */
TAPCLIP_API int tapclip_create(const tapclip_config* cfg, tapclip_handle* out);
TAPCLIP_API int tapclip_destroy(tapclip_handle h);
TAPCLIP_API const char* tapclip_last_error(void);
TAPCLIP_API const char* tapclip_version(void);

/* Replaces model.load_state_dict (clip_wrapper.py:14-15): one call per open_clip state-dict entry
 * (`visual.conv1.weight`, `visual.transformer.resblocks.3.attn.in_proj_weight`, `text_projection`, ...).
 * `data` is an fp32 device tensor of the given shape; the engine keeps its own converted copy.
 * token_embedding.weight, positional_embedding and ln_final.* are kept for tapclip_encode_text only; logit_scale is
 * accepted and ignored (FullModel owns its own logit_scale parameter, model_wrapper.py:26).  Returns non-zero for unknown names or wrong shapes. */
TAPCLIP_API int tapclip_load_weight(tapclip_handle h, const char* name, const float* data, int32_t ndim, const int64_t* shape,
                        void* stream);
/* Non-zero (with the list of missing entries in last_error) until every weight has been loaded. */
TAPCLIP_API int tapclip_weights_complete(tapclip_handle h);

/* ---- hot path ------------------------------------------------------------------------------------ */
/* CLIPWrapper.encode_image (clip_wrapper.py:46-47; row A4): images [B,3,R,R] -> out_feat [B,E] (NOT
 * normalised).  North-star extensions (not in the reference; nullable): out_cls_rows = per-layer per-head CLS-row
 * attention probabilities [B, L, H, N] emitted by the attention kernel's probe epilogue; out_rollout = attention
 * rollout of the CLS token over all L layers, [B, N-1] (Abnar & Zuidema: prod_l (0.5*mean_h P_l + 0.5*I), CLS row). */
TAPCLIP_API int tapclip_encode_image(tapclip_handle h, const float* images, int32_t B, float* out_feat, float* out_cls_rows,
                         float* out_rollout, void* stream);

/* Rows A2,A6-A10 for `C` class prompts at once (model_wrapper.py:32-35,47-75; prompt_learner.py:45-66;
 * attribution_monitor.py:24-34; prompt_adjustor.py:35-36):
 *   ctx [C,P,D] learnable context vectors, tok [C,L,D] frozen token embeddings (L = context_length)
 *   mode INTENDED: attribution pass on the un-adjusted prompt, a = softmax_P(mean_h P_last[h, 0:P, T-1]),
 *                  out_attr_raw / out_attr [C,P];   mode LITERAL: a == 1, out_attr [C,1] (raw not produced)
 *   feature pass on [ctx*a | tok], last position, @ text_projection, L2-norm -> out_text_feat [C,E].
 *   mode 2 (ATTRIBUTION_ONLY): the INTENDED attribution pass alone (out_attr_raw / out_attr), no feature pass: used by the
 *   'gate' / 'residual' adjustors (prompt_adjustor.py:38-44), whose small networks run on the host side between the passes.
 * save_for_backward != 0 keeps the activations `tapclip_text_backward` needs (ONE saved forward per handle: the next
 * tapclip_text_forward / tapclip_encode_text replaces them).
 * out_token (nullable): identifies the activations this call saved (0 if none); pass it to tapclip_text_backward.
 * gather_epoch > 0 (after tapclip_text_gather_config; multi-GPU): the C rows are this rank's classes [gather_row_lo, gather_row_lo+C)
 * of n_cls_total; the head kernel ALSO stores them into every rank's symmetric buffer (peer stores over NVLink) and publishes the
 * epoch there -- the text-feature all-gather of model_wrapper.py:79,83's contraction, fused into the producing kernel.  0 = off. */
TAPCLIP_API int tapclip_text_forward(tapclip_handle h, const float* ctx, const float* tok, int32_t C, int32_t P, int32_t mode,
                         int32_t save_for_backward, float* out_attr_raw, float* out_attr, float* out_text_feat,
                         int64_t* out_token, int64_t gather_row_lo, int32_t gather_epoch, void* stream);

/* Fused text-feature all-gather: peer_bufs is a HOST array of `world` device pointers, peer_bufs[r] = rank r's symmetric
 * buffer as mapped into this process (e.g. torch.distributed._symmetric_memory: hdl.buffer_ptrs), each of
 * 2 * align128(n_cls_total * E * 4) + 32 bytes, zero-initialised: two feature slots used by epoch parity, then 8 int32 flags.
 * world = 0 switches the feature off.  Epochs are chosen by the caller: the same strictly increasing sequence on every rank. */
TAPCLIP_API int tapclip_text_gather_config(tapclip_handle h, void* const* peer_bufs, int32_t world, int32_t rank, int64_t n_cls_total);

/* CLIPWrapper.encode_text (clip_wrapper.py:49-51 -> open_clip CLIP.encode_text; SURVEY 8f rank 1 — FullModel never calls
 * it): token_ids int64 [S, context_length] (device) -> out_feat [S,E] (NOT normalised): token + positional embedding,
 * causal transformer, ln_final, pooling at the EOT position (argmax id), @ text_projection. */
TAPCLIP_API int tapclip_encode_text(tapclip_handle h, const int64_t* token_ids, int32_t S, float* out_feat, void* stream);

/* Rows A5,A11,A12 (model_wrapper.py:41,79,83,90-93): out_img_norm [B,E] = L2-normalised image features,
 * out_logits [B,C] = exp(*logit_scale) * img_norm . text_feat^T.  If labels (int64 [B], device) is
 * non-null: out_loss[0] = sum_b CE_b * inv_batch_total and out_dlogits [B,C] = dloss/dlogits.  A label outside [0, C) yields NaN.
 * gather_epoch > 0: text_feat is this rank's symmetric slot of that epoch (tapclip_text_gather_config); the kernel first waits
 * (bounded) until every rank has published the epoch, so no separate collective or barrier precedes it. */
TAPCLIP_API int tapclip_logits(tapclip_handle h, const float* img_feat, const float* text_feat, const float* logit_scale,
                   const int64_t* labels, int32_t B, int32_t C, float inv_batch_total, float* out_img_norm,
                   float* out_logits, float* out_loss, float* out_dlogits, int32_t gather_epoch, void* stream);

/* Backward of the logit contraction: out_d_text [C,E] = exp(s) * dlogits^T . img_norm,
 * out_d_logit_scale[0] = sum dlogits*logits. */
TAPCLIP_API int tapclip_logits_backward(tapclip_handle h, const float* dlogits, const float* logits, const float* img_norm,
                            const float* logit_scale, int32_t B, int32_t C, float* out_d_text,
                            float* out_d_logit_scale, void* stream);

/* Row A13 (autograd of model_wrapper.py:68-75 down to prompt_learner.context_bank): d_text_feat [C,E]
 * -> out_dctx [C,P,D]; activation gradients only (frozen weights), attribution treated as constant.
 * Must follow a tapclip_text_forward(..., save_for_backward=1) with the same C, P.  token = the forward's out_token: if a
 * later forward on this handle has replaced the saved activations the call FAILS (stale backward) instead of differentiating
 * the wrong forward; token 0 skips that check.  C, P (> 0) are checked against the saved forward. */
TAPCLIP_API int tapclip_text_backward(tapclip_handle h, const float* d_text_feat, float* out_dctx, int64_t token, int32_t C, int32_t P,
                          void* stream);

/* train.py:65-67,105: torch.optim.AdamW step fused over a flat fp32 bank of n elements. */
TAPCLIP_API int tapclip_adamw_step(tapclip_handle h, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                       float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream);

/* utils/eval_metrics.py:19-29,58-63: argmax over classes into out_pred (int64 [B], nullable; first maximum wins, NaN counts as
 * the maximum, as torch.argmax) and, with labels, the accuracy counters ACCUMULATED on the device (all int32, nullable):
 * out_correct[0] += #(pred == label); out_class_total[t] += 1 and out_class_correct[t] += (pred == t) for every sample whose
 * label t is in [0, C) ([C] arrays: the reference's per_class_total / per_class_correct dicts).  One device->host read per
 * epoch replaces the reference's per-sample .item() calls. */
TAPCLIP_API int tapclip_argmax_count(tapclip_handle h, const float* logits, const int64_t* labels, int32_t B, int32_t C,
                         int64_t* out_pred, int32_t* out_correct, int32_t* out_class_correct, int32_t* out_class_total,
                         void* stream);

/* Workspace bytes currently held by the handle (for memory accounting). */
TAPCLIP_API int64_t tapclip_workspace_bytes(tapclip_handle h);
/* Number of kernel launches issued by this handle since creation (bench.py `gpu_launches`). */
TAPCLIP_API int64_t tapclip_launch_count(tapclip_handle h);

/* Per-launch timing of the tensor-core kernels (GEMM, attention) with CUDA events on the launching stream.
 * tapclip_profile(h, 1) starts recording, tapclip_profile(h, 0) stops; tapclip_profile_report synchronises the
 * recorded events and returns a JSON summary (string owned by the handle, valid until the next call). */
TAPCLIP_API int tapclip_profile(tapclip_handle h, int32_t enable);
TAPCLIP_API const char* tapclip_profile_report(tapclip_handle h);

/* ---- single-kernel entry points (used by the per-kernel parity tests and micro-benchmarks) -------- */
/* The pre-LN residual block without LayerNorm kernels (north-star "fused ... pre-LN epilogues"): replaces `x = x + attn(ln_1(x))`,
 * `x = x + mlp(ln_2(x))` of open_clip's ResidualAttentionBlock (invoked from models/model_wrapper.py:58,72 and
 * models/clip_wrapper.py:47).
 *   tapclip_op_gemm_resid : x_out[M,N] (fp32, ld_out) = x_in[M,N] (fp32, ld_in; may alias x_out) + A[M,K].W[N,K]^T + bias; optionally
 *       xb[M,N] = x_out - shift in the 16-bit `dtype` (dense), stats[M][parts][2] = per-row partial (sum, sum of squares) of the
 *       shifted rows (parts = tapclip_op_gemm_stats_parts(N)) and shift[M] = the per-row shift: the mean of the row of x_in,
 *       recovered from the statistics describing x_in (stats_prev [M][prev_parts][2] + shift_prev [M]; null: shift 0).  LayerNorm is
 *       invariant to the shift; it keeps the rounded 16-bit values centred.  ld 0 = dense; N % 128 == 0.
 *   tapclip_op_gemm_fold  : out[M,N] (16-bit) = act(LayerNorm(x; gamma, beta).W^T + b) computed from the UN-normalised 16-bit rows xb,
 *       their statistics partials and the folded operands (w_fold = W diag(gamma) with every row centred -- the centring subtracts
 *       the mean of x inside the contraction --, bias_fold = b + W beta):  LN(x) W^T + b = rstd (xb w_fold^T) + bias_fold.
 *       out_pre (nullable, needs act): pre-activation copy.
 *   tapclip_op_fold_ln_weight / tapclip_op_row_stats_cast : build the folded operands / the (xb, stats) pair of arbitrary rows. */
TAPCLIP_API int32_t tapclip_op_gemm_stats_parts(int64_t N);
TAPCLIP_API int tapclip_op_gemm_resid(const void* a, const void* w, const float* bias, const float* x_in, int64_t ld_in, float* x_out,
                          int64_t ld_out, void* xb, float* stats, float* shift, const float* stats_prev, const float* shift_prev,
                          int32_t prev_parts, int64_t M, int64_t N, int64_t K, int32_t dtype, void* stream);
TAPCLIP_API int tapclip_op_gemm_fold(const void* xb, const float* stats, int32_t stats_parts, const void* w_fold, const float* bias_fold,
                         void* out, void* out_pre, int64_t M, int64_t N, int64_t K, int32_t dtype, int32_t act, void* stream);
TAPCLIP_API int tapclip_op_fold_ln_weight(const float* w, const float* bias, const float* gamma, const float* beta, void* w_fold,
                              int32_t dtype, float* bias_fold, int32_t N, int32_t K, void* stream);
TAPCLIP_API int tapclip_op_row_stats_cast(const float* x, void* xb, int32_t dtype, float* stats, float* shift, int64_t rows, int32_t d,
                              void* stream);

/* Image preprocessing on the device: what `CLIPWrapper.get_preprocess()` (models/clip_wrapper.py:64-65,
 * open_clip's inference transform, applied per image at dataset.py:31) does on the host with Pillow/torchvision:
 * Resize(R, BICUBIC) of the shorter side -> CenterCrop(R) -> ToTensor -> Normalize(mean, std).
 *   image_hwc [H, W, 3] uint8 RGB (device), out_chw [3, R, R] fp32 (device); (crop_top, crop_left) = the crop window's
 *   origin in the resized image (torchvision: int(round((size - R) / 2.0))); mean3 / std3: HOST float[3].
 * Bit-exact with Pillow's 8-bit antialiased bicubic resampling (two fixed-point passes, uint8 intermediate). */
TAPCLIP_API int tapclip_op_preprocess(const uint8_t* image_hwc, int32_t H, int32_t W, float* out_chw, int32_t R, int32_t crop_top,
                          int32_t crop_left, const float* mean3, const float* std3, void* stream);

/* out[M,N] = epilogue(A[M,K] . W[N,K]^T + bias).  dtype BF16/FP16: A,W (and epi-0 out) 16-bit, tcgen05 path; FP32: SIMT path.
 * epi: 0 = store activation type (+act, optional out_pre), 1 = store fp32, 2 = fp32 += ,
 *      3 = out = (A.W^T + bias) * act'(out_pre) with out_pre READ as the saved pre-activations (same 16-bit type as A),
 *      4 = as 3 with bf16 A/W/out and fp16 out_pre (the mixed mode's MLP dgrad).  3/4: bf16 A/W, block_n 0|128|256 only.
 * block_n: 0 = choose | 128 | 256 (single-CTA tiles) | 512 (2-CTA cta_group::2 pairs on 256x256 tiles) */
TAPCLIP_API int tapclip_op_gemm(const void* a, const void* w, const float* bias, void* out, void* out_pre, int64_t M, int64_t N,
                    int64_t K, int32_t dtype, int32_t epi, int32_t act, int32_t block_n, void* stream);
TAPCLIP_API int tapclip_op_layernorm(const float* x, int64_t x_row_stride, const float* gamma, const float* beta, void* out,
                         int32_t out_dtype, float* x_copy, int64_t rows, int32_t d, void* stream);
TAPCLIP_API int tapclip_op_layernorm_bwd(const float* dy, const float* x, const float* gamma, float* dx_acc, void* dx_cast,
                             int32_t cast_dtype, int64_t rows, int32_t d, void* stream);
/* probe_mode: 0 none, 1 text column (out [S,H,P]), 2 CLS row (out [S,H,N] with probe_seq_stride) */
TAPCLIP_API int tapclip_op_attention(const void* qkv, void* out, int32_t dtype, int32_t S, int32_t N, int32_t H, int32_t probe_mode,
                         float* probe_out, int32_t probe_P, int64_t probe_seq_stride, void* stream);
TAPCLIP_API int tapclip_op_attention_bwd(const void* qkv, const void* d_out, void* dqkv, int32_t dtype, int32_t S, int32_t N, int32_t H,
                             void* stream);
/* Attention-rollout extension (image side; not in the reference, BASELINE north_star / configs[3]).
 * tapclip_op_attention_lse: lse [S,H,N] = log2 sum_j 2^(c q_i.k_j), c = log2(e)/8, the softmax statistics of a packed qkv;
 *      attn_out NULL: the statistics kernel alone; attn_out [S*N, H*64]: the attention forward itself, emitting lse from its
 *      softmax registers where the selected kernel can (as the engine runs it when the rollout is requested).
 * tapclip_op_rollout_step:  one layer of the CLS-row propagation r_out = 0.5 r_in + (0.5/H) sum_h r_in^T softmax(Q_h K_h^T / 8),
 *      probabilities recomputed from qkv and lse (no N x N map); r_in NULL = e_0 (first call, LAST layer); last != 0 (FIRST
 *      layer) drops the CLS column: r_out is [S, N-1], otherwise [S, N]. */
TAPCLIP_API int tapclip_op_attention_lse(const void* qkv, void* attn_out, float* lse, int32_t dtype, int32_t S, int32_t N, int32_t H,
                             void* stream);
TAPCLIP_API int tapclip_op_rollout_step(const void* qkv, const float* lse, const float* r_in, float* r_out, int32_t dtype, int32_t S,
                            int32_t N, int32_t H, int32_t last, void* stream);
TAPCLIP_API int tapclip_op_attribution(const float* probe, float* raw, float* attr, int32_t C, int32_t H, int32_t P, void* stream);
TAPCLIP_API int tapclip_op_cast(const float* src, void* dst, int32_t dst_dtype, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TAPCLIP_H_ */
