// Launchers for the non-GEMM kernels of the TAP-CLIP hot path (all on the caller's stream).
#pragma once
#include "common.cuh"

namespace tapclip {

// ---- norm.cu ---------------------------------------------------------------------------------------
// LayerNorm over the last dim (eps 1e-5, torch semantics): out[r,:] = (x[r,:]-mean)*rstd*gamma+beta.
// x rows are `x_row_stride` floats apart (lets ln_post read only the CLS rows); out is dense [rows,d].
// out_dt (DType) selects the activation type.  x_copy (optional, dense fp32) receives the unnormalised row
// (saved for the backward pass).  In-place fp32 (out == x) is allowed.
void layernorm_fwd(const float* x, int64_t x_row_stride, const float* gamma, const float* beta, void* out,
                   int out_dt, float* x_copy, int64_t rows, int d, cudaStream_t stream);
// x[b,t,:] = LayerNorm(cat([cls, patch_out[b]])[t,:] + pos[t,:]) with (gamma, beta) = ln_pre: the vision tower's prologue
// xb / stats / shift (optional, together): shift[r] = mean of row r, xb = the rows minus their shift in bf16, stats[r] = (sum, sum of
// squares) of the shifted row -- the inputs of the first block's folded-LayerNorm QKV GEMM (gemm.h)
void assemble_ln_pre(const float* patch_out, const float* cls, const float* pos, const float* gamma, const float* beta, float* x,
                     int B, int n_tokens, int d, cudaStream_t stream, void* xb = nullptr, float* stats = nullptr, float* shift = nullptr);
// shift[r] = mean_k x[r,k];  xb[r,:] = x[r,:] - shift[r] in the 16-bit type xb_dt;  stats[r] = (sum, sum of squares) of the shifted
// row: the inputs of a folded-LayerNorm GEMM (gemm.h, GemmArgs::stats_in with stats_parts = 1) for rows that no EPI_F32_RESID GEMM
// produced
void row_stats_cast(const float* x, void* xb, int xb_dt, float* stats, float* shift, int64_t rows, int d, cudaStream_t stream);
// LayerNorm(gamma, beta) followed by Linear(W [N,K], bias) folded into one GEMM over the un-normalised rows (gemm.h):
// Wf = W diag(gamma) with every row centred (its mean over k subtracted), in the 16-bit type dt;  fb = bias + W beta
void fold_ln_weight(const float* W, const float* bias, const float* gamma, const float* beta, void* Wf, int dt, float* fb, int N, int K,
                    cudaStream_t stream);
// dx_acc[r,:] += LN'(dy[r,:]; x[r,:], gamma)  (mean/rstd recomputed from x);  dx_cast (optional, activation
// type) receives the updated dx_acc row cast to the activation type.
// dy and x are dense [rows, d]; dx_acc / dx_cast rows are dx_row_stride elements apart (0 = dense).
void layernorm_bwd(const float* dy, const float* x, const float* gamma, float* dx_acc, void* dx_cast,
                   int cast_dt, int64_t rows, int d, cudaStream_t stream, int64_t dx_row_stride = 0);
// out[r,:] = x[r,:] / ||x[r,:]||_2 ; inv_norm[r] (optional) = 1/||x||
void l2norm_fwd(const float* x, float* out, float* inv_norm, int64_t rows, int d, cudaStream_t stream);
// dx = (g - xhat * <xhat, g>) * inv_norm ; optional cast copy in the activation type
void l2norm_bwd(const float* g, const float* xhat, const float* inv_norm, float* dx, void* dx_cast, int cast_dt,
                int64_t rows, int d, cudaStream_t stream);

// ---- head.cu (K4, K5) ---------------------------------------------------------------------------------
// K5: where text_head also stores its rows (peer-mapped symmetric buffers of all ranks, this rank included) and publishes an epoch
struct PeerScatter {
    float* dst[8];        // per rank: base of the [n_cls_total, E] slot of this epoch's parity
    int* flag[8];         // per rank: &flags[this rank] in that rank's buffer
    int world;            // 0 = no peer stores
    int epoch;
    int64_t row_lo;       // global class index of this rank's first row
    int* ticket;          // one int, zero before the first use (self-resetting)
};
// tfeat[c,:] = l2norm(x[(c*row_stride + row_offset),:] @ w_proj), w_proj [D,E] (text_projection as stored by open_clip) of type w_dt;
// inv_norm[c] = 1/||.||;
// tfeat_copy (optional) receives the same rows (the caller's output tensor)
void text_head(const float* x, int64_t row_stride, int64_t row_offset, const void* w_proj, int w_dt, float* tfeat, float* inv_norm,
               float* tfeat_copy, int C, int D, int E, cudaStream_t stream, const PeerScatter* peers = nullptr);
// backward of text_head: dx[(c*row_stride + row_offset),:] = l2norm'(g[c,:]) @ wt_proj (wt_proj [E,D] = text_projection^T, type w_dt)
// + copy in the gradient type cast_dt
void text_head_bwd(const float* g, const float* tfeat, const float* inv_norm, const void* wt_proj, int w_dt, float* dx, void* dx_cast,
                   int cast_dt, int64_t row_stride, int64_t row_offset, int C, int D, int E, cudaStream_t stream);
// img_norm = l2norm(img); logits = exp(*logit_scale) * img_norm . txt^T; with labels: loss[0] = sum_b CE_b * inv_batch_total (summed
// by the last CTA in row order), dlogits = dloss/dlogits.  row_scratch [B] floats, ticket: one int, zero before the first use
// wait_world > 0 (K5): txt is this rank's symmetric slot; the kernel first waits until wait_flags[r] >= wait_epoch for every rank r
void logits_ce(const float* img, const float* txt, const float* logit_scale, const int64_t* labels, float* img_norm, float* logits,
               float* loss, float* dlogits, float* row_scratch, int* ticket, int B, int C, int E, float inv_batch_total, cudaStream_t stream,
               const int* wait_flags = nullptr, int wait_world = 0, int wait_epoch = 0);
// d_txt[c,:] = exp(s) * sum_b dlogits[b,c] * img[b,:];  d_scale[0] = sum dlogits * logits.  class_scratch [C] floats, ticket as above
void logits_bwd_fused(const float* dlogits, const float* logits, const float* img, const float* logit_scale, float* d_txt, float* d_scale,
                      float* class_scratch, int* ticket, int B, int C, int E, cudaStream_t stream);

// ---- attention.cu ----------------------------------------------------------------------------------
enum ProbeMode : int { PROBE_NONE = 0, PROBE_TEXT_COL = 1, PROBE_CLS_ROW = 2 };
struct AttnProbe {
    int mode = PROBE_NONE;
    float* out = nullptr;        // TEXT_COL: [S,H,P] prob(query r<P, key N-1);  CLS_ROW: prob(query 0, key n) at
                                 // out[s*seq_stride + h*N + n]
    int P = 0;
    int64_t seq_stride = 0;
    bool causal = false;         // key j visible to query i only if j <= i (standard CLIP text tower, encode_text)
    int live_q_rows = 0;         // > 0: only query rows [0, live_q_rows) are consumed downstream (last block: the CLS row); the
                                 // tcgen05 kernel then skips whole 128-row query tiles past them, other rows of `out` are unspecified
    float* lse_out = nullptr;    // rollout extension: [S,H,N] softmax statistics log2 sum_j 2^(c q_i.k_j), c = log2(e)/8, of every
                                 // computed query row (non-causal only); rows past live_q_rows may be left unwritten
};
// qkv [S*N, 3*H*64] (packed in_proj output, activation type) -> out [S*N, H*64]; softmax(QK^T/8)V, no mask.
// bf16 / fp16: mma.sync tensor-core flash kernel; fp32: SIMT kernel.  Probabilities asked for by `probe` are
// emitted from the softmax registers; the N x N map is never written.
void attention_fwd(const void* qkv, void* out, int dt, int S, int N, int H, const AttnProbe& probe,
                   cudaStream_t stream);
// tcgen05 version of attention_fwd for 16-bit inputs (attention_tc.cu); same contract, except that probe.lse_out is only
// written by the variants that return true (attention_fwd completes it with attention_lse otherwise)
bool attention_fwd_tc_supported(int dt, int N);
bool attention_fwd_tc(const void* qkv, void* out, int dt, int S, int N, int H, const AttnProbe& probe, cudaStream_t stream);
// rollout extension (rollout.cu): attention_lse writes the softmax statistics lse [S,H,N] of a packed qkv (see
// AttnProbe::lse_out); rollout_step propagates the CLS row of the attention rollout through ONE layer,
//   r_out[s,j] = 0.5 r_in[s,j] + (0.5/H) sum_h sum_i r_in[s,i] softmax_j(q_i.k_j/8),
// recomputing the probabilities from qkv and lse (no N x N map in memory).  r_in = nullptr: e_0 (start, at the LAST layer);
// last = true (the FIRST layer): the CLS column is dropped, r_out is [S, N-1].
void attention_lse(const void* qkv, float* lse, int dt, int S, int N, int H, cudaStream_t stream);
void rollout_step(const void* qkv, const float* lse, const float* r_in, float* r_out, int dt, int S, int N, int H, bool last,
                  cudaStream_t stream);
// tcgen05 version for long 16-bit sequences (rollout_tc.cu); rollout_step dispatches to it (TAPCLIP_ROLLOUT_IMPL=1: mma.sync)
bool rollout_step_tc_supported(int dt, int N);
int rollout_step_tc_ctas(int S, int N);            // CTAs the tcgen05 kernel would launch
void rollout_step_tc(const void* qkv, const float* lse, const float* r_in, float* r_out, int dt, int S, int N, int H, bool last,
                     cudaStream_t stream);
// dqkv [S*N, 3*H*64] from d_out [S*N, H*64] and the saved qkv (probabilities recomputed). N <= 128.
// qkv may be fp16 (mixed mode) while gradients are bf16; fp32 mode: everything fp32.
void attention_bwd(const void* qkv, int qkv_dt, const void* d_out, void* dqkv, int grad_dt, int S, int N, int H,
                   cudaStream_t stream);

// ---- elementwise.cu --------------------------------------------------------------------------------
// images [B,3,R,R] fp32 NCHW -> patches [B*g*g, kpad] (k = c*p*p + py*p + px, zero padded to kpad)
void patchify(const float* images, void* out, int out_dt, int B, int R, int p, int kpad, cudaStream_t stream);
// x[b,0,:] = cls + pos[0];  x[b,1+i,:] = patch_out[b*g2+i,:] + pos[1+i]
void assemble_tokens(const float* patch_out, const float* cls, const float* pos, float* x, int B, int n_tokens, int d,
                     cudaStream_t stream);
// x[c,t,:] = t<P ? ctx[c,t,:]*attr[c, attr_p==1?0:t] : tok[c,t-P,:]   (attr == nullptr -> 1)
void splice_prompts(const float* ctx, const float* tok, const float* attr, int attr_p, float* x, int C, int P, int L,
                    int D, cudaStream_t stream);
// dctx[c,t,:] = dx[c,t,:] * attr[c,...]  for t<P
void splice_bwd(const float* dx, const float* attr, int attr_p, float* dctx, int C, int P, int T, int D,
                cudaStream_t stream);
// x[s,t,:] = token_embedding[ids[s,t],:] + pos[t,:];  eot[s] = argmax_t ids[s,t]   (open_clip CLIP.encode_text prologue)
void embed_tokens(const int64_t* ids, const float* token_embedding, const float* pos, float* x, int32_t* eot, int S, int L, int D,
                  cudaStream_t stream);
// out[r,:] = x[(r*row_stride + idx[r]),:]   (fp32 gather of one row per sequence)
void gather_rows_indexed(const float* x, const int32_t* idx, float* out, int64_t rows, int64_t row_stride, int d, cudaStream_t stream);
// raw[c,p] = mean_h probe[c,h,p];  attr[c,:] = softmax_p(raw[c,:])      (K3)
void attribution_reduce(const float* probe, float* raw, float* attr, int C, int H, int P, cudaStream_t stream);
// out[r,:] = cast(x[(r*row_stride + row_offset),:])
void gather_rows(const float* x, void* out, int out_dt, int64_t rows, int64_t row_stride, int64_t row_offset, int d,
                 cudaStream_t stream);
// dst[(r*row_stride+row_offset),:] = src[r,:] (dst fp32 pre-zeroed) + optional cast copy into dst_cast
void scatter_rows(const float* src, float* dst, void* dst_cast, int cast_dt, int64_t rows, int64_t row_stride,
                  int64_t row_offset, int d, cudaStream_t stream);
void cast_f32(const float* src, void* dst, int dst_dt, int64_t n, cudaStream_t stream);
// dh = dh * act'(h_pre)   (activation type, in place)
void act_bwd_inplace(void* dh, int dh_dt, const void* h_pre, int h_dt, int act, int64_t n, cudaStream_t stream);
// logits[b,c] = exp(*logit_scale) * <img[b,:], txt[c,:]>
void cosine_logits(const float* img, const float* txt, const float* logit_scale, float* logits, int B, int C, int E,
                   cudaStream_t stream);
// loss = mean_b CE(logits[b,:], labels[b]) * B/B_total ... see elementwise.cu; dlogits = dloss/dlogits
void cross_entropy(const float* logits, const int64_t* labels, float* loss, float* dlogits, float* row_scratch, int B, int C,
                   float inv_batch_total, cudaStream_t stream);
// d_txt[c,:] = exp(s) * sum_b dlogits[b,c]*img[b,:];  d_scale = sum_{b,c} dlogits[b,c]*logits[b,c]
void logits_bwd(const float* dlogits, const float* logits, const float* img, const float* logit_scale, float* d_txt,
                float* d_scale, float* class_scratch /*[C]*/, int B, int C, int E, cudaStream_t stream);
// p -= lr*(m_hat/(sqrt(v_hat)+eps) + wd*p)  — torch.optim.AdamW semantics, one fused pass over a flat bank
void adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                float wd, int step, cudaStream_t stream);
// argmax over classes + counters kept on the device (utils/eval_metrics.py:19-29,58-63 without per-sample .item()): correct[0] +=
// #(argmax == label); class_total[t] += 1 and class_correct[t] += (argmax == t) for every sample with label t (int32 [C], nullable)
void argmax_count(const float* logits, const int64_t* labels, int64_t* pred, int* correct, int* class_correct, int* class_total, int B,
                  int C, cudaStream_t stream);

// ---- preprocess.cu ---------------------------------------------------------------------------------
// img [H, W, 3] uint8 RGB (device) -> out [3, R, R] fp32: bicubic resize of the shorter side to R (Pillow's antialiased
// fixed-point resampling), crop window at (top, left) of the resized image, /255, (x - mean) / std.  mean / stdv: host float[3].
void preprocess_image(const uint8_t* img, int H, int W, float* out, int R, int top, int left, const float* mean, const float* stdv,
                      cudaStream_t stream);

}  // namespace tapclip
