// extern "C" surface of libtapclip (see include/tapclip.h for the contract and the reference lines replaced).
#include "engine.h"

using namespace tapclip;

#define TC_API_BEGIN try {
#define TC_API_END                                                     \
    return 0;                                                          \
    }                                                                  \
    catch (const tapclip::Error& e) { set_error(e.msg); return 1; }    \
    catch (const std::exception& e) { set_error(e.what()); return 2; } \
    catch (...) { set_error("unknown error"); return 3; }

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
#define NEED(h) TC_CHECK((h) != nullptr, "null tapclip handle")

extern "C" {

TAPCLIP_API const char* tapclip_last_error(void) { return get_error(); }
TAPCLIP_API const char* tapclip_version(void) { return "tapclip-b200 0.1 (sm_100a: tcgen05/TMEM/TMA GEMM, attention and rollout)"; }

TAPCLIP_API int tapclip_create(const tapclip_config* cfg, tapclip_handle* out) {
    TC_API_BEGIN
    TC_CHECK(cfg != nullptr && out != nullptr, "null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    TC_CHECK(e == cudaSuccess && n > 0, "no CUDA device available: libtapclip has no CPU fallback");
    *out = new tapclip_engine(*cfg);
    TC_API_END
}

TAPCLIP_API int tapclip_destroy(tapclip_handle h) {
    TC_API_BEGIN
    delete h;
    TC_API_END
}

TAPCLIP_API int tapclip_load_weight(tapclip_handle h, const char* name, const float* data, int32_t ndim, const int64_t* shape, void* stream) {
    TC_API_BEGIN
    NEED(h);
    TC_CHECK(name != nullptr && shape != nullptr, "null argument");
    h->impl.load_weight(name, data, ndim, shape, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_weights_complete(tapclip_handle h) {
    TC_API_BEGIN
    NEED(h);
    const std::string m = h->impl.missing_weights();
    TC_CHECK(m.empty(), "weights missing: %s", m.c_str());
    TC_API_END
}

TAPCLIP_API int tapclip_encode_image(tapclip_handle h, const float* images, int32_t B, float* out_feat, float* out_cls_rows,
                                     float* out_rollout, void* stream) {
    TC_API_BEGIN
    NEED(h);
    TC_CHECK(B == 0 || (images != nullptr && out_feat != nullptr), "null argument");
    h->impl.encode_image(images, B, out_feat, out_cls_rows, out_rollout, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_text_forward(tapclip_handle h, const float* ctx, const float* tok, int32_t C, int32_t P, int32_t mode,
                         int32_t save_for_backward, float* out_attr_raw, float* out_attr, float* out_text_feat, int64_t* out_token,
                         int64_t gather_row_lo, int32_t gather_epoch, void* stream) {
    TC_API_BEGIN
    NEED(h);
    TC_CHECK(C == 0 || (ctx != nullptr && tok != nullptr), "null argument");
    const int64_t token = h->impl.text_forward(ctx, tok, C, P, mode, save_for_backward != 0, out_attr_raw, out_attr, out_text_feat, S(stream),
                                               gather_row_lo, gather_epoch);
    if (out_token) *out_token = token;
    TC_API_END
}

TAPCLIP_API int tapclip_text_gather_config(tapclip_handle h, void* const* peer_bufs, int32_t world, int32_t rank, int64_t n_cls_total) {
    TC_API_BEGIN
    NEED(h);
    TC_CHECK(world == 0 || peer_bufs != nullptr, "null argument");
    h->impl.set_text_gather(peer_bufs, world, rank, n_cls_total);
    TC_API_END
}

TAPCLIP_API int tapclip_encode_text(tapclip_handle h, const int64_t* token_ids, int32_t n_seq, float* out_feat, void* stream) {
    TC_API_BEGIN
    NEED(h);
    TC_CHECK(n_seq == 0 || (token_ids != nullptr && out_feat != nullptr), "null argument");
    h->impl.encode_text(token_ids, n_seq, out_feat, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_logits(tapclip_handle h, const float* img_feat, const float* text_feat, const float* logit_scale, const int64_t* labels,
                   int32_t B, int32_t C, float inv_batch_total, float* out_img_norm, float* out_logits, float* out_loss,
                   float* out_dlogits, int32_t gather_epoch, void* stream) {
    TC_API_BEGIN
    NEED(h);
    TC_CHECK(img_feat && text_feat && logit_scale && out_img_norm && out_logits, "null argument");
    h->impl.logits(img_feat, text_feat, logit_scale, labels, B, C, inv_batch_total, out_img_norm, out_logits, out_loss, out_dlogits, S(stream),
                   gather_epoch);
    TC_API_END
}

TAPCLIP_API int tapclip_logits_backward(tapclip_handle h, const float* dlogits, const float* logits, const float* img_norm, const float* logit_scale,
                            int32_t B, int32_t C, float* out_d_text, float* out_d_logit_scale, void* stream) {
    TC_API_BEGIN
    NEED(h);
    TC_CHECK(dlogits && logits && img_norm && logit_scale && out_d_text && out_d_logit_scale, "null argument");
    h->impl.logits_backward(dlogits, logits, img_norm, logit_scale, B, C, out_d_text, out_d_logit_scale, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_text_backward(tapclip_handle h, const float* d_text_feat, float* out_dctx, int64_t token, int32_t C, int32_t P,
                          void* stream) {
    TC_API_BEGIN
    NEED(h);
    TC_CHECK(d_text_feat && out_dctx, "null argument");
    h->impl.text_backward(d_text_feat, out_dctx, S(stream), token, C, P);
    TC_API_END
}

TAPCLIP_API int tapclip_adamw_step(tapclip_handle h, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                       float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream) {
    TC_API_BEGIN
    NEED(h);
    TC_CHECK(step >= 1, "AdamW step counter starts at 1");
    adamw_step(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, S(stream));
    ++h->impl.launches;
    TC_API_END
}

TAPCLIP_API int tapclip_argmax_count(tapclip_handle h, const float* logits, const int64_t* labels, int32_t B, int32_t C, int64_t* out_pred,
                         int32_t* out_correct, int32_t* out_class_correct, int32_t* out_class_total, void* stream) {
    TC_API_BEGIN
    NEED(h);
    TC_CHECK(B == 0 || logits != nullptr, "null argument");
    TC_CHECK(labels != nullptr || (out_correct == nullptr && out_class_correct == nullptr && out_class_total == nullptr), "counters need labels");
    argmax_count(logits, labels, out_pred, out_correct, out_class_correct, out_class_total, B, C, S(stream));
    ++h->impl.launches;
    TC_API_END
}

TAPCLIP_API int64_t tapclip_workspace_bytes(tapclip_handle h) { return h ? h->impl.workspace_bytes() : -1; }
TAPCLIP_API int64_t tapclip_launch_count(tapclip_handle h) { return h ? h->impl.launches : -1; }

TAPCLIP_API int tapclip_profile(tapclip_handle h, int32_t enable) {
    TC_API_BEGIN
    NEED(h);
    h->impl.profiling = enable != 0;
    TC_API_END
}
TAPCLIP_API const char* tapclip_profile_report(tapclip_handle h) { return h ? h->impl.profile_report() : "{}"; }

// ---- single-kernel entry points ------------------------------------------------------------------------
TAPCLIP_API int tapclip_op_gemm(const void* a, const void* w, const float* bias, void* out, void* out_pre, int64_t M, int64_t N, int64_t K,
                    int32_t dtype, int32_t epi, int32_t act, int32_t block_n, void* stream) {
    TC_API_BEGIN
    GemmArgs g;
    g.a = a; g.w = w; g.bias = bias; g.out = out; g.out_pre = out_pre;
    g.M = M; g.N = N; g.K = K; g.lda = K; g.ldw = K; g.ldo = N; g.epi = epi; g.act = act; g.block_n = block_n; g.dt = dtype; g.aux_dt = dtype;
    if (epi == 4) { g.epi = EPI_BF16_ACTGRAD; g.aux_dt = DT_F16; }   // bf16 gradients against fp16 saved pre-activations (mixed mode)
    if (dtype != DT_F32) gemm_tc(g, S(stream));
    else gemm_simt_f32(g, S(stream));
    TC_API_END
}

TAPCLIP_API int32_t tapclip_op_gemm_stats_parts(int64_t N) { return gemm_stats_parts(N); }

TAPCLIP_API int tapclip_op_gemm_resid(const void* a, const void* w, const float* bias, const float* x_in, int64_t ld_in, float* x_out,
                          int64_t ld_out, void* xb, float* stats, float* shift, const float* stats_prev, const float* shift_prev,
                          int32_t prev_parts, int64_t M, int64_t N, int64_t K, int32_t dtype, void* stream) {
    TC_API_BEGIN
    GemmArgs g;
    g.a = a; g.w = w; g.bias = bias; g.out = x_out;
    g.M = M; g.N = N; g.K = K; g.lda = K; g.ldw = K; g.ldo = ld_out ? ld_out : N; g.epi = EPI_F32_RESID; g.act = ACT_NONE; g.dt = dtype;
    g.resid_in = x_in; g.ld_in = ld_in ? ld_in : N; g.xb = xb; g.stats_out = stats; g.shift_out = shift;
    g.stats_prev = stats_prev; g.shift_prev = shift_prev; g.prev_parts = prev_parts;
    gemm_tc(g, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_op_gemm_fold(const void* xb, const float* stats, int32_t stats_parts, const void* w_fold, const float* bias_fold,
                         void* out, void* out_pre, int64_t M, int64_t N, int64_t K, int32_t dtype, int32_t act, void* stream) {
    TC_API_BEGIN
    GemmArgs g;
    g.a = xb; g.w = w_fold; g.bias = bias_fold; g.out = out; g.out_pre = out_pre;
    g.M = M; g.N = N; g.K = K; g.lda = K; g.ldw = K; g.ldo = N; g.epi = EPI_BF16; g.act = act; g.dt = dtype;
    g.stats_in = stats; g.stats_parts = stats_parts;
    TC_CHECK(stats != nullptr, "stats is required");
    gemm_tc(g, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_op_fold_ln_weight(const float* w, const float* bias, const float* gamma, const float* beta, void* w_fold,
                              int32_t dtype, float* bias_fold, int32_t N, int32_t K, void* stream) {
    TC_API_BEGIN
    fold_ln_weight(w, bias, gamma, beta, w_fold, dtype, bias_fold, N, K, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_op_row_stats_cast(const float* x, void* xb, int32_t dtype, float* stats, float* shift, int64_t rows, int32_t d,
                              void* stream) {
    TC_API_BEGIN
    row_stats_cast(x, xb, dtype, stats, shift, rows, d, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_op_preprocess(const uint8_t* image_hwc, int32_t H, int32_t W, float* out_chw, int32_t R, int32_t crop_top,
                          int32_t crop_left, const float* mean3, const float* std3, void* stream) {
    TC_API_BEGIN
    TC_CHECK(image_hwc != nullptr && out_chw != nullptr && mean3 != nullptr && std3 != nullptr, "null argument");
    preprocess_image(image_hwc, H, W, out_chw, R, crop_top, crop_left, mean3, std3, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_op_layernorm(const float* x, int64_t x_row_stride, const float* gamma, const float* beta, void* out, int32_t out_dtype,
                         float* x_copy, int64_t rows, int32_t d, void* stream) {
    TC_API_BEGIN
    layernorm_fwd(x, x_row_stride, gamma, beta, out, out_dtype, x_copy, rows, d, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_op_layernorm_bwd(const float* dy, const float* x, const float* gamma, float* dx_acc, void* dx_cast, int32_t cast_dtype,
                             int64_t rows, int32_t d, void* stream) {
    TC_API_BEGIN
    layernorm_bwd(dy, x, gamma, dx_acc, dx_cast, cast_dtype, rows, d, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_op_attention(const void* qkv, void* out, int32_t dtype, int32_t S_, int32_t N, int32_t H, int32_t probe_mode,
                         float* probe_out, int32_t probe_P, int64_t probe_seq_stride, void* stream) {
    TC_API_BEGIN
    AttnProbe p;
    p.mode = probe_mode; p.out = probe_out; p.P = probe_P; p.seq_stride = probe_seq_stride;
    attention_fwd(qkv, out, dtype, S_, N, H, p, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_op_attention_bwd(const void* qkv, const void* d_out, void* dqkv, int32_t dtype, int32_t S_, int32_t N, int32_t H,
                             void* stream) {
    TC_API_BEGIN
    // dtype = type of the saved qkv; gradients are fp32 in fp32 mode and bf16 otherwise
    attention_bwd(qkv, dtype, d_out, dqkv, dtype == DT_F32 ? DT_F32 : DT_BF16, S_, N, H, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_op_attention_lse(const void* qkv, void* attn_out, float* lse, int32_t dtype, int32_t S_, int32_t N, int32_t H,
                             void* stream) {
    TC_API_BEGIN
    if (attn_out) {
        AttnProbe p;
        p.lse_out = lse;
        attention_fwd(qkv, attn_out, dtype, S_, N, H, p, S(stream));
    } else {
        attention_lse(qkv, lse, dtype, S_, N, H, S(stream));
    }
    TC_API_END
}

TAPCLIP_API int tapclip_op_rollout_step(const void* qkv, const float* lse, const float* r_in, float* r_out, int32_t dtype, int32_t S_,
                            int32_t N, int32_t H, int32_t last, void* stream) {
    TC_API_BEGIN
    rollout_step(qkv, lse, r_in, r_out, dtype, S_, N, H, last != 0, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_op_attribution(const float* probe, float* raw, float* attr, int32_t C, int32_t H, int32_t P, void* stream) {
    TC_API_BEGIN
    attribution_reduce(probe, raw, attr, C, H, P, S(stream));
    TC_API_END
}

TAPCLIP_API int tapclip_op_cast(const float* src, void* dst, int32_t dst_dtype, int64_t n, void* stream) {
    TC_API_BEGIN
    cast_f32(src, dst, dst_dtype, n, S(stream));
    TC_API_END
}

}  // extern "C"
