// Engine state behind the opaque tapclip_handle.
#pragma once
#include <cstdlib>
#include "../../include/tapclip.h"
#include "gemm.h"
#include "kernels.h"

#include <map>
#include <string>
#include <vector>

namespace tapclip {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    void ensure(size_t n);      // grow-only
    void release();
};

struct BlockWeights {
    float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
    float *b_qkv = nullptr, *b_o = nullptr, *b_fc = nullptr, *b_proj = nullptr;
    void *w_qkv = nullptr, *w_o = nullptr, *w_fc = nullptr, *w_proj = nullptr;        // [N,K], activation type
    void *wt_qkv = nullptr, *wt_o = nullptr, *wt_fc = nullptr, *wt_proj = nullptr;    // [K,N] dgrad operands (text tower)
    // LayerNorm folded into the consumer GEMM (16-bit modes; gemm.h GemmArgs::stats_in): W" = W diag(gamma), rows centred, in the
    // operand type; b' = b + W beta; built from fp32 copies of W as soon as W, b, gamma and beta of a group are loaded
    float *f32_qkv = nullptr, *f32_fc = nullptr;
    void *wf_qkv = nullptr, *wf_fc = nullptr;
    float *fb_qkv = nullptr, *fb_fc = nullptr;
};

struct Engine {
    tapclip_config cfg;
    int vdt = DT_BF16;          // activation / operand type of the image tower
    int tdt = DT_BF16;          // ... of the text-tower forward (fp16 in mixed mode: see DESIGN.md "Precision")
    int gdt = DT_BF16;          // ... of the backward pass (gradients are never fp16)
    int esz = 2;                // bytes per activation element (same for all three)
    int grid = 0, n_tok = 0, kpatch = 0, kpatch_pad = 0, vocab = 0;
    int64_t launches = 0;

    std::map<std::string, void*> weights;     // owning storage, keyed by state-dict name (+"#T" for transposes)
    void* w_patch = nullptr; float* cls_emb = nullptr; float* pos_emb = nullptr;
    float *ln_pre_g = nullptr, *ln_pre_b = nullptr, *ln_post_g = nullptr, *ln_post_b = nullptr;
    void *w_vproj = nullptr, *w_tproj = nullptr, *wt_tproj = nullptr, *wde_tproj = nullptr;
    float *tok_emb = nullptr, *text_pos = nullptr, *ln_final_g = nullptr, *ln_final_b = nullptr;   // standard encode_text only
    std::vector<BlockWeights> vis, txt;

    // workspaces (grow-only)
    DevBuf v_patches, v_patch_out, v_x, v_ln, v_qkv, v_attn, v_h, v_pooled, v_roll_qkv, v_lse, v_roll;
    DevBuf t_x, t_ln, t_qkv, t_attn, t_h, t_pooled, t_feat, t_tfeat, t_inv_norm, t_probe, t_attr, t_attr_raw;
    // folded-LayerNorm path: 16-bit (row-shifted) residual copy, live rows of the last block, and the row statistics + shifts in
    // two alternating sets (a residual GEMM reads the set describing its input while it writes the set describing its output)
    DevBuf v_xb, v_xlive, t_xb, t_xlive;
    struct RowStats {
        DevBuf stats[2], shift[2];
        int cur = 0, parts = 1;                                     // set describing the current residual rows; partials per row in it
        void ensure(int64_t rows, int max_parts) { for (int i = 0; i < 2; ++i) { stats[i].ensure((size_t)rows * max_parts * 2 * sizeof(float)); shift[i].ensure((size_t)rows * sizeof(float)); } }
        float* s(int which) { return (float*)stats[which].p; }
        float* h(int which) { return (float*)shift[which].p; }
    } v_rs, t_rs;
    DevBuf t_save_x, t_save_qkv, t_save_h;
    DevBuf b_dx, b_dxc, b_dh, b_dln, b_dattn, b_dqkv, b_dfeat, b_dfeatc, b_dpool;
    DevBuf s_rows, s_cls, s_ticket, e_eot, e_pool;
    int* tickets();
    // K5: text-feature all-gather fused into the head kernel.  peers[r] = rank r's symmetric buffer as mapped into THIS
    // process: [2 slots][n_cls * E] floats, then 8 int flags (flag[r] = last epoch rank r published here)
    struct { int world = 0, rank = 0; int64_t n_cls = 0; uint8_t* peers[8] = {}; } gather;
    size_t gather_slot_bytes() const { return ((size_t)gather.n_cls * cfg.embed_dim * sizeof(float) + 127) & ~(size_t)127; }
    void set_text_gather(void* const* peer_bufs, int world, int rank, int64_t n_cls_total);
    // TAPCLIP_FUSE_HEAD=0: the head (pool + projection + L2-norm, logits + CE, their backward) as separate launches (A/B parity)
    bool fuse_head = !(getenv("TAPCLIP_FUSE_HEAD") && atoi(getenv("TAPCLIP_FUSE_HEAD")) == 0);
    // activations kept by text_forward(save) for text_backward; `token` identifies the forward that wrote them (a later
    // text_forward / encode_text on this handle overwrites the slot: its backward must then fail, not use the wrong tensors)
    struct { bool valid = false; int C = 0, P = 0, T = 0, PA = 1; bool has_attr = false; bool dead_last = false; int64_t token = 0; } saved;
    int64_t forward_seq = 0;

    // optional per-launch CUDA-event timing of the tensor-core kernels (bench.py roofline numbers)
    struct ProfRec { cudaEvent_t a, b; double flops; int kind; int64_t M, N, K; int epi; };
    bool profiling = false;
    // last-block dead-row elimination (TAPCLIP_DEAD_ROWS=0 disables it: measurement / A-B parity only)
    bool dead_rows = !(getenv("TAPCLIP_DEAD_ROWS") && atoi(getenv("TAPCLIP_DEAD_ROWS")) == 0);
    std::vector<ProfRec> prof;
    std::string prof_report;
    void prof_begin(ProfRec& r, cudaStream_t st);
    void prof_end(ProfRec& r, cudaStream_t st);
    const char* profile_report();

    explicit Engine(const tapclip_config& c);
    ~Engine();
    std::vector<DevBuf*> all_bufs();
    int64_t workspace_bytes();

    void* store(const std::string& key, const float* src, int R, int C, int dst_ld, bool transpose, int dt, cudaStream_t st);
    void load_weight(const std::string& name, const float* data, int ndim, const int64_t* shape, cudaStream_t st);
    std::string missing_weights() const;

    void gemm(const void* a, const void* w, const float* bias, void* out, void* out_pre, int64_t M, int64_t N, int64_t K, int epi,
              int act, int dt, cudaStream_t st, int aux_dt = DT_BF16, int64_t lda = 0, int64_t ldo = 0);   // lda/ldo 0 = dense
    void attn_fwd(const void* qkv, void* out, int dt, int S, int N, int H, const AttnProbe& probe, cudaStream_t st);
    void attn_bwd(const void* qkv, const void* d_out, void* dqkv, int S, int N, int H, cudaStream_t st);
    void block_forward(const BlockWeights& b, float* x, int S, int N, int d, int H, int dt, DevBuf& ln, DevBuf& qkv, DevBuf& attn,
                       DevBuf& hbuf, const AttnProbe& probe, bool probs_only, int save_slot, cudaStream_t st, void* rollout_qkv = nullptr,
                       int live_row = -1);
    // The pre-LN block without LayerNorm kernels (north-star "pre-LN epilogue", gemm.h): the residual GEMMs (EPI_F32_RESID) emit the
    // updated rows in 16 bits plus per-row (sum, sum of squares); the QKV / c_fc GEMMs consume them with LayerNorm folded into the
    // weights.  `x` follows the residual stream (it hops through the save slots when save_slot >= 0; after a last block with
    // live_row >= 0 it points at the compact [S, d] live rows); `parts` = statistics partials per row currently in `stats`.
    void block_forward_fused(const BlockWeights& b, float*& x, RowStats& rs, float* scratch, int S, int N, int d, int H, int dt, DevBuf& xb,
                             DevBuf& xlive, DevBuf& ln, DevBuf& qkv, DevBuf& attn, DevBuf& hbuf, const AttnProbe& probe,
                             bool probs_only, int save_slot, bool has_next, cudaStream_t st, void* rollout_qkv = nullptr, int live_row = -1);
    void gemm_fold(const void* xb, const float* stats, int parts, const void* wf, const float* fb, void* out, void* out_pre,
                   int64_t M, int64_t N, int64_t K, int act, int dt, cudaStream_t st);
    // rs != nullptr: also emit the shifted 16-bit copy into xb and the next statistics set (rs flips to it)
    void gemm_resid(const void* a, int64_t lda, const void* w, const float* bias, const float* x_in, int64_t ld_in, float* x_out, int64_t ldo,
                    void* xb, RowStats* rs, int64_t M, int64_t N, int64_t K, int dt, cudaStream_t st);
    void fold_group(BlockWeights& b, int group, const std::string& prefix, int d, int dt, cudaStream_t st);
    // TAPCLIP_FUSE_LN: 0 = LayerNorm as its own kernel (default: measured fastest); 1 = folded into the GEMMs of the
    // text tower; 2 = of both towers.  All three are parity-tested; the folded form has 2 launches and 77 MB of HBM traffic less per
    // LayerNorm, but the step time follows the bytes crossing between the SMs and L2, which the folding does not change.
    int fuse_ln = getenv("TAPCLIP_FUSE_LN") ? atoi(getenv("TAPCLIP_FUSE_LN")) : 0;
    bool use_fold(bool vision) const { return cfg.dtype != DT_F32 && fuse_ln >= (vision ? 2 : 1); }
    void encode_image(const float* images, int B, float* out_feat, float* out_cls_rows, float* out_rollout, cudaStream_t st);
    // returns the token of the saved activations (0 when nothing was saved)
    // gather_epoch > 0 (needs set_text_gather): the normalised features of this rank's classes [gather_row_lo, gather_row_lo + C) are
    // also stored into every rank's symmetric buffer and the epoch is published to every rank
    int64_t text_forward(const float* ctx, const float* tok, int C, int P, int mode, bool save, float* out_attr_raw, float* out_attr,
                         float* out_text_feat, cudaStream_t st, int64_t gather_row_lo = 0, int gather_epoch = 0);
    // token: as returned by the text_forward whose activations are to be used (0 = whatever was saved last); C, P: checked when > 0
    void text_backward(const float* d_text_feat, float* out_dctx, cudaStream_t st, int64_t token = 0, int C = 0, int P = 0);
    void encode_text(const int64_t* ids, int S, float* out_feat, cudaStream_t st);
    // gather_epoch > 0: text_feat is this rank's symmetric slot of that epoch; the kernel first waits for every rank's flag
    void logits(const float* img_feat, const float* text_feat, const float* logit_scale, const int64_t* labels, int B, int C,
                float inv_batch_total, float* out_img_norm, float* out_logits, float* out_loss, float* out_dlogits, cudaStream_t st,
                int gather_epoch = 0);
    void logits_backward(const float* dlogits, const float* logits_, const float* img_norm, const float* logit_scale, int B, int C,
                         float* out_d_text, float* out_d_scale, cudaStream_t st);
};

}  // namespace tapclip

struct tapclip_engine {
    tapclip::Engine impl;
    explicit tapclip_engine(const tapclip_config& c) : impl(c) {}
};
