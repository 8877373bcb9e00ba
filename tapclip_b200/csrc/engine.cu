// Engine behind the C ABI (include/tapclip.h): owns converted frozen weights + workspaces and enqueues the
// kernel sequence of the attribution-instrumented CLIP forward and of the ctx-only backward.
//
// Schedule (SURVEY.md 3.2 de-duplicated): the reference runs 2*B*n_cls text-transformer calls per forward
// (models/model_wrapper.py:48-75); none of them depends on the sample index b, so the engine runs ONE image
// pass and at most TWO passes over [C,T,D] (attribution pass on the raw prompt, feature pass on the adjusted
// prompt).  Residual streams stay fp32; bf16 is used only for tensor-core operands.
#include "engine.h"

#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>

namespace tapclip {

namespace {
thread_local std::string g_error;
}
void set_error(const std::string& msg) { g_error = msg; }
bool pdl_enabled() {
    static const bool on = !(getenv("TAPCLIP_PDL") && atoi(getenv("TAPCLIP_PDL")) == 0);
    return on;
}
const char* get_error() { return g_error.c_str(); }

void ensure_dynamic_smem(const void* kernel, size_t bytes) {
    static std::map<std::pair<int, const void*>, size_t> configured;          // single host thread per process by contract
    int dev;
    TC_CUDA(cudaGetDevice(&dev));
    size_t& have = configured[{dev, kernel}];
    if (bytes > have) {
        TC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have = bytes;
    }
}
int device_sm_count() {
    static int sms[64] = {};
    int dev;
    TC_CUDA(cudaGetDevice(&dev));
    TC_CHECK(dev >= 0 && dev < 64, "device ordinal %d out of range", dev);
    if (sms[dev] == 0) TC_CUDA(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev));
    return sms[dev];
}

namespace {

template <typename T>
__global__ void convert_weight_kernel(const float* __restrict__ src, T* __restrict__ dst, int R, int C, int dst_ld, int transpose) {
    pdl_wait_and_trigger();
    // dst is [R, dst_ld] (transpose == 0, dst[r][c] = src[r][c]) or [C, dst_ld] (transpose, dst[c][r] = src[r][c])
    const int64_t total = (int64_t)R * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / C), c = (int)(i % C);
        const float v = src[i];
        if (transpose) dst[(int64_t)c * dst_ld + r] = from_f32<T>(v);
        else dst[(int64_t)r * dst_ld + c] = from_f32<T>(v);
    }
}

__global__ void fill_kernel(float* p, float v, int64_t n) {
    pdl_wait_and_trigger();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

// --------------------------------------------------------------------------------------------------
void DevBuf::ensure(size_t n) {
    if (n <= bytes) return;
    if (p) TC_CUDA(cudaFree(p));
    p = nullptr; bytes = 0;
    const size_t want = (n + 255) & ~(size_t)255;
    TC_CUDA(cudaMalloc(&p, want));
    bytes = want;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
}

Engine::Engine(const tapclip_config& c) : cfg(c) {
    TC_CHECK(c.dtype == DT_F32 || c.dtype == DT_BF16 || c.dtype == DT_F16, "dtype must be 0 (fp32), 1 (bf16) or 2 (mixed bf16/fp16)");
    TC_CHECK(c.act == ACT_GELU_ERF || c.act == ACT_QUICK_GELU, "act must be 0 (gelu_erf) or 1 (quick_gelu)");
    TC_CHECK(c.image_size > 0 && c.patch_size > 0 && c.image_size % c.patch_size == 0, "image_size %% patch_size != 0");
    TC_CHECK(c.vision_width == c.vision_heads * 64 && c.text_width == c.text_heads * 64,
             "head dim must be 64 (vision %d/%d, text %d/%d)", c.vision_width, c.vision_heads, c.text_width, c.text_heads);
    TC_CHECK(c.vision_width % 128 == 0 && c.text_width % 128 == 0 && c.embed_dim % 128 == 0 && c.vision_width <= 1024 &&
                 c.text_width <= 1024 && c.embed_dim <= 1024,
             "widths must be multiples of 128 and <= 1024");
    TC_CHECK(c.vision_layers >= 1 && c.text_layers >= 1 && c.context_length >= 1, "bad layer count / context length");
    int dev;
    TC_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    TC_CUDA(cudaGetDeviceProperties(&prop, dev));
    TC_CHECK(prop.major == 10, "libtapclip needs an sm_100 (B200) device, found sm_%d%d; there is no fallback path", prop.major, prop.minor);
    vdt = (c.dtype == DT_F32) ? DT_F32 : DT_BF16;
    tdt = (c.dtype == DT_F32) ? DT_F32 : (c.dtype == DT_F16 ? DT_F16 : DT_BF16);
    gdt = vdt;
    esz = dtype_size(vdt);
    grid = c.image_size / c.patch_size;
    n_tok = grid * grid + 1;
    kpatch = 3 * c.patch_size * c.patch_size;
    kpatch_pad = (int)round_up(kpatch, 64);
    vis.resize(c.vision_layers);
    txt.resize(c.text_layers);
}

Engine::~Engine() {
    for (auto& kv : weights) cudaFree(kv.second);
    for (DevBuf* b : all_bufs()) b->release();
}

std::vector<DevBuf*> Engine::all_bufs() {
    return {&v_patches, &v_patch_out, &v_x, &v_ln, &v_qkv, &v_attn, &v_h, &v_pooled, &v_roll_qkv, &v_lse, &v_roll,
            &t_x, &t_ln, &t_qkv, &t_attn, &t_h, &t_pooled, &t_feat, &t_tfeat, &t_inv_norm, &t_probe, &t_attr, &t_attr_raw,
            &t_save_x, &t_save_qkv, &t_save_h, &b_dx, &b_dxc, &b_dh, &b_dln, &b_dattn, &b_dqkv, &b_dfeat, &b_dfeatc, &b_dpool,
            &s_rows, &s_cls, &e_eot, &e_pool, &v_xb, &v_xlive, &t_xb, &t_xlive, &s_ticket, &v_rs.stats[0], &v_rs.stats[1], &v_rs.shift[0], &v_rs.shift[1],
            &t_rs.stats[0], &t_rs.stats[1], &t_rs.shift[0], &t_rs.shift[1]};
}

// two self-resetting tickets of the head kernels' last-CTA reductions (zeroed once, synchronously, when first needed)
void Engine::set_text_gather(void* const* peer_bufs, int world, int rank, int64_t n_cls_total) {
    TC_CHECK(world >= 0 && world <= 8 && rank >= 0 && (world == 0 || rank < world) && n_cls_total >= 0, "bad text-gather configuration");
    gather.world = world; gather.rank = rank; gather.n_cls = n_cls_total;
    for (int r = 0; r < 8; ++r) gather.peers[r] = (r < world) ? (uint8_t*)peer_bufs[r] : nullptr;
    for (int r = 0; r < world; ++r) TC_CHECK(gather.peers[r] != nullptr && ((uintptr_t)gather.peers[r] & 127) == 0, "peer buffer %d missing or misaligned", r);
}

int* Engine::tickets() {
    if (s_ticket.p == nullptr) {
        s_ticket.ensure(4 * sizeof(int));
        TC_CUDA(cudaMemset(s_ticket.p, 0, 4 * sizeof(int)));
        TC_CUDA(cudaDeviceSynchronize());
    }
    return (int*)s_ticket.p;
}

int64_t Engine::workspace_bytes() {
    int64_t n = 0;
    for (DevBuf* b : all_bufs()) n += (int64_t)b->bytes;
    return n;
}

// ---- weights ---------------------------------------------------------------------------------------
void* Engine::store(const std::string& key, const float* src, int R, int C, int dst_ld, bool transpose, int dt, cudaStream_t st) {
    const int rows = transpose ? C : R;
    const size_t bytes = (size_t)rows * dst_ld * dtype_size(dt);
    void* dst = nullptr;
    auto it = weights.find(key);
    if (it != weights.end()) { cudaFree(it->second); weights.erase(it); }
    TC_CUDA(cudaMalloc(&dst, bytes));
    weights[key] = dst;
    TC_CUDA(cudaMemsetAsync(dst, 0, bytes, st));
    const int64_t total = (int64_t)R * C;
    const unsigned g = (unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 16);
    if (dt == DT_BF16) launch_pdl(convert_weight_kernel<bf16>, g, 256, 0, st, src, (bf16*)dst, R, C, dst_ld, transpose ? 1 : 0);
    else if (dt == DT_F16) launch_pdl(convert_weight_kernel<f16>, g, 256, 0, st, src, (f16*)dst, R, C, dst_ld, transpose ? 1 : 0);
    else launch_pdl(convert_weight_kernel<float>, g, 256, 0, st, src, (float*)dst, R, C, dst_ld, transpose ? 1 : 0);
    TC_LAUNCH_CHECK();
    ++launches;
    return dst;
}

static bool shape_is(int ndim, const int64_t* s, std::initializer_list<int64_t> want) {
    if (ndim != (int)want.size()) return false;
    int i = 0;
    for (int64_t w : want)
        if (s[i++] != w) return false;
    return true;
}

void Engine::load_weight(const std::string& name, const float* data, int ndim, const int64_t* shape, cudaStream_t st) {
    TC_CHECK(data != nullptr, "null weight pointer for %s", name.c_str());
    const int dv = cfg.vision_width, dt = cfg.text_width, E = cfg.embed_dim;
    auto bad_shape = [&]() {
        std::ostringstream os;
        os << "weight " << name << " has unexpected shape [";
        for (int i = 0; i < ndim; ++i) os << (i ? "," : "") << shape[i];
        os << "]";
        throw Error{os.str()};
    };
    // entries only the standard CLIP text path (encode_text) uses
    if (name == "token_embedding.weight") { if (ndim != 2 || shape[1] != dt) bad_shape(); vocab = (int)shape[0]; tok_emb = (float*)store(name, data, (int)shape[0], dt, dt, false, DT_F32, st); return; }
    if (name == "positional_embedding") { if (!shape_is(ndim, shape, {cfg.context_length, dt})) bad_shape(); text_pos = (float*)store(name, data, cfg.context_length, dt, dt, false, DT_F32, st); return; }
    if (name == "ln_final.weight") { if (!shape_is(ndim, shape, {dt})) bad_shape(); ln_final_g = (float*)store(name, data, 1, dt, dt, false, DT_F32, st); return; }
    if (name == "ln_final.bias") { if (!shape_is(ndim, shape, {dt})) bad_shape(); ln_final_b = (float*)store(name, data, 1, dt, dt, false, DT_F32, st); return; }
    if (name == "logit_scale" || name == "attn_mask") return;        // not used by any path
    if (name == "visual.conv1.weight") {
        if (!shape_is(ndim, shape, {dv, 3, cfg.patch_size, cfg.patch_size})) bad_shape();
        w_patch = store(name, data, dv, kpatch, kpatch_pad, false, vdt, st);
        return;
    }
    if (name == "visual.class_embedding") { if (!shape_is(ndim, shape, {dv})) bad_shape(); cls_emb = (float*)store(name, data, 1, dv, dv, false, DT_F32, st); return; }
    if (name == "visual.positional_embedding") { if (!shape_is(ndim, shape, {n_tok, dv})) bad_shape(); pos_emb = (float*)store(name, data, n_tok, dv, dv, false, DT_F32, st); return; }
    if (name == "visual.ln_pre.weight") { if (!shape_is(ndim, shape, {dv})) bad_shape(); ln_pre_g = (float*)store(name, data, 1, dv, dv, false, DT_F32, st); return; }
    if (name == "visual.ln_pre.bias") { if (!shape_is(ndim, shape, {dv})) bad_shape(); ln_pre_b = (float*)store(name, data, 1, dv, dv, false, DT_F32, st); return; }
    if (name == "visual.ln_post.weight") { if (!shape_is(ndim, shape, {dv})) bad_shape(); ln_post_g = (float*)store(name, data, 1, dv, dv, false, DT_F32, st); return; }
    if (name == "visual.ln_post.bias") { if (!shape_is(ndim, shape, {dv})) bad_shape(); ln_post_b = (float*)store(name, data, 1, dv, dv, false, DT_F32, st); return; }
    if (name == "visual.proj") { if (!shape_is(ndim, shape, {dv, E})) bad_shape(); w_vproj = store(name, data, dv, E, dv, true, vdt, st); return; }
    if (name == "text_projection") {
        if (!shape_is(ndim, shape, {dt, E})) bad_shape();
        w_tproj = store(name, data, dt, E, dt, true, tdt, st);                  // [E, D] forward operand
        wt_tproj = store(name + "#T", data, dt, E, E, false, gdt, st);          // [D, E] dgrad operand
        wde_tproj = store(name + "#DE", data, dt, E, E, false, tdt, st);        // [D, E] in the forward type: the head kernel's layout
        return;
    }
    // transformer blocks
    const std::string vp = "visual.transformer.resblocks.", tp = "transformer.resblocks.";
    bool is_vis = name.compare(0, vp.size(), vp) == 0, is_txt = name.compare(0, tp.size(), tp) == 0;
    TC_CHECK(is_vis || is_txt, "unknown weight name '%s'", name.c_str());
    const std::string rest = name.substr(is_vis ? vp.size() : tp.size());
    const size_t dot = rest.find('.');
    TC_CHECK(dot != std::string::npos, "unknown weight name '%s'", name.c_str());
    const int layer = atoi(rest.substr(0, dot).c_str());
    const std::string leaf = rest.substr(dot + 1);
    auto& blocks = is_vis ? vis : txt;
    TC_CHECK(layer >= 0 && layer < (int)blocks.size(), "layer index out of range in '%s'", name.c_str());
    BlockWeights& b = blocks[layer];
    const int d = is_vis ? dv : dt;
    const bool need_t = is_txt;          // dgrad copies only for the text tower (the image tower gets no gradient)
    auto vec = [&](float*& slot, int n) { if (!shape_is(ndim, shape, {n})) bad_shape(); slot = (float*)store(name, data, 1, n, n, false, DT_F32, st); };
    auto mat = [&](void*& w, void*& wt, int N, int K) {
        if (!shape_is(ndim, shape, {N, K})) bad_shape();
        w = store(name, data, N, K, K, false, is_vis ? vdt : tdt, st);
        if (need_t) wt = store(name + "#T", data, N, K, N, true, gdt, st);      // [K, N], gradient type
    };
    const bool fold = cfg.dtype != DT_F32;                                       // 16-bit modes: LayerNorm folded into QKV / c_fc
    auto keep_f32 = [&](float*& slot, int N, int K) { if (fold) slot = (float*)store(name + "#F32", data, N, K, K, false, DT_F32, st); };
    const std::string prefix = (is_vis ? vp : tp) + std::to_string(layer) + ".";
    const int odt = is_vis ? vdt : tdt;
    int group = -1;                                                              // fold group this entry belongs to: 0 = ln_1 -> QKV, 1 = ln_2 -> c_fc
    if (leaf == "ln_1.weight") { vec(b.ln1_g, d); group = 0; }
    else if (leaf == "ln_1.bias") { vec(b.ln1_b, d); group = 0; }
    else if (leaf == "ln_2.weight") { vec(b.ln2_g, d); group = 1; }
    else if (leaf == "ln_2.bias") { vec(b.ln2_b, d); group = 1; }
    else if (leaf == "attn.in_proj_weight") { mat(b.w_qkv, b.wt_qkv, 3 * d, d); keep_f32(b.f32_qkv, 3 * d, d); group = 0; }
    else if (leaf == "attn.in_proj_bias") { vec(b.b_qkv, 3 * d); group = 0; }
    else if (leaf == "attn.out_proj.weight") mat(b.w_o, b.wt_o, d, d);
    else if (leaf == "attn.out_proj.bias") vec(b.b_o, d);
    else if (leaf == "mlp.c_fc.weight") { mat(b.w_fc, b.wt_fc, 4 * d, d); keep_f32(b.f32_fc, 4 * d, d); group = 1; }
    else if (leaf == "mlp.c_fc.bias") { vec(b.b_fc, 4 * d); group = 1; }
    else if (leaf == "mlp.c_proj.weight") mat(b.w_proj, b.wt_proj, d, 4 * d);
    else if (leaf == "mlp.c_proj.bias") vec(b.b_proj, d);
    else TC_CHECK(false, "unknown weight name '%s'", name.c_str());
    if (fold && group >= 0) fold_group(b, group, prefix, d, odt, st);
}

// (Re)builds the folded operands of one LayerNorm -> Linear pair once its four tensors are present (on the loading stream: the
// caller synchronises it before the first forward, engine.py load_state_dict).
void Engine::fold_group(BlockWeights& b, int group, const std::string& prefix, int d, int dt, cudaStream_t st) {
    const float* W = group == 0 ? b.f32_qkv : b.f32_fc;
    const float* bias = group == 0 ? b.b_qkv : b.b_fc;
    const float* gamma = group == 0 ? b.ln1_g : b.ln2_g;
    const float* beta = group == 0 ? b.ln1_b : b.ln2_b;
    if (!W || !bias || !gamma || !beta) return;
    const int N = (group == 0 ? 3 : 4) * d, K = d;
    const std::string key = prefix + (group == 0 ? "fold.qkv" : "fold.fc");
    auto alloc = [&](const std::string& k, size_t bytes) {
        auto it = weights.find(k);
        if (it != weights.end()) return it->second;          // same shape as before: reuse
        void* p = nullptr;
        TC_CUDA(cudaMalloc(&p, bytes));
        weights[k] = p;
        return p;
    };
    void* wf = alloc(key + ".w", (size_t)N * K * dtype_size(dt));
    float* fb = (float*)alloc(key + ".b", (size_t)N * 4);
    fold_ln_weight(W, bias, gamma, beta, wf, dt, fb, N, K, st);
    ++launches;
    if (group == 0) { b.wf_qkv = wf; b.fb_qkv = fb; }
    else { b.wf_fc = wf; b.fb_fc = fb; }
}

std::string Engine::missing_weights() const {
    std::ostringstream os;
    auto need = [&](const void* p, const std::string& n) { if (!p) os << n << " "; };
    need(w_patch, "visual.conv1.weight"); need(cls_emb, "visual.class_embedding"); need(pos_emb, "visual.positional_embedding");
    need(ln_pre_g, "visual.ln_pre.weight"); need(ln_pre_b, "visual.ln_pre.bias"); need(ln_post_g, "visual.ln_post.weight");
    need(ln_post_b, "visual.ln_post.bias"); need(w_vproj, "visual.proj"); need(w_tproj, "text_projection");
    for (int t = 0; t < 2; ++t) {
        const auto& blocks = t ? txt : vis;
        const std::string pre = t ? "transformer.resblocks." : "visual.transformer.resblocks.";
        for (size_t i = 0; i < blocks.size(); ++i) {
            const BlockWeights& b = blocks[i];
            const std::string p = pre + std::to_string(i) + ".";
            need(b.ln1_g, p + "ln_1.weight"); need(b.ln1_b, p + "ln_1.bias"); need(b.ln2_g, p + "ln_2.weight"); need(b.ln2_b, p + "ln_2.bias");
            need(b.w_qkv, p + "attn.in_proj_weight"); need(b.b_qkv, p + "attn.in_proj_bias"); need(b.w_o, p + "attn.out_proj.weight");
            need(b.b_o, p + "attn.out_proj.bias"); need(b.w_fc, p + "mlp.c_fc.weight"); need(b.b_fc, p + "mlp.c_fc.bias");
            need(b.w_proj, p + "mlp.c_proj.weight"); need(b.b_proj, p + "mlp.c_proj.bias");
        }
    }
    return os.str();
}

// ---- per-launch profiling (CUDA events on the launching stream) -----------------------------------------
void Engine::prof_begin(ProfRec& r, cudaStream_t st) {
    TC_CUDA(cudaEventCreate(&r.a));
    TC_CUDA(cudaEventCreate(&r.b));
    TC_CUDA(cudaEventRecord(r.a, st));
}
void Engine::prof_end(ProfRec& r, cudaStream_t st) {
    TC_CUDA(cudaEventRecord(r.b, st));
    prof.push_back(r);
}
const char* Engine::profile_report() {
    struct Agg { int64_t n = 0; double ms = 0, flops = 0; };
    std::map<std::string, Agg> kinds;
    std::map<std::string, Agg> shapes;
    static const char* kind_name[] = {"gemm", "attention_fwd", "attention_bwd"};
    for (ProfRec& r : prof) {
        float ms = 0.f;
        cudaEventSynchronize(r.b);
        cudaEventElapsedTime(&ms, r.a, r.b);
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
        Agg& k = kinds[kind_name[r.kind]];
        k.n++; k.ms += ms; k.flops += r.flops;
        char key[96];
        snprintf(key, sizeof(key), "%s M=%lld N=%lld K=%lld epi=%d", kind_name[r.kind], (long long)r.M, (long long)r.N, (long long)r.K, r.epi);
        Agg& s = shapes[key];
        s.n++; s.ms += ms; s.flops += r.flops;
    }
    prof.clear();
    std::ostringstream os;
    os << "{";
    bool first = true;
    for (auto& kv : kinds) {
        os << (first ? "" : ", ") << "\"" << kv.first << "\": {\"launches\": " << kv.second.n << ", \"ms\": " << kv.second.ms
           << ", \"flops\": " << kv.second.flops << "}";
        first = false;
    }
    os << (first ? "" : ", ") << "\"shapes\": {";
    first = true;
    for (auto& kv : shapes) {
        os << (first ? "" : ", ") << "\"" << kv.first << "\": {\"launches\": " << kv.second.n << ", \"ms\": " << kv.second.ms
           << ", \"flops\": " << kv.second.flops << "}";
        first = false;
    }
    os << "}}";
    prof_report = os.str();
    return prof_report.c_str();
}

// ---- primitive wrappers ------------------------------------------------------------------------------
void Engine::gemm(const void* a, const void* w, const float* bias, void* out, void* out_pre, int64_t M, int64_t N, int64_t K,
                  int epi, int act, int dt, cudaStream_t st, int aux_dt, int64_t lda, int64_t ldo) {
    GemmArgs g;
    g.a = a; g.w = w; g.bias = bias; g.out = out; g.out_pre = out_pre;
    g.M = M; g.N = N; g.K = K; g.lda = lda ? lda : K; g.ldw = K; g.ldo = ldo ? ldo : N; g.epi = epi; g.act = act; g.dt = dt; g.aux_dt = aux_dt;
    ProfRec r{nullptr, nullptr, 2.0 * (double)M * (double)N * (double)K, 0, M, N, K, epi};
    if (profiling) prof_begin(r, st);
    if (dt != DT_F32) gemm_tc(g, st);
    else gemm_simt_f32(g, st);
    if (profiling) prof_end(r, st);
    ++launches;
}

void Engine::attn_fwd(const void* qkv, void* out, int dt, int S, int N, int H, const AttnProbe& probe, cudaStream_t st) {
    ProfRec r{nullptr, nullptr, 4.0 * (double)S * H * (double)N * N * 64.0, 1, S, N, H, probe.mode};
    if (profiling) prof_begin(r, st);
    attention_fwd(qkv, out, dt, S, N, H, probe, st);
    if (profiling) prof_end(r, st);
    ++launches;
}

void Engine::attn_bwd(const void* qkv, const void* d_out, void* dqkv, int S, int N, int H, cudaStream_t st) {
    ProfRec r{nullptr, nullptr, 10.0 * (double)S * H * (double)N * N * 64.0, 2, S, N, H, 0};
    if (profiling) prof_begin(r, st);
    attention_bwd(qkv, tdt, d_out, dqkv, gdt, S, N, H, st);
    if (profiling) prof_end(r, st);
    ++launches;
}

void Engine::gemm_fold(const void* xb, const float* stats, int parts, const void* wf, const float* fb, void* out,
                       void* out_pre, int64_t M, int64_t N, int64_t K, int act, int dt, cudaStream_t st) {
    TC_CHECK(wf && fb, "folded LayerNorm weights are missing (load ln_*, in_proj_* and c_fc.* of every block)");
    GemmArgs g;
    g.a = xb; g.w = wf; g.bias = fb; g.out = out; g.out_pre = out_pre;
    g.M = M; g.N = N; g.K = K; g.lda = K; g.ldw = K; g.ldo = N; g.epi = EPI_BF16; g.act = act; g.dt = dt;
    g.stats_in = stats; g.stats_parts = parts;
    ProfRec r{nullptr, nullptr, 2.0 * (double)M * (double)N * (double)K, 0, M, N, K, 6};
    if (profiling) prof_begin(r, st);
    gemm_tc(g, st);
    if (profiling) prof_end(r, st);
    ++launches;
}

void Engine::gemm_resid(const void* a, int64_t lda, const void* w, const float* bias, const float* x_in, int64_t ld_in, float* x_out,
                        int64_t ldo, void* xb, RowStats* rs, int64_t M, int64_t N, int64_t K, int dt, cudaStream_t st) {
    GemmArgs g;
    g.a = a; g.w = w; g.bias = bias; g.out = x_out;
    g.M = M; g.N = N; g.K = K; g.lda = lda ? lda : K; g.ldw = K; g.ldo = ldo ? ldo : N; g.epi = EPI_F32_RESID; g.act = ACT_NONE; g.dt = dt;
    g.resid_in = x_in; g.ld_in = ld_in ? ld_in : N;
    if (rs != nullptr) {
        // the set describing x_in is read (mean of the old rows = the shift of the new 16-bit copy), the other set is written
        g.xb = xb;
        g.stats_prev = rs->s(rs->cur); g.shift_prev = rs->h(rs->cur); g.prev_parts = rs->parts;
        rs->cur ^= 1; rs->parts = gemm_stats_parts(N);
        g.stats_out = rs->s(rs->cur); g.shift_out = rs->h(rs->cur);
    }
    ProfRec r{nullptr, nullptr, 2.0 * (double)M * (double)N * (double)K, 0, M, N, K, 7};
    if (profiling) prof_begin(r, st);
    gemm_tc(g, st);
    if (profiling) prof_end(r, st);
    ++launches;
}

// one residual attention block on x [S*N, d] (fp32, updated in place), LayerNorm as its own kernel (fp32 mode, encode_text,
// TAPCLIP_FUSE_LN=0).
//   probe      : attention probe for this layer (or PROBE_NONE)
//   probs_only : attribution pass, last block: only the probabilities are needed
//   save_slot  : keep x copies / qkv / h_pre for the backward pass (slot = layer)
void Engine::block_forward(const BlockWeights& b, float* x, int S, int N, int d, int H, int dt, DevBuf& ln, DevBuf& qkv, DevBuf& attn,
                           DevBuf& hbuf, const AttnProbe& probe, bool probs_only, int save_slot, cudaStream_t st, void* rollout_qkv,
                           int live_row) {
    const int64_t M = (int64_t)S * N;
    float* sx0 = nullptr; float* sx1 = nullptr; void* sqkv = qkv.p; void* shpre = nullptr;
    if (save_slot >= 0) {
        sx0 = (float*)t_save_x.p + (int64_t)(2 * save_slot) * M * d;
        sx1 = (float*)t_save_x.p + (int64_t)(2 * save_slot + 1) * M * d;
        sqkv = (uint8_t*)t_save_qkv.p + (int64_t)save_slot * M * 3 * d * esz;
        shpre = (uint8_t*)t_save_h.p + (int64_t)save_slot * M * 4 * d * esz;
    }
    if (rollout_qkv) sqkv = rollout_qkv;                       // rollout extension: this layer's Q and K are re-read by rollout_step
    layernorm_fwd(x, d, b.ln1_g, b.ln1_b, ln.p, dt, sx0, M, d, st); ++launches;
    gemm(ln.p, b.w_qkv, b.b_qkv, sqkv, nullptr, M, 3 * d, d, EPI_BF16, ACT_NONE, dt, st);
    attn_fwd(sqkv, attn.p, dt, S, N, H, probe, st);
    if (probs_only) return;
    if (live_row >= 0) {
        // Dead-row elimination: after the LAST block only token `live_row` of every sequence is read (ln_post(x[:,0])
        // / pooling), and past the attention every row is independent.  The out-projection and the MLP therefore run on the
        // S live rows only, addressed in place through the GEMM's leading dimensions (no gather, no copy).
        // With save_slot >= 0 the LN2 input and the MLP pre-activation of the live rows are kept COMPACT ([S, d] / [S, 4d]) at
        // the start of the layer's save slots: text_backward treats the last block the same way.
        const int64_t ld = (int64_t)N * d;
        float* xl = x + (int64_t)live_row * d;
        gemm((const uint8_t*)attn.p + (int64_t)live_row * d * esz, b.w_o, b.b_o, xl, nullptr, S, d, d, EPI_F32_ADD, ACT_NONE, dt, st, DT_BF16, ld, ld);
        layernorm_fwd(xl, ld, b.ln2_g, b.ln2_b, ln.p, dt, sx1, S, d, st); ++launches;
        gemm(ln.p, b.w_fc, b.b_fc, hbuf.p, shpre, S, 4 * d, d, EPI_BF16, cfg.act, dt, st);
        gemm(hbuf.p, b.w_proj, b.b_proj, xl, nullptr, S, d, 4 * d, EPI_F32_ADD, ACT_NONE, dt, st, DT_BF16, 0, ld);
        return;
    }
    gemm(attn.p, b.w_o, b.b_o, x, nullptr, M, d, d, EPI_F32_ADD, ACT_NONE, dt, st);
    layernorm_fwd(x, d, b.ln2_g, b.ln2_b, ln.p, dt, sx1, M, d, st); ++launches;
    gemm(ln.p, b.w_fc, b.b_fc, hbuf.p, shpre, M, 4 * d, d, EPI_BF16, cfg.act, dt, st);
    gemm(hbuf.p, b.w_proj, b.b_proj, x, nullptr, M, d, 4 * d, EPI_F32_ADD, ACT_NONE, dt, st);
}

// The same block with no LayerNorm kernel (16-bit modes; see engine.h).  On entry xb / stats describe the rows at `x`.
void Engine::block_forward_fused(const BlockWeights& b, float*& x, RowStats& rs, float* scratch, int S, int N, int d, int H, int dt,
                                 DevBuf& xb, DevBuf& xlive, DevBuf& ln, DevBuf& qkv, DevBuf& attn, DevBuf& hbuf,
                                 const AttnProbe& probe, bool probs_only, int save_slot, bool has_next, cudaStream_t st, void* rollout_qkv,
                                 int live_row) {
    const int64_t M = (int64_t)S * N;
    float* sx1 = nullptr; float* next_sx0 = nullptr; void* sqkv = qkv.p; void* shpre = nullptr;
    if (save_slot >= 0) {
        // the residual stream itself hops through the save slots: slot 2l holds the input of ln_1 of layer l (written by the
        // previous block's c_proj, or by the splice kernel for l = 0), slot 2l+1 the input of ln_2 (written by the out-projection)
        sx1 = (float*)t_save_x.p + (int64_t)(2 * save_slot + 1) * M * d;
        if (has_next) next_sx0 = (float*)t_save_x.p + (int64_t)(2 * save_slot + 2) * M * d;
        sqkv = (uint8_t*)t_save_qkv.p + (int64_t)save_slot * M * 3 * d * esz;
        shpre = (uint8_t*)t_save_h.p + (int64_t)save_slot * M * 4 * d * esz;
    }
    if (rollout_qkv) sqkv = rollout_qkv;
    gemm_fold(xb.p, rs.s(rs.cur), rs.parts, b.wf_qkv, b.fb_qkv, sqkv, nullptr, M, 3 * d, d, ACT_NONE, dt, st);
    attn_fwd(sqkv, attn.p, dt, S, N, H, probe, st);
    if (probs_only) return;
    if (live_row >= 0) {
        // last block, dead-row form (see block_forward): the live rows leave the [S*N, d] stream here and continue as a compact
        // [S, d] matrix; the out-projection reads row `live_row` of every sequence through its leading dimensions
        const int64_t ld = (int64_t)N * d;
        float* xl = (float*)xlive.p;
        gemm_resid((const uint8_t*)attn.p + (int64_t)live_row * d * esz, ld, b.w_o, b.b_o, x + (int64_t)live_row * d, ld, xl, d, nullptr, nullptr,
                   S, d, d, dt, st);
        layernorm_fwd(xl, d, b.ln2_g, b.ln2_b, ln.p, dt, sx1, S, d, st); ++launches;
        gemm(ln.p, b.w_fc, b.b_fc, hbuf.p, shpre, S, 4 * d, d, EPI_BF16, cfg.act, dt, st);
        gemm(hbuf.p, b.w_proj, b.b_proj, xl, nullptr, S, d, 4 * d, EPI_F32_ADD, ACT_NONE, dt, st);
        x = xl;
        return;
    }
    float* x1 = sx1 ? sx1 : x;
    gemm_resid(attn.p, 0, b.w_o, b.b_o, x, d, x1, d, xb.p, &rs, M, d, d, dt, st);
    gemm_fold(xb.p, rs.s(rs.cur), rs.parts, b.wf_fc, b.fb_fc, hbuf.p, shpre, M, 4 * d, d, cfg.act, dt, st);
    float* x2 = next_sx0 ? next_sx0 : (save_slot >= 0 ? scratch : x1);
    gemm_resid(hbuf.p, 0, b.w_proj, b.b_proj, x1, d, x2, d, xb.p, &rs, M, d, 4 * d, dt, st);
    x = x2;
}

// ---- image tower (row A4) ---------------------------------------------------------------------------
void Engine::encode_image(const float* images, int B, float* out_feat, float* out_cls_rows, float* out_rollout, cudaStream_t st) {
    TC_CHECK(B >= 0, "negative batch");
    if (B == 0) return;
    const std::string miss = missing_weights();
    TC_CHECK(miss.empty(), "weights missing: %s", miss.c_str());
    const int d = cfg.vision_width, H = cfg.vision_heads, N = n_tok, L = cfg.vision_layers, E = cfg.embed_dim;
    const int64_t Mp = (int64_t)B * grid * grid, M = (int64_t)B * N;
    v_patches.ensure(Mp * kpatch_pad * esz);
    v_patch_out.ensure(Mp * d * 4);
    v_x.ensure(M * d * 4);
    v_ln.ensure(M * d * esz);
    v_qkv.ensure(M * 3 * d * esz);
    v_attn.ensure(M * d * esz);
    v_h.ensure(M * 4 * d * esz);
    v_pooled.ensure((int64_t)B * d * esz);
    if (out_rollout) {
        // rollout extension: every layer's packed qkv and softmax statistics are kept, the CLS row is propagated from the last
        // layer to the first after the tower (rollout.cu); no N x N map exists anywhere
        v_roll_qkv.ensure((size_t)L * M * 3 * d * esz);
        v_lse.ensure((size_t)L * B * H * N * sizeof(float));
        v_roll.ensure((size_t)2 * B * N * sizeof(float));
    }

    const bool fused = use_fold(true);
    if (fused) {
        v_xb.ensure(M * d * esz);
        v_rs.ensure(M, gemm_stats_parts(d));
        v_xlive.ensure((size_t)B * d * sizeof(float));
        v_rs.cur = 0; v_rs.parts = 1;
    }
    patchify(images, v_patches.p, vdt, B, cfg.image_size, cfg.patch_size, kpatch_pad, st); ++launches;
    gemm(v_patches.p, w_patch, nullptr, v_patch_out.p, nullptr, Mp, d, kpatch_pad, EPI_F32, ACT_NONE, vdt, st);
    assemble_ln_pre((const float*)v_patch_out.p, cls_emb, pos_emb, ln_pre_g, ln_pre_b, (float*)v_x.p, B, N, d, st,
                    fused ? v_xb.p : nullptr, fused ? v_rs.s(0) : nullptr, fused ? v_rs.h(0) : nullptr); ++launches;
    float* xcur = (float*)v_x.p;                      // where the residual stream lives (fused path: may move to the compact live rows)
    int64_t x_stride = (int64_t)N * d;                // distance between the CLS rows of consecutive images
    for (int l = 0; l < L; ++l) {
        AttnProbe probe;
        if (out_cls_rows) {
            probe.mode = PROBE_CLS_ROW;
            probe.out = out_cls_rows + (int64_t)l * H * N;
            probe.seq_stride = (int64_t)L * H * N;
        }
        if (l == L - 1 && dead_rows) probe.live_q_rows = 1;
        if (out_rollout) probe.lse_out = (float*)v_lse.p + (int64_t)l * B * H * N;
        void* roll_qkv = out_rollout ? (uint8_t*)v_roll_qkv.p + (int64_t)l * M * 3 * d * esz : nullptr;
        const int live_row = (l == L - 1 && dead_rows) ? 0 : -1;                       // only the CLS row feeds ln_post
        if (fused) {
            block_forward_fused(vis[l], xcur, v_rs, (float*)v_x.p, B, N, d, H, vdt, v_xb, v_xlive, v_ln, v_qkv, v_attn, v_h, probe,
                                false, -1, l + 1 < L, st, roll_qkv, live_row);
            if (live_row >= 0) x_stride = d;
        } else {
            block_forward(vis[l], xcur, B, N, d, H, vdt, v_ln, v_qkv, v_attn, v_h, probe, false, -1, st, roll_qkv, live_row);
        }
    }
    layernorm_fwd(xcur, x_stride, ln_post_g, ln_post_b, v_pooled.p, vdt, nullptr, B, d, st); ++launches;
    gemm(v_pooled.p, w_vproj, nullptr, out_feat, nullptr, B, E, d, EPI_F32, ACT_NONE, vdt, st);
    if (out_rollout) {
        // r_L = e_0;  r_{l} = 0.5 r_{l+1} + 0.5 r_{l+1}^T mean_h P_l;  out = r_0 without the CLS column.  At the last layer only
        // r_0 != 0, so the statistics of its dead query rows (live_q_rows) are never read.
        for (int l = L - 1; l >= 0; --l) {
            const float* r_in = (l == L - 1) ? nullptr : (const float*)v_roll.p + (int64_t)((l + 1) & 1) * B * N;
            float* r_out = (l == 0) ? out_rollout : (float*)v_roll.p + (int64_t)(l & 1) * B * N;
            rollout_step((const uint8_t*)v_roll_qkv.p + (int64_t)l * M * 3 * d * esz, (const float*)v_lse.p + (int64_t)l * B * H * N,
                         r_in, r_out, vdt, B, N, H, l == 0, st);
            ++launches;
        }
    }
}

// ---- text side (rows A2, A6-A10) ----------------------------------------------------------------------
int64_t Engine::text_forward(const float* ctx, const float* tok, int C, int P, int mode, bool save, float* out_attr_raw,
                             float* out_attr, float* out_text_feat, cudaStream_t st, int64_t gather_row_lo, int gather_epoch) {
    TC_CHECK(C >= 0 && P >= 1, "bad class count / prompt length");
    TC_CHECK(mode >= 0 && mode <= 2, "attribution mode must be 0 (literal), 1 (intended) or 2 (intended attribution pass only)");
    const bool attr_only = (mode == 2);
    if (attr_only) { mode = 1; TC_CHECK(!save, "the attribution-only pass keeps nothing for backward"); }
    saved.valid = false;
    if (C == 0) return 0;
    const std::string miss = missing_weights();
    TC_CHECK(miss.empty(), "weights missing: %s", miss.c_str());
    const int D = cfg.text_width, H = cfg.text_heads, Lc = cfg.context_length, T = P + Lc, L = cfg.text_layers, E = cfg.embed_dim;
    const bool fused = use_fold(false);
    const int64_t M = (int64_t)C * T;
    t_x.ensure(M * D * 4);
    t_ln.ensure(M * D * esz);
    t_qkv.ensure(M * 3 * D * esz);
    t_attn.ensure(M * D * esz);
    t_h.ensure(M * 4 * D * esz);
    t_pooled.ensure((int64_t)C * D * esz);
    t_feat.ensure((int64_t)C * E * 4);
    t_tfeat.ensure((int64_t)C * E * 4);
    t_inv_norm.ensure((int64_t)C * 4);
    const int PA = (mode == 1) ? P : 1;
    t_attr.ensure((int64_t)C * PA * 4);
    t_attr_raw.ensure((int64_t)C * PA * 4);
    if (save) {
        TC_CHECK(T <= 141, "training needs prompt_len + context_length <= 141 (attention backward), got %d", T);
        t_save_x.ensure((int64_t)2 * L * M * D * 4);
        t_save_qkv.ensure((int64_t)L * M * 3 * D * esz);
        t_save_h.ensure((int64_t)L * M * 4 * D * esz);
    }
    if (fused) {
        t_xb.ensure(M * D * esz);
        t_rs.ensure(M, gemm_stats_parts(D));
        t_xlive.ensure((size_t)C * D * sizeof(float));
    }
    float* x = (float*)t_x.p;
    const float* attr = nullptr;
    // one pass of the text transformer over the prompts already spliced into `xp` (fused path: xp may be save slot 0)
    auto run_blocks = [&](float* xp, bool attribution_pass, bool keep, float*& x_end, int64_t& pool_stride, int64_t& pool_offset) {
        float* xc = xp;
        pool_stride = T; pool_offset = T - 1;
        if (fused) { t_rs.cur = 0; t_rs.parts = 1; row_stats_cast(xc, t_xb.p, tdt, t_rs.s(0), t_rs.h(0), M, D, st); ++launches; }
        for (int l = 0; l < L; ++l) {
            AttnProbe probe;
            const bool last = (l == L - 1);
            if (attribution_pass && last) { probe.mode = PROBE_TEXT_COL; probe.out = (float*)t_probe.p; probe.P = P; }
            // feature pass, last block: only position T-1 is pooled (model_wrapper.py:73) -> out-projection and MLP on C rows
            const int live_row = (!attribution_pass && last && dead_rows) ? T - 1 : -1;
            if (fused) {
                block_forward_fused(txt[l], xc, t_rs, (float*)t_x.p, C, T, D, H, tdt, t_xb, t_xlive, t_ln, t_qkv, t_attn, t_h, probe,
                                    attribution_pass && last, keep ? l : -1, l + 1 < L, st, nullptr, live_row);
                if (live_row >= 0) { pool_stride = 1; pool_offset = 0; }
            } else {
                block_forward(txt[l], xc, C, T, D, H, tdt, t_ln, t_qkv, t_attn, t_h, probe, attribution_pass && last, keep ? l : -1, st, nullptr,
                              live_row);
            }
        }
        x_end = xc;
    };
    float* x_end = nullptr; int64_t pool_stride = T, pool_offset = T - 1;
    if (mode == 1) {
        // attribution pass (rows A7/A8): un-adjusted prompt, probabilities of the last block only
        TC_CHECK(T <= 256, "prompt_len + context_length = %d unsupported (<= 256: the attention kernels' probe)", T);
        t_probe.ensure((int64_t)C * H * P * 4);
        splice_prompts(ctx, tok, nullptr, 1, x, C, P, Lc, D, st); ++launches;
        run_blocks(x, true, false, x_end, pool_stride, pool_offset);
        attribution_reduce((const float*)t_probe.p, (float*)t_attr_raw.p, (float*)t_attr.p, C, H, P, st); ++launches;
        attr = (const float*)t_attr.p;
        if (out_attr_raw) TC_CUDA(cudaMemcpyAsync(out_attr_raw, t_attr_raw.p, (size_t)C * P * 4, cudaMemcpyDeviceToDevice, st));
        if (out_attr) TC_CUDA(cudaMemcpyAsync(out_attr, t_attr.p, (size_t)C * P * 4, cudaMemcpyDeviceToDevice, st));
    } else if (out_attr) {
        // literal mode: the reference's attribution is identically 1.0; ctx*1 == ctx
        launch_pdl(fill_kernel, (unsigned)ceil_div(C, 256), 256, 0, st, out_attr, 1.0f, C);
        TC_LAUNCH_CHECK(); ++launches;
    }
    if (attr_only) return 0;        // 'gate' / 'residual' adjustors: the host applies its small network to out_attr (prompt_adjustor.py:38-44)
    // feature pass (rows A9/A10).  Fused path with save: the spliced prompts are written straight into save slot 0 (the input of
    // ln_1 of layer 0) and the residual stream hops through the save slots from there.
    float* xp = (fused && save) ? (float*)t_save_x.p : x;
    splice_prompts(ctx, tok, attr, PA, xp, C, P, Lc, D, st); ++launches;
    run_blocks(xp, false, save, x_end, pool_stride, pool_offset);
    PeerScatter ps = {};
    if (gather_epoch > 0) {
        TC_CHECK(gather.world > 0 && fuse_head, "fused text-feature gather needs tapclip_text_gather_config and the fused head kernels");
        TC_CHECK(gather_row_lo >= 0 && gather_row_lo + C <= gather.n_cls, "gathered rows [%lld, %lld) outside the %lld configured classes",
                 (long long)gather_row_lo, (long long)(gather_row_lo + C), (long long)gather.n_cls);
        const size_t slot = gather_slot_bytes();
        for (int r = 0; r < gather.world; ++r) {
            ps.dst[r] = (float*)(gather.peers[r] + (size_t)(gather_epoch & 1) * slot);
            ps.flag[r] = (int*)(gather.peers[r] + 2 * slot) + gather.rank;
        }
        ps.world = gather.world; ps.epoch = gather_epoch; ps.row_lo = gather_row_lo; ps.ticket = tickets() + 2;
    }
    if (fuse_head) {
        // K4 (head.cu): pool position T-1, @ text_projection, L2-normalise -- one launch, fp32 rows against the 16-bit weight
        // (+ K5: the rows also go straight into every rank's symmetric buffer)
        text_head(x_end, pool_stride, pool_offset, wde_tproj, tdt, (float*)t_tfeat.p, (float*)t_inv_norm.p, out_text_feat, C, D, E, st,
                  gather_epoch > 0 ? &ps : nullptr); ++launches;
    } else {
        gather_rows(x_end, t_pooled.p, tdt, C, pool_stride, pool_offset, D, st); ++launches;
        gemm(t_pooled.p, w_tproj, nullptr, t_feat.p, nullptr, C, E, D, EPI_F32, ACT_NONE, tdt, st);
        l2norm_fwd((const float*)t_feat.p, (float*)t_tfeat.p, (float*)t_inv_norm.p, C, E, st); ++launches;
        if (out_text_feat) TC_CUDA(cudaMemcpyAsync(out_text_feat, t_tfeat.p, (size_t)C * E * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (save) {
        saved.valid = true; saved.C = C; saved.P = P; saved.T = T; saved.PA = PA; saved.has_attr = (mode == 1); saved.dead_last = dead_rows;
        saved.token = ++forward_seq;
        return saved.token;
    }
    return 0;
}

// ---- standard CLIP text path (CLIPWrapper.encode_text, clip_wrapper.py:49-51; never called by FullModel) ---------------
// open_clip CLIP.encode_text: token + positional embedding, causal transformer, ln_final, EOT (argmax id) pooling, projection.
void Engine::encode_text(const int64_t* ids, int S, float* out_feat, cudaStream_t st) {
    if (S == 0) return;
    const std::string miss = missing_weights();
    TC_CHECK(miss.empty(), "weights missing: %s", miss.c_str());
    TC_CHECK(tok_emb && text_pos && ln_final_g && ln_final_b, "encode_text needs token_embedding.weight, positional_embedding and ln_final.*");
    const int D = cfg.text_width, H = cfg.text_heads, T = cfg.context_length, L = cfg.text_layers, E = cfg.embed_dim;
    const int64_t M = (int64_t)S * T;
    t_x.ensure(M * D * 4);
    t_ln.ensure(M * D * esz);
    t_qkv.ensure(M * 3 * D * esz);
    t_attn.ensure(M * D * esz);
    t_h.ensure(M * 4 * D * esz);
    t_pooled.ensure((int64_t)S * D * esz);
    e_eot.ensure((size_t)S * 4);
    e_pool.ensure((size_t)S * D * 4);
    saved.valid = false;                                  // shares the text workspaces with text_forward
    float* x = (float*)t_x.p;
    embed_tokens(ids, tok_emb, text_pos, x, (int32_t*)e_eot.p, S, T, D, st); ++launches;
    for (int l = 0; l < L; ++l) {
        AttnProbe causal;
        causal.causal = true;
        block_forward(txt[l], x, S, T, D, H, tdt, t_ln, t_qkv, t_attn, t_h, causal, false, -1, st);
    }
    gather_rows_indexed(x, (const int32_t*)e_eot.p, (float*)e_pool.p, S, T, D, st); ++launches;
    layernorm_fwd((const float*)e_pool.p, D, ln_final_g, ln_final_b, t_pooled.p, tdt, nullptr, S, D, st); ++launches;
    gemm(t_pooled.p, w_tproj, nullptr, out_feat, nullptr, S, E, D, EPI_F32, ACT_NONE, tdt, st);
}

// ---- backward to ctx (row A13) ---------------------------------------------------------------------------
void Engine::text_backward(const float* d_text_feat, float* out_dctx, cudaStream_t st, int64_t token, int C_expect, int P_expect) {
    TC_CHECK(saved.valid, "tapclip_text_backward needs a preceding tapclip_text_forward(save_for_backward=1) whose activations are still "
                          "held (a later text_forward / encode_text on this handle replaces them)");
    TC_CHECK(token == 0 || token == saved.token, "stale backward: the activations of forward #%lld were overwritten by forward #%lld on this handle "
             "(one saved forward per handle: run backward before the next forward, or use one CLIPWrapper per model)", (long long)token, (long long)saved.token);
    TC_CHECK((C_expect <= 0 || C_expect == saved.C) && (P_expect <= 0 || P_expect == saved.P), "backward for C=%d, P=%d but the saved forward had C=%d, P=%d",
             C_expect, P_expect, saved.C, saved.P);
    const int C = saved.C, P = saved.P, T = saved.T;
    const int D = cfg.text_width, H = cfg.text_heads, L = cfg.text_layers, E = cfg.embed_dim;
    const int64_t M = (int64_t)C * T;
    b_dx.ensure(M * D * 4);
    b_dxc.ensure(M * D * esz);
    b_dh.ensure(M * 4 * D * esz);
    b_dln.ensure(M * D * 4);
    b_dattn.ensure(M * D * esz);
    b_dqkv.ensure(M * 3 * D * esz);
    b_dfeat.ensure((int64_t)C * E * 4);
    b_dfeatc.ensure((int64_t)C * E * esz);
    b_dpool.ensure((int64_t)C * D * 4);
    // L2-norm and projection backward (model_wrapper.py:74-75), scattered into the last position (:73)
    TC_CUDA(cudaMemsetAsync(b_dx.p, 0, (size_t)M * D * 4, st));
    TC_CUDA(cudaMemsetAsync(b_dxc.p, 0, (size_t)M * D * esz, st));
    if (fuse_head) {
        text_head_bwd(d_text_feat, (const float*)t_tfeat.p, (const float*)t_inv_norm.p, w_tproj, tdt, (float*)b_dx.p, b_dxc.p, gdt, T, T - 1,
                      C, D, E, st); ++launches;
    } else {
        l2norm_bwd(d_text_feat, (const float*)t_tfeat.p, (const float*)t_inv_norm.p, (float*)b_dfeat.p, b_dfeatc.p, gdt, C, E, st); ++launches;
        gemm(b_dfeatc.p, wt_tproj, nullptr, b_dpool.p, nullptr, C, D, E, EPI_F32, ACT_NONE, gdt, st);
        scatter_rows((const float*)b_dpool.p, (float*)b_dx.p, b_dxc.p, gdt, C, T, T - 1, D, st); ++launches;
    }
    for (int l = L - 1; l >= 0; --l) {
        const BlockWeights& b = txt[l];
        const float* x0 = (const float*)t_save_x.p + (int64_t)(2 * l) * M * D;
        const float* x1 = (const float*)t_save_x.p + (int64_t)(2 * l + 1) * M * D;
        const void* qkv = (const uint8_t*)t_save_qkv.p + (int64_t)l * M * 3 * D * esz;
        const void* hpre = (const uint8_t*)t_save_h.p + (int64_t)l * M * 4 * D * esz;
        if (l == L - 1 && saved.dead_last) {
            // Last block, dead-row form (see block_forward): the incoming gradient is non-zero at position T-1 only, so the MLP
            // and out-projection backward run on C rows addressed in place (leading dimension T*D); the saved LN2 input and
            // MLP pre-activation of those rows are compact.
            const int64_t ld = (int64_t)T * D;
            const int64_t off = (int64_t)(T - 1) * D;
            float* dxl = (float*)b_dx.p + off;
            uint8_t* dxcl = (uint8_t*)b_dxc.p + off * esz;
            if (gdt != DT_F32) {
                gemm(dxcl, b.wt_proj, nullptr, b_dh.p, const_cast<void*>(hpre), C, 4 * D, D, EPI_BF16_ACTGRAD, cfg.act, gdt, st, tdt, ld, 0);
            } else {
                gemm(dxcl, b.wt_proj, nullptr, b_dh.p, nullptr, C, 4 * D, D, EPI_BF16, ACT_NONE, gdt, st, DT_BF16, ld, 0);
                act_bwd_inplace(b_dh.p, gdt, hpre, tdt, cfg.act, (int64_t)C * 4 * D, st); ++launches;
            }
            gemm(b_dh.p, b.wt_fc, nullptr, b_dln.p, nullptr, C, D, 4 * D, EPI_F32, ACT_NONE, gdt, st);
            layernorm_bwd((const float*)b_dln.p, x1, b.ln2_g, dxl, dxcl, gdt, C, D, st, ld); ++launches;
            TC_CUDA(cudaMemsetAsync(b_dattn.p, 0, (size_t)M * D * esz, st));                  // rows other than T-1 carry no gradient
            gemm(dxcl, b.wt_o, nullptr, (uint8_t*)b_dattn.p + off * esz, nullptr, C, D, D, EPI_BF16, ACT_NONE, gdt, st, DT_BF16, ld, ld);
            attn_bwd(qkv, b_dattn.p, b_dqkv.p, C, T, H, st);
            gemm(b_dqkv.p, b.wt_qkv, nullptr, b_dln.p, nullptr, M, D, 3 * D, EPI_F32, ACT_NONE, gdt, st);
            layernorm_bwd((const float*)b_dln.p, x0, b.ln1_g, (float*)b_dx.p, b_dxc.p, gdt, M, D, st); ++launches;
            continue;
        }
        // MLP branch
        if (gdt != DT_F32) {
            // dh = (dx . W_proj) * act'(h_pre): the activation derivative is applied in the dgrad GEMM's store stage
            gemm(b_dxc.p, b.wt_proj, nullptr, b_dh.p, const_cast<void*>(hpre), M, 4 * D, D, EPI_BF16_ACTGRAD, cfg.act, gdt, st, tdt);
        } else {
            gemm(b_dxc.p, b.wt_proj, nullptr, b_dh.p, nullptr, M, 4 * D, D, EPI_BF16, ACT_NONE, gdt, st);
            act_bwd_inplace(b_dh.p, gdt, hpre, tdt, cfg.act, M * 4 * D, st); ++launches;
        }
        gemm(b_dh.p, b.wt_fc, nullptr, b_dln.p, nullptr, M, D, 4 * D, EPI_F32, ACT_NONE, gdt, st);
        layernorm_bwd((const float*)b_dln.p, x1, b.ln2_g, (float*)b_dx.p, b_dxc.p, gdt, M, D, st); ++launches;
        // attention branch
        gemm(b_dxc.p, b.wt_o, nullptr, b_dattn.p, nullptr, M, D, D, EPI_BF16, ACT_NONE, gdt, st);
        attn_bwd(qkv, b_dattn.p, b_dqkv.p, C, T, H, st);
        gemm(b_dqkv.p, b.wt_qkv, nullptr, b_dln.p, nullptr, M, D, 3 * D, EPI_F32, ACT_NONE, gdt, st);
        layernorm_bwd((const float*)b_dln.p, x0, b.ln1_g, (float*)b_dx.p, b_dxc.p, gdt, M, D, st); ++launches;
    }
    splice_bwd((const float*)b_dx.p, saved.has_attr ? (const float*)t_attr.p : nullptr, saved.PA, out_dctx, C, P, T, D, st); ++launches;
}

// ---- logits / loss (rows A5, A11, A12) ------------------------------------------------------------------------
void Engine::logits(const float* img_feat, const float* text_feat, const float* logit_scale, const int64_t* labels, int B, int C,
                    float inv_batch_total, float* out_img_norm, float* out_logits, float* out_loss, float* out_dlogits,
                    cudaStream_t st, int gather_epoch) {
    if (B == 0 || C == 0) return;
    if (gather_epoch > 0) {
        TC_CHECK(gather.world > 0 && fuse_head && (size_t)(cfg.embed_dim + C + 32) * sizeof(float) <= 48 * 1024 && C == gather.n_cls,
                 "fused text-feature gather: not configured for %d classes", C);
        TC_CHECK(text_feat == (const float*)(gather.peers[gather.rank] + (size_t)(gather_epoch & 1) * gather_slot_bytes()),
                 "text_feat must be this rank's symmetric slot of the epoch");
    }
    const int E = cfg.embed_dim;
    if (labels) TC_CHECK(out_loss != nullptr, "out_loss is required when labels are given");
    if (fuse_head && (size_t)(E + C + 32) * sizeof(float) <= 48 * 1024) {
        // K4 (head.cu): image L2-norm, logits, cross-entropy, its gradient and the batch mean in ONE launch
        s_rows.ensure((size_t)B * 4);
        logits_ce(img_feat, text_feat, logit_scale, labels, out_img_norm, out_logits, out_loss, out_dlogits, (float*)s_rows.p, tickets(),
                  B, C, E, inv_batch_total, st,
                  gather_epoch > 0 ? (const int*)(gather.peers[gather.rank] + 2 * gather_slot_bytes()) : nullptr,
                  gather_epoch > 0 ? gather.world : 0, gather_epoch);
        ++launches;
        return;
    }
    l2norm_fwd(img_feat, out_img_norm, nullptr, B, E, st); ++launches;
    cosine_logits(out_img_norm, text_feat, logit_scale, out_logits, B, C, E, st); ++launches;
    if (labels) {
        TC_CHECK(out_loss != nullptr, "out_loss is required when labels are given");
        s_rows.ensure((size_t)B * 4);
        cross_entropy(out_logits, labels, out_loss, out_dlogits, (float*)s_rows.p, B, C, inv_batch_total, st); launches += 2;
    }
}

void Engine::logits_backward(const float* dlogits, const float* logits_, const float* img_norm, const float* logit_scale, int B,
                             int C, float* out_d_text, float* out_d_scale, cudaStream_t st) {
    if (C == 0) return;
    s_cls.ensure((size_t)C * 4);
    if (fuse_head && (size_t)(B + 32) * sizeof(float) <= 48 * 1024) {
        logits_bwd_fused(dlogits, logits_, img_norm, logit_scale, out_d_text, out_d_scale, (float*)s_cls.p, tickets() + 1, B, C, cfg.embed_dim, st);
        ++launches;
        return;
    }
    logits_bwd(dlogits, logits_, img_norm, logit_scale, out_d_text, out_d_scale, (float*)s_cls.p, B, C, cfg.embed_dim, st);
    launches += 2;
}

}  // namespace tapclip
