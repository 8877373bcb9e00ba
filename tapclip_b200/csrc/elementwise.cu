// HBM-bound glue kernels of the TAP-CLIP hot path: patch gather, token assembly, ctx splice (K4), the
// attribution reduction (K3), pooled-row gather/scatter, cosine logits + cross-entropy and their backward,
// fused AdamW over the ctx bank, argmax/accuracy.  Reference lines are cited per kernel.
#include "kernels.h"

namespace tapclip {
namespace {

template <typename T> __device__ __forceinline__ void store_val(T* p, float v) { *p = from_f32<T>(v); }

// open_clip VisionTransformer.conv1 (kernel = stride = patch, no bias) is a GEMM over non-overlapping patches:
// this gathers [B,3,R,R] into the GEMM's A operand [B*g*g, kpad], k = c*p*p + py*p + px.
template <typename T>
__global__ void patchify_kernel(const float* __restrict__ img, T* __restrict__ out, int B, int R, int p, int g, int kdim,
                                int kpad) {
    pdl_wait_and_trigger();
    const int64_t total = (int64_t)B * g * g * kpad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % kpad);
        const int64_t row = i / kpad;
        float v = 0.f;
        if (k < kdim) {
            const int px = k % p, py = (k / p) % p, c = k / (p * p);
            const int gx = (int)(row % g), gy = (int)((row / g) % g);
            const int64_t b = row / ((int64_t)g * g);
            v = __ldg(img + ((b * 3 + c) * R + (gy * p + py)) * (int64_t)R + gx * p + px);
        }
        store_val<T>(out + i, v);
    }
}

// Fast path for patch sizes that are multiples of 4 (16, 32): one thread moves 4 consecutive pixels of one patch row
// (16-byte load, 8-byte store); consecutive threads cover consecutive pixels, then patch rows, then channels, so both
// sides are coalesced.  116 MB of traffic at B=128: HBM-bound.
template <typename T>
__global__ void patchify4_kernel(const float* __restrict__ img, T* __restrict__ out, int B, int R, int p, int g, int kdim, int kpad) {
    pdl_wait_and_trigger();
    const int k4 = kpad >> 2;
    const int64_t total = (int64_t)B * g * g * k4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % k4) * 4;
        const int64_t row = i / k4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kdim) {
            const int px = k % p, py = (k / p) % p, c = k / (p * p);
            const int gx = (int)(row % g), gy = (int)((row / g) % g);
            const int64_t b = row / ((int64_t)g * g);
            v = __ldg(reinterpret_cast<const float4*>(img + ((b * 3 + c) * R + (gy * p + py)) * (int64_t)R + gx * p + px));
        }
        T* o = out + row * kpad + k;
        if constexpr (sizeof(T) == 4) {
            *reinterpret_cast<float4*>(o) = v;
        } else {
            uint2 u;
            u.x = pack2<T>(v.x, v.y);
            u.y = pack2<T>(v.z, v.w);
            *reinterpret_cast<uint2*>(o) = u;
        }
    }
}

// open_clip VisionTransformer.forward: cat([class_embedding, patches]) + positional_embedding
__global__ void assemble_tokens_kernel(const float* __restrict__ patch_out, const float* __restrict__ cls,
                                       const float* __restrict__ pos, float* __restrict__ x, int B, int n_tokens, int d) {
    pdl_wait_and_trigger();
    const int d4 = d >> 2;
    const int64_t total = (int64_t)B * n_tokens * d4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % d4) * 4;
        const int64_t tokrow = i / d4;
        const int t = (int)(tokrow % n_tokens);
        const int64_t b = tokrow / n_tokens;
        float4 v = (t == 0) ? __ldg(reinterpret_cast<const float4*>(cls + c))
                            : __ldg(reinterpret_cast<const float4*>(patch_out + (b * (n_tokens - 1) + (t - 1)) * d + c));
        const float4 pe = __ldg(reinterpret_cast<const float4*>(pos + (int64_t)t * d + c));
        v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
        *reinterpret_cast<float4*>(x + tokrow * d + c) = v;
    }
}

// models/prompt_learner.py:45-66 (cat ctx | token embeddings) fused with models/prompt_adjustor.py:35-36
// (ctx * attribution) and the expand/cat of models/model_wrapper.py:49-51,68-69 — one row per class, not per sample.
__global__ void splice_kernel(const float* __restrict__ ctx, const float* __restrict__ tok, const float* __restrict__ attr,
                              int attr_p, float* __restrict__ x, int C, int P, int L, int D) {
    pdl_wait_and_trigger();
    const int T = P + L, d4 = D >> 2;
    const int64_t total = (int64_t)C * T * d4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % d4) * 4;
        const int64_t r = i / d4;
        const int t = (int)(r % T);
        const int64_t c = r / T;
        float4 v;
        if (t < P) {
            v = __ldg(reinterpret_cast<const float4*>(ctx + (c * P + t) * D + c4));
            if (attr) {
                const float a = __ldg(attr + c * attr_p + (attr_p == 1 ? 0 : t));
                v.x *= a; v.y *= a; v.z *= a; v.w *= a;
            }
        } else {
            v = __ldg(reinterpret_cast<const float4*>(tok + (c * L + (t - P)) * D + c4));
        }
        *reinterpret_cast<float4*>(x + r * D + c4) = v;
    }
}

// backward of the splice: only the ctx rows are learnable; attribution is detached (clip_wrapper.py:36)
__global__ void splice_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ attr, int attr_p,
                                  float* __restrict__ dctx, int C, int P, int T, int D) {
    pdl_wait_and_trigger();
    const int d4 = D >> 2;
    const int64_t total = (int64_t)C * P * d4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % d4) * 4;
        const int64_t r = i / d4;
        const int t = (int)(r % P);
        const int64_t c = r / P;
        float4 v = *reinterpret_cast<const float4*>(dx + (c * T + t) * D + c4);
        if (attr) {
            const float a = __ldg(attr + c * attr_p + (attr_p == 1 ? 0 : t));
            v.x *= a; v.y *= a; v.z *= a; v.w *= a;
        }
        *reinterpret_cast<float4*>(dctx + r * D + c4) = v;
    }
}

// K3: head-mean (clip_wrapper.py:36) of the probed column, then attribution_monitor.py:29-32 softmax over P.
// One warp per class; warp-shuffle reductions; P <= 192.
constexpr int ATTR_PER_LANE = 6;              // P <= 192 (forward: P + 77 <= 256 tokens)
__global__ void attribution_kernel(const float* __restrict__ probe, float* __restrict__ raw, float* __restrict__ attr,
                                   int C, int H, int P) {
    pdl_wait_and_trigger();
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    const int lane = threadIdx.x & 31;
    float r[ATTR_PER_LANE];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < ATTR_PER_LANE; ++j) {
        const int p = lane + 32 * j;
        float s = 0.f;
        if (p < P)
            for (int h = 0; h < H; ++h) s += probe[((int64_t)c * H + h) * P + p];
        r[j] = s / (float)H;
        if (p < P) m = fmaxf(m, r[j]);
    }
    m = warp_max(m);
    float e[ATTR_PER_LANE], sum = 0.f;
#pragma unroll
    for (int j = 0; j < ATTR_PER_LANE; ++j) {
        e[j] = (lane + 32 * j < P) ? expf(r[j] - m) : 0.f;
        sum += e[j];
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int j = 0; j < ATTR_PER_LANE; ++j) {
        const int p = lane + 32 * j;
        if (p < P) { raw[(int64_t)c * P + p] = r[j]; attr[(int64_t)c * P + p] = e[j] / sum; }
    }
}

template <typename T>
__global__ void gather_rows_kernel(const float* __restrict__ x, T* __restrict__ out, int64_t rows, int64_t row_stride,
                                   int64_t row_offset, int d) {
    pdl_wait_and_trigger();
    const int64_t total = rows * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d;
        const int c = (int)(i % d);
        store_val<T>(out + i, x[(r * row_stride + row_offset) * d + c]);
    }
}

// open_clip CLIP.encode_text prologue: x = token_embedding(text) + positional_embedding; EOT position = argmax of the ids
__global__ void embed_tokens_kernel(const int64_t* __restrict__ ids, const float* __restrict__ emb, const float* __restrict__ pos,
                                    float* __restrict__ x, int32_t* __restrict__ eot, int S, int L, int D) {
    pdl_wait_and_trigger();
    const int s = blockIdx.x;
    const int d4 = D >> 2;
    for (int i = threadIdx.x; i < L * d4; i += blockDim.x) {
        const int t = i / d4, c = (i % d4) * 4;
        const int64_t id = ids[(int64_t)s * L + t];
        float4 v = __ldg(reinterpret_cast<const float4*>(emb + id * D + c));
        const float4 pe = __ldg(reinterpret_cast<const float4*>(pos + (int64_t)t * D + c));
        v.x += pe.x; v.y += pe.y; v.z += pe.z; v.w += pe.w;
        *reinterpret_cast<float4*>(x + ((int64_t)s * L + t) * D + c) = v;
    }
    if (threadIdx.x == 0) {
        int best = 0; int64_t bv = ids[(int64_t)s * L];
        for (int t = 1; t < L; ++t) { const int64_t v = ids[(int64_t)s * L + t]; if (v > bv) { bv = v; best = t; } }   // first max, as torch.argmax
        eot[s] = best;
    }
}

__global__ void gather_rows_indexed_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx, float* __restrict__ out,
                                           int64_t rows, int64_t row_stride, int d) {
    pdl_wait_and_trigger();
    const int64_t total = rows * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d;
        out[i] = x[(r * row_stride + idx[r]) * d + (i % d)];
    }
}

template <typename T>
__global__ void scatter_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, T* __restrict__ dst_cast,
                                    int64_t rows, int64_t row_stride, int64_t row_offset, int d) {
    pdl_wait_and_trigger();
    const int64_t total = rows * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d;
        const int c = (int)(i % d);
        const int64_t o = (r * row_stride + row_offset) * d + c;
        dst[o] = src[i];
        if (dst_cast) store_val<T>(dst_cast + o, src[i]);
    }
}

template <typename T>
__global__ void cast_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t n) {
    pdl_wait_and_trigger();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        store_val<T>(dst + i, src[i]);
}

template <typename T, typename TH, int ACT>
__global__ void act_bwd_kernel(T* __restrict__ dh, const TH* __restrict__ h_pre, int64_t n) {
    pdl_wait_and_trigger();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dh[i] = from_f32<T>(to_f32<T>(dh[i]) * act_bwd<ACT>(to_f32<TH>(h_pre[i])));
}

// models/model_wrapper.py:79,83 for all (b, c) at once: logits = exp(logit_scale) * I_hat . T_hat^T.  One warp per pair.
__global__ void cosine_logits_kernel(const float* __restrict__ img, const float* __restrict__ txt,
                                     const float* __restrict__ logit_scale, float* __restrict__ logits, int B, int C, int E) {
    pdl_wait_and_trigger();
    const int64_t pair = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= (int64_t)B * C) return;
    const int lane = threadIdx.x & 31;
    const int64_t b = pair / C, c = pair % C;
    float s = 0.f;
    for (int e = lane * 4; e < E; e += 128) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(img + b * E + e));
        const float4 t = __ldg(reinterpret_cast<const float4*>(txt + c * E + e));
        s += (a.x * t.x + a.y * t.y) + (a.z * t.z + a.w * t.w);
    }
    s = warp_sum(s);
    if (lane == 0) logits[pair] = expf(__ldg(logit_scale)) * s;
}

// models/model_wrapper.py:91 F.cross_entropy (mean reduction); one warp per sample; also dloss/dlogits.
__global__ void ce_rows_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, float* __restrict__ row_loss,
                               float* __restrict__ dlogits, int B, int C, float inv_batch_total) {
    pdl_wait_and_trigger();
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int lane = threadIdx.x & 31;
    const float* row = logits + (int64_t)b * C;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(row[c] - m);
    s = warp_sum(s);
    const float lse = m + logf(s);
    // a label outside [0, C) (F.cross_entropy's ignore_index included) poisons the loss with NaN instead of reading out of bounds
    const int64_t label = labels[b];
    const bool ok = label >= 0 && label < C;
    const float nan = __int_as_float(0x7fc00000);
    if (lane == 0) row_loss[b] = ok ? (lse - row[label]) * inv_batch_total : nan;
    if (dlogits)
        for (int c = lane; c < C; c += 32)
            dlogits[(int64_t)b * C + c] = ok ? (expf(row[c] - lse) - (c == label ? 1.f : 0.f)) * inv_batch_total : nan;
}
// deterministic single-block sum
__global__ void sum_kernel(const float* __restrict__ v, float* __restrict__ out, int n) {
    pdl_wait_and_trigger();
    __shared__ float sh[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
        s = warp_sum(s);
        if (threadIdx.x == 0) *out = s;
    }
}

// d_txt[c,e] = exp(s) * sum_b dlogits[b,c] * img[b,e]  (one block per class) ; partial d_scale per class
__global__ void logits_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ logits,
                                  const float* __restrict__ img, const float* __restrict__ logit_scale,
                                  float* __restrict__ d_txt, float* __restrict__ d_scale_part, int B, int C, int E) {
    pdl_wait_and_trigger();
    const int c = blockIdx.x;
    const float es = expf(__ldg(logit_scale));
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s = fmaf(__ldg(dlogits + (int64_t)b * C + c), __ldg(img + (int64_t)b * E + e), s);
        d_txt[(int64_t)c * E + e] = es * s;
    }
    if (threadIdx.x < 32) {
        float s = 0.f;
        for (int b = threadIdx.x; b < B; b += 32) s += dlogits[(int64_t)b * C + c] * logits[(int64_t)b * C + c];
        s = warp_sum(s);
        if (threadIdx.x == 0) d_scale_part[c] = s;
    }
}

// torch.optim.AdamW (train.py:65-67): decoupled weight decay, bias-corrected moments
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             int64_t n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt) {
    pdl_wait_and_trigger();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        float pi = p[i] * (1.f - lr * wd);
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= (lr / bc1) * (mi / denom);
        p[i] = pi;
    }
}

// utils/eval_metrics.py:19-29,58-63: argmax over classes (first max wins, as torch.argmax; NaN counts as the maximum, as in
// torch) + overall and per-class correct / total counters kept on the device (the reference pulls every sample to the host)
__global__ void argmax_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int64_t* __restrict__ pred,
                              int* __restrict__ correct, int* __restrict__ class_correct, int* __restrict__ class_total, int B, int C) {
    pdl_wait_and_trigger();
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int lane = threadIdx.x & 31;
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
        float v = logits[(int64_t)b * C + c];
        if (v != v) v = INFINITY;                     // torch.argmax treats NaN as the largest value (first one wins)
        if (v > best || bi == 0x7fffffff) { best = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) {
        if (pred) pred[b] = bi;
        if (labels) {
            const int64_t t = labels[b];
            const bool hit = (t == bi);
            if (correct && hit) atomicAdd(correct, 1);
            if (t >= 0 && t < C) {                    // eval_metrics.py:26-29: per_class_total[t] += 1; per_class_correct[t] += (t == p)
                if (class_total) atomicAdd(class_total + t, 1);
                if (class_correct && hit) atomicAdd(class_correct + t, 1);
            }
        }
    }
}

inline unsigned grid_for(int64_t n, int threads = 256) { return (unsigned)std::min<int64_t>(ceil_div(n, threads), 148 * 16); }

}  // namespace

void patchify(const float* images, void* out, int out_dt, int B, int R, int p, int kpad, cudaStream_t stream) {
    TC_CHECK(R % p == 0, "image size %d not divisible by patch %d", R, p);
    const int g = R / p, kdim = 3 * p * p;
    const int64_t total = (int64_t)B * g * g * kpad;
    if (total == 0) return;
    if (p % 4 == 0 && R % 4 == 0 && kpad % 4 == 0 && out_dt != DT_F32) {
        const int64_t t4 = total / 4;
        if (out_dt == DT_BF16) launch_pdl(patchify4_kernel<bf16>, grid_for(t4), 256, 0, stream, images, (bf16*)out, B, R, p, g, kdim, kpad);
        else launch_pdl(patchify4_kernel<f16>, grid_for(t4), 256, 0, stream, images, (f16*)out, B, R, p, g, kdim, kpad);
        TC_LAUNCH_CHECK();
        return;
    }
    if (out_dt == DT_BF16) launch_pdl(patchify_kernel<bf16>, grid_for(total), 256, 0, stream, images, (bf16*)out, B, R, p, g, kdim, kpad);
    else if (out_dt == DT_F16) launch_pdl(patchify_kernel<f16>, grid_for(total), 256, 0, stream, images, (f16*)out, B, R, p, g, kdim, kpad);
    else launch_pdl(patchify_kernel<float>, grid_for(total), 256, 0, stream, images, (float*)out, B, R, p, g, kdim, kpad);
    TC_LAUNCH_CHECK();
}

void assemble_tokens(const float* patch_out, const float* cls, const float* pos, float* x, int B, int n_tokens, int d,
                     cudaStream_t stream) {
    const int64_t total = (int64_t)B * n_tokens * (d / 4);
    if (total == 0) return;
    launch_pdl(assemble_tokens_kernel, grid_for(total), 256, 0, stream, patch_out, cls, pos, x, B, n_tokens, d);
    TC_LAUNCH_CHECK();
}

void splice_prompts(const float* ctx, const float* tok, const float* attr, int attr_p, float* x, int C, int P, int L, int D,
                    cudaStream_t stream) {
    const int64_t total = (int64_t)C * (P + L) * (D / 4);
    if (total == 0) return;
    launch_pdl(splice_kernel, grid_for(total), 256, 0, stream, ctx, tok, attr, attr_p, x, C, P, L, D);
    TC_LAUNCH_CHECK();
}

void splice_bwd(const float* dx, const float* attr, int attr_p, float* dctx, int C, int P, int T, int D, cudaStream_t stream) {
    const int64_t total = (int64_t)C * P * (D / 4);
    if (total == 0) return;
    launch_pdl(splice_bwd_kernel, grid_for(total), 256, 0, stream, dx, attr, attr_p, dctx, C, P, T, D);
    TC_LAUNCH_CHECK();
}

void attribution_reduce(const float* probe, float* raw, float* attr, int C, int H, int P, cudaStream_t stream) {
    TC_CHECK(P >= 1 && P <= 32 * ATTR_PER_LANE, "prompt_len %d unsupported by the attribution kernel (1..192)", P);
    if (C == 0) return;
    launch_pdl(attribution_kernel, (unsigned)ceil_div(C, 4), 128, 0, stream, probe, raw, attr, C, H, P);
    TC_LAUNCH_CHECK();
}

void gather_rows(const float* x, void* out, int out_dt, int64_t rows, int64_t row_stride, int64_t row_offset, int d,
                 cudaStream_t stream) {
    if (rows == 0) return;
    if (out_dt == DT_BF16) launch_pdl(gather_rows_kernel<bf16>, grid_for(rows * d), 256, 0, stream, x, (bf16*)out, rows, row_stride, row_offset, d);
    else if (out_dt == DT_F16) launch_pdl(gather_rows_kernel<f16>, grid_for(rows * d), 256, 0, stream, x, (f16*)out, rows, row_stride, row_offset, d);
    else launch_pdl(gather_rows_kernel<float>, grid_for(rows * d), 256, 0, stream, x, (float*)out, rows, row_stride, row_offset, d);
    TC_LAUNCH_CHECK();
}

void embed_tokens(const int64_t* ids, const float* token_embedding, const float* pos, float* x, int32_t* eot, int S, int L, int D,
                  cudaStream_t stream) {
    if (S == 0) return;
    launch_pdl(embed_tokens_kernel, S, 256, 0, stream, ids, token_embedding, pos, x, eot, S, L, D);
    TC_LAUNCH_CHECK();
}

void gather_rows_indexed(const float* x, const int32_t* idx, float* out, int64_t rows, int64_t row_stride, int d, cudaStream_t stream) {
    if (rows == 0) return;
    launch_pdl(gather_rows_indexed_kernel, grid_for(rows * d), 256, 0, stream, x, idx, out, rows, row_stride, d);
    TC_LAUNCH_CHECK();
}

void scatter_rows(const float* src, float* dst, void* dst_cast, int cast_dt, int64_t rows, int64_t row_stride,
                  int64_t row_offset, int d, cudaStream_t stream) {
    if (rows == 0) return;
    TC_CHECK(cast_dt != DT_F16, "gradients are never fp16");
    if (cast_dt == DT_BF16) launch_pdl(scatter_rows_kernel<bf16>, grid_for(rows * d), 256, 0, stream, src, dst, (bf16*)dst_cast, rows, row_stride, row_offset, d);
    else launch_pdl(scatter_rows_kernel<float>, grid_for(rows * d), 256, 0, stream, src, dst, (float*)dst_cast, rows, row_stride, row_offset, d);
    TC_LAUNCH_CHECK();
}

void cast_f32(const float* src, void* dst, int dst_dt, int64_t n, cudaStream_t stream) {
    if (n == 0) return;
    if (dst_dt == DT_BF16) launch_pdl(cast_kernel<bf16>, grid_for(n), 256, 0, stream, src, (bf16*)dst, n);
    else if (dst_dt == DT_F16) launch_pdl(cast_kernel<f16>, grid_for(n), 256, 0, stream, src, (f16*)dst, n);
    else launch_pdl(cast_kernel<float>, grid_for(n), 256, 0, stream, src, (float*)dst, n);
    TC_LAUNCH_CHECK();
}

void act_bwd_inplace(void* dh, int dh_dt, const void* h_pre, int h_dt, int act, int64_t n, cudaStream_t stream) {
    if (n == 0) return;
    const unsigned grid = grid_for(n);
#define TC_ACT_BWD(T, TH)                                                                                             \
    do {                                                                                                              \
        if (act == ACT_GELU_ERF) launch_pdl(act_bwd_kernel<T, TH, ACT_GELU_ERF>, grid, 256, 0, stream, (T*)dh, (const TH*)h_pre, n); \
        else launch_pdl(act_bwd_kernel<T, TH, ACT_QUICK_GELU>, grid, 256, 0, stream, (T*)dh, (const TH*)h_pre, n);             \
    } while (0)
    if (dh_dt == DT_F32 && h_dt == DT_F32) TC_ACT_BWD(float, float);
    else if (dh_dt == DT_BF16 && h_dt == DT_BF16) TC_ACT_BWD(bf16, bf16);
    else if (dh_dt == DT_BF16 && h_dt == DT_F16) TC_ACT_BWD(bf16, f16);
    else TC_CHECK(false, "unsupported dtype combination for act_bwd (%d, %d)", dh_dt, h_dt);
#undef TC_ACT_BWD
    TC_LAUNCH_CHECK();
}

void cosine_logits(const float* img, const float* txt, const float* logit_scale, float* logits, int B, int C, int E,
                   cudaStream_t stream) {
    TC_CHECK(E % 4 == 0, "embed dim must be a multiple of 4");
    const int64_t pairs = (int64_t)B * C;
    if (pairs == 0) return;
    launch_pdl(cosine_logits_kernel, (unsigned)ceil_div(pairs, 8), 256, 0, stream, img, txt, logit_scale, logits, B, C, E);
    TC_LAUNCH_CHECK();
}

void cross_entropy(const float* logits, const int64_t* labels, float* loss, float* dlogits, float* row_scratch, int B, int C,
                   float inv_batch_total, cudaStream_t stream) {
    if (B == 0) return;
    launch_pdl(ce_rows_kernel, (unsigned)ceil_div(B, 8), 256, 0, stream, logits, labels, row_scratch, dlogits, B, C, inv_batch_total);
    TC_LAUNCH_CHECK();
    launch_pdl(sum_kernel, 1, 256, 0, stream, row_scratch, loss, B);
    TC_LAUNCH_CHECK();
}

void logits_bwd(const float* dlogits, const float* logits, const float* img, const float* logit_scale, float* d_txt,
                float* d_scale, float* class_scratch, int B, int C, int E, cudaStream_t stream) {
    if (C == 0) return;
    launch_pdl(logits_bwd_kernel, C, 256, 0, stream, dlogits, logits, img, logit_scale, d_txt, class_scratch, B, C, E);
    TC_LAUNCH_CHECK();
    launch_pdl(sum_kernel, 1, 256, 0, stream, class_scratch, d_scale, C);
    TC_LAUNCH_CHECK();
}

void adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                float wd, int step, cudaStream_t stream) {
    if (n == 0) return;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    launch_pdl(adamw_kernel, grid_for(n), 256, 0, stream, p, g, m, v, n, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt);
    TC_LAUNCH_CHECK();
}

void argmax_count(const float* logits, const int64_t* labels, int64_t* pred, int* correct, int* class_correct, int* class_total, int B,
                  int C, cudaStream_t stream) {
    if (B == 0) return;
    launch_pdl(argmax_kernel, (unsigned)ceil_div(B, 8), 256, 0, stream, logits, labels, pred, correct, class_correct, class_total, B, C);
    TC_LAUNCH_CHECK();
}

}  // namespace tapclip
