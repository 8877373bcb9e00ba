// Row-wise normalisation kernels (HBM-bound): LayerNorm fwd/bwd and L2-norm fwd/bwd.
// One warp per row, the whole row lives in registers (d <= 1024, d % 128 == 0), float4 loads,
// statistics in fp32 via warp shuffles.  Replaces ATen layer_norm and the
// `x / x.norm(dim=-1, keepdim=True)` pairs of models/model_wrapper.py:41,75.
#include "kernels.h"

namespace tapclip {
namespace {

constexpr int MAXV = 8;          // float4 per lane -> d <= 1024
constexpr int WARPS = 8;

template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void store4<bf16>(bf16* p, float4 v) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
}
template <> __device__ __forceinline__ void store4<f16>(f16* p, float4 v) {
    uint2 u;
    u.x = pack_f16x2(v.x, v.y);
    u.y = pack_f16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
layernorm_fwd_kernel(const float* x, int64_t x_row_stride, const float* __restrict__ gamma,
                     const float* __restrict__ beta, T* out, float* x_copy, int64_t rows, int d) {
    pdl_wait_and_trigger();
    const int64_t row = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31, nv = d >> 7;
    const float* xr = x + row * x_row_stride;
    float4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
            q += (a * a + b * b) + (c * c + e * e);
        }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const int c = (i * 32 + lane) * 4;
            if (x_copy) *reinterpret_cast<float4*>(x_copy + row * d + c) = v[i];
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
            float4 o;
            o.x = (v[i].x - mean) * rstd * g.x + b.x;
            o.y = (v[i].y - mean) * rstd * g.y + b.y;
            o.z = (v[i].z - mean) * rstd * g.z + b.z;
            o.w = (v[i].w - mean) * rstd * g.w + b.w;
            store4<T>(out + row * d + c, o);
        }
}

// open_clip VisionTransformer.forward prologue in one pass: x[b,t,:] = ln_pre(cat([class_embedding, patches])[b,t,:] + pos[t,:])
// (replaces assemble_tokens + an in-place LayerNorm: the fp32 token matrix is written once instead of three times)
__global__ void __launch_bounds__(WARPS * 32)
assemble_ln_pre_kernel(const float* __restrict__ patch_out, const float* __restrict__ cls, const float* __restrict__ pos,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ x, int64_t rows,
                       int n_tokens, int d, bf16* __restrict__ xb, float2* __restrict__ stats, float* __restrict__ shift) {
    pdl_wait_and_trigger();
    const int64_t row = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31, nv = d >> 7;
    const int t = (int)(row % n_tokens);
    const int64_t b = row / n_tokens;
    const float* src = (t == 0) ? cls : patch_out + (b * (n_tokens - 1) + (t - 1)) * d;
    float4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const int c = (i * 32 + lane) * 4;
            v[i] = __ldg(reinterpret_cast<const float4*>(src + c));
            const float4 pe = __ldg(reinterpret_cast<const float4*>(pos + (int64_t)t * d + c));
            v[i].x += pe.x; v[i].y += pe.y; v[i].z += pe.z; v[i].w += pe.w;
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const float a = v[i].x - mean, b2 = v[i].y - mean, c2 = v[i].z - mean, e = v[i].w - mean;
            q += (a * a + b2 * b2) + (c2 * c2 + e * e);
        }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const int c = (i * 32 + lane) * 4;
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + c));
            float4 o;
            o.x = (v[i].x - mean) * rstd * g.x + bb.x;
            o.y = (v[i].y - mean) * rstd * g.y + bb.y;
            o.z = (v[i].z - mean) * rstd * g.z + bb.z;
            o.w = (v[i].w - mean) * rstd * g.w + bb.w;
            *reinterpret_cast<float4*>(x + row * d + c) = o;
            v[i] = o;
        }
    if (xb != nullptr) {
        // what the first block's folded-LayerNorm QKV GEMM consumes (gemm.h, GemmArgs::stats_in): the row, shifted by its own
        // mean, in 16 bits; its (sum, sum of squares) as a single partial; and the shift for the next producer
        float s0 = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i)
            if (i < nv) s0 += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        const float m = warp_sum(s0) / (float)d;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i)
            if (i < nv) {
                const float4 z = make_float4(v[i].x - m, v[i].y - m, v[i].z - m, v[i].w - m);
                store4<bf16>(xb + row * d + (i * 32 + lane) * 4, z);
                s1 += (z.x + z.y) + (z.z + z.w);
                s2 += (z.x * z.x + z.y * z.y) + (z.z * z.z + z.w * z.w);
            }
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (lane == 0) { stats[row] = make_float2(s1, s2); shift[row] = m; }
    }
}

// x [rows, d] fp32 -> shift[row] = mean of the row, xb = x - shift in the 16-bit operand type, stats[row] = (sum, sum of squares) of
// the shifted row: the inputs of a folded-LayerNorm GEMM for a residual stream that was not produced by an EPI_F32_RESID GEMM
// (the spliced prompts of the text tower)
template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
row_stats_cast_kernel(const float* __restrict__ x, T* __restrict__ xb, float2* __restrict__ stats, float* __restrict__ shift, int64_t rows,
                      int d) {
    pdl_wait_and_trigger();
    const int64_t row = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31, nv = d >> 7;
    float4 v[MAXV];
    float s0 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            v[i] = *reinterpret_cast<const float4*>(x + row * d + (i * 32 + lane) * 4);
            s0 += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    const float m = warp_sum(s0) / (float)d;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const float4 z = make_float4(v[i].x - m, v[i].y - m, v[i].z - m, v[i].w - m);
            store4<T>(xb + row * d + (i * 32 + lane) * 4, z);
            s1 += (z.x + z.y) + (z.z + z.w);
            s2 += (z.x * z.x + z.y * z.y) + (z.z * z.z + z.w * z.w);
        }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { stats[row] = make_float2(s1, s2); shift[row] = m; }
}

// NV = float4 per lane the instance holds (4: d <= 512, the text tower - half the registers, twice the resident warps; 8: d <= 1024)
template <typename T, int NV>
__global__ void __launch_bounds__(WARPS * 32)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     float* dx_acc, T* dx_cast, int64_t rows, int d, int64_t dx_row_stride) {
    pdl_wait_and_trigger();
    const int64_t row = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31, nv = d >> 7;
    // All three operand rows (x, dy, the dx accumulator) are requested before the first reduction: the rows fill one wave of
    // CTAs, so the kernel's duration is the latency chain of a single warp (three dependent round trips before this).
    float4 v[NV], g[NV], a[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
        if (i < nv) {
            const int c = (i * 32 + lane) * 4;
            v[i] = *reinterpret_cast<const float4*>(x + row * d + c);
            g[i] = *reinterpret_cast<const float4*>(dy + row * d + c);
            a[i] = *reinterpret_cast<const float4*>(dx_acc + row * dx_row_stride + c);
        }
#pragma unroll
    for (int i = 0; i < NV; ++i)
        if (i < nv) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
        if (i < nv) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
    // g = dy * gamma ; xhat = (x-mean)*rstd ; dx = rstd * (g - mean(g) - xhat*mean(g*xhat))
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
        if (i < nv) {
            const int c = (i * 32 + lane) * 4;
            const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
            g[i].x *= gm.x; g[i].y *= gm.y; g[i].z *= gm.z; g[i].w *= gm.w;
            v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;
            sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
            sgx += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
        }
    const float mg = warp_sum(sg) / (float)d, mgx = warp_sum(sgx) / (float)d;
#pragma unroll
    for (int i = 0; i < NV; ++i)
        if (i < nv) {
            const int c = (i * 32 + lane) * 4;
            a[i].x += rstd * (g[i].x - mg - v[i].x * mgx);
            a[i].y += rstd * (g[i].y - mg - v[i].y * mgx);
            a[i].z += rstd * (g[i].z - mg - v[i].z * mgx);
            a[i].w += rstd * (g[i].w - mg - v[i].w * mgx);
            *reinterpret_cast<float4*>(dx_acc + row * dx_row_stride + c) = a[i];
            if (dx_cast) store4<T>(dx_cast + row * dx_row_stride + c, a[i]);
        }
}

__global__ void __launch_bounds__(WARPS * 32)
l2norm_fwd_kernel(const float* __restrict__ x, float* out, float* inv_norm, int64_t rows, int d) {
    pdl_wait_and_trigger();
    const int64_t row = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31, nv = d >> 7;
    float4 v[MAXV];
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            v[i] = *reinterpret_cast<const float4*>(x + row * d + (i * 32 + lane) * 4);
            q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
    const float inv = 1.0f / sqrtf(warp_sum(q));
    if (inv_norm && lane == 0) inv_norm[row] = inv;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            float4 o = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
            *reinterpret_cast<float4*>(out + row * d + (i * 32 + lane) * 4) = o;
        }
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
l2norm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ xhat, const float* __restrict__ inv_norm,
                  float* dx, T* dx_cast, int64_t rows, int d) {
    pdl_wait_and_trigger();
    const int64_t row = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31, nv = d >> 7;
    float4 gv[MAXV], xv[MAXV];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const int c = (i * 32 + lane) * 4;
            gv[i] = *reinterpret_cast<const float4*>(g + row * d + c);
            xv[i] = *reinterpret_cast<const float4*>(xhat + row * d + c);
            dot += (gv[i].x * xv[i].x + gv[i].y * xv[i].y) + (gv[i].z * xv[i].z + gv[i].w * xv[i].w);
        }
    dot = warp_sum(dot);
    const float inv = inv_norm[row];
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
        if (i < nv) {
            const int c = (i * 32 + lane) * 4;
            float4 o;
            o.x = (gv[i].x - xv[i].x * dot) * inv; o.y = (gv[i].y - xv[i].y * dot) * inv;
            o.z = (gv[i].z - xv[i].z * dot) * inv; o.w = (gv[i].w - xv[i].w * dot) * inv;
            *reinterpret_cast<float4*>(dx + row * d + c) = o;
            if (dx_cast) store4<T>(dx_cast + row * d + c, o);
        }
}

// LayerNorm folded into a Linear (gemm.h, GemmArgs::stats_in): one warp per output row n of W [N,K]:
//   Wf[n,k] = T(W[n,k] * gamma[k] - m_n),  m_n = (1/K) sum_k W[n,k] * gamma[k]   (row-centred: x Wf^T = (x - mean x)(W diag gamma)^T)
//   fb[n]   = bias[n] + sum_k W[n,k] * beta[k]
template <typename T>
__global__ void fold_ln_weight_kernel(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, T* __restrict__ Wf, float* __restrict__ fb, int N, int K) {
    pdl_wait_and_trigger();
    const int n = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    float s = 0.f, bb = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float w = W[(int64_t)n * K + k];
        s += w * gamma[k];
        bb += w * beta[k];
    }
    s = warp_sum(s); bb = warp_sum(bb);
    const float m = s / (float)K;
    for (int k = lane; k < K; k += 32) Wf[(int64_t)n * K + k] = from_f32<T>(W[(int64_t)n * K + k] * gamma[k] - m);
    if (lane == 0) fb[n] = (bias ? bias[n] : 0.f) + bb;
}

void check_d(int d) { TC_CHECK(d % 128 == 0 && d >= 128 && d <= 128 * MAXV, "row width %d unsupported (need d %% 128 == 0, d <= 1024)", d); }

}  // namespace

void layernorm_fwd(const float* x, int64_t x_row_stride, const float* gamma, const float* beta, void* out, int out_dt,
                   float* x_copy, int64_t rows, int d, cudaStream_t stream) {
    check_d(d);
    if (rows == 0) return;
    const unsigned grid = (unsigned)ceil_div(rows, WARPS);
    if (out_dt == DT_BF16) launch_pdl(layernorm_fwd_kernel<bf16>, grid, WARPS * 32, 0, stream, x, x_row_stride, gamma, beta, (bf16*)out, x_copy, rows, d);
    else if (out_dt == DT_F16) launch_pdl(layernorm_fwd_kernel<f16>, grid, WARPS * 32, 0, stream, x, x_row_stride, gamma, beta, (f16*)out, x_copy, rows, d);
    else launch_pdl(layernorm_fwd_kernel<float>, grid, WARPS * 32, 0, stream, x, x_row_stride, gamma, beta, (float*)out, x_copy, rows, d);
    TC_LAUNCH_CHECK();
}

void assemble_ln_pre(const float* patch_out, const float* cls, const float* pos, const float* gamma, const float* beta, float* x,
                     int B, int n_tokens, int d, cudaStream_t stream, void* xb, float* stats, float* shift) {
    check_d(d);
    const int64_t rows = (int64_t)B * n_tokens;
    if (rows == 0) return;
    TC_CHECK((xb == nullptr) == (stats == nullptr) && (xb == nullptr) == (shift == nullptr), "assemble_ln_pre: xb, stats and shift go together");
    launch_pdl(assemble_ln_pre_kernel, (unsigned)ceil_div(rows, WARPS), WARPS * 32, 0, stream, patch_out, cls, pos, gamma, beta, x, rows,
               n_tokens, d, (bf16*)xb, (float2*)stats, shift);
    TC_LAUNCH_CHECK();
}

void fold_ln_weight(const float* W, const float* bias, const float* gamma, const float* beta, void* Wf, int dt, float* fb, int N, int K,
                    cudaStream_t stream) {
    TC_CHECK(dt == DT_BF16 || dt == DT_F16, "folded weights are 16-bit operands");
    const unsigned grid = (unsigned)ceil_div((int64_t)N * 32, 256);
    if (dt == DT_F16) launch_pdl(fold_ln_weight_kernel<f16>, grid, 256, 0, stream, W, bias, gamma, beta, (f16*)Wf, fb, N, K);
    else launch_pdl(fold_ln_weight_kernel<bf16>, grid, 256, 0, stream, W, bias, gamma, beta, (bf16*)Wf, fb, N, K);
    TC_LAUNCH_CHECK();
}

void row_stats_cast(const float* x, void* xb, int xb_dt, float* stats, float* shift, int64_t rows, int d, cudaStream_t stream) {
    check_d(d);
    TC_CHECK(xb_dt == DT_BF16 || xb_dt == DT_F16, "row_stats_cast writes a 16-bit copy");
    if (rows == 0) return;
    const unsigned grid = (unsigned)ceil_div(rows, WARPS);
    if (xb_dt == DT_BF16) launch_pdl(row_stats_cast_kernel<bf16>, grid, WARPS * 32, 0, stream, x, (bf16*)xb, (float2*)stats, shift, rows, d);
    else launch_pdl(row_stats_cast_kernel<f16>, grid, WARPS * 32, 0, stream, x, (f16*)xb, (float2*)stats, shift, rows, d);
    TC_LAUNCH_CHECK();
}

void layernorm_bwd(const float* dy, const float* x, const float* gamma, float* dx_acc, void* dx_cast, int cast_dt,
                   int64_t rows, int d, cudaStream_t stream, int64_t dx_row_stride) {
    check_d(d);
    TC_CHECK(cast_dt != DT_F16, "gradients are never fp16");
    if (rows == 0) return;
    const unsigned grid = (unsigned)ceil_div(rows, WARPS);
    const int64_t ld = dx_row_stride ? dx_row_stride : (int64_t)d;
    if (d <= 512) {
        if (cast_dt == DT_BF16) launch_pdl(layernorm_bwd_kernel<bf16, 4>, grid, WARPS * 32, 0, stream, dy, x, gamma, dx_acc, (bf16*)dx_cast, rows, d, ld);
        else launch_pdl(layernorm_bwd_kernel<float, 4>, grid, WARPS * 32, 0, stream, dy, x, gamma, dx_acc, (float*)dx_cast, rows, d, ld);
    } else {
        if (cast_dt == DT_BF16) launch_pdl(layernorm_bwd_kernel<bf16, MAXV>, grid, WARPS * 32, 0, stream, dy, x, gamma, dx_acc, (bf16*)dx_cast, rows, d, ld);
        else launch_pdl(layernorm_bwd_kernel<float, MAXV>, grid, WARPS * 32, 0, stream, dy, x, gamma, dx_acc, (float*)dx_cast, rows, d, ld);
    }
    TC_LAUNCH_CHECK();
}

void l2norm_fwd(const float* x, float* out, float* inv_norm, int64_t rows, int d, cudaStream_t stream) {
    check_d(d);
    if (rows == 0) return;
    launch_pdl(l2norm_fwd_kernel, (unsigned)ceil_div(rows, WARPS), WARPS * 32, 0, stream, x, out, inv_norm, rows, d);
    TC_LAUNCH_CHECK();
}

void l2norm_bwd(const float* g, const float* xhat, const float* inv_norm, float* dx, void* dx_cast, int cast_dt,
                int64_t rows, int d, cudaStream_t stream) {
    check_d(d);
    TC_CHECK(cast_dt != DT_F16, "gradients are never fp16");
    if (rows == 0) return;
    const unsigned grid = (unsigned)ceil_div(rows, WARPS);
    if (cast_dt == DT_BF16) launch_pdl(l2norm_bwd_kernel<bf16>, grid, WARPS * 32, 0, stream, g, xhat, inv_norm, dx, (bf16*)dx_cast, rows, d);
    else launch_pdl(l2norm_bwd_kernel<float>, grid, WARPS * 32, 0, stream, g, xhat, inv_norm, dx, (float*)dx_cast, rows, d);
    TC_LAUNCH_CHECK();
}

}  // namespace tapclip
