// K4 — the head of the hot path as two kernels (+ their two backward kernels), HBM/latency-bound, everything fp32 except the
// frozen projection weight:
//   text_head      rows A10 (models/model_wrapper.py:73-75): gather position T-1 of every class sequence, @ text_projection,
//                  L2-normalise -> T^ [C,E] (+ 1/||.|| for the backward).  One CTA per class, two output columns per thread:
//                  the projection is read as [D,E] (the reference's own layout) so every k-step is one coalesced row.
//   logits_ce      rows A5, A11, A12 (model_wrapper.py:41,79,83,90-93): L2-normalise the image row, logit_scale * I^ . T^^T for all
//                  classes, cross-entropy row loss and dloss/dlogits, and the batch mean -- one CTA per image; the LAST CTA to finish
//                  (atomic ticket) sums the per-row losses in a fixed order, so the loss is deterministic without a second launch.
//   logits_bwd     d T^ [C,E] = exp(s) * dlogits^T . I^ and d logit_scale = sum dlogits * logits (same last-CTA reduction).
//   text_head_bwd  L2-norm backward, @ text_projection^T, scattered into position T-1 of the (pre-zeroed) dx stream.
// These replace gather_rows + a [C,D]x[D,E] GEMM launch + l2norm (text side), l2norm + cosine_logits + ce_rows + sum (logit side),
// logits_bwd + sum and l2norm_bwd + GEMM + scatter_rows of the first version: 7 -> 2 launches forward, 5 -> 2 backward.
#include "kernels.h"

namespace tapclip {
namespace {

constexpr int HEAD_THREADS = 256;

template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<bf16>(const bf16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
template <> __device__ __forceinline__ void load8<f16>(const f16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i])); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// block-wide sum (all threads get the result); `sh` holds one float per warp
__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();                                   // sh may still be read from a previous call
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += sh[i];           // fixed order: deterministic
    return t;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
    v = warp_max(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float t = -INFINITY;
    for (int i = 0; i < nw; ++i) t = fmaxf(t, sh[i]);
    return t;
}

// out[n] = sum_k vec[k] * W[k, n] for n in [0, N), W row-major [K, N] (N % 8 == 0).  The CTA's 256 threads form 4 groups of 64; a
// group takes a quarter of the k range, a thread 8 adjacent columns (one 16-byte / 32-byte load per k step), and 8 k steps are
// loaded before any of them is used, so 8 independent loads per thread are in flight; the 4 partial results meet in shared memory.
// (History: walking rows of a [N, K] matrix warp by warp -- two dependent L2 round trips per output -- took 70 us for a 512 x 512
// projection of 65 rows; two columns per thread with 4-byte loads was still 80 us: the kernel is pure load latency.)
// part_s: [4][N] floats of scratch; out_s may alias part_s[0].  Needs blockDim.x == 256.
template <typename TW>
__device__ __forceinline__ void matvec_cols(const float* vec_s, const TW* __restrict__ W, float* out_s, float* part_s, int N, int K) {
    const int grp = threadIdx.x >> 6, j = threadIdx.x & 63;
    const int k_lo = (K * grp) / 4, k_hi = (K * (grp + 1)) / 4;
    for (int n0 = j * 8; n0 < N; n0 += 64 * 8) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        int k = k_lo;
        for (; k + 8 <= k_hi; k += 8) {
            float w[8][8];
#pragma unroll
            for (int u = 0; u < 8; ++u) load8<TW>(W + (int64_t)(k + u) * N + n0, w[u]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float v = vec_s[k + u];
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(v, w[u][i], acc[i]);
            }
        }
        for (; k < k_hi; ++k) {
            float w[8];
            load8<TW>(W + (int64_t)k * N + n0, w);
            const float v = vec_s[k];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(v, w[i], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) part_s[grp * N + n0 + i] = acc[i];
    }
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x)
        out_s[n] = (part_s[n] + part_s[N + n]) + (part_s[2 * N + n] + part_s[3 * N + n]);
    __syncthreads();
}

// K5: the text-feature all-gather fused into the kernel that produces the features.  Every rank owns a symmetric buffer
// (torch symmetric memory: the same allocation mapped into every peer's address space over NVLink); the head kernel stores its
// normalised rows straight into EVERY rank's buffer at the rows' global class offsets (peer stores), then the last CTA publishes
// `epoch` in flag slot [rank] of every peer (system-scope release).  The consumer (logits_ce) spins on its local flags until every
// rank's epoch has arrived (system-scope acquire).  Two slots, used by epoch parity: a rank that runs one step ahead never
// overwrites rows a slower peer is still reading.
template <typename TW>
__global__ void __launch_bounds__(HEAD_THREADS)
text_head_kernel(const float* __restrict__ x, int64_t row_stride, int64_t row_offset, const TW* __restrict__ w_proj /*[D,E]*/,
                 float* __restrict__ tfeat, float* __restrict__ inv_norm, float* __restrict__ tfeat_copy, int D, int E, const PeerScatter ps) {
    pdl_wait_and_trigger();
    extern __shared__ float sm[];
    float* xs = sm;                 // [D]
    float* fs = sm + D;             // [4][E] partial results of the four k groups; fs[0..E) then holds the projected row
    float* red = fs + 4 * E;        // [32]
    const int c = blockIdx.x;
    const float* xr = x + ((int64_t)c * row_stride + row_offset) * D;
    for (int k = threadIdx.x * 4; k < D; k += HEAD_THREADS * 4) *reinterpret_cast<float4*>(xs + k) = *reinterpret_cast<const float4*>(xr + k);
    __syncthreads();
    matvec_cols<TW>(xs, w_proj, fs, fs, E, D);
    float q = 0.f;
    for (int e = threadIdx.x; e < E; e += HEAD_THREADS) q += fs[e] * fs[e];
    const float inv = 1.0f / sqrtf(block_sum(q, red));
    if (threadIdx.x == 0) inv_norm[c] = inv;
    for (int e = threadIdx.x; e < E; e += HEAD_THREADS) {
        const float v = fs[e] * inv;
        tfeat[(int64_t)c * E + e] = v;
        if (tfeat_copy) tfeat_copy[(int64_t)c * E + e] = v;
        for (int r = 0; r < ps.world; ++r) ps.dst[r][(ps.row_lo + c) * E + e] = v;          // peer stores (NVLink), incl. this rank's own buffer
    }
    if (ps.world > 0) {
        __shared__ int is_last;
        __threadfence_system();                          // this CTA's peer stores are ordered before its ticket
        __syncthreads();
        if (threadIdx.x == 0) is_last = (atomicAdd(ps.ticket, 1) == (int)gridDim.x - 1);
        __syncthreads();
        if (is_last && threadIdx.x < ps.world) {
            __threadfence_system();
            asm volatile("st.release.sys.global.b32 [%0], %1;" ::"l"(ps.flag[threadIdx.x]), "r"(ps.epoch) : "memory");
            if (threadIdx.x == 0) *ps.ticket = 0;
        }
    }
}

template <typename TW, typename TG>
__global__ void __launch_bounds__(HEAD_THREADS)
text_head_bwd_kernel(const float* __restrict__ g, const float* __restrict__ tfeat, const float* __restrict__ inv_norm,
                     const TW* __restrict__ wt_proj /*[E,D]*/, float* __restrict__ dx, TG* __restrict__ dx_cast, int64_t row_stride,
                     int64_t row_offset, int D, int E) {
    pdl_wait_and_trigger();
    extern __shared__ float sm[];
    float* gs = sm;                 // [E]  d feat (before the projection)
    float* ds = sm + E;             // [4][D] partials; ds[0..D) then holds the result
    float* red = ds + 4 * D;
    const int c = blockIdx.x;
    float dot = 0.f;
    for (int e = threadIdx.x; e < E; e += HEAD_THREADS) dot += g[(int64_t)c * E + e] * tfeat[(int64_t)c * E + e];
    dot = block_sum(dot, red);
    const float inv = inv_norm[c];
    for (int e = threadIdx.x; e < E; e += HEAD_THREADS)
        gs[e] = (g[(int64_t)c * E + e] - tfeat[(int64_t)c * E + e] * dot) * inv;          // d/dx of x / ||x||
    __syncthreads();
    matvec_cols<TW>(gs, wt_proj, ds, ds, D, E);
    const int64_t o = ((int64_t)c * row_stride + row_offset) * D;
    for (int k = threadIdx.x; k < D; k += HEAD_THREADS) {
        dx[o + k] = ds[k];
        if (dx_cast) dx_cast[o + k] = from_f32<TG>(ds[k]);
    }
}

// ticket[0]: CTAs finished so far (self-resetting); the last CTA reduces `partial[0..n)` into out[0] in index order
__device__ __forceinline__ void last_block_sum(const float* partial, int n, float* out, int* ticket, float* red) {
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += __ldcg(partial + i);
    s = block_sum(s, red);
    if (threadIdx.x == 0) { *out = s; *ticket = 0; }
}

__global__ void __launch_bounds__(HEAD_THREADS)
logits_ce_kernel(const float* __restrict__ img, const float* txt, const float* __restrict__ logit_scale,
                 const int64_t* __restrict__ labels, float* __restrict__ img_norm, float* __restrict__ logits, float* __restrict__ loss,
                 float* __restrict__ dlogits, float* __restrict__ row_loss, int* __restrict__ ticket, int B, int C, int E,
                 float inv_batch_total, const int* wait_flags, int wait_world, int wait_epoch) {
    pdl_wait_and_trigger();
    if (wait_world > 0) {
        // K5 consumer side: the text features in `txt` (this rank's symmetric buffer) are complete once every rank has published
        // the epoch.  Bounded spin: a peer that never arrives becomes a trap, not a hung GPU.
        if (threadIdx.x < wait_world) {
            const long long t0 = clock64();
            int v;
            do {
                asm volatile("ld.acquire.sys.global.b32 %0, [%1];" : "=r"(v) : "l"(wait_flags + threadIdx.x) : "memory");
                if (v < wait_epoch && clock64() - t0 > 20000000000LL) {
                    printf("tapclip: rank %d never published text-feature epoch %d (flag %d)\n", (int)threadIdx.x, wait_epoch, v);
                    __trap();
                }
            } while (v < wait_epoch);
        }
        __syncthreads();
    }
    extern __shared__ float sm[];
    float* is = sm;                 // [E] normalised image row
    float* ls = sm + E;             // [C] logits of this row
    float* red = ls + C;            // [32]
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = HEAD_THREADS / 32;
    float q = 0.f;
    for (int e = threadIdx.x; e < E; e += HEAD_THREADS) { const float v = img[(int64_t)b * E + e]; is[e] = v; q += v * v; }
    const float inv = 1.0f / sqrtf(block_sum(q, red));
    for (int e = threadIdx.x; e < E; e += HEAD_THREADS) { const float v = is[e] * inv; is[e] = v; img_norm[(int64_t)b * E + e] = v; }
    __syncthreads();
    const float es = expf(__ldg(logit_scale));
    // one warp per class, two classes in flight per warp and the E loop unrolled: the loads of a step do not wait for each other
    for (int c = warp; c < C; c += 2 * nw) {
        const int c2 = c + nw;
        float s = 0.f, s2 = 0.f;
#pragma unroll 4
        for (int e = lane * 4; e < E; e += 128) {
            const float4 t = *reinterpret_cast<const float4*>(txt + (int64_t)c * E + e);      // may be peer-written: no read-only path
            s += (is[e] * t.x + is[e + 1] * t.y) + (is[e + 2] * t.z + is[e + 3] * t.w);
            if (c2 < C) {
                const float4 u = *reinterpret_cast<const float4*>(txt + (int64_t)c2 * E + e);
                s2 += (is[e] * u.x + is[e + 1] * u.y) + (is[e + 2] * u.z + is[e + 3] * u.w);
            }
        }
        s = warp_sum(s); s2 = warp_sum(s2);
        if (lane == 0) {
            const float l = es * s; ls[c] = l; logits[(int64_t)b * C + c] = l;
            if (c2 < C) { const float l2 = es * s2; ls[c2] = l2; logits[(int64_t)b * C + c2] = l2; }
        }
    }
    if (labels == nullptr) return;
    __syncthreads();
    // models/model_wrapper.py:91 F.cross_entropy, mean over the GLOBAL batch (inv_batch_total = 1 / sum of per-rank batches)
    float m = -INFINITY;
    for (int c = threadIdx.x; c < C; c += HEAD_THREADS) m = fmaxf(m, ls[c]);
    m = block_max(m, red);
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += HEAD_THREADS) s += expf(ls[c] - m);
    const float lse = m + logf(block_sum(s, red));
    const int64_t label = labels[b];
    // a label outside [0, C) (F.cross_entropy's ignore_index included: the reference never uses it) poisons the loss with NaN
    // instead of reading out of bounds: torch raises a device-side assert there
    const bool ok = label >= 0 && label < C;
    if (threadIdx.x == 0) row_loss[b] = ok ? (lse - ls[label]) * inv_batch_total : __int_as_float(0x7fc00000);
    if (dlogits)
        for (int c = threadIdx.x; c < C; c += HEAD_THREADS)
            dlogits[(int64_t)b * C + c] = ok ? (expf(ls[c] - lse) - (c == label ? 1.f : 0.f)) * inv_batch_total : __int_as_float(0x7fc00000);
    last_block_sum(row_loss, B, loss, ticket, red);
}

// d_txt[c,e] = exp(s) * sum_b dlogits[b,c] * img[b,e]  (one CTA per class);  d_scale = sum_{b,c} dlogits * logits
__global__ void __launch_bounds__(HEAD_THREADS)
logits_bwd_fused_kernel(const float* __restrict__ dlogits, const float* __restrict__ logits, const float* __restrict__ img,
                        const float* __restrict__ logit_scale, float* __restrict__ d_txt, float* __restrict__ d_scale,
                        float* __restrict__ class_part, int* __restrict__ ticket, int B, int C, int E) {
    pdl_wait_and_trigger();
    extern __shared__ float sm[];
    float* dl = sm;                 // [B] column c of dlogits
    float* red = sm + B;
    const int c = blockIdx.x;
    float p = 0.f;
    for (int b = threadIdx.x; b < B; b += HEAD_THREADS) {
        const float d = dlogits[(int64_t)b * C + c];
        dl[b] = d;
        p += d * logits[(int64_t)b * C + c];
    }
    p = block_sum(p, red);          // also orders the dl[] writes before the reads below
    if (threadIdx.x == 0) class_part[c] = p;
    const float es = expf(__ldg(logit_scale));
    for (int e = threadIdx.x * 2; e < E; e += HEAD_THREADS * 2) {     // two adjacent columns per thread; 8 rows loaded before any is used
        float s0 = 0.f, s1 = 0.f;
        int b = 0;
        for (; b + 8 <= B; b += 8) {
            float2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const float2*>(img + (int64_t)(b + u) * E + e));
#pragma unroll
            for (int u = 0; u < 8; ++u) { s0 = fmaf(dl[b + u], v[u].x, s0); s1 = fmaf(dl[b + u], v[u].y, s1); }
        }
        for (; b < B; ++b) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(img + (int64_t)b * E + e));
            s0 = fmaf(dl[b], v.x, s0); s1 = fmaf(dl[b], v.y, s1);
        }
        d_txt[(int64_t)c * E + e] = es * s0; d_txt[(int64_t)c * E + e + 1] = es * s1;
    }
    last_block_sum(class_part, C, d_scale, ticket, red);
}

}  // namespace

void text_head(const float* x, int64_t row_stride, int64_t row_offset, const void* w_proj, int w_dt, float* tfeat, float* inv_norm,
               float* tfeat_copy, int C, int D, int E, cudaStream_t stream, const PeerScatter* peers) {
    PeerScatter ps = {};
    if (peers) ps = *peers;
    if (C == 0) return;
    TC_CHECK(D % 8 == 0 && E % 8 == 0, "text_head needs D %% 8 == 0 and E %% 8 == 0");
    const size_t smem = (size_t)(D + 4 * E + 32) * sizeof(float);
    if (w_dt == DT_BF16) launch_pdl(text_head_kernel<bf16>, C, HEAD_THREADS, smem, stream, x, row_stride, row_offset, (const bf16*)w_proj, tfeat, inv_norm, tfeat_copy, D, E, ps);
    else if (w_dt == DT_F16) launch_pdl(text_head_kernel<f16>, C, HEAD_THREADS, smem, stream, x, row_stride, row_offset, (const f16*)w_proj, tfeat, inv_norm, tfeat_copy, D, E, ps);
    else launch_pdl(text_head_kernel<float>, C, HEAD_THREADS, smem, stream, x, row_stride, row_offset, (const float*)w_proj, tfeat, inv_norm, tfeat_copy, D, E, ps);
    TC_LAUNCH_CHECK();
}

void text_head_bwd(const float* g, const float* tfeat, const float* inv_norm, const void* wt_proj, int w_dt, float* dx, void* dx_cast,
                   int cast_dt, int64_t row_stride, int64_t row_offset, int C, int D, int E, cudaStream_t stream) {
    if (C == 0) return;
    TC_CHECK(D % 8 == 0 && E % 8 == 0, "text_head_bwd needs D %% 8 == 0 and E %% 8 == 0");
    TC_CHECK(cast_dt != DT_F16 && (w_dt == DT_F32) == (cast_dt == DT_F32), "text_head_bwd: gradients are bf16 or fp32, the projection 16-bit or fp32");
    const size_t smem = (size_t)(4 * D + E + 32) * sizeof(float);
    if (w_dt == DT_BF16) launch_pdl(text_head_bwd_kernel<bf16, bf16>, C, HEAD_THREADS, smem, stream, g, tfeat, inv_norm, (const bf16*)wt_proj, dx, (bf16*)dx_cast, row_stride, row_offset, D, E);
    else if (w_dt == DT_F16) launch_pdl(text_head_bwd_kernel<f16, bf16>, C, HEAD_THREADS, smem, stream, g, tfeat, inv_norm, (const f16*)wt_proj, dx, (bf16*)dx_cast, row_stride, row_offset, D, E);
    else launch_pdl(text_head_bwd_kernel<float, float>, C, HEAD_THREADS, smem, stream, g, tfeat, inv_norm, (const float*)wt_proj, dx, (float*)dx_cast, row_stride, row_offset, D, E);
    TC_LAUNCH_CHECK();
}

void logits_ce(const float* img, const float* txt, const float* logit_scale, const int64_t* labels, float* img_norm, float* logits,
               float* loss, float* dlogits, float* row_scratch, int* ticket, int B, int C, int E, float inv_batch_total, cudaStream_t stream,
               const int* wait_flags, int wait_world, int wait_epoch) {
    if (B == 0) return;
    TC_CHECK(E % 4 == 0, "embed dim must be a multiple of 4");
    const size_t smem = (size_t)(E + C + 32) * sizeof(float);
    TC_CHECK(smem <= 48 * 1024, "too many classes for the fused logits kernel (%d)", C);
    launch_pdl(logits_ce_kernel, B, HEAD_THREADS, smem, stream, img, txt, logit_scale, labels, img_norm, logits, loss, dlogits, row_scratch,
               ticket, B, C, E, inv_batch_total, wait_flags, wait_world, wait_epoch);
    TC_LAUNCH_CHECK();
}

void logits_bwd_fused(const float* dlogits, const float* logits, const float* img, const float* logit_scale, float* d_txt, float* d_scale,
                      float* class_scratch, int* ticket, int B, int C, int E, cudaStream_t stream) {
    if (C == 0) return;
    const size_t smem = (size_t)(B + 32) * sizeof(float);
    TC_CHECK(smem <= 48 * 1024, "batch too large for the fused logits backward kernel (%d)", B);
    launch_pdl(logits_bwd_fused_kernel, C, HEAD_THREADS, smem, stream, dlogits, logits, img, logit_scale, d_txt, d_scale, class_scratch,
               ticket, B, C, E);
    TC_LAUNCH_CHECK();
}

}  // namespace tapclip
