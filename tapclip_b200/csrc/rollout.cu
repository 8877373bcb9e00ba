// Attention rollout of the CLS token (north-star extension, SURVEY 8a row A-ext; Abnar & Zuidema 2020) without ever
// forming an N x N map.
//
// Only the CLS row of R = A_L ... A_1, A_l = 0.5*mean_h P_l + 0.5*I, is asked for, so a row vector is propagated from the
// LAST layer to the first:   r <- 0.5 * r + (0.5/H) * sum_h r^T P_{l,h}.
// r^T P_h is the same contraction as dV = P^T dO in a flash-attention backward pass with dO one column wide: the
// probabilities are RECOMPUTED tile by tile from the layer's saved Q and K and the per-row softmax statistics
//      lse[s,h,i] = log2 sum_j 2^(c * q_i.k_j),  c = log2(e)/8
// (emitted by the forward attention kernels, or by attention_lse below), weighted by r_i and summed over the queries in
// registers.  Per layer that is ONE QK^T pass and one exp2 per (head, query, key) - the bound is the MUFU pipe - instead of
// the head-mean map's four QK^T passes plus a write and a read of B*N*N floats (16 GB at ViT-L/14@336, B=512, 24 layers).
// r_i is folded into the exponent: r_i * 2^(x - lse_i) = 2^(x - (lse_i - log2 r_i)); r_i = 0 gives +inf and a zero term.
//
//  * rollout_step_mma_kernel  16-bit Q/K: S^T = K Q^T on mma.sync m16n8k16; a CTA owns up to 256 keys of one image (32 per
//                             warp), loops over heads and 32-query steps; no atomics, deterministic.
//                             Small problems and N <= 128; rollout_step() sends everything else to the tcgen05 kernel in
//                             rollout_tc.cu (same contract).
//  * rollout_step_simt_kernel fp32 parity mode (one warp per key).
//  * attn_lse_mma_kernel / attn_lse_simt_kernel  the statistics alone, for forward kernels that do not emit them.
#include "kernels.h"
#include <cstdlib>

namespace tapclip {
namespace {

constexpr int DH = 64;
constexpr int QC = 32;                 // queries (rollout step) / keys (statistics) per inner step
constexpr float SCALE_LOG2 = 0.125f * 1.4426950408889634f;

// stage `rows` rows of 64 16-bit values (row r of the tile = global row r0 + r, zero beyond N) into XOR-swizzled smem
template <typename T>
__device__ __forceinline__ void stage_rows(uint8_t* dst, const T* src, int64_t row_stride, int r0, int rows, int N) {
    for (int idx = threadIdx.x; idx < rows * 8; idx += blockDim.x) {
        const int row = idx >> 3, ch = idx & 7, grow = r0 + row;
        const bool ok = grow < N;
        cp_async_16(smem_u32(dst + row * 128 + ((ch ^ (row & 7)) << 4)), src + (int64_t)(ok ? grow : 0) * row_stride + ch * 8, ok);
    }
}

// sc[nb][e]: rows g (e = 0,1) and g+8 (e = 2,3) of the warp's 16 A rows x columns c0 + nb*8 + tq*2 + (e&1) of the B tile
template <typename T>
__device__ __forceinline__ void tile_16x32(float (&sc)[4][4], const uint32_t (&af)[4][4], const uint8_t* Bs, int c0, int mat, int l7) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int nbp = 0; nbp < 2; ++nbp) {
            const int row = c0 + (nbp * 2 + (mat >> 1)) * 8 + l7, ch = ks * 2 + (mat & 1);
            uint32_t b[4];
            ldmatrix_x4(b, smem_u32(Bs + row * 128 + ((ch ^ (row & 7)) << 4)));
            mma_16816<T>(sc[nbp * 2], af[ks], b[0], b[1]);
            mma_16816<T>(sc[nbp * 2 + 1], af[ks], b[2], b[3]);
        }
    }
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// r_out[s, key - skip] = 0.5 * r_in[s, key] + (0.5/H) * sum_h sum_i r_in[s, i] * 2^(c q_i.k_key - lse[s,h,i])
// r_in == nullptr: r_in = e_0 (the CLS row of the identity).  skip = 1 on the last step drops the CLS column.
// A warp owns 32 keys (two 16-row A fragments kept in registers for the head), so every ldmatrix of a Q fragment feeds
// four MMAs: the shared-memory pipe was the busiest unit (59 %) with 16 keys per warp.
constexpr int RS_KEYS = 32;            // keys per warp

template <typename T>
__global__ void __launch_bounds__(256, 2)
rollout_step_mma_kernel(const T* __restrict__ qkv, const float* __restrict__ lse, const float* __restrict__ r_in,
                        float* __restrict__ r_out, int N, int H, int npad, int skip) {
    pdl_wait_and_trigger();
    extern __shared__ __align__(128) uint8_t smem[];
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x, d = H * DH;                     // (image-major launch order measured faster than tile-major: 1.53 vs 1.65 ms)
    const int kt = nwarps * RS_KEYS, k0 = blockIdx.y * kt;
    uint8_t* Ks = smem;                                       // [kt][128 B]    this CTA's keys of the current head
    uint8_t* Qs = Ks + kt * 128;                              // [npad][128 B]  every query of the current head
    float* lr = reinterpret_cast<float*>(Qs + npad * 128);    // [npad] log2 r_in (head independent)
    float* lw = lr + npad;                                    // [npad] lse - log2 r_in of the current head
    const int g = lane >> 2, tq = lane & 3, mat = lane >> 3, l7 = lane & 7;
    for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        float v = -INFINITY;
        if (i < N) v = r_in ? log2f(r_in[(int64_t)s * N + i]) : (i == 0 ? 0.f : -INFINITY);
        lr[i] = v;
    }
    float acc[2][2][2];                                       // [key fragment][row half][partial sum]
#pragma unroll
    for (int f = 0; f < 2; ++f) { acc[f][0][0] = acc[f][0][1] = acc[f][1][0] = acc[f][1][1] = 0.f; }
    for (int h = 0; h < H; ++h) {
        __syncthreads();                                      // previous head's tiles are no longer read; lr is visible
        const T* base = qkv + (int64_t)s * N * 3 * d + h * DH;
        stage_rows<T>(Ks, base + d, 3 * d, k0, kt, N);
        stage_rows<T>(Qs, base, 3 * d, 0, npad, N);
        cp_async_commit();
        const float* lse_h = lse + ((int64_t)s * H + h) * N;
        for (int i = threadIdx.x; i < npad; i += blockDim.x)
            lw[i] = (i < N && lr[i] != -INFINITY) ? lse_h[i] - lr[i] : INFINITY;     // r_i = 0: the row's statistics may be unwritten
        cp_async_wait<0>();
        __syncthreads();
        if (k0 + warp * RS_KEYS >= N) continue;               // warp-uniform; the barriers above are still reached
        uint32_t kf[2][4][4];
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int row = warp * RS_KEYS + f * 16 + (mat & 1) * 8 + l7, ch = ks * 2 + (mat >> 1);
                ldmatrix_x4(kf[f][ks], smem_u32(Ks + row * 128 + ((ch ^ (row & 7)) << 4)));
            }
        for (int qc = 0; qc < npad; qc += QC) {
            float sc[2][4][4];
#pragma unroll
            for (int f = 0; f < 2; ++f)
#pragma unroll
                for (int i = 0; i < 4; ++i) { sc[f][i][0] = sc[f][i][1] = sc[f][i][2] = sc[f][i][3] = 0.f; }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                for (int nbp = 0; nbp < 2; ++nbp) {
                    const int row = qc + (nbp * 2 + (mat >> 1)) * 8 + l7, ch = ks * 2 + (mat & 1);
                    uint32_t b[4];
                    ldmatrix_x4(b, smem_u32(Qs + row * 128 + ((ch ^ (row & 7)) << 4)));
#pragma unroll
                    for (int f = 0; f < 2; ++f) {
                        mma_16816<T>(sc[f][nbp * 2], kf[f][ks], b[0], b[1]);
                        mma_16816<T>(sc[f][nbp * 2 + 1], kf[f][ks], b[2], b[3]);
                    }
                }
            }
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const float2 w = *reinterpret_cast<const float2*>(lw + qc + nb * 8 + tq * 2);
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    acc[f][0][0] += ex2_approx(fmaf(sc[f][nb][0], SCALE_LOG2, -w.x));
                    acc[f][0][1] += ex2_approx(fmaf(sc[f][nb][1], SCALE_LOG2, -w.y));
                    acc[f][1][0] += ex2_approx(fmaf(sc[f][nb][2], SCALE_LOG2, -w.x));
                    acc[f][1][1] += ex2_approx(fmaf(sc[f][nb][3], SCALE_LOG2, -w.y));
                }
            }
        }
    }
    const float wh = 0.5f / (float)H;
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float a = acc[f][half][0] + acc[f][half][1];
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            const int key = k0 + warp * RS_KEYS + f * 16 + g + half * 8;
            if (tq == 0 && key < N && key >= skip) {
                const float rk = r_in ? r_in[(int64_t)s * N + key] : (key == 0 ? 1.f : 0.f);
                r_out[(int64_t)s * (N - skip) + key - skip] = fmaf(wh, a, 0.5f * rk);
            }
        }
}

// fp32 parity mode: one warp per key, lanes over the queries
__global__ void __launch_bounds__(128)
rollout_step_simt_kernel(const float* __restrict__ qkv, const float* __restrict__ lse, const float* __restrict__ r_in,
                         float* __restrict__ r_out, int N, int H, int skip) {
    pdl_wait_and_trigger();
    __shared__ float ks[4][DH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.y, key = blockIdx.x * 4 + warp;
    const int d = H * DH;
    if (key >= N) return;
    float acc = 0.f;
    for (int h = 0; h < H; ++h) {
        const float* base = qkv + (int64_t)s * N * 3 * d + h * DH;
        __syncwarp();
        ks[warp][lane] = base[(int64_t)key * 3 * d + d + lane];
        ks[warp][lane + 32] = base[(int64_t)key * 3 * d + d + lane + 32];
        __syncwarp();
        for (int i = lane; i < N; i += 32) {
            const float ri = r_in ? r_in[(int64_t)s * N + i] : (i == 0 ? 1.f : 0.f);
            if (ri == 0.f) continue;
            const float* qp = base + (int64_t)i * 3 * d;
            float a = 0.f;
#pragma unroll 8
            for (int j = 0; j < DH; ++j) a = fmaf(ks[warp][j], qp[j], a);
            acc = fmaf(ri, exp2f(a * SCALE_LOG2 - lse[((int64_t)s * H + h) * N + i]), acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0 && key >= skip) {
        const float rk = r_in ? r_in[(int64_t)s * N + key] : (key == 0 ? 1.f : 0.f);
        r_out[(int64_t)s * (N - skip) + key - skip] = fmaf(0.5f / (float)H, acc, 0.5f * rk);
    }
}

// lse[s,h,i] = log2 sum_j 2^(c q_i.k_j): a CTA owns 16 query rows per warp of one (image, head) and walks all keys
template <typename T>
__global__ void __launch_bounds__(256, 2)
attn_lse_mma_kernel(const T* __restrict__ qkv, float* __restrict__ lse, int N, int H, int npad) {
    pdl_wait_and_trigger();
    extern __shared__ __align__(128) uint8_t smem[];
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x / H, h = blockIdx.x % H, d = H * DH;
    const int qt = nwarps * 16, q0 = blockIdx.y * qt;
    uint8_t* Qs = smem;
    uint8_t* Ks = Qs + qt * 128;
    const int g = lane >> 2, tq = lane & 3, mat = lane >> 3, l7 = lane & 7;
    const T* base = qkv + (int64_t)s * N * 3 * d + h * DH;
    stage_rows<T>(Qs, base, 3 * d, q0, qt, N);
    stage_rows<T>(Ks, base + d, 3 * d, 0, npad, N);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int r0 = q0 + warp * 16;
    if (r0 >= N) return;
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int row = warp * 16 + (mat & 1) * 8 + l7, ch = ks * 2 + (mat >> 1);
        ldmatrix_x4(qf[ks], smem_u32(Qs + row * 128 + ((ch ^ (row & 7)) << 4)));
    }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    for (int kc = 0; kc < npad; kc += QC) {
        float sc[4][4];
        tile_16x32<T>(sc, qf, Ks, kc, mat, l7);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const bool pad = kc + nb * 8 + tq * 2 + e >= N;
                sc[nb][e] = pad ? -INFINITY : sc[nb][e] * SCALE_LOG2;
                sc[nb][e + 2] = pad ? -INFINITY : sc[nb][e + 2] * SCALE_LOG2;
                mx0 = fmaxf(mx0, sc[nb][e]);
                mx1 = fmaxf(mx1, sc[nb][e + 2]);
            }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);          // finite: every 32-key step holds a real key
        l0 *= exp2f(m0 - mn0); l1 *= exp2f(m1 - mn1);
        m0 = mn0; m1 = mn1;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            l0 += exp2f(sc[nb][0] - mn0) + exp2f(sc[nb][1] - mn0);
            l1 += exp2f(sc[nb][2] - mn1) + exp2f(sc[nb][3] - mn1);
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    if (tq == 0) {
        float* o = lse + ((int64_t)s * H + h) * N;
        if (r0 + g < N) o[r0 + g] = m0 + log2f(l0);
        if (r0 + g + 8 < N) o[r0 + g + 8] = m1 + log2f(l1);
    }
}

__global__ void __launch_bounds__(128)
attn_lse_simt_kernel(const float* __restrict__ qkv, float* __restrict__ lse, int N, int H) {
    pdl_wait_and_trigger();
    __shared__ float qs[4][DH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.y / H, h = blockIdx.y % H, row = blockIdx.x * 4 + warp;
    const int d = H * DH;
    if (row >= N) return;
    const float* base = qkv + (int64_t)s * N * 3 * d + h * DH;
    qs[warp][lane] = base[(int64_t)row * 3 * d + lane];
    qs[warp][lane + 32] = base[(int64_t)row * 3 * d + lane + 32];
    __syncwarp();
    float m = -INFINITY, l = 0.f;
    for (int key = lane; key < N; key += 32) {
        const float* kp = base + (int64_t)key * 3 * d + d;
        float a = 0.f;
#pragma unroll 8
        for (int j = 0; j < DH; ++j) a = fmaf(qs[warp][j], kp[j], a);
        const float v = a * SCALE_LOG2, mn = fmaxf(m, v);
        l = l * exp2f(m - mn) + exp2f(v - mn);
        m = mn;
    }
    const float mw = warp_max(m);
    l = warp_sum(l * exp2f(m - mw));                           // lanes without a key: l = 0, m = -inf -> 0 * 0
    if (lane == 0) lse[((int64_t)s * H + h) * N + row] = mw + log2f(l);
}

// warps per CTA so that ceil(N/16) 16-row blocks split evenly over the fewest CTAs of at most 8 warps
inline int warps_for(int N, int& ntiles) {
    const int nrb = (int)ceil_div(N, 16);
    ntiles = (int)ceil_div(nrb, 8);
    return (int)ceil_div(nrb, ntiles);
}

}  // namespace

void attention_lse(const void* qkv, float* lse, int dt, int S, int N, int H, cudaStream_t stream) {
    if (S == 0) return;
    TC_CHECK(N >= 1 && H >= 1 && lse, "bad attention_lse arguments");
    if (dt == DT_BF16 || dt == DT_F16) {
        int ntiles;
        const int nwarps = warps_for(N, ntiles), npad = (int)round_up(N, QC);
        const size_t smem = (size_t)(nwarps * 16 + npad) * 128;
        TC_CHECK(smem <= 227 * 1024, "sequence length %d too long for the attention statistics kernel", N);
        const int which = dt == DT_F16;
        if (which) ensure_dynamic_smem((const void*)attn_lse_mma_kernel<f16>, smem);
        else ensure_dynamic_smem((const void*)attn_lse_mma_kernel<bf16>, smem);
        dim3 grid((unsigned)(S * H), (unsigned)ntiles);
        if (which) launch_pdl(attn_lse_mma_kernel<f16>, grid, nwarps * 32, smem, stream, (const f16*)qkv, lse, N, H, npad);
        else launch_pdl(attn_lse_mma_kernel<bf16>, grid, nwarps * 32, smem, stream, (const bf16*)qkv, lse, N, H, npad);
    } else {
        dim3 grid((unsigned)ceil_div(N, 4), (unsigned)(S * H));
        launch_pdl(attn_lse_simt_kernel, grid, 128, 0, stream, (const float*)qkv, lse, N, H);
    }
    TC_LAUNCH_CHECK();
}

void rollout_step(const void* qkv, const float* lse, const float* r_in, float* r_out, int dt, int S, int N, int H, bool last,
                  cudaStream_t stream) {
    if (S == 0) return;
    TC_CHECK(N >= 2 && H >= 1 && lse && r_out, "bad rollout_step arguments");
    const int skip = last ? 1 : 0;
    static const int impl = getenv("TAPCLIP_ROLLOUT_IMPL") ? atoi(getenv("TAPCLIP_ROLLOUT_IMPL")) : 0;   // 0 auto, 1 mma.sync, 2 tcgen05
    // tcgen05 kernel when its (image, key-tile group) CTAs fill at least ~2/3 of the SMs
    if (impl != 1 && rollout_step_tc_supported(dt, N) && (impl == 2 || rollout_step_tc_ctas(S, N) >= 96)) {
        rollout_step_tc(qkv, lse, r_in, r_out, dt, S, N, H, last, stream);
        return;
    }
    if (dt == DT_BF16 || dt == DT_F16) {
        // 32-key blocks split evenly over the fewest CTAs of at most 8 warps
        const int nkb = (int)ceil_div(N, RS_KEYS), ntiles = (int)ceil_div(nkb, 8), nwarps = (int)ceil_div(nkb, ntiles);
        const int npad = (int)round_up(N, QC);
        const size_t smem = (size_t)(nwarps * RS_KEYS + npad) * 128 + (size_t)2 * npad * sizeof(float);
        TC_CHECK(smem <= 227 * 1024, "sequence length %d too long for the rollout kernel", N);
        const int which = dt == DT_F16;
        if (which) ensure_dynamic_smem((const void*)rollout_step_mma_kernel<f16>, smem);
        else ensure_dynamic_smem((const void*)rollout_step_mma_kernel<bf16>, smem);
        dim3 grid((unsigned)S, (unsigned)ntiles);
        if (which) launch_pdl(rollout_step_mma_kernel<f16>, grid, nwarps * 32, smem, stream, (const f16*)qkv, lse, r_in, r_out, N, H, npad, skip);
        else launch_pdl(rollout_step_mma_kernel<bf16>, grid, nwarps * 32, smem, stream, (const bf16*)qkv, lse, r_in, r_out, N, H, npad, skip);
    } else {
        dim3 grid((unsigned)ceil_div(N, 4), (unsigned)S);
        launch_pdl(rollout_step_simt_kernel, grid, 128, 0, stream, (const float*)qkv, lse, r_in, r_out, N, H, skip);
    }
    TC_LAUNCH_CHECK();
}

}  // namespace tapclip
