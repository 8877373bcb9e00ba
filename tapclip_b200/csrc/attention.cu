// K2 — fused attention (head dim 64, no mask) with a probe epilogue.
//
// Replaces nn.MultiheadAttention's scaled-dot-product attention inside open_clip's residual blocks
//  AND the forward hook of models/clip_wrapper.py:29-40: the probabilities the
// attribution needs (text: column T-1 for the P ctx query rows; vision extension: the CLS row) are
// emitted from the softmax registers, so the [S,H,N,N] maps the hook would capture never reach HBM.
//
//  * attn_fwd_mma_kernel   bf16, mma.sync m16n8k16 tensor-core path, flash-style online softmax over
//                          32-key steps, K/V/Q staged in XOR-swizzled smem by cp.async.
//  * attn_fwd_simt_kernel  fp32 parity mode (one warp per query row).
//  * attn_bwd_mma_kernel   bf16 dQ,dK,dV for the text tower (N <= 128) on mma.sync, probabilities recomputed.
//  * attn_bwd_kernel       fp32 parity-mode backward (SIMT, shared memory).
#include "kernels.h"
#include <type_traits>
#include <cstdlib>

namespace tapclip {
namespace {

constexpr int DH = 64;
constexpr int KC = 32;

// ------------------------------------------------------------------------------------------------
// bf16 tensor-core forward
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
attn_fwd_mma_kernel(const T* __restrict__ qkv, T* __restrict__ out, int N, int H, int npad, float scale_log2,
                    int probe_mode, float* __restrict__ probe_out, int probe_P, int64_t probe_seq_stride, int causal) {
    pdl_wait_and_trigger();
    extern __shared__ __align__(128) uint8_t smem[];
    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x / H, h = blockIdx.x % H;
    const int d = H * DH;
    const int qrows = nwarps * 16;
    const int q0 = blockIdx.y * qrows;
    uint8_t* Qs = smem;
    uint8_t* Ks = Qs + qrows * 128;
    uint8_t* Vs = Ks + npad * 128;

    const T* base = qkv + (int64_t)s * N * 3 * d + h * DH;
    for (int idx = threadIdx.x; idx < qrows * 8; idx += blockDim.x) {
        const int row = idx >> 3, ch = idx & 7, grow = q0 + row;
        const bool ok = grow < N;
        cp_async_16(smem_u32(Qs + row * 128 + ((ch ^ (row & 7)) << 4)), base + (int64_t)(ok ? grow : 0) * 3 * d + ch * 8, ok);
    }
    for (int idx = threadIdx.x; idx < npad * 8; idx += blockDim.x) {
        const int row = idx >> 3, ch = idx & 7;
        const bool ok = row < N;
        const T* src = base + (int64_t)(ok ? row : 0) * 3 * d + ch * 8;
        const uint32_t off = row * 128 + ((ch ^ (row & 7)) << 4);
        cp_async_16(smem_u32(Ks + off), src + d, ok);
        cp_async_16(smem_u32(Vs + off), src + 2 * d, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    const int r0 = q0 + warp * 16;          // first query row of this warp
    if (r0 >= N) return;
    const int g = lane >> 2, tq = lane & 3, mat = lane >> 3, l7 = lane & 7;

    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int row = warp * 16 + (mat & 1) * 8 + l7, ch = ks * 2 + (mat >> 1);
        ldmatrix_x4(qf[ks], smem_u32(Qs + row * 128 + ((ch ^ (row & 7)) << 4)));
    }

    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float ps0 = 0.f, ps1 = 0.f;             // probed (scaled, log2-domain) score of key N-1 for rows g / g+8
    const bool cls_any = (probe_mode == PROBE_CLS_ROW) && r0 == 0;            // warp-uniform
    const bool cls_probe = cls_any && g == 0;
    float* cls_out = probe_out ? probe_out + (int64_t)s * probe_seq_stride + (int64_t)h * N : nullptr;

    // causal: keys beyond this warp's last query row are all masked, so the key loop stops at the diagonal block
    const int kend = causal ? min(npad, ((r0 + 16 + KC - 1) / KC) * KC) : npad;
    for (int kc = 0; kc < kend; kc += KC) {
        float sc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int nbp = 0; nbp < 2; ++nbp) {
                const int key = kc + (nbp * 2 + (mat >> 1)) * 8 + l7, ch = ks * 2 + (mat & 1);
                uint32_t b[4];
                ldmatrix_x4(b, smem_u32(Ks + key * 128 + ((ch ^ (key & 7)) << 4)));
                mma_16816<T>(sc[nbp * 2], qf[ks], b[0], b[1]);
                mma_16816<T>(sc[nbp * 2 + 1], qf[ks], b[2], b[3]);
            }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
        if (kc + KC < N && !cls_any && !(causal && kc + KC > r0)) {
            // interior chunk (warp-uniform): no masking, no probe bookkeeping
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                sc[nb][0] *= scale_log2; sc[nb][1] *= scale_log2; sc[nb][2] *= scale_log2; sc[nb][3] *= scale_log2;
                mx0 = fmaxf(mx0, fmaxf(sc[nb][0], sc[nb][1]));
                mx1 = fmaxf(mx1, fmaxf(sc[nb][2], sc[nb][3]));
            }
        } else {
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int key = kc + nb * 8 + tq * 2 + e;
                    float v0 = sc[nb][e] * scale_log2, v1 = sc[nb][e + 2] * scale_log2;
                    if (key >= N) { v0 = -INFINITY; v1 = -INFINITY; }
                    if (causal) { if (key > r0 + g) v0 = -INFINITY; if (key > r0 + g + 8) v1 = -INFINITY; }
                    if (key == N - 1) { ps0 = v0; ps1 = v1; }
                    if (cls_probe && key < N) cls_out[key] = v0;
                    sc[nb][e] = v0; sc[nb][e + 2] = v1;
                    mx0 = fmaxf(mx0, v0); mx1 = fmaxf(mx1, v1);
                }
            }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float a0 = exp2f(m0 - mn0), a1 = exp2f(m1 - mn1);
        m0 = mn0; m1 = mn1;
        l0 *= a0; l1 *= a1;
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i][0] *= a0; o[i][1] *= a0; o[i][2] *= a1; o[i][3] *= a1; }
        uint32_t pa[2][4];
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            const float p0 = exp2f(sc[nb][0] - mn0), p1 = exp2f(sc[nb][1] - mn0);
            const float p2 = exp2f(sc[nb][2] - mn1), p3 = exp2f(sc[nb][3] - mn1);
            l0 += p0 + p1; l1 += p2 + p3;
            pa[nb >> 1][(nb & 1) * 2 + 0] = pack2<T>(p0, p1);
            pa[nb >> 1][(nb & 1) * 2 + 1] = pack2<T>(p2, p3);
        }
#pragma unroll
        for (int ks2 = 0; ks2 < 2; ++ks2) {
#pragma unroll
            for (int dbp = 0; dbp < 4; ++dbp) {
                const int key = kc + ks2 * 16 + (mat & 1) * 8 + l7, ch = dbp * 2 + (mat >> 1);
                uint32_t b[4];
                ldmatrix_x4_trans(b, smem_u32(Vs + key * 128 + ((ch ^ (key & 7)) << 4)));
                mma_16816<T>(o[dbp * 2], pa[ks2], b[0], b[1]);
                mma_16816<T>(o[dbp * 2 + 1], pa[ks2], b[2], b[3]);
            }
        }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;

    // ---- probe epilogue -------------------------------------------------------------------------
    if (probe_mode == PROBE_TEXT_COL) {
        if (tq == (((N - 1) & 7) >> 1)) {
            // ps0/ps1 were captured from element e = (N-1)&1 of this thread's column pair
            const int ra = r0 + g, rb = r0 + g + 8;
            float* po = probe_out + ((int64_t)s * H + h) * probe_P;
            if (ra < probe_P) po[ra] = exp2f(ps0 - m0) * inv0;
            if (rb < probe_P) po[rb] = exp2f(ps1 - m1) * inv1;
        }
    } else if (cls_probe) {
        for (int k8 = 0; k8 < N; k8 += 8) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int key = k8 + tq * 2 + e;
                if (key < N) cls_out[key] = exp2f(cls_out[key] - m0) * inv0;      // own earlier writes
            }
        }
    }

    // ---- output: registers -> this warp's (now free) Q rows in smem -> 16-byte global stores ------
    __syncwarp();
    uint8_t* Ws = Qs + warp * 16 * 128;
#pragma unroll
    for (int db = 0; db < 8; ++db) {
        *reinterpret_cast<uint32_t*>(Ws + g * 128 + ((db ^ (g & 7)) << 4) + tq * 4) = pack2<T>(o[db][0] * inv0, o[db][1] * inv0);
        *reinterpret_cast<uint32_t*>(Ws + (g + 8) * 128 + ((db ^ ((g + 8) & 7)) << 4) + tq * 4) = pack2<T>(o[db][2] * inv1, o[db][3] * inv1);
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int idx = it * 32 + lane, row = idx >> 3, ch = idx & 7;
        if (r0 + row < N) {
            const uint4 v = *reinterpret_cast<const uint4*>(Ws + row * 128 + ((ch ^ (row & 7)) << 4));
            *reinterpret_cast<uint4*>(out + ((int64_t)s * N + r0 + row) * d + h * DH + ch * 8) = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fp32 parity-mode forward: one warp per query row, lanes own keys for the scores and dims for P*V
// ------------------------------------------------------------------------------------------------
constexpr int KMAX = 19;   // keys per lane -> N <= 608

template <typename T>
__global__ void __launch_bounds__(128)
attn_fwd_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, int N, int H, float scale, int probe_mode,
                     float* __restrict__ probe_out, int probe_P, int64_t probe_seq_stride, int causal) {
    pdl_wait_and_trigger();
    __shared__ float qs[4][DH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.y / H, h = blockIdx.y % H;
    const int row = blockIdx.x * 4 + warp;
    const int d = H * DH;
    if (row >= N) return;
    const T* base = qkv + (int64_t)s * N * 3 * d + h * DH;
    qs[warp][lane] = to_f32<T>(base[(int64_t)row * 3 * d + lane]);
    qs[warp][lane + 32] = to_f32<T>(base[(int64_t)row * 3 * d + lane + 32]);
    __syncwarp();
    float pr[KMAX];
    float mx = -INFINITY;
#pragma unroll
    for (int kk = 0; kk < KMAX; ++kk) {
        const int key = kk * 32 + lane;
        float v = -INFINITY;
        if (key < N && !(causal && key > row)) {
            const T* kp = base + (int64_t)key * 3 * d + d;
            float acc = 0.f;
#pragma unroll 8
            for (int j = 0; j < DH; ++j) acc = fmaf(qs[warp][j], to_f32<T>(kp[j]), acc);
            v = acc * scale;
        }
        pr[kk] = v;
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int kk = 0; kk < KMAX; ++kk) {
        const float e = (kk * 32 + lane < N && !(causal && kk * 32 + lane > row)) ? expf(pr[kk] - mx) : 0.f;
        pr[kk] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int kk = 0; kk < KMAX; ++kk) {
        pr[kk] *= inv;
        if (kk * 32 < N) {
            for (int src = 0; src < 32; ++src) {
                const float p = __shfl_sync(0xffffffffu, pr[kk], src);
                const int key = kk * 32 + src;
                if (key < N) {
                    const T* vp = base + (int64_t)key * 3 * d + 2 * d;
                    o0 = fmaf(p, to_f32<T>(vp[lane]), o0);
                    o1 = fmaf(p, to_f32<T>(vp[lane + 32]), o1);
                }
            }
        }
    }
    T* op = out + ((int64_t)s * N + row) * d + h * DH;
    op[lane] = from_f32<T>(o0);
    op[lane + 32] = from_f32<T>(o1);
    if (probe_mode == PROBE_TEXT_COL && row < probe_P) {
#pragma unroll
        for (int kk = 0; kk < KMAX; ++kk)
            if (kk * 32 + lane == N - 1) probe_out[((int64_t)s * H + h) * probe_P + row] = pr[kk];
    } else if (probe_mode == PROBE_CLS_ROW && row == 0) {
#pragma unroll
        for (int kk = 0; kk < KMAX; ++kk)
            if (kk * 32 + lane < N) probe_out[(int64_t)s * probe_seq_stride + (int64_t)h * N + kk * 32 + lane] = pr[kk];
    }
}

// ------------------------------------------------------------------------------------------------
// backward, SIMT form (fp32 parity mode; 16-bit modes only for 128 < N <= 141, where the tensor-core kernel below does not
// apply): one CTA per (sequence, head), fp32 math in shared memory, TQ = type of the saved qkv, TG = type of the gradients
//   P = softmax(scale*QK^T); dV = P^T dO; dP = dO V^T; dS = P o (dP - rowsum(P o dP)) * scale;
//   dQ = dS K; dK = dS^T Q
// ------------------------------------------------------------------------------------------------
constexpr int LDH = DH + 1;

constexpr int BWD_KEYS_PER_LANE = 5;          // dS pass: one warp per query row, up to 160 keys
template <typename TQ, typename TG>
__global__ void __launch_bounds__(256)
attn_bwd_kernel(const TQ* __restrict__ qkv, const TG* __restrict__ d_out, TG* __restrict__ dqkv, int N, int H, float scale) {
    pdl_wait_and_trigger();
    extern __shared__ __align__(16) float sm[];
    float* Q = sm;
    float* K = Q + N * LDH;
    float* V = K + N * LDH;
    float* dO = V + N * LDH;
    float* Pm = dO + N * LDH;            // [N][N+1]
    const int LDP = N + 1;
    const int s = blockIdx.x / H, h = blockIdx.x % H;
    const int d = H * DH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const TQ* base = qkv + (int64_t)s * N * 3 * d + h * DH;
    for (int i = tid; i < N * DH; i += blockDim.x) {
        const int r = i / DH, c = i % DH;
        Q[r * LDH + c] = to_f32<TQ>(base[(int64_t)r * 3 * d + c]);
        K[r * LDH + c] = to_f32<TQ>(base[(int64_t)r * 3 * d + d + c]);
        V[r * LDH + c] = to_f32<TQ>(base[(int64_t)r * 3 * d + 2 * d + c]);
        dO[r * LDH + c] = to_f32<TG>(d_out[((int64_t)s * N + r) * d + h * DH + c]);
    }
    __syncthreads();
    // scores
    for (int i = tid; i < N * N; i += blockDim.x) {
        const int r = i / N, c = i % N;
        float acc = 0.f;
#pragma unroll 16
        for (int j = 0; j < DH; ++j) acc = fmaf(Q[r * LDH + j], K[c * LDH + j], acc);
        Pm[r * LDP + c] = acc * scale;
    }
    __syncthreads();
    // row softmax
    for (int r = warp; r < N; r += nwarps) {
        float mx = -INFINITY;
        for (int c = lane; c < N; c += 32) mx = fmaxf(mx, Pm[r * LDP + c]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int c = lane; c < N; c += 32) { const float e = expf(Pm[r * LDP + c] - mx); Pm[r * LDP + c] = e; sum += e; }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        for (int c = lane; c < N; c += 32) Pm[r * LDP + c] *= inv;
    }
    __syncthreads();
    // dV[j][c] = sum_i P[i][j] dO[i][c]
    TG* dbase = dqkv + (int64_t)s * N * 3 * d + h * DH;
    for (int i = tid; i < N * DH; i += blockDim.x) {
        const int j = i / DH, c = i % DH;
        float acc = 0.f;
        for (int r = 0; r < N; ++r) acc = fmaf(Pm[r * LDP + j], dO[r * LDH + c], acc);
        dbase[(int64_t)j * 3 * d + 2 * d + c] = from_f32<TG>(acc);
    }
    __syncthreads();
    // dS in place of P (one warp per row; up to BWD_KEYS_PER_LANE keys per lane)
    for (int r = warp; r < N; r += nwarps) {
        float dp[BWD_KEYS_PER_LANE], pv[BWD_KEYS_PER_LANE];
        float dsum = 0.f;
#pragma unroll
        for (int u = 0; u < BWD_KEYS_PER_LANE; ++u) {
            const int c = lane + 32 * u;
            dp[u] = 0.f; pv[u] = 0.f;
            if (c < N) {
                float acc = 0.f;
#pragma unroll 16
                for (int j = 0; j < DH; ++j) acc = fmaf(dO[r * LDH + j], V[c * LDH + j], acc);
                dp[u] = acc;
                pv[u] = Pm[r * LDP + c];
                dsum = fmaf(pv[u], acc, dsum);
            }
        }
        dsum = warp_sum(dsum);
#pragma unroll
        for (int u = 0; u < BWD_KEYS_PER_LANE; ++u) {
            const int c = lane + 32 * u;
            if (c < N) Pm[r * LDP + c] = pv[u] * (dp[u] - dsum) * scale;
        }
    }
    __syncthreads();
    // dQ[i][c] = sum_j dS[i][j] K[j][c] ; dK[j][c] = sum_i dS[i][j] Q[i][c]
    for (int i = tid; i < N * DH; i += blockDim.x) {
        const int r = i / DH, c = i % DH;
        float aq = 0.f, ak = 0.f;
        for (int j = 0; j < N; ++j) {
            aq = fmaf(Pm[r * LDP + j], K[j * LDH + c], aq);
            ak = fmaf(Pm[j * LDP + r], Q[j * LDH + c], ak);
        }
        dbase[(int64_t)r * 3 * d + c] = from_f32<TG>(aq);
        dbase[(int64_t)r * 3 * d + d + c] = from_f32<TG>(ak);
    }
}


// ------------------------------------------------------------------------------------------------
// bf16 tensor-core backward (text tower: N <= 128).  One CTA per (sequence, head), one warp per 16-row block.
//   phase 1 (warp owns 16 query rows):  S = Q K^T, P = softmax(scale*S) (whole row in registers),
//            dP = dO V^T, D = rowsum(P o dP), dS = P o (dP - D) * scale, dQ = dS K; P and dS -> smem (bf16)
//   phase 2 (warp owns 16 key rows):    dV = P^T dO, dK = dS^T Q   (A operands via ldmatrix.trans on P / dS)
// ------------------------------------------------------------------------------------------------
template <int NB16, typename TQ>
__global__ void __launch_bounds__(NB16 * 32, NB16 <= 6 ? 2 : 1)     // two CTAs per SM up to N = 96 (the kernel is latency-bound)
attn_bwd_mma_kernel(const TQ* __restrict__ qkv, const bf16* __restrict__ d_out, bf16* __restrict__ dqkv, int N, int H,
                    float scale) {
    pdl_wait_and_trigger();
    constexpr int NP = NB16 * 16;               // padded sequence length
    constexpr int NB8 = NB16 * 2;               // 8-wide key blocks
    constexpr int PLD = NP * 2 + 16;            // P / dS row pitch in bytes (odd multiple of 16 B: conflict-free ldmatrix)
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* Qs = smem;
    uint8_t* Ks = Qs + NP * 128;
    uint8_t* Vs = Ks + NP * 128;
    uint8_t* Os = Vs + NP * 128;                // dO
    uint8_t* Ps = Os + NP * 128;
    uint8_t* Ss = Ps + NP * PLD;                // dS
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x / H, h = blockIdx.x % H;
    const int d = H * DH;
    const TQ* base = qkv + (int64_t)s * N * 3 * d + h * DH;
    const bf16* dobase = d_out + (int64_t)s * N * d + h * DH;
    // NP * 8 sixteen-byte chunks per operand, NB16 * 32 threads: exactly 4 chunks per thread
    if constexpr (std::is_same<TQ, bf16>::value) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int idx = threadIdx.x + it * NB16 * 32;
            const int row = idx >> 3, ch = idx & 7;
            const bool ok = row < N;
            const TQ* src = base + (int64_t)(ok ? row : 0) * 3 * d + ch * 8;
            const uint32_t off = row * 128 + ((ch ^ (row & 7)) << 4);
            cp_async_16(smem_u32(Qs + off), src, ok);
            cp_async_16(smem_u32(Ks + off), src + d, ok);
            cp_async_16(smem_u32(Vs + off), src + 2 * d, ok);
            cp_async_16(smem_u32(Os + off), dobase + (int64_t)(ok ? row : 0) * d + ch * 8, ok);
        }
    } else {
        // saved activations are fp16 (mixed mode) while gradients are bf16: Q/K/V are converted once here.  All twelve
        // global loads of a thread are issued before the first conversion so that their latencies overlap.
        uint4 raw[4][3];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int idx = threadIdx.x + it * NB16 * 32;
            const int row = idx >> 3, ch = idx & 7;
            const bool ok = row < N;
            const TQ* src = base + (int64_t)(ok ? row : 0) * 3 * d + ch * 8;
            cp_async_16(smem_u32(Os + row * 128 + ((ch ^ (row & 7)) << 4)), dobase + (int64_t)(ok ? row : 0) * d + ch * 8, ok);
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                raw[it][part] = make_uint4(0u, 0u, 0u, 0u);
                if (ok) raw[it][part] = __ldg(reinterpret_cast<const uint4*>(src + part * d));
            }
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int idx = threadIdx.x + it * NB16 * 32;
            const int row = idx >> 3, ch = idx & 7;
            const uint32_t off = row * 128 + ((ch ^ (row & 7)) << 4);
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const __half2* hp = reinterpret_cast<const __half2*>(&raw[it][part]);
                uint4 cv;
                uint32_t* cp = reinterpret_cast<uint32_t*>(&cv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(hp[e]);
                    cp[e] = pack_bf16x2(f.x, f.y);
                }
                uint8_t* dst = part == 0 ? Qs : (part == 1 ? Ks : Vs);
                *reinterpret_cast<uint4*>(dst + off) = cv;
            }
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    const int g = lane >> 2, tq = lane & 3, mat = lane >> 3, l7 = lane & 7;
    const int r0 = warp * 16;
    bf16* dbase = dqkv + (int64_t)s * N * 3 * d + h * DH;
    {
        // ---------------- phase 1 ----------------
        uint32_t af[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int row = r0 + (mat & 1) * 8 + l7, ch = ks * 2 + (mat >> 1);
            ldmatrix_x4(af[ks], smem_u32(Qs + row * 128 + ((ch ^ (row & 7)) << 4)));
        }
        float sc[NB8][4];
#pragma unroll
        for (int i = 0; i < NB8; ++i) { sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int nbp = 0; nbp < NB16; ++nbp) {
                const int key = (nbp * 2 + (mat >> 1)) * 8 + l7, ch = ks * 2 + (mat & 1);
                uint32_t b[4];
                ldmatrix_x4(b, smem_u32(Ks + key * 128 + ((ch ^ (key & 7)) << 4)));
                mma_bf16_16816(sc[nbp * 2], af[ks], b[0], b[1]);
                mma_bf16_16816(sc[nbp * 2 + 1], af[ks], b[2], b[3]);
            }
        }
        // softmax over the full row (rows g and g+8 of this warp's block)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nb = 0; nb < NB8; ++nb) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int key = nb * 8 + tq * 2 + e;
                float v0 = sc[nb][e] * scale, v1 = sc[nb][e + 2] * scale;
                if (key >= N) { v0 = -INFINITY; v1 = -INFINITY; }
                sc[nb][e] = v0; sc[nb][e + 2] = v1;
                mx0 = fmaxf(mx0, v0); mx1 = fmaxf(mx1, v1);
            }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int nb = 0; nb < NB8; ++nb) {
            sc[nb][0] = __expf(sc[nb][0] - mx0); sc[nb][1] = __expf(sc[nb][1] - mx0);
            sc[nb][2] = __expf(sc[nb][2] - mx1); sc[nb][3] = __expf(sc[nb][3] - mx1);
            l0 += sc[nb][0] + sc[nb][1]; l1 += sc[nb][2] + sc[nb][3];
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
        for (int nb = 0; nb < NB8; ++nb) { sc[nb][0] *= i0; sc[nb][1] *= i0; sc[nb][2] *= i1; sc[nb][3] *= i1; }
        // dP = dO V^T
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int row = r0 + (mat & 1) * 8 + l7, ch = ks * 2 + (mat >> 1);
            ldmatrix_x4(af[ks], smem_u32(Os + row * 128 + ((ch ^ (row & 7)) << 4)));
        }
        float dp[NB8][4];
#pragma unroll
        for (int i = 0; i < NB8; ++i) { dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int nbp = 0; nbp < NB16; ++nbp) {
                const int key = (nbp * 2 + (mat >> 1)) * 8 + l7, ch = ks * 2 + (mat & 1);
                uint32_t b[4];
                ldmatrix_x4(b, smem_u32(Vs + key * 128 + ((ch ^ (key & 7)) << 4)));
                mma_bf16_16816(dp[nbp * 2], af[ks], b[0], b[1]);
                mma_bf16_16816(dp[nbp * 2 + 1], af[ks], b[2], b[3]);
            }
        }
        float D0 = 0.f, D1 = 0.f;
#pragma unroll
        for (int nb = 0; nb < NB8; ++nb) {
            D0 += sc[nb][0] * dp[nb][0] + sc[nb][1] * dp[nb][1];
            D1 += sc[nb][2] * dp[nb][2] + sc[nb][3] * dp[nb][3];
        }
        D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
        D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
        // P, dS -> smem (bf16) and dS -> A fragments for dQ = dS K
        uint32_t da[NB16][4];
#pragma unroll
        for (int nb = 0; nb < NB8; ++nb) {
            const float s0 = sc[nb][0] * (dp[nb][0] - D0) * scale, s1 = sc[nb][1] * (dp[nb][1] - D0) * scale;
            const float s2 = sc[nb][2] * (dp[nb][2] - D1) * scale, s3 = sc[nb][3] * (dp[nb][3] - D1) * scale;
            const uint32_t plo = pack_bf16x2(sc[nb][0], sc[nb][1]), phi = pack_bf16x2(sc[nb][2], sc[nb][3]);
            const uint32_t slo = pack_bf16x2(s0, s1), shi = pack_bf16x2(s2, s3);
            const uint32_t coff = (nb * 8 + tq * 2) * 2;
            *reinterpret_cast<uint32_t*>(Ps + (r0 + g) * PLD + coff) = plo;
            *reinterpret_cast<uint32_t*>(Ps + (r0 + g + 8) * PLD + coff) = phi;
            *reinterpret_cast<uint32_t*>(Ss + (r0 + g) * PLD + coff) = slo;
            *reinterpret_cast<uint32_t*>(Ss + (r0 + g + 8) * PLD + coff) = shi;
            da[nb >> 1][(nb & 1) * 2 + 0] = slo;
            da[nb >> 1][(nb & 1) * 2 + 1] = shi;
        }
        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll
        for (int ks2 = 0; ks2 < NB16; ++ks2) {
#pragma unroll
            for (int dbp = 0; dbp < 4; ++dbp) {
                const int key = ks2 * 16 + (mat & 1) * 8 + l7, ch = dbp * 2 + (mat >> 1);
                uint32_t b[4];
                ldmatrix_x4_trans(b, smem_u32(Ks + key * 128 + ((ch ^ (key & 7)) << 4)));
                mma_bf16_16816(o[dbp * 2], da[ks2], b[0], b[1]);
                mma_bf16_16816(o[dbp * 2 + 1], da[ks2], b[2], b[3]);
            }
        }
#pragma unroll
        for (int db = 0; db < 8; ++db) {
            if (r0 + g < N) *reinterpret_cast<uint32_t*>(dbase + (int64_t)(r0 + g) * 3 * d + db * 8 + tq * 2) = pack_bf16x2(o[db][0], o[db][1]);
            if (r0 + g + 8 < N) *reinterpret_cast<uint32_t*>(dbase + (int64_t)(r0 + g + 8) * 3 * d + db * 8 + tq * 2) = pack_bf16x2(o[db][2], o[db][3]);
        }
    }
    __syncthreads();
    // ---------------- phase 2: this warp's 16 KEY rows j0..j0+15 ----------------
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        const uint8_t* Am = which == 0 ? Ps : Ss;       // [i][j] ; A[m=j][k=i] via ldmatrix.trans
        const uint8_t* Bm = which == 0 ? Os : Qs;       // dO (for dV) or Q (for dK), B[k=i][n=dim] via ldmatrix.trans
        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < NB16; ++ks) {
            uint32_t a[4];
            {
                const int i = ks * 16 + (mat >> 1) * 8 + l7, j = r0 + (mat & 1) * 8;
                ldmatrix_x4_trans(a, smem_u32(Am + i * PLD + j * 2));
            }
#pragma unroll
            for (int dbp = 0; dbp < 4; ++dbp) {
                const int i = ks * 16 + (mat & 1) * 8 + l7, ch = dbp * 2 + (mat >> 1);
                uint32_t b[4];
                ldmatrix_x4_trans(b, smem_u32(Bm + i * 128 + ((ch ^ (i & 7)) << 4)));
                mma_bf16_16816(o[dbp * 2], a, b[0], b[1]);
                mma_bf16_16816(o[dbp * 2 + 1], a, b[2], b[3]);
            }
        }
        const int coloff = which == 0 ? 2 * d : d;      // dV or dK column block of dqkv
#pragma unroll
        for (int db = 0; db < 8; ++db) {
            if (r0 + g < N) *reinterpret_cast<uint32_t*>(dbase + (int64_t)(r0 + g) * 3 * d + coloff + db * 8 + tq * 2) = pack_bf16x2(o[db][0], o[db][1]);
            if (r0 + g + 8 < N) *reinterpret_cast<uint32_t*>(dbase + (int64_t)(r0 + g + 8) * 3 * d + coloff + db * 8 + tq * 2) = pack_bf16x2(o[db][2], o[db][3]);
        }
    }
}

template <int NB16, typename TQ>
void launch_attn_bwd_mma(const TQ* qkv, const bf16* d_out, bf16* dqkv, int S, int N, int H, cudaStream_t stream) {
    constexpr int NP = NB16 * 16;
    const size_t smem = (size_t)4 * NP * 128 + (size_t)2 * NP * (NP * 2 + 16);
    ensure_dynamic_smem((const void*)attn_bwd_mma_kernel<NB16, TQ>, smem);
    launch_pdl(attn_bwd_mma_kernel<NB16, TQ>, S * H, NB16 * 32, smem, stream, qkv, d_out, dqkv, N, H, 0.125f);
}

template <typename TQ>
void dispatch_attn_bwd_mma(const TQ* q, const bf16* g, bf16* o, int S, int N, int H, cudaStream_t stream) {
    switch ((N + 15) / 16) {
        case 1: launch_attn_bwd_mma<1, TQ>(q, g, o, S, N, H, stream); break;
        case 2: launch_attn_bwd_mma<2, TQ>(q, g, o, S, N, H, stream); break;
        case 3: launch_attn_bwd_mma<3, TQ>(q, g, o, S, N, H, stream); break;
        case 4: launch_attn_bwd_mma<4, TQ>(q, g, o, S, N, H, stream); break;
        case 5: launch_attn_bwd_mma<5, TQ>(q, g, o, S, N, H, stream); break;
        case 6: launch_attn_bwd_mma<6, TQ>(q, g, o, S, N, H, stream); break;
        case 7: launch_attn_bwd_mma<7, TQ>(q, g, o, S, N, H, stream); break;
        default: launch_attn_bwd_mma<8, TQ>(q, g, o, S, N, H, stream); break;
    }
}

}  // namespace

void attention_fwd(const void* qkv, void* out, int dt, int S, int N, int H, const AttnProbe& probe, cudaStream_t stream) {
    if (S == 0) return;
    TC_CHECK(N >= 1 && H >= 1, "bad attention shape");
    if (probe.mode == PROBE_TEXT_COL) TC_CHECK(probe.out && probe.P >= 1 && probe.P <= N, "bad text probe");
    if (probe.mode == PROBE_CLS_ROW) TC_CHECK(probe.out && probe.seq_stride >= (int64_t)H * N, "bad CLS probe");
    static const int impl = getenv("TAPCLIP_ATTN_IMPL") ? atoi(getenv("TAPCLIP_ATTN_IMPL")) : 0;   // 0 auto, 1 mma.sync, 2 tcgen05
    // auto: the persistent tcgen05 kernel for N <= 208 when there are enough (sequence, head, q-tile) items to keep its
    // two softmax groups per SM busy (ViT-B/16 at B=128: 3072 items, 42 us vs 99 us; the C=65 text tower, 520 items at
    // N=93: 9.7 us vs 12.4 us); the mma.sync flash kernel for 208 < N <= 256, causal launches and smaller problems
    if (!probe.causal && (impl == 2 || (impl == 0 && attention_fwd_tc_supported(dt, N) && (N <= 208 || N > 256) && (int64_t)S * H * ((N + 127) / 128) >= 512))) {   // (threshold on the full problem: a live_q_rows launch keeps the kernel choice)
        if (!attention_fwd_tc(qkv, out, dt, S, N, H, probe, stream) && probe.lse_out) attention_lse(qkv, probe.lse_out, dt, S, N, H, stream);
        return;
    }
    if (dt == DT_BF16 || dt == DT_F16) {
        const int npad = (int)round_up(N, KC);
        const int nrb = (int)ceil_div(N, 16);
        int nwarps = nrb <= 8 ? nrb : (int)ceil_div(nrb, ceil_div(nrb, 8));
        const int nz = (int)ceil_div(nrb, nwarps);
        const size_t smem = (size_t)(nwarps * 16 + 2 * npad) * 128;
        TC_CHECK(smem <= 227 * 1024, "sequence length %d too long for the attention kernel", N);
        const int which = dt == DT_F16 ? 1 : 0;
        if (which) ensure_dynamic_smem((const void*)attn_fwd_mma_kernel<f16>, smem);
        else ensure_dynamic_smem((const void*)attn_fwd_mma_kernel<bf16>, smem);
        dim3 grid((unsigned)(S * H), (unsigned)nz);
        const float sl2 = 0.125f * 1.4426950408889634f;
        if (which) launch_pdl(attn_fwd_mma_kernel<f16>, grid, nwarps * 32, smem, stream, (const f16*)qkv, (f16*)out, N, H, npad, sl2, probe.mode, probe.out, probe.P, probe.seq_stride, (int)probe.causal);
        else launch_pdl(attn_fwd_mma_kernel<bf16>, grid, nwarps * 32, smem, stream, (const bf16*)qkv, (bf16*)out, N, H, npad, sl2, probe.mode, probe.out, probe.P, probe.seq_stride, (int)probe.causal);
    } else {
        TC_CHECK(N <= KMAX * 32, "sequence length %d too long for the fp32 attention kernel", N);
        dim3 grid((unsigned)ceil_div(N, 4), (unsigned)(S * H));
        launch_pdl(attn_fwd_simt_kernel<float>, grid, 128, 0, stream, (const float*)qkv, (float*)out, N, H, 0.125f, probe.mode, probe.out,
                   probe.P, probe.seq_stride, (int)probe.causal);
    }
    TC_LAUNCH_CHECK();
    if (probe.lse_out) attention_lse(qkv, probe.lse_out, dt, S, N, H, stream);       // these kernels do not emit the statistics
}

void attention_bwd(const void* qkv, int qkv_dt, const void* d_out, void* dqkv, int grad_dt, int S, int N, int H, cudaStream_t stream) {
    if (S == 0) return;
    const size_t smem_simt = ((size_t)4 * N * LDH + (size_t)N * (N + 1)) * sizeof(float);
    TC_CHECK(N <= 32 * BWD_KEYS_PER_LANE && smem_simt <= 227 * 1024, "attention backward supports sequence length <= 141 (got %d)", N);
    if (grad_dt == DT_BF16 && N > 128) {
        // beyond the tensor-core kernel's 128 tokens (prompt_len 52..64): the SIMT form, ~20x slower per layer, still exact math
        ensure_dynamic_smem(qkv_dt == DT_F16 ? (const void*)attn_bwd_kernel<f16, bf16> : (const void*)attn_bwd_kernel<bf16, bf16>, smem_simt);
        if (qkv_dt == DT_F16) launch_pdl(attn_bwd_kernel<f16, bf16>, S * H, 256, smem_simt, stream, (const f16*)qkv, (const bf16*)d_out, (bf16*)dqkv, N, H, 0.125f);
        else if (qkv_dt == DT_BF16) launch_pdl(attn_bwd_kernel<bf16, bf16>, S * H, 256, smem_simt, stream, (const bf16*)qkv, (const bf16*)d_out, (bf16*)dqkv, N, H, 0.125f);
        else TC_CHECK(false, "unsupported saved-activation dtype %d", qkv_dt);
    } else if (grad_dt == DT_BF16) {
        if (qkv_dt == DT_BF16) dispatch_attn_bwd_mma<bf16>((const bf16*)qkv, (const bf16*)d_out, (bf16*)dqkv, S, N, H, stream);
        else if (qkv_dt == DT_F16) dispatch_attn_bwd_mma<f16>((const f16*)qkv, (const bf16*)d_out, (bf16*)dqkv, S, N, H, stream);
        else TC_CHECK(false, "unsupported saved-activation dtype %d", qkv_dt);
    } else {
        TC_CHECK(grad_dt == DT_F32 && qkv_dt == DT_F32, "unsupported dtype combination for attention backward");
        ensure_dynamic_smem((const void*)attn_bwd_kernel<float, float>, smem_simt);
        launch_pdl(attn_bwd_kernel<float, float>, S * H, 256, smem_simt, stream, (const float*)qkv, (const float*)d_out, (float*)dqkv, N, H, 0.125f);
    }
    TC_LAUNCH_CHECK();
}

}  // namespace tapclip
