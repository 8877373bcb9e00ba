// K1-LN — residual GEMM with a fused "pre-LN" epilogue for sm_100a:
//
//     x[M,N] += A[M,K] * W[N,K]^T + bias            (fp32 residual stream, updated in place)
//     ln_out[M,N] = LayerNorm(x_new; gamma, beta)   (16-bit operand of the NEXT GEMM)        x_copy[M,N] = x_new (optional)
//
// Replaces `x = x + attn(...)` / `x = x + mlp(...)` followed by `ln_2(x)` / the next block's `ln_1(x)` of open_clip's
// ResidualAttentionBlock: one launch instead of two, and the fp32 residual stream is not re-read by a
// LayerNorm kernel.  LayerNorm needs whole rows, a tcgen05 accumulator tile is 128 x 256: the N / 256 CTAs that own the
// column slices of one 128-row block form a thread-block CLUSTER and exchange their per-row partial sums (sum, sum of
// squares) through distributed shared memory, so every CTA keeps one 128 x 256 tile like the plain GEMM (same parallelism).
//
// Per CTA (320 threads): warp 0 TMA producer, warp 1 tcgen05.mma issuer (warp-uniform, elect.sync), warps 2..9 epilogue
// (two per TMEM lane quarter, alternating 32-column chunks).  Epilogue of a tile:
//   pass A  TMEM -> registers, + bias + x_old (coalesced loads one chunk ahead, transposed through smem), partial row sums, x_new written back
//           INTO TMEM (the accumulator doubles as the row buffer) and, through a swizzled smem transpose, to x / x_copy
//   exchange  the two warps of a lane quarter combine through smem; one thread per row sends the CTA's partial sums to every
//           peer CTA (st.shared::cluster) and arrives on the peer's mbarrier (release.cluster); everyone waits on its own
//   pass B  TMEM -> registers, (x - mean) * rstd * gamma + beta -> 16 bit -> smem transpose -> coalesced stores to ln_out
#include "gemm.h"
#include <type_traits>

namespace tapclip {
namespace {

constexpr int BLOCK_M = 128, BLOCK_N = 256, BLOCK_K = 64, UMMA_K = 16;
constexpr int EPI_WARPS = 8, NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int STAGES = 3;
constexpr int STAGE_A_BYTES = BLOCK_M * BLOCK_K * 2, STAGE_B_BYTES = BLOCK_N * BLOCK_K * 2, STAGE_BYTES = STAGE_A_BYTES + STAGE_B_BYTES;
constexpr int EPI_BUF_BYTES = 32 * 128, EPI_BYTES = EPI_WARPS * EPI_BUF_BYTES;
constexpr int MAX_CS = 4;                                         // cluster size = N / 256 (2: D=512, 3: d=768, 4: d=1024)
// smem after the operand ring and the staging buffers
constexpr int OFF_BARS = 0;                                       // full[3], empty[3], tmem_full[2], tmem_empty[2], stats[2], tmem slot
constexpr int OFF_VEC = 256;                                      // bias, gamma, beta of this CTA's 256 columns
constexpr int OFF_PART = OFF_VEC + 3 * BLOCK_N * 4;               // float2 part[2 buf][2 column halves][128 rows]
constexpr int OFF_PEER = OFF_PART + 2 * 2 * 128 * 8;              // float2 peer[2 buf][MAX_CS ranks][128 rows]
constexpr int TAIL_BYTES = OFF_PEER + 2 * MAX_CS * 128 * 8;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + TAIL_BYTES + 1024;

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                                      // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool f16) {
    const uint32_t fmt = f16 ? 0u : 1u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
          "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]),
          "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]),
          "f"(v[30]), "f"(v[31])
        : "memory");
}
// wait on a barrier whose arrivals come from other CTAs of the cluster (their st.shared::cluster data must be visible)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 4000000000LL) { printf("tapclip: cluster barrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}

template <int CS, bool F16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_resid_ln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, float* x, int ldx,
                     const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta, void* ln_out,
                     float* x_copy, int M, int N, int K) {
    using T16 = typename std::conditional<F16, f16, bf16>::type;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * STAGE_A_BYTES;
    uint8_t* smem_epi = smem + STAGES * STAGE_BYTES;
    uint8_t* tail = smem_epi + EPI_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail + OFF_BARS);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* stats_bar = tmem_empty + 2;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(stats_bar + 2);
    float* bias_s = reinterpret_cast<float*>(tail + OFF_VEC);
    float* gamma_s = bias_s + BLOCK_N;
    float* beta_s = gamma_s + BLOCK_N;
    float2* part = reinterpret_cast<float2*>(tail + OFF_PART);          // [buf][half][row]
    float2* peer = reinterpret_cast<float2*>(tail + OFF_PEER);          // [buf][rank][row]

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = (int)blockIdx.x / CS, n_clusters = (int)gridDim.x / CS;
    const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
    const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
    const int n0 = (int)rank * BLOCK_N;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], EPI_WARPS); mbar_init(&stats_bar[i], (CS - 1) * 128); }
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_base_slot, 2 * BLOCK_N);
    tc_fence_before();
    cluster_sync_all();                        // the peers' barriers must be initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_slot, 0);
    pdl_trigger();
    pdl_wait();
    // this CTA's slice of bias / gamma / beta
    for (int i = threadIdx.x; i < BLOCK_N; i += NUM_THREADS) {
        bias_s[i] = bias ? __ldg(bias + n0 + i) : 0.f;
        gamma_s[i] = __ldg(gamma + n0 + i);
        beta_s[i] = __ldg(beta + n0 + i);
    }
    __syncthreads();

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int mb = cluster_id; mb < m_tiles; mb += n_clusters) {
                const int m0 = mb * BLOCK_M;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                    tma_load_2d(smem_a + stage * STAGE_A_BYTES, &tmap_a, kb * BLOCK_K, m0, &full_bar[stage]);
                    tma_load_2d(smem_b + stage * STAGE_B_BYTES, &tmap_b, kb * BLOCK_K, n0, &full_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (warp-uniform, elect.sync) ================================
        constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, F16);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int mb = cluster_id; mb < m_tiles; mb += n_clusters) {
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint64_t adesc = make_smem_desc(smem_u32(smem_a + stage * STAGE_A_BYTES));
                const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + stage * STAGE_B_BYTES));
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_ss_elect(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                umma_commit_elect(&empty_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit_elect(&tmem_full[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ================================ epilogue warps ==============================
        const int q = warp & 3, ew = warp - 2, half = ew >> 2;          // TMEM lane quarter; chunks with (c & 1) == half
        const int row = q * 32 + lane;
        const uint32_t stage_u32 = smem_u32(smem_epi + ew * EPI_BUF_BYTES);
        const int rd_row = lane >> 3, rd_ch = lane & 7;                // fp32 stores: 8 lanes cover one 128-byte row segment
        const int r2_row = lane >> 2, r2_ch = lane & 3;                // 16-bit stores: 4 lanes cover one 64-byte row segment
        const float inv_n = 1.f / (float)N;
        int acc = 0; uint32_t acc_phase = 0;
        int it_local = 0;
        for (int mb = cluster_id; mb < m_tiles; mb += n_clusters, ++it_local) {
            const int m0 = mb * BLOCK_M;
            const int grow = m0 + row;
            const int row0 = m0 + q * 32;
            const int buf = it_local & 1;
            const uint32_t sparity = (uint32_t)((it_local >> 1) & 1);
            // x_old of a chunk = this warp's 32 x 32 block, loaded coalesced (8 lanes per 128-byte row segment); the loads of chunk
            // k+1 are issued while chunk k is processed, the first ones before the accumulator is even ready
            float4 xin[8];
            auto load_old = [&](int c) {
                const int gcol_in = n0 + c * 32 + rd_ch * 4;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + rd_row;
                    xin[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row0 + rr < M && c < BLOCK_N / 32) xin[it] = *reinterpret_cast<const float4*>(x + (int64_t)(row0 + rr) * ldx + gcol_in);
                }
            };
            load_old(half);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
            float s1 = 0.f, s2 = 0.f;
            // ---------------- pass A ----------------
#pragma unroll 1
            for (int c = half; c < BLOCK_N / 32; c += 2) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c * 32, r);
                __syncwarp();                                          // the previous chunk's read-back of the staging buffer is complete
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + rd_row;
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stage_u32 + (uint32_t)rr * 128u + (((uint32_t)rd_ch ^ ((uint32_t)rr & 7u)) << 4)),
                                 "f"(xin[it].x), "f"(xin[it].y), "f"(xin[it].z), "f"(xin[it].w) : "memory");
                }
                __syncwarp();
                load_old(c + 2);                                        // next chunk's old values: in flight during this chunk's work
                float v[32];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4 * j]), "=f"(v[4 * j + 1]), "=f"(v[4 * j + 2]), "=f"(v[4 * j + 3])
                                 : "r"(stage_u32 + (uint32_t)lane * 128u + (((uint32_t)j ^ ((uint32_t)lane & 7u)) << 4)) : "memory");
                tmem_ld_wait();
                if (grow < M) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        v[j] += __uint_as_float(r[j]) + bias_s[c * 32 + j];
                        s1 += v[j];
                        s2 = fmaf(v[j], v[j], s2);
                    }
                }
                tmem_st_32x32(taddr + c * 32, v);                       // the accumulator tile doubles as the row buffer for pass B
                // x_new -> swizzled staging -> coalesced stores to x (and the saved copy)
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stage_u32 + (uint32_t)lane * 128u + (((uint32_t)j ^ ((uint32_t)lane & 7u)) << 4)),
                                 "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
                __syncwarp();
                const int gcol = n0 + c * 32 + rd_ch * 4;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + rd_row;
                    float4 a;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                                 : "r"(stage_u32 + (uint32_t)rr * 128u + (((uint32_t)rd_ch ^ ((uint32_t)rr & 7u)) << 4)) : "memory");
                    if (row0 + rr < M) {
                        *reinterpret_cast<float4*>(x + (int64_t)(row0 + rr) * ldx + gcol) = a;
                        if (x_copy) *reinterpret_cast<float4*>(x_copy + (int64_t)(row0 + rr) * N + gcol) = a;
                    }
                }
            }
            tmem_st_wait();
            // ---------------- row statistics: the two warps of the quarter, then the CTAs of the cluster ----------------
            part[(buf * 2 + half) * 128 + row] = make_float2(s1, s2);
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            {
                const float2 o = part[(buf * 2 + (half ^ 1)) * 128 + row];
                s1 += o.x; s2 += o.y;
            }
            if (half == 0) {
#pragma unroll
                for (int p = 0; p < CS; ++p) {
                    if (p != (int)rank) {
                        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(mapa_u32(&peer[(buf * MAX_CS + (int)rank) * 128 + row], (uint32_t)p)),
                                     "f"(s1), "f"(s2) : "memory");
                        mbar_arrive_cluster(mapa_u32(&stats_bar[buf], (uint32_t)p));
                    }
                }
            }
            mbar_wait_cluster(&stats_bar[buf], sparity);
#pragma unroll
            for (int p = 0; p < CS; ++p) {
                if (p != (int)rank) {
                    const float2 o = peer[(buf * MAX_CS + p) * 128 + row];
                    s1 += o.x; s2 += o.y;
                }
            }
            const float mean = s1 * inv_n;
            const float rstd = rsqrtf(fmaxf(s2 * inv_n - mean * mean, 0.f) + 1e-5f);
            // ---------------- pass B ----------------
#pragma unroll 1
            for (int c = half; c < BLOCK_N / 32; c += 2) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + c * 32, r);
                tmem_ld_wait();
                if (c >= BLOCK_N / 32 - 2) {                            // this warp's last TMEM read of the tile
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) {                           // 32 columns -> 64 bytes of 16-bit values per row
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int col = c * 32 + j * 8 + e * 2;
                        const float y0 = (__uint_as_float(r[j * 8 + e * 2]) - mean) * rstd * gamma_s[col] + beta_s[col];
                        const float y1 = (__uint_as_float(r[j * 8 + e * 2 + 1]) - mean) * rstd * gamma_s[col + 1] + beta_s[col + 1];
                        pk[e] = pack2<T16>(y0, y1);
                    }
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_u32 + (uint32_t)lane * 64u + (((uint32_t)j ^ (((uint32_t)lane >> 1) & 3u)) << 4)),
                                 "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                }
                __syncwarp();
                const int gcol = n0 + c * 32 + r2_ch * 8;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int rr = it * 8 + r2_row;
                    uint4 a;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w)
                                 : "r"(stage_u32 + (uint32_t)rr * 64u + (((uint32_t)r2_ch ^ (((uint32_t)rr >> 1) & 3u)) << 4)) : "memory");
                    if (row0 + rr < M)
                        asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<uint8_t*>(ln_out) + ((int64_t)(row0 + rr) * N + gcol) * 2),
                                     "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w) : "memory");
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    cluster_sync_all();                        // no CTA exits (or frees TMEM) while a peer can still write into its shared memory
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 2 * BLOCK_N);
}

int g_num_sms = 0;

template <int CS, bool F16>
void launch(const GemmLnArgs& g, cudaStream_t stream) {
    auto kern = gemm_resid_ln_kernel<CS, F16>;
    static bool configured = false;
    if (!configured) {
        TC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    if (g_num_sms == 0) {
        int dev;
        TC_CUDA(cudaGetDevice(&dev));
        TC_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    constexpr CUtensorMapDataType DT16 = F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const CUtensorMap& ta = make_tmap(g.a, DT16, 2, g.M, g.K, g.K, BLOCK_M, BLOCK_K);
    const CUtensorMap& tb = make_tmap(g.w, DT16, 2, g.N, g.K, g.K, BLOCK_N, BLOCK_K);
    const int64_t m_tiles = ceil_div(g.M, BLOCK_M);
    // clusters that can be resident at once (a cluster must fit inside one GPC, so this can be below #SM / CS)
    static int max_clusters = 0;
    if (max_clusters == 0) {
        cudaLaunchConfig_t qc = {};
        qc.gridDim = dim3(CS * (g_num_sms / CS)); qc.blockDim = dim3(NUM_THREADS); qc.dynamicSmemBytes = SMEM_BYTES;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = CS; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
        qc.attrs = qa; qc.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &qc) != cudaSuccess || n <= 0) { cudaGetLastError(); n = g_num_sms / CS; }
        max_clusters = std::min(n, g_num_sms / CS);
    }
    const int clusters = (int)std::min<int64_t>(m_tiles, max_clusters);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * clusters);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    TC_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, g.x, (int)g.ldx, g.bias, g.gamma, g.beta, g.ln_out, g.x_copy, (int)g.M, (int)g.N, (int)g.K));
    TC_LAUNCH_CHECK();
}

}  // namespace

bool gemm_resid_ln_supported(int64_t N, int64_t K, int dt) {
    return (dt == DT_BF16 || dt == DT_F16) && N % BLOCK_N == 0 && N / BLOCK_N >= 2 && N / BLOCK_N <= MAX_CS && K % 8 == 0;
}

void gemm_resid_ln(const GemmLnArgs& g, cudaStream_t stream) {
    TC_CHECK(g.M > 0 && gemm_resid_ln_supported(g.N, g.K, g.dt), "fused residual+LayerNorm GEMM: unsupported shape N=%lld K=%lld dt=%d", (long long)g.N, (long long)g.K, g.dt);
    TC_CHECK(g.x && g.gamma && g.beta && g.ln_out, "fused residual+LayerNorm GEMM: null argument");
    TC_CHECK((reinterpret_cast<uintptr_t>(g.x) & 15) == 0 && (g.ldx * 4) % 16 == 0 && (reinterpret_cast<uintptr_t>(g.ln_out) & 15) == 0, "fused residual+LayerNorm GEMM: alignment");
    const bool f16 = g.dt == DT_F16;
    switch ((int)(g.N / BLOCK_N)) {
        case 2: if (f16) launch<2, true>(g, stream); else launch<2, false>(g, stream); break;
        case 3: if (f16) launch<3, true>(g, stream); else launch<3, false>(g, stream); break;
        default: if (f16) launch<4, true>(g, stream); else launch<4, false>(g, stream); break;
    }
}

}  // namespace tapclip
