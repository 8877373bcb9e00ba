// K1 — tcgen05/TMEM bf16 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] * W[N,K]^T + bias).
//
// Replaces the cuBLAS `addmm` calls behind nn.Linear / nn.MultiheadAttention in_proj/out_proj of the
// reference's third-party model (SURVEY.md 2.3 rows k2,k5,k8,k10 + patch-embed + the two projections).
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: 128B-swizzled A (128x64) and W (BLOCK_N x 64) tiles into a 4-stage smem ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16, fp32 accumulate)
//   warps 2..5  epilogue: tcgen05.ld TMEM -> registers -> bias / activation -> swizzled smem -> TMA store
//               (or TMA reduce-add for the fp32 residual stream)
// Accumulators are double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1.  M/N/K tails are handled by TMA zero-fill on loads and clipping on stores.
#include "gemm.h"
#include <type_traits>

namespace tapclip {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;           // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int STAGE_A_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int EPI_BUF_BYTES = 32 * 128;             // 32 rows x 128 B, one TMA-store box
constexpr int EPI_BYTES = 4 * 2 * EPI_BUF_BYTES;    // 4 warps x 2 buffers

template <int BLOCK_N> struct Cfg {
    static constexpr int STAGE_B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = STAGE_A_BYTES + STAGE_B_BYTES;
    static constexpr int STAGES = (BLOCK_N == 256) ? 4 : 6;
    static constexpr int TMEM_COLS = 2 * BLOCK_N;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 256 /*barriers*/ + 1024 /*align slack*/;
};

// UMMA shared-memory descriptor: K-major operand, 128B swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major) = 1
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset = 1024 B between 8-row groups
    d |= (uint64_t)1 << 46;                        // descriptor version 1 (sm_100)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
// UMMA instruction descriptor: (bf16|fp16) x same -> fp32, both operands K-major, M=128, N=BLOCK_N
__host__ __device__ constexpr uint32_t make_idesc(int n, bool f16) {
    const uint32_t fmt = f16 ? 0u : 1u;             // F16F32Format: 0 = F16, 1 = BF16
    return (1u << 4) /*C=f32*/ | (fmt << 7) /*A*/ | (fmt << 10) /*B*/ | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(BLOCK_M >> 4) << 24);
}

template <int BLOCK_N, int EPI, int ACT, bool STORE_PRE, bool F16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_c2,
               const float* __restrict__ bias, int M, int N, int K) {
    using C = Cfg<BLOCK_N>;
    using T16 = typename std::conditional<F16, f16, bf16>::type;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + C::STAGES * STAGE_A_BYTES;
    uint8_t* smem_epi = smem + C::STAGES * C::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + EPI_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + C::STAGES;
    uint64_t* tmem_full = bars + 2 * C::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M, n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        tma_prefetch_desc(&tmap_c);
        if (STORE_PRE) tma_prefetch_desc(&tmap_c2);
        for (int i = 0; i < C::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_base_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * BLOCK_M, n0 = (tile % n_tiles) * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                    tma_load_2d(smem_a + stage * STAGE_A_BYTES, &tmap_a, kb * BLOCK_K, m0, &full_bar[stage]);
                    tma_load_2d(smem_b + stage * C::STAGE_B_BYTES, &tmap_b, kb * BLOCK_K, n0, &full_bar[stage]);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BLOCK_N, F16);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(smem_u32(smem_a + stage * STAGE_A_BYTES));
                    const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + stage * C::STAGE_B_BYTES));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 32 B (= 16 bf16) inside the 128B swizzle row: +2 in the >>4 address field
                        umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);          // frees this smem stage once the MMAs retire
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);                // accumulator ready for the epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ================================ epilogue warps ==============================
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        const int ew = warp - 2;
        uint8_t* stage_buf = smem_epi + ew * 2 * EPI_BUF_BYTES;
        const uint32_t row_off = (uint32_t)lane * 128u;
        const uint32_t sw = (uint32_t)(lane & 7);
        int acc = 0; uint32_t acc_phase = 0;
        int buf = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile / n_tiles) * BLOCK_M, n0 = (tile % n_tiles) * BLOCK_N;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
            const int row0 = m0 + q * 32;

            if constexpr (EPI == EPI_BF16) {
                constexpr int CH = 64;                       // 64 bf16 columns = 128 B per row per store box
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / CH; ++c) {
                    uint32_t r0[32], r1[32];
                    tmem_ld_32x32(taddr + c * CH, r0);
                    tmem_ld_32x32(taddr + c * CH + 32, r1);
                    tmem_ld_wait();
                    if (c == BLOCK_N / CH - 1) {             // all TMEM reads of this tile are done
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                    }
                    const int col0 = n0 + c * CH;
                    float v[64];
#pragma unroll
                    for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
                    if (bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 64; j += 4) {
                            if (col0 + j < N) {
                                float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
                                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                            }
                        }
                    }
                    if (row0 < M && col0 < N) {
                        if constexpr (STORE_PRE) {
                            // both staging buffers per chunk: [0] pre-activation, [1] activated
                            if (lane == 0) tma_store_wait_read<0>();
                            __syncwarp();
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                uint4 p;
                                p.x = pack2<T16>(v[8 * j + 0], v[8 * j + 1]); p.y = pack2<T16>(v[8 * j + 2], v[8 * j + 3]);
                                p.z = pack2<T16>(v[8 * j + 4], v[8 * j + 5]); p.w = pack2<T16>(v[8 * j + 6], v[8 * j + 7]);
                                *reinterpret_cast<uint4*>(stage_buf + row_off + (((uint32_t)j ^ sw) << 4)) = p;
                            }
#pragma unroll
                            for (int j = 0; j < 64; ++j) v[j] = act_fwd_fast<ACT>(v[j]);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                uint4 p;
                                p.x = pack2<T16>(v[8 * j + 0], v[8 * j + 1]); p.y = pack2<T16>(v[8 * j + 2], v[8 * j + 3]);
                                p.z = pack2<T16>(v[8 * j + 4], v[8 * j + 5]); p.w = pack2<T16>(v[8 * j + 6], v[8 * j + 7]);
                                *reinterpret_cast<uint4*>(stage_buf + EPI_BUF_BYTES + row_off + (((uint32_t)j ^ sw) << 4)) = p;
                            }
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_2d(&tmap_c2, stage_buf, col0, row0);
                                tma_store_2d(&tmap_c, stage_buf + EPI_BUF_BYTES, col0, row0);
                                tma_store_commit();
                            }
                        } else {
                            if (lane == 0) tma_store_wait_read<1>();     // buffer used two chunks ago is free
                            __syncwarp();
                            uint8_t* sb = stage_buf + buf * EPI_BUF_BYTES;
#pragma unroll
                            for (int j = 0; j < 64; ++j) v[j] = act_fwd_fast<ACT>(v[j]);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                uint4 p;
                                p.x = pack2<T16>(v[8 * j + 0], v[8 * j + 1]); p.y = pack2<T16>(v[8 * j + 2], v[8 * j + 3]);
                                p.z = pack2<T16>(v[8 * j + 4], v[8 * j + 5]); p.w = pack2<T16>(v[8 * j + 6], v[8 * j + 7]);
                                *reinterpret_cast<uint4*>(sb + row_off + (((uint32_t)j ^ sw) << 4)) = p;
                            }
                            fence_proxy_async_smem();
                            __syncwarp();
                            if (lane == 0) { tma_store_2d(&tmap_c, sb, col0, row0); tma_store_commit(); }
                            buf ^= 1;
                        }
                    }
                }
            } else {
                constexpr int CH = 32;                       // 32 fp32 columns = 128 B per row per store box
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / CH; ++c) {
                    uint32_t r0[32];
                    tmem_ld_32x32(taddr + c * CH, r0);
                    tmem_ld_wait();
                    if (c == BLOCK_N / CH - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                    }
                    const int col0 = n0 + c * CH;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
                    if (bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (col0 + j < N) {
                                float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
                                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                            }
                        }
                    }
                    if (row0 < M && col0 < N) {
                        if (lane == 0) tma_store_wait_read<1>();
                        __syncwarp();
                        uint8_t* sb = stage_buf + buf * EPI_BUF_BYTES;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 p = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                            *reinterpret_cast<float4*>(sb + row_off + (((uint32_t)j ^ sw) << 4)) = p;
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if constexpr (EPI == EPI_F32_ADD) tma_reduce_add_2d(&tmap_c, sb, col0, row0);
                            else tma_store_2d(&tmap_c, sb, col0, row0);
                            tma_store_commit();
                        }
                        buf ^= 1;
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        TC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        TC_CHECK(qres == cudaDriverEntryPointSuccess && p != nullptr, "cuTensorMapEncodeTiled not available");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D row-major tensor [rows, cols] with leading dimension ld (elements); box = [box_rows, box_cols]; 128B swizzle
CUtensorMap make_tmap(const void* ptr, CUtensorMapDataType dt, int elem_bytes, int64_t rows, int64_t cols, int64_t ld,
                      int box_rows, int box_cols) {
    TC_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base pointer must be 16-byte aligned");
    TC_CHECK((ld * elem_bytes) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes (ld=%lld)", (long long)ld);
    TC_CHECK(box_cols * elem_bytes == 128, "box inner extent must be 128 bytes");
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(ld * elem_bytes)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_fn()(&m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TC_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return m;
}

int g_num_sms = 0;

template <int BLOCK_N, int EPI, int ACT, bool STORE_PRE, bool F16>
void launch(const GemmArgs& g, cudaStream_t stream) {
    using C = Cfg<BLOCK_N>;
    auto kern = gemm_tc_kernel<BLOCK_N, EPI, ACT, STORE_PRE, F16>;
    constexpr CUtensorMapDataType DT16 = F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    static bool configured = false;
    if (!configured) {
        TC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    if (g_num_sms == 0) {
        int dev;
        TC_CUDA(cudaGetDevice(&dev));
        TC_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    CUtensorMap ta = make_tmap(g.a, DT16, 2, g.M, g.K, g.lda, BLOCK_M, BLOCK_K);
    CUtensorMap tb = make_tmap(g.w, DT16, 2, g.N, g.K, g.ldw, BLOCK_N, BLOCK_K);
    CUtensorMap tc, tc2;
    if (EPI == EPI_BF16) tc = make_tmap(g.out, DT16, 2, g.M, g.N, g.ldo, 32, 64);
    else tc = make_tmap(g.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.M, g.N, g.ldo, 32, 32);
    tc2 = tc;
    if (STORE_PRE) tc2 = make_tmap(g.out_pre, DT16, 2, g.M, g.N, g.ldo, 32, 64);
    const int64_t tiles = ceil_div(g.M, BLOCK_M) * ceil_div(g.N, BLOCK_N);
    const int grid = (int)std::min<int64_t>(tiles, g_num_sms);
    kern<<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, tc, tc2, g.bias, (int)g.M, (int)g.N, (int)g.K);
    TC_LAUNCH_CHECK();
}

template <int BLOCK_N, bool F16>
void dispatch(const GemmArgs& g, cudaStream_t stream) {
    if (g.epi == EPI_F32) return launch<BLOCK_N, EPI_F32, ACT_NONE, false, F16>(g, stream);
    if (g.epi == EPI_F32_ADD) return launch<BLOCK_N, EPI_F32_ADD, ACT_NONE, false, F16>(g, stream);
    TC_CHECK(g.epi == EPI_BF16, "unknown epilogue %d", g.epi);
    const bool pre = g.out_pre != nullptr;
    if (g.act == ACT_NONE) { TC_CHECK(!pre, "out_pre needs an activation"); return launch<BLOCK_N, EPI_BF16, ACT_NONE, false, F16>(g, stream); }
    if (g.act == ACT_GELU_ERF) return pre ? launch<BLOCK_N, EPI_BF16, ACT_GELU_ERF, true, F16>(g, stream)
                                          : launch<BLOCK_N, EPI_BF16, ACT_GELU_ERF, false, F16>(g, stream);
    if (g.act == ACT_QUICK_GELU) return pre ? launch<BLOCK_N, EPI_BF16, ACT_QUICK_GELU, true, F16>(g, stream)
                                            : launch<BLOCK_N, EPI_BF16, ACT_QUICK_GELU, false, F16>(g, stream);
    TC_CHECK(false, "unknown activation %d", g.act);
}

}  // namespace

void gemm_tc(const GemmArgs& g, cudaStream_t stream) {
    TC_CHECK(g.M > 0 && g.N > 0 && g.K > 0, "empty GEMM %lldx%lldx%lld", (long long)g.M, (long long)g.N, (long long)g.K);
    TC_CHECK(g.K % 8 == 0 && g.N % 8 == 0, "tcgen05 GEMM needs K%%8==0 and N%%8==0 (K=%lld N=%lld)", (long long)g.K, (long long)g.N);
    int bn = g.block_n;
    if (bn == 0) {
        // wide tiles when they already fill the machine, narrow ones otherwise
        if (g_num_sms == 0) { int dev; TC_CUDA(cudaGetDevice(&dev)); TC_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev)); }
        const int64_t tiles256 = ceil_div(g.M, BLOCK_M) * ceil_div(g.N, 256);
        bn = (g.N % 256 == 0 && tiles256 >= 2 * g_num_sms) ? 256 : 128;
    }
    TC_CHECK(g.dt == DT_BF16 || g.dt == DT_F16, "tcgen05 GEMM operands must be bf16 or fp16");
    if (g.dt == DT_F16) { if (bn == 256) dispatch<256, true>(g, stream); else dispatch<128, true>(g, stream); }
    else { if (bn == 256) dispatch<256, false>(g, stream); else dispatch<128, false>(g, stream); }
}

}  // namespace tapclip
