// Micro-benchmark: issue cost and completion time of a burst of small tcgen05.mma (the P.V step of attention:
// 13 x [128x64x16], A from TMEM) vs the same math as SS MMAs, from a single elected lane.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_issue_bench tools/micro/mma_issue_bench.cu -I tapclip_b200/csrc
#include "common.cuh"
#include <cstdio>
using namespace tapclip;
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    uint64_t d = 0; d |= (uint64_t)((a & 0x3FFFF) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
__host__ __device__ constexpr uint32_t idesc(int m, int n, bool bmn) { return (1u << 4) | (1u << 7) | (1u << 10) | ((bmn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

__device__ __forceinline__ void umma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc_, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc_), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(64, 1) k(int mode, int n_mma, int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 44 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); fence_proxy_async_smem(); }
    if (warp == 1) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = slot;
    if (mode == 3) {
        // warp-uniform variant: warp index and TMEM base made provably uniform with shuffles, all 32 lanes run the loop
        const int uwarp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
        const uint32_t utb = __shfl_sync(0xffffffffu, tb, 0);
        if (uwarp == 0) {
            const uint64_t bd = desc_sw128(smem_u32(smem + 16384));
            long long t_issue = 0, t_done = 0;
            for (int it = 0; it < iters; ++it) {
                const long long t0 = clock64();
#pragma unroll
                for (int ks = 0; ks < 13; ++ks) if (ks < n_mma) umma_ts_elect(utb + 128, utb + ks * 8, bd + (uint64_t)(ks * 128), idesc(128, 64, true), ks != 0);
                umma_commit_elect(&bar);
                const long long t1 = clock64();
                mbar_wait(&bar, it & 1);
                const long long t2 = clock64();
                t_issue += t1 - t0; t_done += t2 - t0;
            }
            if (blockIdx.x == 0 && lane == 0) { out[0] = t_issue / iters; out[1] = t_done / iters; }
        }
    } else if (warp == 0 && lane == 0) {
        const uint64_t ad = desc_sw128(smem_u32(smem)), bd = desc_sw128(smem_u32(smem + 16384));
        long long t_issue = 0, t_done = 0;
        for (int it = 0; it < iters; ++it) {
            const long long t0 = clock64();
            if (mode == 0) {          // TS: A = TMEM columns (P), B = MN-major V tile, N = 64
#pragma unroll
                for (int ks = 0; ks < 13; ++ks) if (ks < n_mma) umma_ts(tb + 128, tb + ks * 8, bd + (uint64_t)(ks * 128), idesc(128, 64, true), ks != 0);
            } else if (mode == 1) {   // SS, N = 64, K-major operands
#pragma unroll
                for (int ks = 0; ks < 13; ++ks) if (ks < n_mma) umma_bf16(tb + 128, ad + 2 * (ks & 3), bd + 2 * (ks & 3), idesc(128, 64, false), ks != 0);
            } else {                  // SS, N = 208 (the S = Q K^T step), 4 MMAs
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_bf16(tb, ad + 2 * ks, bd + 2 * ks, idesc(128, 208, false), ks != 0);
            }
            umma_commit(&bar);
            const long long t1 = clock64();
            mbar_wait(&bar, it & 1);
            const long long t2 = clock64();
            t_issue += t1 - t0; t_done += t2 - t0;
        }
        if (blockIdx.x == 0) { out[0] = t_issue / iters; out[1] = t_done / iters; }
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 1) tmem_dealloc(tb, 512);
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    const char* names[] = {"TS  128x64x16 (P.V)", "SS  128x64x16", "SS  128x208x16 x4 (Q.K^T)", "TS  uniform warp + elect"};
    for (int mode = 0; mode < 4; ++mode)
        for (int n : {1, 4, 13}) {
            if (mode == 2 && n != 4) continue;
            k<<<148, 64, 45 * 1024>>>(mode, n, 200, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%-28s n=%2d: issue %5lld cycles, issue+complete %5lld cycles [%s]\n", names[mode], n, h[0], h[1], cudaGetErrorString(e));
        }
    return 0;
}
