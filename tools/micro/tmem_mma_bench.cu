// Micro-benchmark: latency of a tcgen05.ld round (7 x 16 columns + wait) (a) on an idle tensor core, (b) right after
// an MMA wrote those columns, (c) while MMAs into the OTHER TMEM half are in flight.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmem_mma_bench tools/micro/tmem_mma_bench.cu -I tapclip_b200/csrc
#include "common.cuh"
#include <cstdio>
using namespace tapclip;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    uint64_t d = 0; d |= (uint64_t)((a & 0x3FFFF) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
__host__ __device__ constexpr uint32_t idesc(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

// mode 0: no MMA.  mode 1: S-like MMA (4 x 128x208x16) into half 0, loads after its commit.  mode 2: as 1 plus a stream of
// MMAs into half 1 issued right before the loads (in flight while they execute).
__global__ void __launch_bounds__(192, 1) k(int mode, int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_s, bar_go, bar_x;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16 + 26) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar_s, 1); mbar_init(&bar_go, 4); mbar_init(&bar_x, 1); fence_mbar_init(); fence_proxy_async_smem(); }
    if (warp == 4) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = slot;
    long long acc_ld = 0, acc_ld2 = 0; float sink = 0.f;
    if (warp == 4) {
        if (lane == 0) {
            const uint64_t qd = desc_sw128(smem_u32(smem)), kd = desc_sw128(smem_u32(smem + 16384));
            for (int it = 0; it < iters; ++it) {
                mbar_wait(&bar_go, (uint32_t)(it & 1));   // readers are ready for round `it`
                tc_fence_after();
                if (mode >= 1) for (int kk = 0; kk < 4; ++kk) umma_bf16(tb, qd + 2 * kk, kd + 2 * kk, idesc(128, 208), kk != 0);
                umma_commit(&bar_s);
                if (mode == 2) {
                    mbar_wait(&bar_s, it & 1);          // S done; now flood the other half while the readers load
                    for (int rep = 0; rep < 3; ++rep)
                        for (int kk = 0; kk < 4; ++kk) umma_bf16(tb + 256, qd + 2 * kk, kd + 2 * kk, idesc(128, 208), 1);
                    umma_commit(&bar_x);
                    mbar_wait(&bar_x, it & 1);
                }
            }
        }
    } else if (warp < 4) {
        const uint32_t trow = tb + ((uint32_t)(warp * 32) << 16);
        for (int it = 0; it < iters; ++it) {
            if (lane == 0) mbar_arrive(&bar_go);
            mbar_wait(&bar_s, it & 1);
            tc_fence_after();
            uint32_t r[7][16];
            const long long t0 = clock64();
#pragma unroll
            for (int u = 0; u < 7; ++u) tmem_ld_32x16(trow + u * 16, r[u]);
            tmem_ld_wait();
            const long long t1 = clock64();
#pragma unroll
            for (int u = 0; u < 7; ++u) sink += __uint_as_float(r[u][0]) + __uint_as_float(r[u][15]);
            // second round on the same (now clean) columns
#pragma unroll
            for (int u = 0; u < 7; ++u) tmem_ld_32x16(trow + u * 16, r[u]);
            tmem_ld_wait();
            const long long t2 = clock64();
#pragma unroll
            for (int u = 0; u < 7; ++u) sink += __uint_as_float(r[u][1]);
            acc_ld += t1 - t0; acc_ld2 += t2 - t1;
            tc_fence_before();
            __syncwarp();
        }
        if (lane == 0 && blockIdx.x == 0) { out[warp * 2] = acc_ld / iters; out[warp * 2 + 1] = acc_ld2 / iters; }
        if (sink == 1.2345f) out[15] = 1;
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 4) tmem_dealloc(tb, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 16 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    for (int mode = 0; mode < 3; ++mode) {
        cudaMemset(d, 0, 128);
        k<<<148, 192, 44 * 1024>>>(mode, 200, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("mode %d (%s): first round %lld %lld %lld %lld cycles, second round %lld %lld %lld %lld  [%s]\n", mode,
               mode == 0 ? "no MMA" : mode == 1 ? "after S-like MMA" : "MMAs in flight on the other half",
               h[0], h[2], h[4], h[6], h[1], h[3], h[5], h[7], cudaGetErrorString(e));
    }
    return 0;
}
