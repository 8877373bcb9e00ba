// Micro-benchmark: tcgen05.ld throughput per SM (how fast can softmax/epilogue warps drain TMEM?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmem_bench tools/micro/tmem_bench.cu -I tapclip_b200/csrc
#include "common.cuh"
#include <cstdio>
using namespace tapclip;

template <int X>
__global__ void __launch_bounds__(256, 1) tmem_rd(int iters, int nwarps, long long* cycles, float* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nwarps) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < 256; c += 4 * X) {
                uint32_t r[4][X];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if constexpr (X == 32) tmem_ld_32x32(base + (warp >> 2) * 256 + c + u * X, r[u]);
                    else tmem_ld_32x16(base + (warp >> 2) * 256 + c + u * X, r[u]);
                }
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 4; ++u) acc += __uint_as_float(r[u][0]) + __uint_as_float(r[u][X - 1]);
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
    long long* cyc; float* sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
    const int iters = 200;
    for (int x : {16, 32})
        for (int nw : {1, 2, 4, 8}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (x == 16) tmem_rd<16><<<148, 256>>>(iters, nw, cyc, sink); else tmem_rd<32><<<148, 256>>>(iters, nw, cyc, sink);
                cudaDeviceSynchronize();
            }
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * 256 * 128 * nw;       // per SM: 256 columns x 32 lanes x 4 B per warp pass
            printf("ld.x%d warps=%d: %lld cycles, %.1f B/cycle/SM (%s)\n", x, nw, h[0], bytes / (double)h[0], cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
