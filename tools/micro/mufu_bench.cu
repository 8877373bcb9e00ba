// Micro-benchmark: MUFU.EX2 issue rate per SM sub-partition (cycles per warp-wide ex2.approx.ftz.f32), alone and mixed
// with the FFMA/FADD/F2FP of a softmax inner loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mufu_bench tools/micro/mufu_bench.cu
#include <cstdio>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(int iters, float seed, long long* cyc, float* sink) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = seed * (threadIdx.x + j);
    float l = 0.f; unsigned pk = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            if (MODE == 0) { v[j] = ex2(v[j]); v[j + 1] = ex2(v[j + 1]); }
            else {
                const float p0 = ex2(fmaf(v[j], 0.18f, -3.f)), p1 = ex2(fmaf(v[j + 1], 0.18f, -3.f));
                l += p0 + p1;
                __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
                pk ^= *reinterpret_cast<unsigned*>(&h);
                v[j] += 1e-3f; v[j + 1] -= 1e-3f;
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    float a = l + __uint_as_float(pk);
#pragma unroll
    for (int j = 0; j < 16; ++j) a += v[j];
    if (a == 1.2345f) sink[0] = a;
}
int main() {
    long long* c; float* s; cudaMalloc(&c, 8); cudaMalloc(&s, 4);
    const int iters = 2000;
    for (int warps : {4, 8, 16}) {
        for (int mode = 0; mode < 2; ++mode) {
            for (int rep = 0; rep < 2; ++rep) { if (mode == 0) k<0><<<148, warps * 32>>>(iters, 1e-3f, c, s); else k<1><<<148, warps * 32>>>(iters, 1e-3f, c, s); cudaDeviceSynchronize(); }
            long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            const double per = (double)h / ((double)iters * 16 * (warps / 4));
            printf("%2d warps/SM, %s: %.2f cycles per warp-wide ex2 per sub-partition (%.1f exp/clk/SM)\n", warps, mode ? "softmax mix" : "ex2 only   ", per, 128.0 / per);
        }
    }
    return 0;
}
