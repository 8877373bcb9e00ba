// Rollout step on tcgen05 (long sequences, e.g. ViT-L/14@336: 577 tokens).  Same contract as rollout_step in rollout.cu:
//   r_out[s,j] = 0.5 r_in[s,j] + (0.5/H) sum_h sum_i r_in[s,i] 2^(c q_i.k_j - lse[s,h,i])
//
// The transposed score tile S^T = K Q^T puts one KEY per TMEM lane and the queries along the columns, so the sum over the
// queries that the propagation needs stays inside one thread (no shuffles, no shared-memory reduction), and the tensor
// work leaves the legacy HMMA path that bounded the mma.sync kernel together with the MUFU pipe.
//
// One CTA per (image, 128-key tile), 20 warps:
//   warp 0      TMA: per head the K tile [128 x 64] and all Q rows [npad x 64] as 32-row 128B-swizzled boxes, 2 head stages
//   warp 1      tcgen05.mma issuer: per head and 128-query block  D[128 x wb] = K_tile . Q_block^T  (M=128, N=wb, K=64) into
//               the TMEM column quarter blk % 4; tcgen05.commit releases the head's stage when its MMAs are done
//   warps 2-3   lw[h][i] = lse[s,h,i] - log2 r_in[s,i] for the next head (double buffered): r_i is folded into the exponent
//   warps 4-19  four groups of four warps (one TMEM lane quarter each); group g drains the blocks with blk % 4 == g:
//               tcgen05.ld 32 columns -> 2^(c x - lw) -> per-thread sum; four warps per scheduler keep the MUFU pipe fed while
//               other groups wait for their next block (two groups of 256-column blocks: 1.28 ms at B=512, N=577)
// The groups' partial sums of a key are combined through shared memory at the end; no atomics, deterministic.
#include "gemm.h"
#include "kernels.h"
#include <cstdlib>

namespace tapclip {
namespace {

constexpr int DH = 64;
constexpr int RT_NG = 4;               // drain groups = TMEM column quarters
constexpr int RT_THREADS = 128 + RT_NG * 128;
constexpr int RT_KEYS = 128;           // keys per CTA = TMEM lanes
constexpr int RT_QB = 512 / RT_NG;     // queries per MMA block = columns of one group's TMEM region
constexpr int RT_NST = 2;              // head stages of K/Q in shared memory
constexpr float SCALE_LOG2 = 0.125f * 1.4426950408889634f;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;              // SBO: 8 rows x 128 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint32_t idesc_kmajor(int m, int n, bool f16) {
    const uint32_t fmt = f16 ? 0u : 1u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool F16>
__global__ void __launch_bounds__(RT_THREADS, 1)
rollout_step_tc_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ lse, const float* __restrict__ r_in,
                       float* __restrict__ r_out, int S, int N, int H, int npad, int skip, int tile_major, int kpc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // kpc = key tiles per CTA (1 or 2): with two, the head's Q rows are fetched once for 256 keys (L2 -> SM traffic -1/3 at N = 577)
    const int stage_bytes = (kpc * RT_KEYS + npad) * 128;     // K tile(s), then the Q rows (a multiple of 1024: npad % 32 == 0)
    float* lr = reinterpret_cast<float*>(smem + RT_NST * stage_bytes);   // [npad] log2 r_in
    float* lw = lr + npad;                                    // [2][npad]
    float* comb = lr;                                         // [RT_NG - 1][2][128] partial sums, over lr / lw once every drain is done
    uint64_t* bars = reinterpret_cast<uint64_t*>(lw + 2 * npad);
    uint64_t* full = bars;                // [RT_NST] TMA transaction barriers
    uint64_t* empty = bars + RT_NST;      // [RT_NST] the head's MMAs are complete
    uint64_t* bar_s = empty + RT_NST;     // [RT_NG] block in TMEM region g is complete
    uint64_t* tfree = bar_s + RT_NG;      // [RT_NG] count 4: the group has drained its region
    uint64_t* lw_full = tfree + RT_NG;    // [2] count 2
    uint64_t* lw_empty = lw_full + 2;     // [2] count 4 * RT_NG: every drain warp, once per head
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lw_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int ntl = (N + RT_KEYS - 1) / RT_KEYS, npi = (ntl + kpc - 1) / kpc;     // key tiles / CTAs per image
    const int s = tile_major ? blockIdx.x / npi : blockIdx.x % S;
    const int t0 = (tile_major ? blockIdx.x % npi : blockIdx.x / S) * kpc;
    const int nkt = min(kpc, ntl - t0), k0 = t0 * RT_KEYS;                        // this CTA's key tiles [t0, t0 + nkt)
    const int d = H * DH;
    const int nb = (npad + RT_QB - 1) / RT_QB;                // query blocks per head (>= 3: checked by the launcher)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap);
        for (int i = 0; i < RT_NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < RT_NG; ++i) { mbar_init(&bar_s[i], 1); mbar_init(&tfree[i], 4); }
        for (int i = 0; i < 2; ++i) { mbar_init(&lw_full[i], 2); mbar_init(&lw_empty[i], 4 * RT_NG); }
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    pdl_wait();                                               // r_in is the previous step's output
    for (int i = threadIdx.x; i < npad; i += RT_THREADS) {
        float v = -INFINITY;
        if (i < N) v = r_in ? log2f(r_in[(int64_t)s * N + i]) : (i == 0 ? 0.f : -INFINITY);
        lr[i] = v;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_trigger();

    if (warp == 0) {
        // ---- loader ----
        if (lane == 0) {
            for (int h = 0; h < H; ++h) {
                const int st = h % RT_NST;
                if (h >= RT_NST) mbar_wait(&empty[st], (uint32_t)((h / RT_NST - 1) & 1));
                uint8_t* base = smem + st * stage_bytes;
                mbar_expect_tx(&full[st], (uint32_t)((nkt * RT_KEYS + npad) * 128));
                for (int r = 0; r < nkt * RT_KEYS / 32; ++r) tma_load_2d(base + r * 32 * 128, &tmap, d + h * DH, s * N + k0 + r * 32, &full[st]);
                for (int r = 0; r < npad / 32; ++r) tma_load_2d(base + (kpc * RT_KEYS + r * 32) * 128, &tmap, h * DH, s * N + r * 32, &full[st]);
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer (warp-uniform control flow, elect.sync issue) ----
        int blk = 0;
        for (int h = 0; h < H; ++h) {
            const int st = h % RT_NST;
            mbar_wait(&full[st], (uint32_t)((h / RT_NST) & 1));
            tc_fence_after();
            const uint32_t kaddr = smem_u32(smem + st * stage_bytes);
            for (int kt = 0; kt < nkt; ++kt) {
                const uint64_t kd = desc_sw128(kaddr + (uint32_t)(kt * RT_KEYS) * 128u);
                for (int qb = 0; qb < nb; ++qb, ++blk) {
                    const int g = blk % RT_NG, use = blk / RT_NG;
                    if (use >= 1) { mbar_wait(&tfree[g], (uint32_t)((use - 1) & 1)); tc_fence_after(); }
                    const int wb = min(RT_QB, npad - qb * RT_QB);
                    const uint64_t qd = desc_sw128(kaddr + (uint32_t)(kpc * RT_KEYS + qb * RT_QB) * 128u);
                    const uint32_t idesc = idesc_kmajor(RT_KEYS, wb, F16);
#pragma unroll
                    for (int k = 0; k < DH / 16; ++k) umma_ss_elect(tmem_base + g * RT_QB, kd + 2 * k, qd + 2 * k, idesc, k != 0);
                    umma_commit_elect(&bar_s[g]);
                }
            }
            umma_commit_elect(&empty[st]);                    // every MMA that reads this stage has completed
        }
    } else if (warp < 4) {
        // ---- lw producers ----
        const int t = (int)threadIdx.x - 64;
        for (int h = 0; h < H; ++h) {
            const int b = h & 1;
            if (h >= 2) mbar_wait(&lw_empty[b], (uint32_t)(((h >> 1) - 1) & 1));
            const float* lse_h = lse + ((int64_t)s * H + h) * N;
            // r_i = 0 (log2 = -inf): the row's statistics may be unwritten (dead query rows of the last layer) - never read them
            for (int i = t; i < npad; i += 64) lw[b * npad + i] = (i < N && lr[i] != -INFINITY) ? lse_h[i] - lr[i] : INFINITY;
            __syncwarp();
            if (lane == 0) mbar_arrive(&lw_full[b]);
        }
    } else {
        // ---- drain group g: keys = TMEM lanes, queries = columns ----
        const int g = (warp - 4) >> 2, q = warp & 3;                        // a group's four warps cover the four TMEM lane quarters
        const uint32_t trow = tmem_base + g * RT_QB + ((uint32_t)(q * 32) << 16);
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;             // of the block being drained
        float acc_kt[2] = {0.f, 0.f};                                       // per key tile, over heads and query blocks
        auto drain32 = [&](const uint32_t (&x)[32], const float* lwp) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 w = *reinterpret_cast<const float4*>(lwp + j);          // same address in every lane: broadcast
                acc0 += ex2_approx(fmaf(__uint_as_float(x[j]), SCALE_LOG2, -w.x));
                acc1 += ex2_approx(fmaf(__uint_as_float(x[j + 1]), SCALE_LOG2, -w.y));
                acc2 += ex2_approx(fmaf(__uint_as_float(x[j + 2]), SCALE_LOG2, -w.z));
                acc3 += ex2_approx(fmaf(__uint_as_float(x[j + 3]), SCALE_LOG2, -w.w));
            }
        };
        int blk = 0, use = 0;
        for (int h = 0; h < H; ++h) {
            bool have_lw = false;
            for (int kt = 0; kt < nkt; ++kt)
            for (int qb = 0; qb < nb; ++qb, ++blk) {
                if (blk % RT_NG != g) continue;
                if (!have_lw) { mbar_wait(&lw_full[h & 1], (uint32_t)((h >> 1) & 1)); have_lw = true; }
                mbar_wait(&bar_s[g], (uint32_t)(use & 1));
                tc_fence_after();
                const int wb = min(RT_QB, npad - qb * RT_QB);
                const float* lwp = lw + (h & 1) * npad + qb * RT_QB;
                uint32_t a[32], b[32];
                tmem_ld_32x32(trow, a);
                for (int c0 = 0; c0 < wb; c0 += 64) {
                    tmem_ld_wait();
                    const bool second = c0 + 32 < wb;
                    if (second) tmem_ld_32x32(trow + c0 + 32, b);
                    drain32(a, lwp + c0);
                    if (second) {
                        tmem_ld_wait();
                        if (c0 + 64 < wb) tmem_ld_32x32(trow + c0 + 64, a);
                        drain32(b, lwp + c0 + 32);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tfree[g]);
                ++use;
                const float sblk = (acc0 + acc1) + (acc2 + acc3);
                acc0 = acc1 = acc2 = acc3 = 0.f;
                if (kt == 0) acc_kt[0] += sblk; else acc_kt[1] += sblk;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&lw_empty[h & 1]);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(RT_NG * 128) : "memory");     // every drain warp is done with lw (and the producers with lr)
        if (g > 0) {
            comb[((g - 1) * 2 + 0) * RT_KEYS + q * 32 + lane] = acc_kt[0];
            comb[((g - 1) * 2 + 1) * RT_KEYS + q * 32 + lane] = acc_kt[1];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(RT_NG * 128) : "memory");
        if (g == 0) {
#pragma unroll
            for (int kt = 0; kt < 2; ++kt) {
                float acc = acc_kt[kt];
#pragma unroll
                for (int o = 0; o < RT_NG - 1; ++o) acc += comb[(o * 2 + kt) * RT_KEYS + q * 32 + lane];
                const int key = k0 + kt * RT_KEYS + q * 32 + lane;
                if (kt < nkt && key < N && key >= skip) {
                    const float rk = r_in ? r_in[(int64_t)s * N + key] : (key == 0 ? 1.f : 0.f);
                    r_out[(int64_t)s * (N - skip) + key - skip] = fmaf(0.5f / (float)H, acc, 0.5f * rk);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

size_t rollout_tc_smem(int npad, int kpc) {
    return (size_t)RT_NST * (kpc * RT_KEYS + npad) * 128 + (size_t)3 * npad * sizeof(float) + 256 + 1024;
}

}  // namespace

// key tiles per CTA: two when they fit (the Q rows of a head then serve 256 keys)
int rollout_tc_kpc(int N) {
    static const int kpc_env = getenv("TAPCLIP_ROLLOUT_KPC") ? atoi(getenv("TAPCLIP_ROLLOUT_KPC")) : 0;   // measurement switch: 1 | 2
    const int npad = (int)round_up(N, 32), ntl = (int)ceil_div(N, RT_KEYS);
    return (kpc_env != 1 && ntl >= 2 && rollout_tc_smem(npad, 2) <= 227 * 1024) ? 2 : 1;
}

bool rollout_step_tc_supported(int dt, int N) {
    const int npad = (int)round_up(N, 32), ntl = (int)ceil_div(N, RT_KEYS), nb = (int)ceil_div(npad, RT_QB), kpc = rollout_tc_kpc(N);
    // every CTA needs >= 3 (key tile, query block) MMA blocks per head: then no drain group is without a block in two consecutive
    // heads (lw_empty hand-off, see the kernel).  The image's last CTA holds ntl % kpc key tiles when that is not zero.
    const int min_tiles = (ntl % kpc) ? ntl % kpc : kpc;
    return (dt == DT_BF16 || dt == DT_F16) && min_tiles * nb >= 3 && rollout_tc_smem(npad, kpc) <= 227 * 1024;
}
int rollout_step_tc_ctas(int S, int N) { return S * (int)ceil_div(ceil_div(N, RT_KEYS), rollout_tc_kpc(N)); }

void rollout_step_tc(const void* qkv, const float* lse, const float* r_in, float* r_out, int dt, int S, int N, int H, bool last,
                     cudaStream_t stream) {
    TC_CHECK(rollout_step_tc_supported(dt, N), "tcgen05 rollout step: 16-bit inputs, 128 < N <= ~700");
    const int d = H * DH, npad = (int)round_up(N, 32);
    const bool f16 = dt == DT_F16;
    const CUtensorMap& tm = make_tmap(qkv, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (int64_t)S * N, 3 * d,
                                      3 * d, 32, 64);
    const int ntl = (int)ceil_div(N, RT_KEYS);
    const int kpc = rollout_tc_kpc(N);
    const size_t smem = rollout_tc_smem(npad, kpc);
    if (f16) ensure_dynamic_smem((const void*)rollout_step_tc_kernel<true>, smem);
    else ensure_dynamic_smem((const void*)rollout_step_tc_kernel<false>, smem);
    const unsigned grid = (unsigned)(S * (int)ceil_div(ntl, kpc));
    // tile-major launch order: the key tiles of an image run at the same time and share its Q rows in L2 (1.28 vs 1.52 ms at
    // B=512, N=577: image-major re-reads every row from HBM, 3.8 GB per launch, as 128-byte pieces)
    static const int tile_major = getenv("TAPCLIP_ROLLOUT_ORDER") ? atoi(getenv("TAPCLIP_ROLLOUT_ORDER")) : 1;
    if (f16) launch_pdl(rollout_step_tc_kernel<true>, grid, RT_THREADS, smem, stream, tm, lse, r_in, r_out, S, N, H, npad, last ? 1 : 0, tile_major, kpc);
    else launch_pdl(rollout_step_tc_kernel<false>, grid, RT_THREADS, smem, stream, tm, lse, r_in, r_out, S, N, H, npad, last ? 1 : 0, tile_major, kpc);
    TC_LAUNCH_CHECK();
}

}  // namespace tapclip
