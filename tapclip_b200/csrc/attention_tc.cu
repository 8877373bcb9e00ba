// K2 on tcgen05: fused attention forward for sequences of up to 256 tokens (ViT-B/16: 197, ViT-B/32: 50, text: P+77),
// head dim 64, no mask, with the same probe epilogue as attention.cu.
//
// Work item = (sequence, head, 128-query tile):
//   TMA   Q tile [128 x 64], K and V [NKP x 64] (NKP = keys padded to 16) into 128B-swizzled smem
//   MMA1  S = Q K^T        tcgen05.mma  M=128, N=NKP, K=64  (A, B K-major from smem)  -> TMEM columns [0, NKP)
//   softmax: one thread per query row reads its S row from TMEM, takes the max, computes exp2, writes the 16-bit
//         probabilities BACK INTO TMEM over the S columns (two keys per 32-bit column) and emits the probe
//   MMA2  O = P V          tcgen05.mma  M=128, N=64 (80 in the persistent kernel: 16 columns of ones give the row sum),
//         K=NKP   (A = P from TMEM, B = V as stored: MN-major smem descriptor)
//   epilogue: O / rowsum -> 16-bit -> global
// S, P and O never touch shared memory or HBM; the N x N map is never materialised.
// Two kernels: attn_fwd_tc2_kernel (NKP <= 208: persistent, warp-specialised, described at its definition) and
// attn_fwd_tc_kernel (208 < NKP <= 256: one CTA per item, the simple form of the same data flow).
#include "gemm.h"
#include "kernels.h"
#include <cstdlib>

namespace tapclip {
namespace {

constexpr int DH = 64;
constexpr int O_COL = 128;            // O accumulator: P (16-bit) needs only columns [0, NKP/2) <= [0,128); S columns beyond are dead by then
constexpr int TMEM_COLS = 256;        // -> two CTAs per SM can hold their accumulators at once
constexpr int ATTN_THREADS = 192;

__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                        // LBO (unused: one 128-byte atom along the contiguous dimension)
    d |= (uint64_t)(1024 >> 4) << 32;              // SBO: 8 rows x 128 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t attn_idesc(int m, int n, bool f16, bool b_mn_major) {
    const uint32_t fmt = f16 ? 0u : 1u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

#ifdef TAPCLIP_ATTN_TRACE
// developer-only phase trace of CTA 0 (tools/micro/attn_trace.py builds a private copy of the library with this enabled)
__device__ long long g_attn_trace[8 * 16 * 16];     // [group][warp quarter][item][slot]
#define TRACE(slot) do { if (tr) tr[slot] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^x for x <= 0 on the FMA/ALU pipes (no MUFU): round-to-nearest split x = n + f, |f| <= 0.5, minimax cubic for 2^f
// (max relative error 7.5e-5, below the rounding of the 16-bit probabilities), n added into the exponent field.
// The softmax pass is MUFU-bound (16 exp2/clk/SM, tools/micro/mufu_bench.cu): moving a third of the exponentials here
// shortens it, as in FlashAttention-4's software exp2.
__device__ __forceinline__ float poly_exp2(float x) {
    x = fmaxf(x, -125.f);
    const float t = x + 12582912.f;                  // 1.5 * 2^23: n = round(x) lands in the low mantissa bits
    const float f = x - (t - 12582912.f);
    float p = fmaf(f, 0.05517084f, 0.24260908f);
    p = fmaf(p, f, 0.69326109f);
    p = fmaf(p, f, 0.99992847f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// Softmax of ONE query row held in TMEM (one thread = one TMEM lane).  Two passes over the S row (max, then exp2/sum),
// read in groups of up to 64 columns with 4 tcgen05.ld in flight per wait; the 16-bit probabilities are written back
// into TMEM over the S columns already consumed (two keys per 32-bit column).  Only the last group can contain padded
// keys, so interior groups are branch-free.  `cls_out` (warp-uniform non-null only for the warp holding the CLS row;
// lane `cls_lane` writes) receives the unnormalised probabilities of that row.  Returns the row sum; *p_last = p[N-1].
template <typename T16>
__device__ __forceinline__ float softmax_to_tmem(uint32_t trow, int N, int nkp, float scale_log2, float* cls_out, bool cls_lane,
                                                 float* p_last, long long* tr = nullptr) {
    const int nch = nkp / 16;
    // pass 1: row maximum, 4 x 16 columns per tcgen05.wait::ld; four independent accumulators keep the FMNMX chain off
    // the critical path
    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
    for (int c0 = 0; c0 < nch; c0 += 4) {
        uint32_t r[4][16];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c0 + u < nch) tmem_ld_32x16(trow + (c0 + u) * 16, r[u]);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if ((c0 + u + 1) * 16 <= N) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    m0 = fmaxf(m0, __uint_as_float(r[u][j])); m1 = fmaxf(m1, __uint_as_float(r[u][j + 1]));
                    m2 = fmaxf(m2, __uint_as_float(r[u][j + 2])); m3 = fmaxf(m3, __uint_as_float(r[u][j + 3]));
                }
            } else if (c0 + u < nch) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if ((c0 + u) * 16 + j < N) m0 = fmaxf(m0, __uint_as_float(r[u][j]));
            }
        }
    }
    const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    const float mneg = -mx * scale_log2;
    TRACE(4);
    float l = 0.f, l1 = 0.f, pl = 0.f;
    for (int c0 = 0; c0 < nch; c0 += 4) {
        uint32_t r[4][16];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c0 + u < nch) tmem_ld_32x16(trow + (c0 + u) * 16, r[u]);
        tmem_ld_wait();
        if ((c0 + 4) * 16 < N && cls_out == nullptr) {
            // interior group: every key valid, no probe bookkeeping
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const float p0 = fast_exp2(fmaf(__uint_as_float(r[u][j]), scale_log2, mneg));
                    const float p1 = fast_exp2(fmaf(__uint_as_float(r[u][j + 1]), scale_log2, mneg));
                    if (j & 2) l1 += p0 + p1; else l += p0 + p1;
                    pk[j >> 1] = pack2<T16>(p0, p1);
                }
                tmem_st_32x8(trow + (c0 + u) * 8, pk);
            }
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (c0 + u < nch) {
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const int key = (c0 + u) * 16 + j;
                        float p0 = fast_exp2(fmaf(__uint_as_float(r[u][j]), scale_log2, mneg));
                        float p1 = fast_exp2(fmaf(__uint_as_float(r[u][j + 1]), scale_log2, mneg));
                        if (key >= N) p0 = 0.f;
                        if (key + 1 >= N) p1 = 0.f;
                        if (key == N - 1) pl = p0;
                        if (key + 1 == N - 1) pl = p1;
                        if (cls_out != nullptr && cls_lane) { if (key < N) cls_out[key] = p0; if (key + 1 < N) cls_out[key + 1] = p1; }
                        l += p0 + p1;
                        pk[j >> 1] = pack2<T16>(p0, p1);
                    }
                    tmem_st_32x8(trow + (c0 + u) * 8, pk);
                }
            }
        }
    }
    *p_last = pl;
    return l + l1;
}

// O row (64 fp32 columns at trow + O_COL) -> registers, scaled by 1/rowsum and packed to 16 bits
template <typename T16>
__device__ __forceinline__ void load_o_row(uint32_t trow, float inv, uint4 (&o)[8]) {
    uint32_t r[4][16];
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld_32x16(trow + O_COL + c * 16, r[c]);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        o[2 * c].x = pack2<T16>(__uint_as_float(r[c][0]) * inv, __uint_as_float(r[c][1]) * inv);
        o[2 * c].y = pack2<T16>(__uint_as_float(r[c][2]) * inv, __uint_as_float(r[c][3]) * inv);
        o[2 * c].z = pack2<T16>(__uint_as_float(r[c][4]) * inv, __uint_as_float(r[c][5]) * inv);
        o[2 * c].w = pack2<T16>(__uint_as_float(r[c][6]) * inv, __uint_as_float(r[c][7]) * inv);
        o[2 * c + 1].x = pack2<T16>(__uint_as_float(r[c][8]) * inv, __uint_as_float(r[c][9]) * inv);
        o[2 * c + 1].y = pack2<T16>(__uint_as_float(r[c][10]) * inv, __uint_as_float(r[c][11]) * inv);
        o[2 * c + 1].z = pack2<T16>(__uint_as_float(r[c][12]) * inv, __uint_as_float(r[c][13]) * inv);
        o[2 * c + 1].w = pack2<T16>(__uint_as_float(r[c][14]) * inv, __uint_as_float(r[c][15]) * inv);
    }
}

// Softmax + epilogue of one row for the simple (non-persistent) kernel: waits S, writes P, signals bar_p, waits O, stores.
template <typename T16>
__device__ __forceinline__ void softmax_row(uint32_t trow, uint64_t* bar_s, uint32_t par_s, uint64_t* bar_p, uint64_t* bar_o,
                                            uint32_t par_o, int lane, int grow, int s, int h, int d, int N, int H, int nkp,
                                            float scale_log2, int probe_mode, float* __restrict__ probe_out, int probe_P,
                                            int64_t probe_seq_stride, void* __restrict__ out_) {
    mbar_wait(bar_s, par_s);
    tc_fence_after();
    const bool cls_warp = (probe_mode == PROBE_CLS_ROW) && (grow - lane) == 0;        // the warp that holds query row 0
    float* cls_out = cls_warp ? probe_out + (int64_t)s * probe_seq_stride + (int64_t)h * N : nullptr;
    float p_last;
    const float l = softmax_to_tmem<T16>(trow, N, nkp, scale_log2, cls_out, lane == 0, &p_last);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    const float inv = 1.f / l;
    if (probe_mode == PROBE_TEXT_COL && grow < probe_P) probe_out[((int64_t)s * H + h) * probe_P + grow] = p_last * inv;
    if (cls_warp && lane == 0)
        for (int key = 0; key < N; ++key) cls_out[key] *= inv;              // own earlier writes
    mbar_wait(bar_o, par_o);
    tc_fence_after();
    uint4 o[8];
    load_o_row<T16>(trow, inv, o);
    if (grow < N) {
        uint4* gp = reinterpret_cast<uint4*>(reinterpret_cast<T16*>(out_) + ((int64_t)s * N + grow) * d + h * DH);
#pragma unroll
        for (int c = 0; c < 8; ++c) gp[c] = o[c];
    }
}

template <bool F16>
__global__ void __launch_bounds__(ATTN_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv, void* __restrict__ out_,
                   int N, int H, int nkp, int nqt, float scale_log2, int probe_mode, float* __restrict__ probe_out, int probe_P,
                   int64_t probe_seq_stride) {
    using T16 = typename std::conditional<F16, f16, bf16>::type;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* Qs = smem;
    uint8_t* Ks = Qs + 128 * 128;
    uint8_t* Vs = Ks + nkp * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + nkp * 128);
    uint64_t* bar_load = bars;
    uint64_t* bar_s = bars + 1;
    uint64_t* bar_p = bars + 2;
    uint64_t* bar_o = bars + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x % nqt, sh = blockIdx.x / nqt;
    const int s = sh / H, h = sh % H;
    const int d = H * DH;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        mbar_init(bar_load, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_p, 4);
        mbar_init(bar_o, 1);
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            const int row_q = s * N + qt * 128, row_kv = s * N;
            mbar_expect_tx(bar_load, (uint32_t)(128 * 128 + 2 * nkp * 128));
            tma_load_2d(Qs, &tmap_q, h * DH, row_q, bar_load);
            tma_load_2d(Ks, &tmap_kv, d + h * DH, row_kv, bar_load);
            tma_load_2d(Vs, &tmap_kv, 2 * d + h * DH, row_kv, bar_load);
            mbar_wait(bar_load, 0);
            tc_fence_after();
            // S = Q K^T
            const uint32_t idesc_s = attn_idesc(128, nkp, F16, false);
            const uint64_t qd = smem_desc_sw128(smem_u32(Qs)), kd = smem_desc_sw128(smem_u32(Ks));
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_base, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
            umma_commit(bar_s);
            // O = P V   (P from TMEM: 8 columns per 16 keys; V rows are keys: advance 16 rows = 2048 B per k-step)
            mbar_wait(bar_p, 0);
            tc_fence_after();
            const uint32_t idesc_o = attn_idesc(128, DH, F16, true);
            const uint64_t vd = smem_desc_sw128(smem_u32(Vs));
            for (int ks = 0; ks < nkp / 16; ++ks)
                umma_ts(tmem_base + O_COL, tmem_base + ks * 8, vd + (uint64_t)(ks * 128), idesc_o, ks != 0);
            umma_commit(bar_o);
        }
    } else if (warp >= 2) {
        const int q = warp & 3;
        const int row = q * 32 + lane;                      // query row inside the tile == TMEM lane
        const int grow = qt * 128 + row;                    // query row inside the sequence
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        softmax_row<T16>(trow, bar_s, 0, bar_p, bar_o, 0, lane, grow, s, h, d, N, H, nkp, scale_log2, probe_mode, probe_out, probe_P,
                         probe_seq_stride, out_);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}


// ------------------------------------------------------------------------------------------------------------------
// Persistent, software-pipelined version (NKP <= 208, e.g. ViT-B/16 and the text tower): one CTA per SM loops over
// (sequence, head, q-tile) items; item i belongs to softmax group i&1 and lives in TMEM half i&1.
//   warp 0      loader: TMA of Q/K/V for up to three items ahead (three smem operand slots)
//   warp 1      TMEM allocator + the only tcgen05.mma issuer.  Its control flow is warp-uniform (warp index and TMEM base
//               come out of shuffles) and the issue is elect.sync-predicated: ptxas then keeps the MMA operands in uniform
//               registers; a single-lane `if (lane == 0)` region wraps every UTCHMMA in an ELECT/R2UR waterfall
//               (tools/micro/mma_issue_bench.cu: 13 P.V MMAs issue in 395 instead of 900 cycles)
//   warps 2,3   idle: they pad warpgroup 0 so that setmaxnreg can hand its registers to the softmax warpgroups
//   warps 4..11 two softmax groups of 4 warps (one thread per query row).  setmaxnreg gives these threads 232 registers,
//               so up to ATTN_NRES 16-key chunks of the S row are read from TMEM once and stay in registers for the max
//               and the exp2/pack pass (the rest are read twice); P goes back into TMEM over the S columns; then the group
//               waits for O = P V (whose extra ones column carries the row sum), scales it, writes it into the warp's private
//               4 KB staging tile and hands it to ONE TMA store.
// The two groups run out of phase, so one group's exp2 pass overlaps the other's TMEM / shared-memory phases.
// All hand-offs are mbarriers (no CTA-wide or group-wide bar.sync inside the loop).
// ------------------------------------------------------------------------------------------------------------------
#ifndef ATTN_NRES
#define ATTN_NRES 8
#endif
constexpr int ATTN2_THREADS = 384;
constexpr int NSLOT = 3;

// One 16-key chunk (keys [16u, 16u+16)) of a query row held in registers.  MAYBE_MASKED = false: the caller guarantees
// every key of the chunk is < N (no per-key tests are even compiled).
template <bool MAYBE_MASKED>
__device__ __forceinline__ void chunk_max(const uint32_t (&c)[16], int u, int N, float& m0, float& m1, float& m2, float& m3) {
    if (!MAYBE_MASKED || (u + 1) * 16 <= N) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            m0 = fmaxf(m0, __uint_as_float(c[j])); m1 = fmaxf(m1, __uint_as_float(c[j + 1]));
            m2 = fmaxf(m2, __uint_as_float(c[j + 2])); m3 = fmaxf(m3, __uint_as_float(c[j + 3]));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (u * 16 + j < N) m0 = fmaxf(m0, __uint_as_float(c[j]));
    }
}
#ifndef ATTN_POLY
#define ATTN_POLY 0
#endif
constexpr int POLY_EVERY = ATTN_POLY;          // n > 0: every n-th exponential of the persistent kernel's softmax runs on the FMA pipe.
                                       // Measured (ViT-B/16 layer, B=128): 0 -> 62 us, 3 -> 77 us (round 1); with the tensor-core row sums, item
                                       // period in cycles: 0 -> 7330, 4 -> 7500, 8 -> 7250: the pass is not MUFU-throughput-bound

// exp2, row-sum, 16-bit pack and write-back of P chunk u (TMEM columns [8u, 8u+8) of the row); probe bookkeeping
// CLS = float* : the CLS row's unnormalised p goes to global memory (scalar stores, any alignment);
// CLS = uint32_t: to a 16-float-aligned shared-memory staging row (address; 0 = none) with four vector stores
// SUM = false: the caller takes the row sum from the tensor core (ones column of the P.V product) instead
template <typename T16, bool MAYBE_MASKED, bool SUM, typename CLS>
__device__ __forceinline__ void chunk_exp(const uint32_t (&c)[16], int u, int N, float scale_log2, float mneg, uint32_t trow,
                                          CLS cls_out, bool cls_lane, float& l0, float& l1, float& p_last) {
    float pv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float x = fmaf(__uint_as_float(c[j]), scale_log2, mneg);
        pv[j] = (POLY_EVERY > 0 && j % POLY_EVERY == 1) ? poly_exp2(x) : fast_exp2(x);
    }
    if (MAYBE_MASKED && (u + 1) * 16 >= N) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (u * 16 + j >= N) pv[j] = 0.f;
            if (u * 16 + j == N - 1) p_last = pv[j];
        }
    }
    if constexpr (std::is_same<CLS, uint32_t>::value) {
        if (cls_out != 0u && cls_lane) {
#pragma unroll
            for (int v = 0; v < 4; ++v)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cls_out + (uint32_t)(u * 16 + v * 4) * 4u), "f"(pv[4 * v]),
                             "f"(pv[4 * v + 1]), "f"(pv[4 * v + 2]), "f"(pv[4 * v + 3]) : "memory");
        }
    } else {
        if (cls_out != nullptr && cls_lane) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (u * 16 + j < N) cls_out[u * 16 + j] = pv[j];
        }
    }
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        if constexpr (SUM) { if (j & 2) l1 += pv[j] + pv[j + 1]; else l0 += pv[j] + pv[j + 1]; }
        pk[j >> 1] = pack2<T16>(pv[j], pv[j + 1]);
    }
    tmem_st_32x8(trow + u * 8, pk);
}

// NCH = compile-time number of 16-key chunks the softmax handles (>= NKP/16: chunks past NKP are fully masked), so that every
// register array below is indexed with constants; chunks below FIRST_MASKABLE are valid for every N this instance serves.
constexpr int CLS_STAGE2 = 208;        // floats per group: the persistent kernel serves N <= 208
// The row sums come out of the tensor core: the P.V product is issued with N = 80, where B columns 64..79 are a second MN-atom
// of the descriptor that points (leading byte offset) at 2 KB of ones, so O column 64 of a row is the fp32 sum of exactly the
// 16-bit probabilities the product used -- and the softmax threads drop their 208 FADDs per row (phase trace: -1000 cycles of
// the exp2 pass per item).  O uses that sum in every launch, so a probe never changes O.  Launches that publish per-row
// statistics -- the text-column probe (attribution) and the rollout's lse -- take the EXACT instance, whose threads also add
// up the unrounded probabilities; the CLS-row probe sums its staged row with the whole warp.
#ifndef ATTN_SUMN
#define ATTN_SUMN 16                   // extra B columns of the P.V product (item period: 16 -> 7330, 32 -> 7470, 64 -> 7580 cycles)
#endif
constexpr int ONES_OFF = 2048;         // bytes behind the operand slots: barriers + CLS staging come first
constexpr int OSTAGE_OFF = 4096;       // then 8 x 4 KB of O staging (one 32-row x 128-byte tile per softmax warp), the source of the TMA stores
template <bool F16, int NCH, bool EXACT>
__global__ void __launch_bounds__(ATTN2_THREADS, 1)
attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const __grid_constant__ CUtensorMap tmap_o, void* __restrict__ out_,
                    int N, int H, int nkp, int nqt, int n_items, float scale_log2, int probe_mode, float* __restrict__ probe_out,
                    int probe_P, int64_t probe_seq_stride, float* __restrict__ lse_out, int pair, int nslot_np, int s_early) {
    using T16 = typename std::conditional<F16, f16, bf16>::type;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // Operand slots.  pair = 0: nslot_np (3, or 2 when three do not fit next to the staging) slots of [Q 16 KB | K | V], item i in slot i % nslot_np.
    // pair = 1 (two q-tiles per head, 128 < N <= 208): the scheduling unit is a head, its q-tiles are items 2p and 2p+1 (one per
    // softmax group) and share ONE slot [Q0 | Q1 | K | V] (slot p % 2): K and V cross L2 -> shared memory once per head.  All
    // three tcgen05 attention-side kernels were found sitting at ~3 TB/s of TMA loads; this removes 38 % of them here.
    const int q_bytes = pair ? 2 * 128 * 128 : 128 * 128;
    const int slot_bytes = q_bytes + 2 * nkp * 128;
    const int nslot = pair ? 2 : nslot_np;
    auto slot_of = [&](int i) { return pair ? ((i >> 1) & 1) : (i % nslot_np); };
    auto q_off = [&](int i) { return pair ? (i & 1) * 128 * 128 : 0; };
    auto load_parity = [&](int i) { return (uint32_t)((pair ? (i >> 2) : (i / nslot_np)) & 1); };
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + nslot * slot_bytes);
    uint64_t* bar_load = bars;            // [3] TMA transaction barriers, one per operand slot
    uint64_t* bar_s = bars + 3;           // [2] S = QK^T complete (per group / TMEM half)
    uint64_t* bar_p = bars + 5;           // [2] P written to TMEM by the group's 4 warps
    uint64_t* bar_o = bars + 7;           // [2] O = PV complete
    uint64_t* bar_tfree = bars + 9;       // [2] O read out of TMEM by the group's 4 warps: the TMEM half may take the next S
    uint64_t* bar_free = bars + 11;       // [2] O stored: the operand slot (its Q tile doubles as store staging) may be refilled
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
    float* cls_stage = reinterpret_cast<float*>(smem + nslot * slot_bytes + 192);   // [2 groups][CLS_STAGE2] unnormalised CLS-row p
    uint8_t* ones = smem + nslot * slot_bytes + ONES_OFF;                           // [16 keys x 128 B] of 1.0

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int d = H * DH;
    // n_items counts scheduling units: items, or heads (= two items) in pair mode
    const int n_units = (n_items > (int)blockIdx.x) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int n_mine = pair ? 2 * n_units : n_units;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        tma_prefetch_desc(&tmap_o);
        for (int i = 0; i < NSLOT; ++i) mbar_init(&bar_load[i], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_s[i], 1); mbar_init(&bar_p[i], 4); mbar_init(&bar_o[i], 1);
            mbar_init(&bar_tfree[i], 4); mbar_init(&bar_free[i], 4);
        }
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    if (warp == 2) {
        const uint32_t one2 = F16 ? 0x3C003C00u : 0x3F803F80u;
        for (int k = lane; k < 2048 / 16; k += 32) reinterpret_cast<uint4*>(ones)[k] = make_uint4(one2, one2, one2, one2);
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_trigger();
    pdl_wait();

    // register reallocation between warpgroups (each setmaxnreg sits at the top of its role's branch: ptxas budgets
    // registers per dominated region)
    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // 4 x 40 + 8 x 232 = 12 x 168
        if (warp == 0) {
            // ---- loader ----
            if (lane == 0 && pair) {
                for (int p = 0; p < n_units; ++p) {
                    // slot p%2 was last used by head p-2: both of its items must have stored their O
                    if (p >= 2) { mbar_wait(&bar_free[0], (uint32_t)((p - 2) & 1)); mbar_wait(&bar_free[1], (uint32_t)((p - 2) & 1)); }
                    const int sh = (int)blockIdx.x + p * (int)gridDim.x, s = sh / H, h = sh % H;
                    uint8_t* Qs = smem + (p & 1) * slot_bytes;
                    uint64_t* bl = &bar_load[p & 1];
                    mbar_expect_tx(bl, (uint32_t)slot_bytes);
                    tma_load_2d(Qs, &tmap_q, h * DH, s * N, bl);
                    tma_load_2d(Qs + 128 * 128, &tmap_q, h * DH, s * N + 128, bl);
                    tma_load_2d(Qs + q_bytes, &tmap_kv, d + h * DH, s * N, bl);
                    tma_load_2d(Qs + q_bytes + nkp * 128, &tmap_kv, 2 * d + h * DH, s * N, bl);
                }
            } else if (lane == 0) {
                for (int i = 0; i < n_mine; ++i) {
                    // slot i%3 was last used by item i-3: operands dead after PV(i-3), staging (Q area) dead once its O is stored
                    if (i >= nslot_np) mbar_wait(&bar_free[(i - nslot_np) & 1], (uint32_t)(((i - nslot_np) >> 1) & 1));
                    const int id = (int)blockIdx.x + i * (int)gridDim.x;
                    const int qt = id % nqt, sh = id / nqt, s = sh / H, h = sh % H;
                    uint8_t* Qs = smem + (i % nslot_np) * slot_bytes;
                    uint64_t* bl = &bar_load[i % nslot_np];
                    mbar_expect_tx(bl, (uint32_t)slot_bytes);
                    tma_load_2d(Qs, &tmap_q, h * DH, s * N + qt * 128, bl);
                    tma_load_2d(Qs + 128 * 128, &tmap_kv, d + h * DH, s * N, bl);
                    tma_load_2d(Qs + 128 * 128 + nkp * 128, &tmap_kv, 2 * d + h * DH, s * N, bl);
                }
            }
        } else if (warp == 1) {
            // ---- MMA issuer (all 32 lanes run the control flow; elect.sync picks the issuing lane) ----
            const uint32_t idesc_o = attn_idesc(128, DH + ATTN_SUMN, F16, true);
            const int nks = nkp / 16;
            // S(i) is issued in two column ranges.  TMEM columns [0, 128) of the half only ever held P(i-2), which P.V(i-2) -- issued
            // before, and the tensor pipe executes in order -- has consumed: these key columns go out as soon as the operands have
            // landed.  Columns [128, NKP) are shared with O(i-2) and wait until the softmax group has read it.  (Before, the whole
            // S waited for O(i-2): 130..480 cycles of "S wait" per item in the phase trace; sequences of up to 128 keys -- the text
            // tower -- now never wait.)
            const int n_a = (s_early && nkp > 128) ? 128 : nkp, n_b = nkp - n_a;    // s_early = 0 (measurement switch): one range, after O(i-2)
            const uint32_t idesc_sa = attn_idesc(128, n_a, F16, false), idesc_sb = attn_idesc(128, n_b > 0 ? n_b : 16, F16, false);
            auto issue_s = [&](int i) {
                const int hf = i & 1;
                uint8_t* slot = smem + slot_of(i) * slot_bytes;
                mbar_wait(&bar_load[slot_of(i)], load_parity(i));
                if (!s_early && i >= 2) mbar_wait(&bar_tfree[hf], (uint32_t)(((i >> 1) - 1) & 1));
                tc_fence_after();
                const uint32_t thalf = tmem_base + hf * 256;
                const uint64_t qd = smem_desc_sw128(smem_u32(slot + q_off(i))), kd = smem_desc_sw128(smem_u32(slot + q_bytes));
#pragma unroll
                for (int k = 0; k < DH / 16; ++k) umma_ss_elect(thalf, qd + 2 * k, kd + 2 * k, idesc_sa, k != 0);
                if (n_b > 0) {
                    if (i >= 2) mbar_wait(&bar_tfree[hf], (uint32_t)(((i >> 1) - 1) & 1));     // O(i-2) has left columns [128, 208)
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < DH / 16; ++k)       // keys 128..: 128 rows x 128 B further into the K tile
                        umma_ss_elect(thalf + 128, qd + 2 * k, kd + 1024 + 2 * k, idesc_sb, k != 0);
                }
                umma_commit_elect(&bar_s[hf]);
            };
            auto issue_pv = [&](int j) {
                const int hf = j & 1;
                const uint32_t thalf = tmem_base + hf * 256;
                mbar_wait(&bar_p[hf], (uint32_t)((j >> 1) & 1));
                tc_fence_after();
                const uint32_t v_addr = smem_u32(smem + slot_of(j) * slot_bytes + q_bytes + nkp * 128);
                const uint64_t vd = smem_desc_sw128(v_addr);
                // the second 64-column atom of B (leading byte offset, 16-byte units, bits 16..29) is the block of ones
                const uint64_t lbo0 = (uint64_t)(((smem_u32(ones) - v_addr) >> 4) - 1u) << 16;   // the descriptor already holds LBO = 1
#pragma unroll
                for (int ks = 0; ks < 13; ++ks)                        // NKP <= 208: at most 13 k-steps
                    if (ks < nks) umma_ts_elect(thalf + O_COL, thalf + ks * 8, vd + (uint64_t)(ks * 128) + lbo0 - ((uint64_t)(ks * 128) << 16),
                                                idesc_o, ks != 0);
                umma_commit_elect(&bar_o[hf]);
            };
            // steady-state event order of the two (out-of-phase) groups: P(j), tfree(j), P(j+1), tfree(j+1), ...
            if (n_mine > 0) issue_s(0);
            if (n_mine > 1) issue_s(1);
            for (int j = 0; j < n_mine; ++j) {
                issue_pv(j);
                if (j + 2 < n_mine) issue_s(j + 2);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ---- softmax group g: owns TMEM half g and every second item ----
        const int g = (warp - 4) >> 2;
        const int q = warp & 3;                                        // TMEM lane quarter
        const int row = q * 32 + lane;
        const uint32_t trow = tmem_base + g * 256 + ((uint32_t)(q * 32) << 16);
        // (q-tile, head, sequence) of the group's current item, advanced incrementally (no divisions inside the loop)
        // pair mode: the group keeps q-tile g and walks the CTA's heads; otherwise item ids b + i * grid, i = g, g + 2, ...
        const int step2 = 2 * (int)gridDim.x;
        const int step_qt = pair ? 0 : step2 % nqt, step_sh = pair ? (int)gridDim.x : step2 / nqt, step_h = step_sh % H, step_s = step_sh / H;
        const int id0 = (int)blockIdx.x + g * (int)gridDim.x;
        int qt = pair ? g : id0 % nqt, h = pair ? (int)blockIdx.x % H : (id0 / nqt) % H, s = pair ? (int)blockIdx.x / H : (id0 / nqt) / H;
        for (int i = g; i < n_mine; i += 2) {
            const uint32_t par = (uint32_t)((i >> 1) & 1);
            const int grow = qt * 128 + row;
            const bool warp_active = qt * 128 + q * 32 < N;            // else: all 32 rows of this warp are padding
            const bool cls_warp = (probe_mode == PROBE_CLS_ROW) && qt == 0 && q == 0;
            // CLS probe: lane 0 (query row 0) stages its unnormalised probabilities in shared memory during the exp2 pass; the
            // whole warp normalises them and writes the row to global memory with coalesced stores once l is known
            const uint32_t cls_out = cls_warp ? smem_u32(cls_stage + g * CLS_STAGE2) : 0u;
#ifdef TAPCLIP_ATTN_TRACE
            long long* tr = (blockIdx.x == 0 && lane == 0 && (i >> 1) < 16) ? g_attn_trace + ((g * 4 + q) * 16 + (i >> 1)) * 16 : nullptr;
#endif
            TRACE(0);
            mbar_wait(&bar_s[g], par);
            TRACE(3);
            tc_fence_after();
            float l = 1.f, p_last = 0.f;
            if (warp_active) {
                // The S row is read from TMEM once and kept in registers for both the max and the exp2 pass -- except, for
                // NCH = 13, its last 3 chunks, which are read twice (208 + ~60 registers is more than a thread can have).
                // P chunk u overwrites S columns [8u, 8u+8), which never reach the re-read chunks [160, 208).
                constexpr int NRES = NCH < ATTN_NRES ? NCH : ATTN_NRES;                       // register-resident chunks
                constexpr int NTAIL = NCH - NRES;                              // chunks read twice
                constexpr int FIRST_MASKABLE = NCH == 13 ? 8 : NCH == 8 ? 4 : 0;   // N > 16 * FIRST_MASKABLE for this instance
                float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
                if constexpr (NTAIL > 0) {
                    uint32_t t[NTAIL > 0 ? NTAIL : 1][16];
#pragma unroll
                    for (int v = 0; v < NTAIL; ++v) tmem_ld_32x16(trow + (NRES + v) * 16, t[v]);
                    tmem_ld_wait();
#pragma unroll
                    for (int v = 0; v < NTAIL; ++v) chunk_max<true>(t[v], NRES + v, N, m0, m1, m2, m3);
                    // scheduling fence: the tail chunks' registers die here, before the resident chunks land
                    asm volatile("" : "+f"(m0), "+f"(m1), "+f"(m2), "+f"(m3) :: "memory");
                }
                uint32_t r[NRES][16];
#pragma unroll
                for (int u = 0; u < NRES; ++u) tmem_ld_32x16(trow + u * 16, r[u]);
                tmem_ld_wait();
                TRACE(4);
#pragma unroll
                for (int u = 0; u < NRES; ++u) {
                    if constexpr (true) {
                        if (u < FIRST_MASKABLE) chunk_max<false>(r[u], u, N, m0, m1, m2, m3);
                        else chunk_max<true>(r[u], u, N, m0, m1, m2, m3);
                    }
                }
                const float mneg = -fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * scale_log2;
                TRACE(6);
                float l0 = 0.f, l1 = 0.f;
#pragma unroll
                for (int u = 0; u < NRES; ++u) {
                    if (u < FIRST_MASKABLE) chunk_exp<T16, false, EXACT>(r[u], u, N, scale_log2, mneg, trow, cls_out, lane == 0, l0, l1, p_last);
                    else chunk_exp<T16, true, EXACT>(r[u], u, N, scale_log2, mneg, trow, cls_out, lane == 0, l0, l1, p_last);
                }
                if constexpr (NTAIL > 0) {
                    uint32_t t[NTAIL > 0 ? NTAIL : 1][16];
#pragma unroll
                    for (int v = 0; v < NTAIL; ++v) tmem_ld_32x16(trow + (NRES + v) * 16, t[v]);
                    tmem_ld_wait();
#pragma unroll
                    for (int v = 0; v < NTAIL; ++v) chunk_exp<T16, true, EXACT>(t[v], NRES + v, N, scale_log2, mneg, trow, cls_out, lane == 0, l0, l1, p_last);
                }
                if constexpr (EXACT) {
                    l = l0 + l1;
                    if (lse_out && grow < N) lse_out[((int64_t)s * H + h) * N + grow] = log2f(l) - mneg;     // rollout statistics
                }
                tmem_st_wait();
            }
            TRACE(5);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_p[g]);                     // this warp's 32 rows of P are in TMEM
            if (warp_active) {
                if (EXACT && probe_mode == PROBE_TEXT_COL && grow < probe_P) probe_out[((int64_t)s * H + h) * probe_P + grow] = p_last * __frcp_rn(l);
                if (cls_warp) {
                    __syncwarp();                                          // lane 0's staged row is visible to the warp
                    const float* st = cls_stage + g * CLS_STAGE2;
                    float part = 0.f;
                    for (int key = lane; key < N; key += 32) part += st[key];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    const float inv0 = __frcp_rn(part);
                    float* cls_gl = probe_out + (int64_t)s * probe_seq_stride + (int64_t)h * N;
                    for (int key = lane; key < N; key += 32) cls_gl[key] = st[key] * inv0;
                    __syncwarp();                                          // staging may be overwritten by the group's next item
                }
            }
            TRACE(7);
            mbar_wait(&bar_o[g], par);
            TRACE(8);
            tc_fence_after();
            uint32_t o[4][16];
            float inv = 1.f;
            if (warp_active) {
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_32x16(trow + O_COL + c * 16, o[c]);
                uint32_t lsum;
                tmem_ld_32x1(trow + O_COL + DH, lsum);                // row sum: the ones column of the product
                tmem_ld_wait();
                inv = __frcp_rn(__uint_as_float(lsum));
            }
            TRACE(10);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tfree[g]);                 // O is in registers: the next S may overwrite this TMEM half
            if (lane == 0) mbar_arrive(&bar_free[g]);                  // P.V has read K/V (and S = Q K^T the Q tile): the operand slot may be refilled
            if (warp_active) {
                // O row (64 fp32 columns) -> scaled 16-bit -> this warp's 4 KB staging tile in the 128-byte-swizzle layout ->
                // ONE TMA store of [32 rows x 128 B] (rows past N are clipped by the [S][N][d] map).  The register -> shared ->
                // register -> global version spent 850 of the drain's 1500 cycles in its 8 LDS + 8 STG per warp, queued in the
                // sub-partition's memory pipe behind the other group's MUFU stream.
                const uint32_t stage = smem_u32(smem + nslot * slot_bytes + OSTAGE_OFF) + (uint32_t)(warp - 4) * 4096u;
                if (lane == 0) tma_store_wait_read<0>();               // the previous item's store has read the tile
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t* rr = &o[c >> 1][(c & 1) * 8];
                    const uint32_t x0 = pack2<T16>(__uint_as_float(rr[0]) * inv, __uint_as_float(rr[1]) * inv);
                    const uint32_t x1 = pack2<T16>(__uint_as_float(rr[2]) * inv, __uint_as_float(rr[3]) * inv);
                    const uint32_t x2 = pack2<T16>(__uint_as_float(rr[4]) * inv, __uint_as_float(rr[5]) * inv);
                    const uint32_t x3 = pack2<T16>(__uint_as_float(rr[6]) * inv, __uint_as_float(rr[7]) * inv);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (uint32_t)lane * 128u + (((uint32_t)c ^ ((uint32_t)lane & 7u)) << 4)),
                                 "r"(x0), "r"(x1), "r"(x2), "r"(x3) : "memory");
                }
                TRACE(11);
                fence_proxy_async_smem();                              // generic-proxy writes -> visible to the TMA engine
                __syncwarp();
                TRACE(12);
                if (lane == 0) {
                    tma_store_3d(&tmap_o, stage, h * DH, qt * 128 + q * 32, s);
                    tma_store_commit();
                }
                TRACE(13);
            }
            TRACE(9);
            qt += step_qt;
            if (qt >= nqt) { qt -= nqt; ++h; }
            h += step_h;
            if (h >= H) { h -= H; ++s; }
            s += step_s;
        }
        if (lane == 0) tma_store_wait_read<0>();                       // shared memory outlives the last store's read
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// Long sequences (N > 208, e.g. ViT-L/14@336: 577 tokens): the same warp-specialised persistent structure with a loop over
// 128-key blocks and an online softmax (flash attention).  Per (item, key block j):
//   MMA   S_j = Q K_j^T (128 x 128) into the group's TMEM half           [after P_{j-1} has been consumed by PV_{j-1}]
//   softmax  one thread per query row: S_j row -> registers, running max m and sum l, alpha = 2^((m_old - m_new) c);
//            P_j = 2^(S_j c - m_new c) -> 16 bit -> TMEM over the S columns;  O *= alpha in TMEM (skipped when no row of the
//            warp moved its max);  l = l alpha + sum(P_j)
//   MMA   O += P_j V_j (TS-MMA, accumulates in TMEM columns [128, 192))
// and O / l is written once per item.  K/V blocks stream through a 5-stage TMA ring in the order the MMA warp consumes them:
// the two groups' items advance in lock step, (A, j), (B, j), (A, j+1), ...
// ------------------------------------------------------------------------------------------------------------------
constexpr int KVB = 128;               // keys per block
constexpr int KV_STAGES = 5;
constexpr int KV_STAGE_BYTES = 2 * KVB * 128;
constexpr int ATTN3_SMEM = 2 * 128 * 128 + KV_STAGES * KV_STAGE_BYTES + 256 + 1024;   // 194.25 KB with the reserved KB: the 196 KB carve-out
constexpr int CLS_STAGE3 = 1024;       // CLS probe launches only: [2 groups][1024] floats of staging behind the barriers (N <= 1024);
constexpr int ATTN3_SMEM_CLS = ATTN3_SMEM + 2 * CLS_STAGE3 * 4;   // measured at N = 577: +7 % without a probe (next carve-out step), so opt-in

template <bool F16>
__global__ void __launch_bounds__(ATTN2_THREADS, 1)
attn_fwd_tc_kv_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv, void* __restrict__ out_,
                      int N, int H, int nqt, int n_items, float scale_log2, int probe_mode, float* __restrict__ probe_out, int probe_P,
                      int64_t probe_seq_stride, float* __restrict__ lse_out) {
    using T16 = typename std::conditional<F16, f16, bf16>::type;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* qbuf = smem;                                   // [2 groups][128 x 128 B]; doubles as the O staging of the group
    uint8_t* kvbuf = smem + 2 * 128 * 128;                  // [KV_STAGES][K block | V block]
    uint64_t* bars = reinterpret_cast<uint64_t*>(kvbuf + KV_STAGES * KV_STAGE_BYTES);
    uint64_t* q_full = bars;              // [2]
    uint64_t* q_free = bars + 2;          // [2] the group's 4 warps are done with the Q buffer (O stored)
    uint64_t* kv_full = bars + 4;         // [KV_STAGES]
    uint64_t* kv_empty = bars + 4 + KV_STAGES;   // [KV_STAGES] PV of the block complete
    uint64_t* bar_s = bars + 4 + 2 * KV_STAGES;  // [2]
    uint64_t* bar_p = bar_s + 2;          // [2] count 4
    uint64_t* bar_o = bar_s + 4;          // [2]
    uint64_t* bar_tfree = bar_s + 6;      // [2] count 4: O of the finished item has left TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 8);
    float* mblk_s = reinterpret_cast<float*>(kvbuf + KV_STAGES * KV_STAGE_BYTES + 192);   // [2 groups][8]: the CLS row's max at each key block
    float* cls_stage = reinterpret_cast<float*>(kvbuf + KV_STAGES * KV_STAGE_BYTES + 256);   // only present (and touched) with PROBE_CLS_ROW

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int d = H * DH;
    const int nkb = (N + KVB - 1) / KVB;
    const int n_mine = (n_items > (int)blockIdx.x) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1); mbar_init(&q_free[i], 4); mbar_init(&bar_s[i], 1); mbar_init(&bar_p[i], 4);
            mbar_init(&bar_o[i], 1); mbar_init(&bar_tfree[i], 4);
        }
        for (int i = 0; i < KV_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_trigger();
    pdl_wait();

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0) {
            // ---- loader: Q of the pair's items, then their K/V blocks in consumption order ----
            if (lane == 0) {
                int stage = 0; uint32_t phase = 0;
                for (int i0 = 0; i0 < n_mine; i0 += 2) {
                    for (int g = 0; g < 2; ++g) {
                        const int i = i0 + g;
                        if (i >= n_mine) break;
                        if (i >= 2) mbar_wait(&q_free[g], (uint32_t)(((i >> 1) - 1) & 1));
                        const int id = (int)blockIdx.x + i * (int)gridDim.x;
                        const int qt = id % nqt, sh = id / nqt, s = sh / H, h = sh % H;
                        mbar_expect_tx(&q_full[g], 128 * 128);
                        tma_load_2d(qbuf + g * 128 * 128, &tmap_q, h * DH, s * N + qt * 128, &q_full[g]);
                    }
                    for (int j = 0; j < nkb; ++j) {
                        for (int g = 0; g < 2; ++g) {
                            const int i = i0 + g;
                            if (i >= n_mine) break;
                            const int id = (int)blockIdx.x + i * (int)gridDim.x;
                            const int sh = id / nqt, s = sh / H, h = sh % H;
                            mbar_wait(&kv_empty[stage], phase ^ 1);
                            mbar_expect_tx(&kv_full[stage], KV_STAGE_BYTES);
                            uint8_t* kb = kvbuf + stage * KV_STAGE_BYTES;
                            tma_load_2d(kb, &tmap_kv, d + h * DH, s * N + j * KVB, &kv_full[stage]);
                            tma_load_2d(kb + KVB * 128, &tmap_kv, 2 * d + h * DH, s * N + j * KVB, &kv_full[stage]);
                            if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        } else if (warp == 1) {
            // ---- MMA issuer (warp-uniform control flow, elect.sync issue) ----
            const uint32_t idesc_s = attn_idesc(128, KVB, F16, false), idesc_o = attn_idesc(128, DH, F16, true);
            int s_stage = 0; uint32_t s_phase = 0;         // cursor of the next S to issue in the K/V ring
            int p_stage = 0;                               // cursor of the next PV
            uint32_t n_blk[2] = {0u, 0u};                  // blocks issued so far per group (parities of bar_s / bar_p / bar_o)
            for (int i0 = 0; i0 < n_mine; i0 += 2) {
                const int npres = (i0 + 1 < n_mine) ? 2 : 1;
                for (int g = 0; g < npres; ++g) {
                    const int i = i0 + g;
                    mbar_wait(&q_full[g], (uint32_t)((i >> 1) & 1));
                    if (i >= 2) mbar_wait(&bar_tfree[g], (uint32_t)(((i >> 1) - 1) & 1));     // the previous item's O has left this TMEM half
                }
                for (int j = 0; j < nkb; ++j) {
                    for (int g = 0; g < npres; ++g) {
                        const uint32_t thalf = tmem_base + g * 256;
                        mbar_wait(&kv_full[s_stage], s_phase);
                        if (j > 0) mbar_wait(&bar_o[g], (n_blk[g] - 1) & 1u);                   // PV_{j-1} done: P_{j-1} consumed, O consistent
                        tc_fence_after();
                        const uint64_t qd = smem_desc_sw128(smem_u32(qbuf + g * 128 * 128));
                        const uint64_t kd = smem_desc_sw128(smem_u32(kvbuf + s_stage * KV_STAGE_BYTES));
#pragma unroll
                        for (int k = 0; k < DH / 16; ++k) umma_ss_elect(thalf, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
                        umma_commit_elect(&bar_s[g]);
                        if (++s_stage == KV_STAGES) { s_stage = 0; s_phase ^= 1; }
                    }
                    for (int g = 0; g < npres; ++g) {
                        const uint32_t thalf = tmem_base + g * 256;
                        mbar_wait(&bar_p[g], n_blk[g] & 1u);
                        tc_fence_after();
                        const uint64_t vd = smem_desc_sw128(smem_u32(kvbuf + p_stage * KV_STAGE_BYTES + KVB * 128));
#pragma unroll
                        for (int ks = 0; ks < KVB / 16; ++ks)
                            umma_ts_elect(thalf + O_COL, thalf + ks * 8, vd + (uint64_t)(ks * 128), idesc_o, (j | ks) != 0);
                        umma_commit_elect(&bar_o[g]);
                        umma_commit_elect(&kv_empty[p_stage]);                                // the ring stage may be refilled
                        if (++p_stage == KV_STAGES) p_stage = 0;
                        ++n_blk[g];
                    }
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        // ---- softmax group g ----
        const int g = (warp - 4) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t trow = tmem_base + g * 256 + ((uint32_t)(q * 32) << 16);
        const int rd_row = lane >> 3, rd_ch = lane & 7;
        const int step2 = 2 * (int)gridDim.x;
        const int step_qt = step2 % nqt, step_sh = step2 / nqt, step_h = step_sh % H, step_s = step_sh / H;
        const int id0 = (int)blockIdx.x + g * (int)gridDim.x;
        int qt = id0 % nqt, h = (id0 / nqt) % H, s = (id0 / nqt) / H;
        uint32_t blk = 0;                                              // blocks processed by this group (barrier parities)
        uint8_t* Qs = qbuf + g * 128 * 128;
        for (int i = g; i < n_mine; i += 2) {
            const int grow = qt * 128 + row;
            const bool warp_active = qt * 128 + q * 32 < N;
            const bool cls_warp = (probe_mode == PROBE_CLS_ROW) && qt == 0 && q == 0;
            const bool cls_thread = cls_warp && lane == 0;
            // CLS probe: lane 0 stages p (relative to each block's running max) in shared memory during the pass; the whole warp
            // rescales, normalises and stores the row coalesced at the end of the item
            const uint32_t cls_out = cls_warp ? smem_u32(cls_stage + g * CLS_STAGE3) : 0u;
            float m_run = -INFINITY, l = 0.f, p_last = 0.f;
            for (int j = 0; j < nkb; ++j, ++blk) {
                mbar_wait(&bar_s[g], blk & 1u);
                tc_fence_after();
                if (warp_active) {
                    const int n_eff = N - j * KVB;                     // valid keys of this block (>= 1; > 128 for interior blocks)
                    uint32_t r[8][16];
#pragma unroll
                    for (int u = 0; u < 8; ++u) tmem_ld_32x16(trow + u * 16, r[u]);
                    tmem_ld_wait();
                    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                    for (int u = 0; u < 8; ++u) chunk_max<true>(r[u], u, n_eff, m0, m1, m2, m3);
                    const float m_new = fmaxf(m_run, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
                    const float alpha = fast_exp2((m_run - m_new) * scale_log2);       // 0 for the first block (m_run = -inf)
                    const float mneg = -m_new * scale_log2;
                    float l0 = 0.f, l1 = 0.f, pl = 0.f;
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        chunk_exp<T16, true, true>(r[u], u, n_eff, scale_log2, mneg, trow, cls_out ? cls_out + (uint32_t)(j * KVB) * 4u : 0u, lane == 0, l0, l1, pl);
                    if (j == nkb - 1) p_last = pl;
                    l = fmaf(l, alpha, l0 + l1);
                    m_run = m_new;
                    if (cls_thread) mblk_s[g * 8 + (j & 7)] = m_new;          // (shared memory: a dynamically indexed register array lives in local memory)
                    if (j > 0 && !__all_sync(0xffffffffu, alpha == 1.f)) {
                        // O *= alpha (PV_{j-1} is complete: S_j was only issued after it)
                        uint32_t o[4][16];
#pragma unroll
                        for (int c = 0; c < 4; ++c) tmem_ld_32x16(trow + O_COL + c * 16, o[c]);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            uint32_t w8[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) w8[e] = __float_as_uint(__uint_as_float(o[c >> 1][(c & 1) * 8 + e]) * alpha);
                            tmem_st_32x8(trow + O_COL + c * 8, w8);
                        }
                    }
                    tmem_st_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_p[g]);
            }
            // ---- the item's O is complete after the last PV ----
            mbar_wait(&bar_o[g], (blk - 1) & 1u);
            tc_fence_after();
            const float inv = __frcp_rn(warp_active ? l : 1.f);
            uint32_t o[4][16];
            if (warp_active) {
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_32x16(trow + O_COL + c * 16, o[c]);
                tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tfree[g]);
            if (warp_active) {
                if (probe_mode == PROBE_TEXT_COL && grow < probe_P) probe_out[((int64_t)s * H + h) * probe_P + grow] = p_last * inv;
                if (lse_out && grow < N) lse_out[((int64_t)s * H + h) * N + grow] = fmaf(m_run, scale_log2, log2f(l));   // rollout statistics
                if (cls_warp) {
                    __syncwarp();                                      // lane 0's staged row is visible to the warp
                    const float inv0 = __shfl_sync(0xffffffffu, inv, 0), m0 = __shfl_sync(0xffffffffu, m_run, 0);
                    float fb[8];                                       // per key block: 2^((m_block - m_final) c) / l
#pragma unroll
                    for (int b = 0; b < 8; ++b) fb[b] = b < nkb ? fast_exp2((mblk_s[g * 8 + b] - m0) * scale_log2) * inv0 : 0.f;
                    float* cls_gl = probe_out + (int64_t)s * probe_seq_stride + (int64_t)h * N;
#pragma unroll
                    for (int b = 0; b < 8; ++b)
                        if (b < nkb)
                            for (int key = b * KVB + lane; key < min(N, (b + 1) * KVB); key += 32) cls_gl[key] = cls_stage[g * CLS_STAGE3 + key] * fb[b];
                    __syncwarp();                                      // staging may be overwritten by the group's next item
                }
                const uint32_t stage = smem_u32(Qs) + (uint32_t)(q * 32) * 128u;      // the Q tile is dead: every S of the item is done
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t* rr = &o[c >> 1][(c & 1) * 8];
                    const uint32_t x0 = pack2<T16>(__uint_as_float(rr[0]) * inv, __uint_as_float(rr[1]) * inv);
                    const uint32_t x1 = pack2<T16>(__uint_as_float(rr[2]) * inv, __uint_as_float(rr[3]) * inv);
                    const uint32_t x2 = pack2<T16>(__uint_as_float(rr[4]) * inv, __uint_as_float(rr[5]) * inv);
                    const uint32_t x3 = pack2<T16>(__uint_as_float(rr[6]) * inv, __uint_as_float(rr[7]) * inv);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (uint32_t)lane * 128u + (((uint32_t)c ^ ((uint32_t)lane & 7u)) << 4)),
                                 "r"(x0), "r"(x1), "r"(x2), "r"(x3) : "memory");
                }
                __syncwarp();
                uint8_t* gbase = reinterpret_cast<uint8_t*>(out_) + (((int64_t)s * N + qt * 128 + q * 32) * d + h * DH) * 2 + rd_ch * 16;
                const int rows_left = N - (qt * 128 + q * 32);
                uint4 x[8];
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + rd_row;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x[it].x), "=r"(x[it].y), "=r"(x[it].z), "=r"(x[it].w)
                                 : "r"(stage + (uint32_t)rr * 128u + (((uint32_t)rd_ch ^ ((uint32_t)rr & 7u)) << 4)) : "memory");
                }
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + rd_row;
                    if (rr < rows_left) *reinterpret_cast<uint4*>(gbase + (int64_t)rr * d * 2) = x[it];
                }
                fence_proxy_async_smem();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&q_free[g]);
            qt += step_qt;
            if (qt >= nqt) { qt -= nqt; ++h; }
            h += step_h;
            if (h >= H) { h -= H; ++s; }
            s += step_s;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

#ifdef TAPCLIP_ATTN_TRACE
extern "C" __attribute__((visibility("default"))) int tapclip_debug_attn_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, g_attn_trace, sizeof(long long) * 8 * 16 * 16);
}
#endif

bool attention_fwd_tc_supported(int dt, int N) { return (dt == DT_BF16 || dt == DT_F16) && N >= 1 && N <= 1024; }

bool attention_fwd_tc(const void* qkv, void* out, int dt, int S, int N, int H, const AttnProbe& probe, cudaStream_t stream) {
    TC_CHECK(attention_fwd_tc_supported(dt, N), "tcgen05 attention supports 16-bit inputs and N <= 1024");

    const int d = H * DH;
    const int nkp = (int)round_up(N, 16);
    const int nqt = (int)ceil_div(probe.live_q_rows > 0 ? std::min(N, probe.live_q_rows) : N, 128);   // query tiles actually computed
    const bool f16 = dt == DT_F16;
    const CUtensorMapDataType tdt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const CUtensorMap& tq = make_tmap(qkv, tdt, 2, (int64_t)S * N, 3 * d, 3 * d, 128, 64);
    const float sl2 = 0.125f * 1.4426950408889634f;
    if (nkp > 256) {
        // flash-style kernel: 128-key blocks through a TMA ring
        const CUtensorMap& tkb = make_tmap(qkv, tdt, 2, (int64_t)S * N, 3 * d, 3 * d, KVB, 64);
        if (f16) ensure_dynamic_smem((const void*)attn_fwd_tc_kv_kernel<true>, ATTN3_SMEM_CLS);
        else ensure_dynamic_smem((const void*)attn_fwd_tc_kv_kernel<false>, ATTN3_SMEM_CLS);
        const size_t smem3 = probe.mode == PROBE_CLS_ROW ? ATTN3_SMEM_CLS : ATTN3_SMEM;
        const int num_sms3 = device_sm_count();
        const int n_items = S * H * nqt;
        const unsigned grid3 = (unsigned)std::min(n_items, num_sms3);
        if (f16) launch_pdl(attn_fwd_tc_kv_kernel<true>, grid3, ATTN2_THREADS, smem3, stream, tq, tkb, out, N, H, nqt, n_items, sl2, probe.mode, probe.out, probe.P, probe.seq_stride, probe.lse_out);
        else launch_pdl(attn_fwd_tc_kv_kernel<false>, grid3, ATTN2_THREADS, smem3, stream, tq, tkb, out, N, H, nqt, n_items, sl2, probe.mode, probe.out, probe.P, probe.seq_stride, probe.lse_out);
        TC_LAUNCH_CHECK();
        return true;
    }
    const CUtensorMap& tkv = make_tmap(qkv, tdt, 2, (int64_t)S * N, 3 * d, 3 * d, nkp, 64);
    if (nkp <= 208) {
        // persistent pipelined kernel: 3 operand slots of (Q 16 KB + K + V) fit in shared memory; with two q-tiles per head
        // (pair mode) 2 slots of (Q0 + Q1 + K + V), so that a head's K and V are loaded once
        static const int pair_env = getenv("TAPCLIP_ATTN_PAIR") ? atoi(getenv("TAPCLIP_ATTN_PAIR")) : 1;    // 0: measurement switch
        const int pair = (nqt == 2 && pair_env != 0) ? 1 : 0;
        static const int s_early = getenv("TAPCLIP_ATTN_SEARLY") ? atoi(getenv("TAPCLIP_ATTN_SEARLY")) : 1;      // 0: measurement switch
        const size_t slot1 = 128 * 128 + 2 * (size_t)nkp * 128, tail = OSTAGE_OFF + 8 * 4096 + 1024;
        const int nslot_np = NSLOT * slot1 + tail <= 227 * 1024 ? NSLOT : 2;      // e.g. the one-q-tile launch of a ViT-B/16 CLS-only last layer (N = 197): 2
        const size_t slots = pair ? 2 * (2 * 128 * 128 + 2 * (size_t)nkp * 128) : nslot_np * slot1;
        const size_t smem2 = slots + tail;   // + barriers/TMEM slot/CLS staging, ones block, O staging, alignment slack
        const CUtensorMap& to = make_tmap_seq(out, tdt, 2, S, N, d, d, 32, 64);   // + barriers/TMEM slot, CLS staging, alignment slack
        const int num_sms = device_sm_count();
        const int n_items = pair ? S * H : S * H * nqt;        // scheduling units
        const unsigned grid2 = (unsigned)std::min(n_items, num_sms);
        // softmax chunk count instance: 4 (N <= 64, ViT-B/32), 8 (N <= 128, the text tower), 13 (N <= 208, ViT-B/16)
        const int nch = nkp / 16;
        auto go = [&](auto kern) {
            ensure_dynamic_smem((const void*)kern, smem2);
            launch_pdl(kern, grid2, ATTN2_THREADS, smem2, stream, tq, tkv, to, out, N, H, nkp, nqt, n_items, sl2, probe.mode, probe.out, probe.P, probe.seq_stride, probe.lse_out, pair, nslot_np, s_early);
        };
        const bool exact = probe.mode == PROBE_TEXT_COL || probe.lse_out != nullptr;     // launches that publish per-row statistics
        auto pick = [&](auto f16_c, auto nch_c) {
            if (exact) go(attn_fwd_tc2_kernel<decltype(f16_c)::value, decltype(nch_c)::value, true>);
            else go(attn_fwd_tc2_kernel<decltype(f16_c)::value, decltype(nch_c)::value, false>);
        };
        using std::integral_constant;
        if (nch <= 4) { if (f16) pick(std::true_type{}, integral_constant<int, 4>{}); else pick(std::false_type{}, integral_constant<int, 4>{}); }
        else if (nch <= 8) { if (f16) pick(std::true_type{}, integral_constant<int, 8>{}); else pick(std::false_type{}, integral_constant<int, 8>{}); }
        else { if (f16) pick(std::true_type{}, integral_constant<int, 13>{}); else pick(std::false_type{}, integral_constant<int, 13>{}); }
        TC_LAUNCH_CHECK();
        return true;
    }
    const size_t smem = 128 * 128 + 2 * (size_t)nkp * 128 + 64 + 1024;
    if (f16) ensure_dynamic_smem((const void*)attn_fwd_tc_kernel<true>, smem);
    else ensure_dynamic_smem((const void*)attn_fwd_tc_kernel<false>, smem);
    const unsigned grid = (unsigned)((int64_t)S * H * nqt);
    if (f16) launch_pdl(attn_fwd_tc_kernel<true>, grid, ATTN_THREADS, smem, stream, tq, tkv, out, N, H, nkp, nqt, sl2, probe.mode, probe.out, probe.P, probe.seq_stride);
    else launch_pdl(attn_fwd_tc_kernel<false>, grid, ATTN_THREADS, smem, stream, tq, tkv, out, N, H, nkp, nqt, sl2, probe.mode, probe.out, probe.P, probe.seq_stride);
    TC_LAUNCH_CHECK();
    return false;                                              // this variant does not emit probe.lse_out
}

}  // namespace tapclip
