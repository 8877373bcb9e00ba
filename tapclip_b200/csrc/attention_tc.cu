// K2 on tcgen05: fused attention forward for sequences of up to 256 tokens (ViT-B/16: 197, ViT-B/32: 50, text: P+77),
// head dim 64, no mask, with the same probe epilogue as attention.cu.
//
// One CTA per (sequence, head, 128-query tile):
//   TMA   Q tile [128 x 64], K and V [NKP x 64] (NKP = keys padded to 16) into 128B-swizzled smem
//   MMA1  S = Q K^T        tcgen05.mma  M=128, N=NKP, K=64  (A, B K-major from smem)  -> TMEM columns [0, NKP)
//   4 softmax warps: one thread per query row reads its S row from TMEM (two passes: max, then exp2/sum), writes
//         the 16-bit probabilities BACK INTO TMEM over the S columns (two keys per 32-bit column) and emits the probe
//   MMA2  O = P V          tcgen05.mma  M=128, N=64, K=NKP   (A = P from TMEM, B = V as stored: MN-major smem descriptor)
//   epilogue: O / rowsum -> 16-bit -> global
// S, P and O never touch shared memory or HBM; the N x N map is never materialised.
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

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Softmax of ONE query row held in TMEM (one thread = one TMEM lane).  Two passes over the S row (max, then exp2/sum),
// read in groups of up to 64 columns with 4 tcgen05.ld in flight per wait; the 16-bit probabilities are written back
// into TMEM over the S columns already consumed (two keys per 32-bit column).  Only the last group can contain padded
// keys, so interior groups are branch-free.  `cls_out` (warp-uniform non-null only for the warp holding the CLS row;
// lane `cls_lane` writes) receives the unnormalised probabilities of that row.  Returns the row sum; *p_last = p[N-1].
template <typename T16>
__device__ __forceinline__ float softmax_to_tmem(uint32_t trow, int N, int nkp, float scale_log2, float* cls_out, bool cls_lane,
                                                 float* p_last) {
    const int nch = nkp / 16;
    float mx = -INFINITY;
    for (int c0 = 0; c0 < nch; c0 += 4) {
        uint32_t r[4][16];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c0 + u < nch) tmem_ld_32x16(trow + (c0 + u) * 16, r[u]);
        tmem_ld_wait();
        if ((c0 + 4) * 16 <= N) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(r[u][j]));
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if ((c0 + u) * 16 + j < N) mx = fmaxf(mx, __uint_as_float(r[u][j]));
        }
    }
    const float mneg = -mx * scale_log2;
    float l = 0.f, pl = 0.f;
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
                    l += p0 + p1;
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
    return l;
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
// (sequence, head, q-tile) items.  Three smem operand slots and two TMEM halves let the control thread prefetch item
// i+1 and issue S(i) while one softmax group still works on item i-1; the two softmax groups (4 warps each) alternate
// items, so the exp/sum work of consecutive items overlaps and no per-item setup (TMEM alloc, barrier init) remains.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void group_sync(int g) {      // named barrier of one 128-thread softmax group
    if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
}

constexpr int ATTN2_THREADS = 320;     // warp 0 loader (TMA), warp 1 TMEM alloc, warps 2..9 two softmax groups of 4 warps
constexpr int NSLOT = 3;

template <bool F16>
__global__ void __launch_bounds__(ATTN2_THREADS, 1)
attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv, void* __restrict__ out_,
                    int N, int H, int nkp, int nqt, int n_items, float scale_log2, int probe_mode, float* __restrict__ probe_out,
                    int probe_P, int64_t probe_seq_stride) {
    using T16 = typename std::conditional<F16, f16, bf16>::type;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int slot_bytes = 128 * 128 + 2 * nkp * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSLOT * slot_bytes);
    uint64_t* bar_load = bars;            // [3] TMA transaction barriers, one per operand slot
    uint64_t* bar_s = bars + 3;           // [2] S = QK^T complete (per softmax group / TMEM half)
    uint64_t* bar_o = bars + 5;           // [2] O = PV complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = H * DH;
    const int n_mine = (n_items > (int)blockIdx.x) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        for (int i = 0; i < NSLOT; ++i) mbar_init(&bar_load[i], 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_s[i], 1); mbar_init(&bar_o[i], 1); }
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        // ---- loader: keeps up to three items of Q/K/V in flight ----
        if (lane == 0) {
            for (int i = 0; i < n_mine; ++i) {
                // slot i%3 was last used by item i-3; its V tile is dead once PV(i-3) has completed
                if (i >= NSLOT) mbar_wait(&bar_o[(i - NSLOT) & 1], (uint32_t)(((i - NSLOT) >> 1) & 1));
                const int id = (int)blockIdx.x + i * (int)gridDim.x;
                const int qt = id % nqt, sh = id / nqt, s = sh / H, h = sh % H;
                uint8_t* Qs = smem + (i % NSLOT) * slot_bytes;
                uint64_t* bl = &bar_load[i % NSLOT];
                mbar_expect_tx(bl, (uint32_t)slot_bytes);
                tma_load_2d(Qs, &tmap_q, h * DH, s * N + qt * 128, bl);
                tma_load_2d(Qs + 128 * 128, &tmap_kv, d + h * DH, s * N, bl);
                tma_load_2d(Qs + 128 * 128 + nkp * 128, &tmap_kv, 2 * d + h * DH, s * N, bl);
            }
        }
    } else if (warp >= 2) {
        // ---- softmax group g: owns TMEM half g and every second item; its first thread also issues the two MMAs ----
        const int g = (warp - 2) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const bool leader = ((warp - 2) & 3) == 0 && lane == 0;
        const uint32_t thalf = tmem_base + g * 256;
        const uint32_t trow = thalf + ((uint32_t)(q * 32) << 16);
        const uint32_t idesc_s = attn_idesc(128, nkp, F16, false), idesc_o = attn_idesc(128, DH, F16, true);
        for (int i = g; i < n_mine; i += 2) {
            const int id = (int)blockIdx.x + i * (int)gridDim.x;
            const int qt = id % nqt, sh = id / nqt, s = sh / H, h = sh % H;
            const uint32_t par = (uint32_t)((i >> 1) & 1);
            uint8_t* Qs = smem + (i % NSLOT) * slot_bytes;
            if (leader) {
                mbar_wait(&bar_load[i % NSLOT], (uint32_t)((i / NSLOT) & 1));
                tc_fence_after();
                const uint64_t qd = smem_desc_sw128(smem_u32(Qs)), kd = smem_desc_sw128(smem_u32(Qs + 128 * 128));
#pragma unroll
                for (int k = 0; k < DH / 16; ++k) umma_bf16(thalf, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
                umma_commit(&bar_s[g]);
            }
            const int grow = qt * 128 + row;
            const bool warp_active = qt * 128 + q * 32 < N;            // else: all 32 rows of this warp are padding
            mbar_wait(&bar_s[g], par);
            tc_fence_after();
            float l = 1.f, p_last = 0.f;
            const bool cls_warp = (probe_mode == PROBE_CLS_ROW) && qt == 0 && q == 0;
            float* cls_out = cls_warp ? probe_out + (int64_t)s * probe_seq_stride + (int64_t)h * N : nullptr;
            if (warp_active) {
                l = softmax_to_tmem<T16>(trow, N, nkp, scale_log2, cls_out, lane == 0, &p_last);
                tmem_st_wait();
            }
            tc_fence_before();
            group_sync(g);                                                     // P of all 128 rows is in TMEM
            if (leader) {
                tc_fence_after();
                const uint64_t vd = smem_desc_sw128(smem_u32(Qs + 128 * 128 + nkp * 128));
                for (int ks = 0; ks < nkp / 16; ++ks) umma_ts(thalf + O_COL, thalf + ks * 8, vd + (uint64_t)(ks * 128), idesc_o, ks != 0);
                umma_commit(&bar_o[g]);
            }
            const float inv = 1.f / l;
            if (warp_active) {
                if (probe_mode == PROBE_TEXT_COL && grow < probe_P) probe_out[((int64_t)s * H + h) * probe_P + grow] = p_last * inv;
                if (cls_warp && lane == 0)
                    for (int key = 0; key < N; ++key) cls_out[key] *= inv;      // own earlier writes
            }
            mbar_wait(&bar_o[g], par);
            tc_fence_after();
            if (warp_active) {
                uint4 o[8];
                load_o_row<T16>(trow, inv, o);
                if (grow < N) {
                    uint4* gp = reinterpret_cast<uint4*>(reinterpret_cast<T16*>(out_) + ((int64_t)s * N + grow) * d + h * DH);
#pragma unroll
                    for (int c = 0; c < 8; ++c) gp[c] = o[c];
                }
            }
            tc_fence_before();
            group_sync(g);                                                     // O drained: the next S may overwrite this half
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool attention_fwd_tc_supported(int dt, int N) { return (dt == DT_BF16 || dt == DT_F16) && N >= 1 && N <= 256; }

void attention_fwd_tc(const void* qkv, void* out, int dt, int S, int N, int H, const AttnProbe& probe, cudaStream_t stream) {
    TC_CHECK(attention_fwd_tc_supported(dt, N), "tcgen05 attention supports 16-bit inputs and N <= 256");

    const int d = H * DH;
    const int nkp = (int)round_up(N, 16), nqt = (int)ceil_div(N, 128);
    const bool f16 = dt == DT_F16;
    const CUtensorMapDataType tdt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const CUtensorMap& tq = make_tmap(qkv, tdt, 2, (int64_t)S * N, 3 * d, 3 * d, 128, 64);
    const CUtensorMap& tkv = make_tmap(qkv, tdt, 2, (int64_t)S * N, 3 * d, 3 * d, nkp, 64);
    const float sl2 = 0.125f * 1.4426950408889634f;
    if (nkp <= 208) {
        // persistent pipelined kernel: 3 operand slots of (Q 16 KB + K + V) fit in shared memory
        const size_t smem2 = NSLOT * (128 * 128 + 2 * (size_t)nkp * 128) + 128 + 1024;
        static size_t conf2[2] = {0, 0};
        if (smem2 > conf2[f16]) {
            if (f16) TC_CUDA(cudaFuncSetAttribute(attn_fwd_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            else TC_CUDA(cudaFuncSetAttribute(attn_fwd_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            conf2[f16] = smem2;
        }
        static int num_sms = 0;
        if (num_sms == 0) { int dev; TC_CUDA(cudaGetDevice(&dev)); TC_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev)); }
        const int n_items = S * H * nqt;
        const unsigned grid2 = (unsigned)std::min(n_items, num_sms);
        if (f16) launch_pdl(attn_fwd_tc2_kernel<true>, grid2, ATTN2_THREADS, smem2, stream, tq, tkv, out, N, H, nkp, nqt, n_items, sl2, probe.mode, probe.out, probe.P, probe.seq_stride);
        else launch_pdl(attn_fwd_tc2_kernel<false>, grid2, ATTN2_THREADS, smem2, stream, tq, tkv, out, N, H, nkp, nqt, n_items, sl2, probe.mode, probe.out, probe.P, probe.seq_stride);
        TC_LAUNCH_CHECK();
        return;
    }
    const size_t smem = 128 * 128 + 2 * (size_t)nkp * 128 + 64 + 1024;
    static size_t conf[2] = {0, 0};
    if (smem > conf[f16]) {
        if (f16) TC_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else TC_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conf[f16] = smem;
    }
    const unsigned grid = (unsigned)((int64_t)S * H * nqt);
    if (f16) launch_pdl(attn_fwd_tc_kernel<true>, grid, ATTN_THREADS, smem, stream, tq, tkv, out, N, H, nkp, nqt, sl2, probe.mode, probe.out, probe.P, probe.seq_stride);
    else launch_pdl(attn_fwd_tc_kernel<false>, grid, ATTN_THREADS, smem, stream, tq, tkv, out, N, H, nkp, nqt, sl2, probe.mode, probe.out, probe.P, probe.seq_stride);
    TC_LAUNCH_CHECK();
}

}  // namespace tapclip
