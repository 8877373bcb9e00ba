// K1 — tcgen05/TMEM bf16 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] * W[N,K]^T + bias).
//
// Replaces the cuBLAS `addmm` calls behind nn.Linear / nn.MultiheadAttention in_proj/out_proj of the
// reference's third-party model (SURVEY.md 2.3 rows k2,k5,k8,k10 + patch-embed + the two projections).
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0      TMA producer: 128B-swizzled A (128x64) and W tiles into a 4..8-stage smem ring
//   warp 1      TMEM allocator + tcgen05.mma issuer (fp32 accumulate in TMEM).  The single-CTA path runs the issue loop
//               warp-converged with elect.sync-predicated instructions, so the UTCHMMA operands live in uniform registers
//   warps 2..9  epilogue, two warps per TMEM lane quarter (alternating 128-byte column chunks): tcgen05.ld TMEM ->
//               registers -> bias (staged in smem) / activation / activation-gradient -> swizzled smem transpose ->
//               coalesced 16-byte streaming stores (red.global.add.v4.f32 for the fp32 residual stream)
// Accumulators are double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1.  M/N/K tails are handled by TMA zero-fill on loads and clipping on stores.
//
// Tile order: static striding (tile = cta, cta + grid, ...) or, with a scheduler slot (GemmExtra::sched), DYNAMIC: the
// producer lane draws tile indices from a global atomic counter and hands them to the MMA / epilogue warps through a
// 4-deep shared-memory ring.  When another stream's kernels hold part of the SMs (the text tower runs beside the image
// tower), late-starting CTAs then simply draw fewer tiles instead of leaving a static share for the end.
//
// LayerNorm never runs as its own kernel around these GEMMs (north-star "pre-LN epilogue"):
//   EPI_F32_RESID  the residual update x_new = x_old + A W^T + b also writes x_new in the 16-bit operand type and per-row
//                  partial (sum, sum of squares) of x_new (one pair per n-tile and epilogue-warp parity: written, never
//                  accumulated, so there is no zeroing pass and the result is deterministic);
//   FOLD           the consumer GEMM multiplies the UN-normalised 16-bit rows with the gamma-scaled, row-centred weight
//                  W" = W diag(gamma) - (row mean of W diag(gamma)) -- centring the weight rows subtracts the mean of x inside the
//                  contraction -- and applies  LN(x) W^T + b = rstd (x W"^T) + b'  per row in its epilogue (b' = b + W beta):
//                  one FMA per element, the cost of a plain bias add.
//
// Two tile shapes:
//   CTA2 = false  cta_group::1, UMMA 128 x BLOCK_N x 16, one CTA per tile (48 KB of operands per k-block).
//   CTA2 = true   cta_group::2, UMMA 256 x BLOCK_N x 16 issued by the leader of a 2-CTA cluster: each CTA holds its
//                 own 128 rows of A and HALF of the W tile (32 KB per k-block -> 1.5x fewer L2->SM bytes per flop,
//                 6 stages in flight).  Measured equal to the 1-CTA shape on the image-tower GEMMs, slower on small ones.
#include "gemm.h"
#include <type_traits>
#include <cstdlib>
#include <map>
#include <tuple>

namespace tapclip {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;           // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;                        // two epilogue warps per TMEM lane quarter, alternating column chunks
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int STAGE_A_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int EPI_BUF_BYTES = 32 * 128;             // 32 rows x 128 B, one TMA-store box
constexpr int EPI_BYTES = EPI_WARPS * EPI_BUF_BYTES; // one staging buffer per epilogue warp

template <int BLOCK_N, bool CTA2> struct Cfg {
    static constexpr int B_ROWS = CTA2 ? BLOCK_N / 2 : BLOCK_N;          // rows of W this CTA stages per k-block
    static constexpr int STAGE_B_BYTES = B_ROWS * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = STAGE_A_BYTES + STAGE_B_BYTES;
    static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;            // 4 (48 KB), 6 (32 KB), 8 (24 KB)
    static constexpr int TMEM_COLS = 2 * BLOCK_N;
    // barriers (ring, TMEM, scheduler ring) + scheduler tiles + TMEM slot: 256 B; bias tile: BLOCK_N floats
    static constexpr int SMEM_USED = STAGES * STAGE_BYTES + EPI_BYTES + 256 + BLOCK_N * 4;
    // 1024 B of slack for the manual 1024-byte alignment of the carve-up (the kernel traps if it would not fit)
    static constexpr int SMEM_BYTES = SMEM_USED + 1024;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert((2 * STAGES + 4 + 2 * 4) * 8 + 4 * 4 + 4 <= 256, "barrier region");
};
constexpr int SCHED_DEPTH = 4;

// per-launch extras of the residual / folded-LayerNorm epilogues and of the dynamic tile scheduler
struct GemmExtra {
    const float* resid_in; int ld_in;      // EPI_F32_RESID
    void* xb; float* stats_out; float* shift_out;
    const float* stats_prev; const float* shift_prev; int prev_parts;
    const float* stats_in; int stats_parts;      // FOLD
    int* sched;                            // {next tile, finished CTAs} or null (static tile order)
};

// consumer side of the scheduler ring (whole warp): returns the next tile index (>= num_tiles: no more work)
__device__ __forceinline__ int sched_fetch(uint64_t* full, uint64_t* empty, const int* ring, int& slot, uint32_t& phase) {
    mbar_wait(&full[slot], phase);
    int t = *reinterpret_cast<const volatile int*>(&ring[slot]);
    t = __shfl_sync(0xffffffffu, t, 0);
    __syncwarp();
    if (elect_one()) mbar_arrive(&empty[slot]);
    if (++slot == SCHED_DEPTH) { slot = 0; phase ^= 1; }
    return t;
}

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
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool f16) {
    const uint32_t fmt = f16 ? 0u : 1u;             // F16F32Format: 0 = F16, 1 = BF16
    return (1u << 4) /*C=f32*/ | (fmt << 7) /*A*/ | (fmt << 10) /*B*/ | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

// two packed 16-bit values x (type T16) times act'(h) of two packed pre-activations h (fp16 if h_f16 else bf16)
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t w);
template <> __device__ __forceinline__ float2 unpack2<bf16>(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
template <> __device__ __forceinline__ float2 unpack2<f16>(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
template <int ACT, typename T16>
__device__ __forceinline__ uint32_t mul_act_grad(uint32_t x, uint32_t h, int h_f16) {
    const float2 xf = unpack2<T16>(x);
    const float2 hf = h_f16 ? unpack2<f16>(h) : unpack2<bf16>(h);
    return pack2<T16>(xf.x * act_bwd_fast<ACT>(hf.x), xf.y * act_bwd_fast<ACT>(hf.y));
}

template <int BLOCK_N, int EPI, int ACT, bool STORE_PRE, bool F16, bool CTA2, bool FOLD>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               void* out, void* out_pre, int ldo,
               const float* __restrict__ bias, int M, int N, int K, int dbg, int aux_f16, const GemmExtra ex) {
    using C = Cfg<BLOCK_N, CTA2>;
    constexpr bool ACTGRAD = (EPI == EPI_BF16_ACTGRAD);      // out = (acc + bias) * act'(aux), aux = out_pre pointer, same layout as out
    constexpr bool RESID = (EPI == EPI_F32_RESID);           // out = resid_in + acc + bias (+ 16-bit copy + row statistics)
    constexpr bool OUT16 = (EPI == EPI_BF16) || ACTGRAD;
    static_assert(!FOLD || EPI == EPI_BF16, "the folded-LayerNorm epilogue exists for 16-bit outputs only");
    using T16 = typename std::conditional<F16, f16, bf16>::type;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + C::STAGES * STAGE_A_BYTES;
    uint8_t* smem_epi = smem + C::STAGES * C::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + EPI_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + C::STAGES;
    uint64_t* tmem_full = bars + 2 * C::STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* sched_full = tmem_empty + 2;
    uint64_t* sched_empty = sched_full + SCHED_DEPTH;
    int* sched_tile = reinterpret_cast<int*>(sched_empty + SCHED_DEPTH);
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(sched_tile + SCHED_DEPTH);
    float* bias_s = reinterpret_cast<float*>(smem_epi + EPI_BYTES + 256);      // bias of the current tile, shared by the epilogue warps

    // warp index (and below the TMEM base) come out of shuffles so that ptxas knows they are warp-uniform
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // work decomposition: a "unit" is one CTA (CTA2 = false) or one 2-CTA cluster owning 256 rows (CTA2 = true)
    constexpr int UNIT_M = CTA2 ? 2 * BLOCK_M : BLOCK_M;
    const uint32_t cta_rank = CTA2 ? __shfl_sync(0xffffffffu, cluster_ctarank(), 0) : 0u;      // through a shuffle: warp-uniform for ptxas
    const int unit = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int num_units = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int m_tiles = (M + UNIT_M - 1) / UNIT_M, n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = m_tiles * n_tiles;
    const int num_kb = (K + BLOCK_K - 1) / BLOCK_K;
    const bool dyn = !CTA2 && ex.sched != nullptr;            // dynamic tile order (kernel parameter: warp-uniform)

    if (warp == 0 && lane == 0) {
        uint32_t dyn_bytes;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
        if ((uint32_t)(smem - smem_raw) + (uint32_t)C::SMEM_USED > dyn_bytes) {
            printf("tapclip: gemm_tc_kernel shared-memory carve-up does not fit (base offset %u)\n", (uint32_t)(smem - smem_raw));
            __trap();
        }
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int i = 0; i < C::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], CTA2 ? 2 * EPI_WARPS : EPI_WARPS); }
        for (int i = 0; i < SCHED_DEPTH; ++i) { mbar_init(&sched_full[i], 1); mbar_init(&sched_empty[i], 1 + EPI_WARPS); }
        fence_mbar_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) {
        if constexpr (CTA2) tmem_alloc_2sm(tmem_base_slot, C::TMEM_COLS);
        else tmem_alloc(tmem_base_slot, C::TMEM_COLS);
    }
    tc_fence_before();
    if constexpr (CTA2) cluster_sync_all();        // the peer's barriers must be initialised before any remote arrive
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_slot, 0);
    // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail (PDL);
    // from here on global memory written by that kernel is read
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        // ================================ TMA producer (+ tile scheduler) ================================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int pslot = 0; uint32_t pphase = 0;
            int tile = unit;
            while (true) {
                if (dyn) {                                  // hand the tile (or the end marker) to the MMA and epilogue warps
                    mbar_wait(&sched_empty[pslot], pphase ^ 1);
                    *reinterpret_cast<volatile int*>(&sched_tile[pslot]) = tile;
                    mbar_arrive(&sched_full[pslot]);
                    if (++pslot == SCHED_DEPTH) { pslot = 0; pphase ^= 1; }
                }
                if (tile >= num_tiles) break;
                // the next tile is drawn now: the atomic's round trip hides behind this tile's loads
                const int next = dyn ? num_units + atomicAdd(ex.sched, 1) : tile + num_units;
                const int m0 = (tile / n_tiles) * UNIT_M + (int)cta_rank * BLOCK_M;
                const int n0 = (tile % n_tiles) * BLOCK_N + (int)cta_rank * (CTA2 ? C::B_ROWS : 0);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if constexpr (CTA2) {
                        // both CTAs' loads complete on the LEADER's full barrier, which expects the pair's bytes
                        if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
                        const uint32_t fb = mapa_u32(&full_bar[stage], 0);
                        tma_load_2d_2sm(smem_a + stage * STAGE_A_BYTES, &tmap_a, kb * BLOCK_K, m0, fb);
                        tma_load_2d_2sm(smem_b + stage * C::STAGE_B_BYTES, &tmap_b, kb * BLOCK_K, n0, fb);
                    } else {
                        mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                        tma_load_2d(smem_a + stage * STAGE_A_BYTES, &tmap_a, kb * BLOCK_K, m0, &full_bar[stage]);
                        tma_load_2d(smem_b + stage * C::STAGE_B_BYTES, &tmap_b, kb * BLOCK_K, n0, &full_bar[stage]);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                tile = next;
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if constexpr (!CTA2) {
            // all 32 lanes run the warp-uniform control flow; elect.sync picks the issuing lane (see common.cuh)
            constexpr uint32_t idesc = make_idesc(UNIT_M, BLOCK_N, F16);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            int cslot = 0; uint32_t cphase = 0;
            int tile = dyn ? sched_fetch(sched_full, sched_empty, sched_tile, cslot, cphase) : unit;
            while (tile < num_tiles) {
                if (RESID && !(dbg & 32)) {
                    // pull the tile's old residual rows into L2 now: the epilogue reads them one mainloop from here
                    const int pm0 = (tile / n_tiles) * BLOCK_M, pn0 = (tile % n_tiles) * BLOCK_N;
                    const uint32_t bytes = (uint32_t)min(BLOCK_N, N - pn0) * 4u;
#pragma unroll
                    for (int rr = 0; rr < BLOCK_M / 32; ++rr) {
                        const int r = pm0 + rr * 32 + lane;
                        if (r < M) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ex.resid_in + (int64_t)r * ex.ld_in + pn0), "r"(bytes) : "memory");
                    }
                }
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(smem_u32(smem_a + stage * STAGE_A_BYTES));
                    const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + stage * C::STAGE_B_BYTES));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)      // advance 32 B (= 16 bf16) inside the 128B swizzle row: +2 in the >>4 address field
                        umma_ss_elect(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_elect(&empty_bar[stage]);           // frees this smem stage once the MMAs retire
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_elect(&tmem_full[acc]);                 // accumulator ready for the epilogue warps
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                tile = dyn ? sched_fetch(sched_full, sched_empty, sched_tile, cslot, cphase) : tile + num_units;
            }
        } else if (cta_rank == 0) {
            // leader CTA of the pair: the same warp-converged issue loop (the pair rank is warp-uniform), cta_group::2 forms
            constexpr uint32_t idesc = make_idesc(UNIT_M, BLOCK_N, F16);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = unit; tile < num_tiles; tile += num_units) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(smem_u32(smem_a + stage * STAGE_A_BYTES));
                    const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + stage * C::STAGE_B_BYTES));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_2sm_elect(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_2sm_elect(&empty_bar[stage]);       // frees this smem stage in both CTAs of the pair
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm_elect(&tmem_full[acc]);             // accumulator ready for the epilogue warps of both CTAs
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ================================ epilogue warps ==============================
        // TMEM -> registers (one accumulator row per thread) -> bias/activation -> XOR-swizzled smem transpose ->
        // coalesced 16-byte global stores (each store instruction covers 4 rows x 128 B).  Stores are fire-and-forget:
        // nothing in this loop waits on the memory system except the tcgen05.ld itself (and, for EPI_F32_RESID, the old
        // residual values, requested in the store mapping before the accumulator is read).
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        const int ew = warp - 2;
        const int chunk_par = ew >> 2;                       // this warp handles the chunks with (c & 1) == chunk_par
        const uint32_t stage_u32 = smem_u32(smem_epi + ew * EPI_BUF_BYTES);
        const uint32_t row_off = (uint32_t)lane * 128u;
        const uint32_t sw = (uint32_t)(lane & 7);
        const int rd_row = lane >> 3, rd_ch = lane & 7;      // read-back mapping: 8 lanes cover one 128-byte row
        int acc = 0; uint32_t acc_phase = 0;
        int cslot = 0; uint32_t cphase = 0;
        int tile = dyn ? sched_fetch(sched_full, sched_empty, sched_tile, cslot, cphase) : unit;
        // FOLD: the first 8 statistics partials of this thread's row, requested one tile ahead (at the end of the previous tile's
        // epilogue), so their latency never sits between two tiles
        [[maybe_unused]] float2 st[8];
        auto request_stats = [&](int t) {
            if constexpr (FOLD) {
                const int row = (t / n_tiles) * UNIT_M + (int)cta_rank * BLOCK_M + q * 32 + lane;
                const float2* sp = reinterpret_cast<const float2*>(ex.stats_in) + (int64_t)row * ex.stats_parts;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    st[p] = make_float2(0.f, 0.f);
                    if (t < num_tiles && row < M && p < ex.stats_parts) st[p] = sp[p];
                }
            }
        };
        request_stats(tile);
        while (tile < num_tiles) {
            const int m0 = (tile / n_tiles) * UNIT_M + (int)cta_rank * BLOCK_M, n0 = (tile % n_tiles) * BLOCK_N;
            const int row0 = m0 + q * 32;
            if (bias != nullptr) {
                // stage this tile's bias in smem while the MMAs of the tile are still running (a global load per
                // chunk inside the epilogue loop exposed its full latency: ncu long_scoreboard on the bias FADDs)
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");   // everyone is done with the previous tile's bias
                const int et = (int)threadIdx.x - 64;
                if (et < BLOCK_N / 4) {
                    const int col = n0 + et * 4;
                    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (col < N) b = __ldg(reinterpret_cast<const float4*>(bias + col));
                    *reinterpret_cast<float4*>(bias_s + et * 4) = b;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
            }
            // FOLD: this thread's row statistics -> out = fa * acc + b'[n],  fa = rstd (the mean is subtracted by the centred weight)
            float fa = 1.f;
            if constexpr (FOLD) {
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int p = 0; p < 8; ++p) { s1 += st[p].x; s2 += st[p].y; }
                const int row = row0 + lane;
                for (int p0 = 8; p0 < ex.stats_parts; p0 += 8) {            // more than 8 partials: rows wider than 1024 / 128-wide tiles
                    const float2* sp = reinterpret_cast<const float2*>(ex.stats_in) + (int64_t)row * ex.stats_parts;
#pragma unroll
                    for (int p = 0; p < 8; ++p)
                        if (row < M && p0 + p < ex.stats_parts) { const float2 t = sp[p0 + p]; s1 += t.x; s2 += t.y; }
                }
                const float inv_k = 1.0f / (float)K;
                const float mean = s1 * inv_k;
                fa = rsqrtf(fmaxf(s2 * inv_k - mean * mean, 0.f) + 1e-5f);
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
            if constexpr (RESID) {
                // ---- x_new = x_old + A W^T + b, its 16-bit copy and its row statistics -------------------------------------------
                // Global memory is touched in the coalesced store mapping only (lane (rd_row, rd_ch) owns 16 bytes of rows
                // it*4 + rd_row; 8 lanes cover a 128-byte row segment): the accumulator chunk goes through the warp's swizzled
                // staging buffer as in the other epilogues.  (Accessing memory straight from the accumulator layout -- thread <->
                // row, every warp-wide 16-byte access spread over 32 rows -- needs a third of the instructions and no shuffles,
                // but ran 1.7x slower: 96 vs 55 us on the 25216x768x768 out-projection.)
                // The old residual values are the only global READ: requested two chunks ahead -- the first two chunks while this
                // tile's MMAs still run (the MMA warp pulled the tile's rows into L2 when it started the tile).  Addresses and row
                // predicates are formed once per tile; the row sums stay lane-local through the chunk loop and cross the 8 lanes
                // of a row once per tile (the first version reduced per chunk and executed 3x the instructions of the red.add
                // epilogue: one epilogue warp's issue rate had become the kernel's bound).
                //
                // The 16-bit copy is SHIFTED by the previous mean of its row, zb = x_new - mean(x_old) (the folded-LayerNorm consumer
                // is invariant to a per-row shift): what gets rounded is then the centred value, as when LayerNorm's output is
                // rounded, and the statistics (sum z, sum z^2) are well conditioned.  mean(x_old) comes from the previous
                // statistics of the row (stats_prev + shift_prev); the new shift is stored for the next producer.
                constexpr int RCH = 32;                              // fp32 columns per 128-byte staging row
                constexpr int NMINE = BLOCK_N / RCH / 2;             // chunks of this warp: c = chunk_par + 2 i
                const int col_l = chunk_par * RCH + rd_ch * 4;       // this lane's first column inside the tile (chunk 0 of the warp)
                const int64_t row_l = row0 + rd_row;                 // this lane's first row (it = 0)
                const float* xin = ex.resid_in + row_l * ex.ld_in + n0 + col_l;
                float* xout = reinterpret_cast<float*>(out) + row_l * ldo + n0 + col_l;
                T16* zout = reinterpret_cast<T16*>(ex.xb) + row_l * N + n0 + col_l;
                const int64_t in_step = 4 * (int64_t)ex.ld_in, out_step = 4 * (int64_t)ldo, z_step = 4 * (int64_t)N;
                // the shift of row (row0 + lane) is formed by lane `lane` -- one row per lane, all of its partials requested at once,
                // consumed after the accumulator wait -- and handed to the lanes that store the row by shuffle
                float2 sp_[8];
                float shp = 0.f;
                const bool have_prev = ex.stats_prev != nullptr && row0 + lane < M;
                if (have_prev) {
                    const float2* sp = reinterpret_cast<const float2*>(ex.stats_prev) + (int64_t)(row0 + lane) * ex.prev_parts;
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        sp_[p] = make_float2(0.f, 0.f);
                        if (p < ex.prev_parts) sp_[p] = sp[p];
                    }
                    shp = ex.shift_prev[row0 + lane];
                }
                // One body, two instances: FULL = all 32 rows of the warp exist (every tile but the last row block) -> no per-row
                // predicates, no divergent branches; the tail instance predicates each row.
                auto body = [&](auto full_tag) {
                    constexpr bool FULLT = decltype(full_tag)::value;
                    uint32_t okmask = 0xffu;
                    if constexpr (!FULLT) {
                        okmask = 0;
#pragma unroll
                        for (int it = 0; it < 8; ++it)
                            if (row_l + it * 4 < M) okmask |= 1u << it;
                    }
                    // all three streams of this epilogue are touched once: streaming (evict-first) accesses keep the A / W operand tiles,
                    // which other CTAs re-read, resident in L2
                    float4 xo[2][8];
                    auto request = [&](int i, float4 (&dst)[8]) {
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            dst[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if ((FULLT || (okmask >> it & 1u)) && !(dbg & 4))
                                asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dst[it].x), "=f"(dst[it].y), "=f"(dst[it].z), "=f"(dst[it].w)
                                             : "l"(xin + it * in_step + i * 2 * RCH) : "memory");
                        }
                    };
                    request(0, xo[0]);
                    if constexpr (NMINE > 1) request(1, xo[1]);
                    float rs[8], rq[8];                              // lane-local (sum, sum of squares) of rows it*4 + rd_row
#pragma unroll
                    for (int it = 0; it < 8; ++it) { rs[it] = 0.f; rq[it] = 0.f; }
                    mbar_wait(&tmem_full[acc], acc_phase);
                    tc_fence_after();
                    float sh[8];                                     // the shifts of rows it*4 + rd_row
                    {
                        float my = 0.f;
                        if (have_prev) {
                            float s1 = 0.f;
#pragma unroll
                            for (int p = 0; p < 8; ++p) s1 += sp_[p].x;
                            for (int p = 8; p < ex.prev_parts; ++p)  // more than 8 partials: 128-wide tiles of a wide row
                                s1 += (reinterpret_cast<const float2*>(ex.stats_prev) + (int64_t)(row0 + lane) * ex.prev_parts)[p].x;
                            my = shp + s1 / (float)N;                // = mean of the row of x_old
                        }
#pragma unroll
                        for (int it = 0; it < 8; ++it) sh[it] = __shfl_sync(0xffffffffu, my, it * 4 + rd_row);
                    }
#pragma unroll
                    for (int i = 0; i < NMINE; ++i) {
                        const int c = chunk_par + 2 * i;
                        uint32_t r0[32];
                        tmem_ld_32x32(taddr + c * RCH, r0);
                        tmem_ld_wait();
                        if (i == NMINE - 1) {                        // this warp's TMEM reads of the tile are done
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                        }
                        __syncwarp();                                // previous read-back of the staging buffer is complete
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_u32 + row_off + (((uint32_t)j ^ sw) << 4)),
                                         "r"(r0[4 * j]), "r"(r0[4 * j + 1]), "r"(r0[4 * j + 2]), "r"(r0[4 * j + 3]) : "memory");
                        // the bias of this lane's four columns (the same for all its rows): added together with the old values
                        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (bias != nullptr) b4 = *reinterpret_cast<const float4*>(bias_s + c * RCH + rd_ch * 4);
                        __syncwarp();
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int r = it * 4 + rd_row;
                            float a0, a1, a2, a3;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3)
                                         : "r"(stage_u32 + (uint32_t)r * 128u + (((uint32_t)rd_ch ^ ((uint32_t)r & 7u)) << 4)) : "memory");
                            const float4 o = xo[i & 1][it];
                            a0 += o.x + b4.x; a1 += o.y + b4.y; a2 += o.z + b4.z; a3 += o.w + b4.w;
                            const bool ok = FULLT || (okmask >> it & 1u);
                            if (ok && !(dbg & 16))
                                asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(xout + it * out_step + i * 2 * RCH),
                                             "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
                            a0 -= sh[it]; a1 -= sh[it]; a2 -= sh[it]; a3 -= sh[it];
                            if (ok && ex.xb != nullptr && !(dbg & 8))
                                asm volatile("st.global.cs.v2.b32 [%0], {%1, %2};" ::"l"(zout + it * z_step + i * 2 * RCH),
                                             "r"(pack2<T16>(a0, a1)), "r"(pack2<T16>(a2, a3)) : "memory");
                            if (!FULLT && !ok) { a0 = 0.f; a1 = 0.f; a2 = 0.f; a3 = 0.f; }
                            rs[it] += (a0 + a1) + (a2 + a3);
                            rq[it] = fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, fmaf(a3, a3, rq[it]))));
                        }
                        if (i + 2 < NMINE) request(i + 2, xo[i & 1]);
                    }
                    if (ex.stats_out != nullptr) {
                        // the 8 lanes that share a row (consecutive lanes) reduce their sums; one (sum, sum of squares) pair per row,
                        // n-tile and warp parity: every slot has exactly one writer
                        const int parts = 2 * n_tiles, part = 2 * (tile % n_tiles) + chunk_par;
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            float ps = rs[it], pq = rq[it];
                            ps += __shfl_xor_sync(0xffffffffu, ps, 1); pq += __shfl_xor_sync(0xffffffffu, pq, 1);
                            ps += __shfl_xor_sync(0xffffffffu, ps, 2); pq += __shfl_xor_sync(0xffffffffu, pq, 2);
                            ps += __shfl_xor_sync(0xffffffffu, ps, 4); pq += __shfl_xor_sync(0xffffffffu, pq, 4);
                            if (rd_ch == 0 && (FULLT || (okmask >> it & 1u))) {
                                const int64_t r = row_l + it * 4;
                                reinterpret_cast<float2*>(ex.stats_out)[r * parts + part] = make_float2(ps, pq);
                                if (part == 0) ex.shift_out[r] = sh[it];
                            }
                        }
                    }
                };
                if (dbg == 1 || dbg == 2) {                          // measurement switches: consume the tile, store nothing
                    mbar_wait(&tmem_full[acc], acc_phase);
                    tc_fence_after();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                } else if (row0 + 32 <= M) body(std::true_type{});
                else body(std::false_type{});
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                tile = dyn ? sched_fetch(sched_full, sched_empty, sched_tile, cslot, cphase) : tile + num_units;
                continue;
            }
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            constexpr int CH = OUT16 ? 64 : 32;  // columns per 128-byte staging row
            constexpr int OUT_ESZ = OUT16 ? 2 : 4;
            constexpr int NCHUNK = BLOCK_N / CH;
            constexpr int LAST_MINE = NCHUNK - 2;            // last chunk index (before adding chunk_par) of each warp
#pragma unroll 1
            for (int c = chunk_par; c < NCHUNK; c += 2) {
                const int col0 = n0 + c * CH;
                const int gcol = col0 + rd_ch * (16 / OUT_ESZ);
                float v[CH];
                {
                    uint32_t r0[32];
                    tmem_ld_32x32(taddr + c * CH, r0);
                    if constexpr (CH == 64) {
                        uint32_t r1[32];
                        tmem_ld_32x32(taddr + c * CH + 32, r1);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
                    } else {
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
                    }
                }
                if (c >= LAST_MINE) {                        // this warp's TMEM reads of the tile are done
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if constexpr (CTA2) mbar_arrive_cluster(mapa_u32(&tmem_empty[acc], 0)); else mbar_arrive(&tmem_empty[acc]); }
                }
                if constexpr (FOLD) {
#pragma unroll
                    for (int j = 0; j < CH; j += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(bias_s + c * CH + j);   // smem broadcast
                        v[j] = fmaf(fa, v[j], b.x); v[j + 1] = fmaf(fa, v[j + 1], b.y);
                        v[j + 2] = fmaf(fa, v[j + 2], b.z); v[j + 3] = fmaf(fa, v[j + 3], b.w);
                    }
                } else if (bias != nullptr) {
#pragma unroll
                    for (int j = 0; j < CH; j += 4) {
                        const float4 b = *reinterpret_cast<const float4*>(bias_s + c * CH + j);   // smem broadcast
                        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                    }
                }
                if (row0 >= M || col0 >= N || dbg == 2) continue;
                constexpr int NPASS = (EPI == EPI_BF16 && STORE_PRE) ? 2 : 1;
                uint4 hq[ACTGRAD ? 8 : 1];
                if constexpr (ACTGRAD) {
                    // the saved pre-activations of this chunk, fetched with the same coalesced (row, 16-byte) mapping the
                    // stores below use; issued now so their latency hides behind the staging round trip
                    const uint8_t* hbase = reinterpret_cast<const uint8_t*>(out_pre);
                    const int hcol = col0 + rd_ch * 8;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = it * 4 + rd_row;
                        hq[it] = make_uint4(0u, 0u, 0u, 0u);
                        if (row0 + r < M && hcol < N)
                            hq[it] = __ldg(reinterpret_cast<const uint4*>(hbase + ((int64_t)(row0 + r) * ldo + hcol) * 2));
                    }
                }
#pragma unroll
                for (int pass = 0; pass < NPASS; ++pass) {
                    const bool is_pre = STORE_PRE && pass == 0;       // first pass of STORE_PRE writes the pre-activation copy
                    if (!is_pre && !ACTGRAD) {
#pragma unroll
                        for (int j = 0; j < CH; ++j) v[j] = act_fwd_fast<ACT>(v[j]);
                    }
                    const uint32_t sb = stage_u32;
                    __syncwarp();                                     // previous read-back of the staging buffer is complete
                    // registers -> swizzled staging (conflict-free 16-byte shared stores)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        uint32_t p0, p1, p2, p3;
                        if constexpr (OUT16) {
                            p0 = pack2<T16>(v[8 * j + 0], v[8 * j + 1]); p1 = pack2<T16>(v[8 * j + 2], v[8 * j + 3]);
                            p2 = pack2<T16>(v[8 * j + 4], v[8 * j + 5]); p3 = pack2<T16>(v[8 * j + 6], v[8 * j + 7]);
                        } else {
                            p0 = __float_as_uint(v[4 * j]); p1 = __float_as_uint(v[4 * j + 1]);
                            p2 = __float_as_uint(v[4 * j + 2]); p3 = __float_as_uint(v[4 * j + 3]);
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + row_off + (((uint32_t)j ^ sw) << 4)),
                                     "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
                    }
                    __syncwarp();
                    // staging -> global: lane (rd_row, rd_ch) moves 16 bytes; 8 lanes = one full 128-byte row segment
                    uint8_t* gbase = reinterpret_cast<uint8_t*>(is_pre ? out_pre : out);
                    if (dbg != 1) {
                        if constexpr (EPI == EPI_F32_ADD) {
                            // residual update x += tile as fire-and-forget vector reductions (measured faster than a plain
                            // read-modify-write: 37 vs 54 us on the 25216x768x768 out-projection); every element has exactly
                            // one contributor, so the result is deterministic
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int r = it * 4 + rd_row;
                                float a0, a1, a2, a3;
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3)
                                             : "r"(sb + (uint32_t)r * 128u + (((uint32_t)rd_ch ^ ((uint32_t)r & 7u)) << 4)) : "memory");
                                if (row0 + r < M && gcol < N)
                                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gbase + ((int64_t)(row0 + r) * ldo + gcol) * 4),
                                                 "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
                            }
                        } else {
#pragma unroll
                            for (int it = 0; it < 8; ++it) {
                                const int r = it * 4 + rd_row;
                                uint32_t x0, x1, x2, x3;
                                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                                             : "r"(sb + (uint32_t)r * 128u + (((uint32_t)rd_ch ^ ((uint32_t)r & 7u)) << 4)) : "memory");
                                if constexpr (ACTGRAD) {
                                    x0 = mul_act_grad<ACT, T16>(x0, hq[it].x, aux_f16); x1 = mul_act_grad<ACT, T16>(x1, hq[it].y, aux_f16);
                                    x2 = mul_act_grad<ACT, T16>(x2, hq[it].z, aux_f16); x3 = mul_act_grad<ACT, T16>(x3, hq[it].w, aux_f16);
                                }
                                if (row0 + r < M && gcol < N)
                                    asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(gbase + ((int64_t)(row0 + r) * ldo + gcol) * OUT_ESZ),
                                                 "r"(x0), "r"(x1), "r"(x2), "r"(x3) : "memory");   // streaming: keep A/W resident in L2
                            }
                        }
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            tile = dyn ? sched_fetch(sched_full, sched_empty, sched_tile, cslot, cphase) : tile + num_units;
            request_stats(tile);
        }
    }

    tc_fence_before();
    if constexpr (CTA2) cluster_sync_all();        // neither CTA may exit (or free TMEM) while its peer can still reach it
    else __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        if constexpr (CTA2) tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
        else tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
    if (dyn && threadIdx.x == 0) {
        // the last CTA to leave re-arms the scheduler slot, so a captured launch can be replayed
        __threadfence();
        if (atomicAdd(ex.sched + 1, 1) == (int)gridDim.x - 1) { ex.sched[0] = 0; ex.sched[1] = 0; }
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
}  // namespace

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
CUtensorMap encode_tmap(const void* ptr, CUtensorMapDataType dt, int elem_bytes, int64_t rows, int64_t cols, int64_t ld,
                        int box_rows, int box_cols);

// cuTensorMapEncodeTiled costs a few microseconds on the host; the engine's operands live in stable buffers, so the
// encoded maps are cached by (pointer, geometry).  Single-threaded per handle/process by contract (include/tapclip.h).
struct TmapKey {
    const void* ptr; int dt; int64_t rows, cols, ld; int box_rows, box_cols; int64_t seqs = 0;      // seqs > 0: [seqs][rows][cols] map
    bool operator<(const TmapKey& o) const {
        return std::tie(ptr, dt, rows, cols, ld, box_rows, box_cols, seqs) < std::tie(o.ptr, o.dt, o.rows, o.cols, o.ld, o.box_rows, o.box_cols, o.seqs);
    }
};
std::map<TmapKey, CUtensorMap> g_tmap_cache;

CUtensorMap make_tmap(const void* ptr, CUtensorMapDataType dt, int elem_bytes, int64_t rows, int64_t cols, int64_t ld,
                             int box_rows, int box_cols) {
    const TmapKey key{ptr, (int)dt, rows, cols, ld, box_rows, box_cols};
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) return it->second;
    if (g_tmap_cache.size() > 4096) g_tmap_cache.clear();
    return g_tmap_cache.emplace(key, encode_tmap(ptr, dt, elem_bytes, rows, cols, ld, box_rows, box_cols)).first->second;
}

CUtensorMap make_tmap_seq(const void* ptr, CUtensorMapDataType dt, int elem_bytes, int64_t S, int64_t N, int64_t cols, int64_t ld,
                          int box_rows, int box_cols) {
    const TmapKey key{ptr, (int)dt, N, cols, ld, box_rows, box_cols, S};
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) return it->second;
    if (g_tmap_cache.size() > 4096) g_tmap_cache.clear();
    TC_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * elem_bytes) % 16 == 0, "TMA base pointer and row pitch must be 16-byte aligned");
    TC_CHECK(box_cols * elem_bytes == 128 && box_rows >= 1 && box_rows <= 256, "box: 128 bytes x [1, 256] rows");
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)N, (cuuint64_t)S};
    cuuint64_t strides[2] = {(cuuint64_t)(ld * elem_bytes), (cuuint64_t)(N * ld * elem_bytes)};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1u};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_fn()(&m, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TC_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed with code %d", (int)r);
    return g_tmap_cache.emplace(key, m).first->second;
}

CUtensorMap encode_tmap(const void* ptr, CUtensorMapDataType dt, int elem_bytes, int64_t rows, int64_t cols, int64_t ld,
                        int box_rows, int box_cols) {
    TC_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA base pointer must be 16-byte aligned");
    TC_CHECK((ld * elem_bytes) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes (ld=%lld)", (long long)ld);
    TC_CHECK(box_cols * elem_bytes == 128, "box inner extent must be 128 bytes");
    TC_CHECK(box_rows >= 1 && box_rows <= 256, "TMA box rows must be in [1, 256]");
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

namespace {

// per-device launch state (a process may drive several GPUs): SM count and the ring of scheduler slots
struct DeviceState {
    int num_sms = 0;
    int* sched = nullptr;          // SCHED_SLOTS x {next tile, finished CTAs}; every slot re-arms itself (see the kernel's last lines)
    uint32_t seq = 0;
};
constexpr int SCHED_SLOTS = 4096;
DeviceState& device_state() {
    static DeviceState st[64];
    int dev;
    TC_CUDA(cudaGetDevice(&dev));
    TC_CHECK(dev >= 0 && dev < 64, "device ordinal %d out of range", dev);
    DeviceState& d = st[dev];
    if (d.num_sms == 0) TC_CUDA(cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev));
    return d;
}
// TAPCLIP_GEMM_DEBUG (measurement only; results are WRONG when set): 1 = skip the epilogue's global stores, 2 = skip the epilogue body;
// residual epilogue ablations (bit flags): 4 = no old-value loads, 8 = no 16-bit copy, 16 = no fp32 store, 32 = no L2 prefetch
int g_debug = getenv("TAPCLIP_GEMM_DEBUG") ? atoi(getenv("TAPCLIP_GEMM_DEBUG")) : 0;
// TAPCLIP_GEMM_SCHED: 1 = dynamic tile order (atomic counter), 0 = static striding
int g_dynamic = getenv("TAPCLIP_GEMM_SCHED") ? atoi(getenv("TAPCLIP_GEMM_SCHED")) : 0;

int choose_block_n(int64_t N) { return (N % 256 == 0) ? 256 : 128; }

template <int BLOCK_N, int EPI, int ACT, bool STORE_PRE, bool F16, bool CTA2, bool FOLD = false>
void launch(const GemmArgs& g, cudaStream_t stream) {
    using C = Cfg<BLOCK_N, CTA2>;
    auto kern = gemm_tc_kernel<BLOCK_N, EPI, ACT, STORE_PRE, F16, CTA2, FOLD>;
    constexpr CUtensorMapDataType DT16 = F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    DeviceState& ds = device_state();
    ensure_dynamic_smem((const void*)kern, C::SMEM_BYTES);
    const CUtensorMap& ta = make_tmap(g.a, DT16, 2, g.M, g.K, g.lda, BLOCK_M, BLOCK_K);
    const CUtensorMap& tb = make_tmap(g.w, DT16, 2, g.N, g.K, g.ldw, C::B_ROWS, BLOCK_K);
    const int out_esz = (EPI == EPI_BF16 || EPI == EPI_BF16_ACTGRAD) ? 2 : 4;
    TC_CHECK((reinterpret_cast<uintptr_t>(g.out) & 15) == 0 && (g.ldo * out_esz) % 16 == 0, "GEMM output must be 16-byte aligned with a 16-byte row pitch");
    if (STORE_PRE || EPI == EPI_BF16_ACTGRAD) TC_CHECK((reinterpret_cast<uintptr_t>(g.out_pre) & 15) == 0, "GEMM pre-activation output must be 16-byte aligned");
    GemmExtra ex = {};
    if (EPI == EPI_F32_RESID) {
        TC_CHECK(g.resid_in != nullptr && (reinterpret_cast<uintptr_t>(g.resid_in) & 15) == 0 && g.ld_in % 4 == 0, "residual input must be 16-byte aligned with a 16-byte row pitch");
        TC_CHECK((reinterpret_cast<uintptr_t>(g.xb) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.stats_out) & 7) == 0 &&
                 (reinterpret_cast<uintptr_t>(g.stats_prev) & 7) == 0, "xb must be 16-byte, the statistics 8-byte aligned");
        TC_CHECK(g.N % BLOCK_N == 0, "the residual epilogue needs N %% %d == 0 (N=%lld)", BLOCK_N, (long long)g.N);
        TC_CHECK((g.stats_out == nullptr) == (g.shift_out == nullptr) && (g.stats_prev == nullptr) == (g.shift_prev == nullptr),
                 "statistics and shifts go together");
        TC_CHECK(g.stats_prev == nullptr || (g.prev_parts >= 1 && g.stats_prev != g.stats_out && g.shift_prev != g.shift_out),
                 "the previous statistics are read while the new ones are written: use two buffers");
        ex.resid_in = g.resid_in; ex.ld_in = (int)g.ld_in; ex.xb = g.xb; ex.stats_out = g.stats_out; ex.shift_out = g.shift_out;
        ex.stats_prev = g.stats_prev; ex.shift_prev = g.shift_prev; ex.prev_parts = g.prev_parts;
    }
    if (FOLD) {
        TC_CHECK(g.stats_in != nullptr && g.stats_parts >= 1 && g.bias != nullptr, "folded-LayerNorm GEMM needs stats_in and the folded bias");
        TC_CHECK((reinterpret_cast<uintptr_t>(g.stats_in) & 7) == 0, "stats_in alignment");
        ex.stats_in = g.stats_in; ex.stats_parts = g.stats_parts;
    }
    if (g.bias) TC_CHECK((reinterpret_cast<uintptr_t>(g.bias) & 15) == 0, "bias must be 16-byte aligned");
    if (CTA2) {
        const int64_t tiles = ceil_div(g.M, 2 * BLOCK_M) * ceil_div(g.N, BLOCK_N);
        const int clusters = (int)std::min<int64_t>(tiles, ds.num_sms / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
        TC_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, g.out, g.out_pre, (int)g.ldo, g.bias, (int)g.M, (int)g.N, (int)g.K, g_debug, (int)(g.aux_dt == DT_F16), ex));
    } else {
        const int64_t tiles = ceil_div(g.M, BLOCK_M) * ceil_div(g.N, BLOCK_N);
        const int grid = (int)std::min<int64_t>(tiles, ds.num_sms);
        if (g_dynamic && tiles > grid) {           // with one tile per CTA there is nothing to schedule
            if (ds.sched == nullptr) {
                TC_CUDA(cudaMalloc(&ds.sched, SCHED_SLOTS * 2 * sizeof(int)));
                TC_CUDA(cudaMemset(ds.sched, 0, SCHED_SLOTS * 2 * sizeof(int)));
                TC_CUDA(cudaDeviceSynchronize());
            }
            ex.sched = ds.sched + 2 * (ds.seq++ % SCHED_SLOTS);
        }
        launch_pdl(kern, grid, NUM_THREADS, C::SMEM_BYTES, stream, ta, tb, g.out, g.out_pre, (int)g.ldo, g.bias, (int)g.M, (int)g.N, (int)g.K, g_debug, (int)(g.aux_dt == DT_F16), ex);
    }
    TC_LAUNCH_CHECK();
}

template <int BLOCK_N, bool F16, bool CTA2>
void dispatch(const GemmArgs& g, cudaStream_t stream) {
    if (g.epi == EPI_F32) return launch<BLOCK_N, EPI_F32, ACT_NONE, false, F16, CTA2>(g, stream);
    if (g.epi == EPI_F32_ADD) return launch<BLOCK_N, EPI_F32_ADD, ACT_NONE, false, F16, CTA2>(g, stream);
    if (g.epi == EPI_F32_RESID) {
        if constexpr (!CTA2) return launch<BLOCK_N, EPI_F32_RESID, ACT_NONE, false, F16, CTA2>(g, stream);
        TC_CHECK(false, "the residual epilogue runs on single-CTA tiles only");
    }
    if (g.epi == EPI_BF16_ACTGRAD) {
        // dgrad through the MLP activation: out = (A W^T) * act'(out_pre); only the shapes the backward uses are built
        TC_CHECK(g.out_pre != nullptr && (g.aux_dt == DT_BF16 || g.aux_dt == DT_F16), "activation-gradient epilogue needs 16-bit pre-activations in out_pre");
        if constexpr (!CTA2 && !F16) {
            if (g.act == ACT_GELU_ERF) return launch<BLOCK_N, EPI_BF16_ACTGRAD, ACT_GELU_ERF, false, F16, CTA2>(g, stream);
            if (g.act == ACT_QUICK_GELU) return launch<BLOCK_N, EPI_BF16_ACTGRAD, ACT_QUICK_GELU, false, F16, CTA2>(g, stream);
        }
        TC_CHECK(false, "activation-gradient epilogue: unsupported activation %d / tile shape / operand type", g.act);
    }
    TC_CHECK(g.epi == EPI_BF16, "unknown epilogue %d", g.epi);
    const bool pre = g.out_pre != nullptr;
    if (g.stats_in != nullptr) {
        // LayerNorm folded into the GEMM (see GemmArgs::stats_in)
        if constexpr (!CTA2) {
            if (g.act == ACT_NONE) { TC_CHECK(!pre, "out_pre needs an activation"); return launch<BLOCK_N, EPI_BF16, ACT_NONE, false, F16, CTA2, true>(g, stream); }
            if (g.act == ACT_GELU_ERF) return pre ? launch<BLOCK_N, EPI_BF16, ACT_GELU_ERF, true, F16, CTA2, true>(g, stream)
                                                  : launch<BLOCK_N, EPI_BF16, ACT_GELU_ERF, false, F16, CTA2, true>(g, stream);
            if (g.act == ACT_QUICK_GELU) return pre ? launch<BLOCK_N, EPI_BF16, ACT_QUICK_GELU, true, F16, CTA2, true>(g, stream)
                                                    : launch<BLOCK_N, EPI_BF16, ACT_QUICK_GELU, false, F16, CTA2, true>(g, stream);
        }
        TC_CHECK(false, "folded-LayerNorm GEMM: unsupported activation %d / tile shape", g.act);
    }
    if (g.act == ACT_NONE) { TC_CHECK(!pre, "out_pre needs an activation"); return launch<BLOCK_N, EPI_BF16, ACT_NONE, false, F16, CTA2>(g, stream); }
    if (g.act == ACT_GELU_ERF) return pre ? launch<BLOCK_N, EPI_BF16, ACT_GELU_ERF, true, F16, CTA2>(g, stream)
                                          : launch<BLOCK_N, EPI_BF16, ACT_GELU_ERF, false, F16, CTA2>(g, stream);
    if (g.act == ACT_QUICK_GELU) return pre ? launch<BLOCK_N, EPI_BF16, ACT_QUICK_GELU, true, F16, CTA2>(g, stream)
                                            : launch<BLOCK_N, EPI_BF16, ACT_QUICK_GELU, false, F16, CTA2>(g, stream);
    TC_CHECK(false, "unknown activation %d", g.act);
}

}  // namespace

int gemm_stats_parts(int64_t N) { return 2 * (int)ceil_div(N, choose_block_n(N)); }

void gemm_tc(const GemmArgs& g, cudaStream_t stream) {
    TC_CHECK(g.M > 0 && g.N > 0 && g.K > 0, "empty GEMM %lldx%lldx%lld", (long long)g.M, (long long)g.N, (long long)g.K);
    TC_CHECK(g.K % 8 == 0 && g.N % 8 == 0, "tcgen05 GEMM needs K%%8==0 and N%%8==0 (K=%lld N=%lld)", (long long)g.K, (long long)g.N);
    TC_CHECK(g.dt == DT_BF16 || g.dt == DT_F16, "tcgen05 GEMM operands must be bf16 or fp16");
    // block_n: 0 = choose; 128 / 256 = one CTA per 128 x block_n tile; 512 = 2-CTA pairs on 256 x 256 tiles
    int bn = g.block_n;
    static const int env_bn = getenv("TAPCLIP_GEMM_BN") ? atoi(getenv("TAPCLIP_GEMM_BN")) : 0;   // measurement override: 128 | 256 | 512
    const bool fixed_shape = g.epi == EPI_F32_RESID || g.stats_in != nullptr;   // the statistics layout follows choose_block_n
    if (bn == 0 && env_bn != 0 && g.N % 256 == 0 && g.epi != EPI_BF16_ACTGRAD && !fixed_shape) bn = env_bn;
    if (g.epi == EPI_F32_RESID) {
        TC_CHECK(bn == 0 || bn == choose_block_n(g.N), "the residual epilogue fixes the tile shape (statistics layout)");
        bn = 0;
    }
    // TAPCLIP_GEMM_QKV2CTA=1: the image tower's QKV projection (large M, N >= 2048, short K, plain 16-bit store) on 2-CTA pairs,
    // the one shape where the pair wins in isolation (69.8 vs 74.2 us, above cuBLAS 71.5).  In the step (three runs each, one box):
    // 8.274 / 8.290 / 8.384 ms with it, 8.290 / 8.312 / 8.382 without -- below the noise, so it stays off.
    static const int qkv2 = getenv("TAPCLIP_GEMM_QKV2CTA") ? atoi(getenv("TAPCLIP_GEMM_QKV2CTA")) : 0;
    if (bn == 0 && qkv2 && g.M >= 16384 && g.N >= 2048 && g.N % 256 == 0 && g.K <= 1024 && g.epi == EPI_BF16 && g.act == ACT_NONE && g.out_pre == nullptr && !fixed_shape) bn = 512;
    if (bn == 0) {
        // Measured (tools/gemm_bench.py, B200): 128x256 single-CTA tiles are within 3 % of the 2-CTA 256x256 tiles on the
        // image-tower shapes (both ~0.95 of cuBLAS: the mainloop is not L2-feed-bound) and 10-15 % faster on the small
        // text-tower shapes, where a 2-CTA pair halves the number of schedulable units.
        bn = choose_block_n(g.N);
    }
    const bool f16 = g.dt == DT_F16;
    if (bn == 512) { if (f16) dispatch<256, true, true>(g, stream); else dispatch<256, false, true>(g, stream); }
    else if (bn == 256) { if (f16) dispatch<256, true, false>(g, stream); else dispatch<256, false, false>(g, stream); }
    else { if (f16) dispatch<128, true, false>(g, stream); else dispatch<128, false, false>(g, stream); }
}

}  // namespace tapclip
