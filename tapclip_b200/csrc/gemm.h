// GEMM entry points shared by the engine:  out[M,N] = epilogue(A[M,K] * W[N,K]^T + bias[N])
#pragma once
#include "common.cuh"
#include <algorithm>

namespace tapclip {

enum Epi : int {
    EPI_BF16 = 0,      // out (bf16) = act(acc + bias);  optional out_pre (bf16) = acc + bias
    EPI_F32 = 1,       // out (f32)  = acc + bias
    EPI_F32_ADD = 2,   // out (f32) += acc + bias        (residual stream update)
    EPI_BF16_ACTGRAD = 3,  // out (bf16) = (acc + bias) * act'(out_pre)   (dgrad through the MLP activation; out_pre is READ:
                           // the saved 16-bit pre-activations, type aux_dt, same [M,N] layout and leading dimension as out)
    EPI_F32_RESID = 4,     // out (f32, ldo) = resid_in (f32, ld_in) + acc + bias: the residual update `x = x + f(...)` of a pre-LN
                           // block, out of place (or in place: resid_in == out).  Optionally also emits what the NEXT LayerNorm
                           // needs: xb = the updated rows, SHIFTED by a per-row constant, in the 16-bit operand type (dense [M,N]),
                           // stats_out = per-row partial (sum, sum of squares) of the shifted rows, one pair per (n-tile,
                           // epilogue-warp parity), shift_out = the shift (the mean of the row BEFORE the update, taken from the
                           // previous statistics; 0 without them).  LayerNorm is invariant to the shift; it keeps the rounded
                           // values centred -- see stats_in
};

struct GemmArgs {
    const void* a = nullptr;   // [M, K] row-major, leading dimension lda (elements); bf16 (tc) or f32 (simt)
    const void* w = nullptr;   // [N, K] row-major (nn.Linear weight layout), leading dimension ldw
    void* out = nullptr;       // [M, N], leading dimension ldo
    void* out_pre = nullptr;   // optional pre-activation copy (same type/ld as out)
    const float* bias = nullptr;
    int64_t M = 0, N = 0, K = 0;
    int64_t lda = 0, ldw = 0, ldo = 0;
    int epi = EPI_F32;
    int act = ACT_NONE;
    int block_n = 0;           // 0 = choose; 128 / 256 = 128 x block_n single-CTA tiles; 512 = 2-CTA pairs, 256 x 256 tiles
    int aux_dt = DT_BF16;      // EPI_BF16_ACTGRAD: type of the pre-activations behind out_pre (DT_BF16 or DT_F16)
    int dt = DT_BF16;          // tcgen05 path: 16-bit type of A, W and of the EPI_BF16 output (DT_BF16 or DT_F16)
    // EPI_F32_RESID
    const float* resid_in = nullptr;   // [M, N] fp32, leading dimension ld_in
    int64_t ld_in = 0;
    void* xb = nullptr;                // optional: 16-bit copy (type dt) of the updated rows minus their shift, dense [M, N]
    float* stats_out = nullptr;        // optional: [M][gemm_stats_parts(N)][2] partial (sum, sum of squares) of the shifted rows
    float* shift_out = nullptr;        // [M] the shift used (goes with stats_out)
    const float* stats_prev = nullptr; // optional: the statistics / shifts describing resid_in (another buffer than stats_out):
    const float* shift_prev = nullptr; //   shift = shift_prev + sum(stats_prev.x) / N = mean of the row of resid_in
    int prev_parts = 0;
    // LayerNorm folded into the GEMM (EPI_BF16 only): A holds the UN-normalised (possibly row-shifted) rows x in 16 bits, W the gamma-scaled and
    // row-centred weight W" = W diag(gamma) - rowmean(W diag(gamma)) (so x W"^T = (x - mean(x)) (W diag(gamma))^T), bias the folded
    // bias b' = b + W beta; with rstd recovered from stats_in the epilogue forms  LN(x) W^T + b = rstd * (x W"^T) + b'.
    const float* stats_in = nullptr;   // [M][stats_parts][2] partial (sum, sum of squares) over the K elements of each row of A
    int stats_parts = 0;
};
// number of (sum, sum of squares) partials per row an EPI_F32_RESID GEMM with N output columns writes (2 per n-tile)
int gemm_stats_parts(int64_t N);

// Cached TMA descriptor of a 2-D row-major tensor [rows, cols] (leading dimension ld, elements of elem_bytes), 128B-swizzled
// boxes of box_rows x box_cols (box_cols * elem_bytes must be 128).  Shared by the GEMM and the tcgen05 attention kernel.
// Returned BY VALUE (128 bytes): a caller may hold two descriptors at once, and the cache evicts everything when it is full.
CUtensorMap make_tmap(const void* ptr, CUtensorMapDataType dt, int elem_bytes, int64_t rows, int64_t cols, int64_t ld,
                             int box_rows, int box_cols);
// [S][N][cols] view (sequence-major rows, pitch ld elements): boxes of box_rows x box_cols inside ONE sequence -- rows past N are
// clipped by the map, so a tile store never reaches the next sequence
CUtensorMap make_tmap_seq(const void* ptr, CUtensorMapDataType dt, int elem_bytes, int64_t S, int64_t N, int64_t cols, int64_t ld,
                          int box_rows, int box_cols);

// tcgen05/TMEM/TMA path: A and W bf16 (or fp16, g.dt); out = same 16-bit type (EPI_BF16) or f32
void gemm_tc(const GemmArgs& g, cudaStream_t stream);

// fp32 SIMT path for the fp32 parity mode: A, W, out all f32; EPI_BF16 means "store in the activation
// type" (f32 here) with the optional activation / pre-activation copy
void gemm_simt_f32(const GemmArgs& g, cudaStream_t stream);

}  // namespace tapclip
