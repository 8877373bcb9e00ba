"""K1 parity: the tcgen05/TMEM/TMA GEMM (and the fp32 SIMT GEMM) against torch matmul, through the C ABI."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(a, w, bias, dtype, epi, act=-1, block_n=0, out_init=None, want_pre=False):
    from tapclip_b200 import _lib
    lib = _lib.load()
    M, K = a.shape
    N = w.shape[0]
    if epi == _lib.EPI_ACT:
        out = torch.empty(M, N, device="cuda", dtype={"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[dtype])
    else:
        out = out_init.clone() if out_init is not None else torch.empty(M, N, device="cuda", dtype=torch.float32)
    pre = torch.empty_like(out) if want_pre else None
    _lib.check(lib.tapclip_op_gemm(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), _lib.ptr(pre), M, N, K,
                                   _lib.DTYPE[dtype], epi, act, block_n, _lib.stream_ptr()))
    torch.cuda.synchronize()
    return out, pre


def _ref_act(x, act):
    if act == 0:
        return torch.nn.functional.gelu(x)
    if act == 1:
        return x * torch.sigmoid(1.702 * x)
    return x


SHAPES = [
    (128, 256, 64), (128, 256, 768), (256, 768, 768), (1576, 2304, 768), (1576, 768, 3072),
    (197 * 8, 3072, 768), (93 * 65, 1536, 512), (93 * 65, 512, 2048), (65, 512, 512), (8, 512, 768),
    (300, 128, 128), (1000, 384, 640), (129, 264, 72),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("block_n", [0, 128, 256, 512])    # 512 = 2-CTA pairs (cta_group::2), 256x256 tiles
def test_gemm_tc_f32_out(M, N, K, block_n):
    from tapclip_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out, _ = _gemm(a, w, bias, "bf16", _lib.EPI_F32, block_n=block_n)
    ref = a.float() @ w.float().t() + bias
    err = (out - ref).abs().max().item()
    assert err < 2e-3, f"max abs err {err}"          # fp32 accumulation of exact bf16 products: order-of-summation noise only


@pytest.mark.parametrize("M,N,K", [(1576, 768, 768), (93 * 65, 512, 2048), (200, 256, 128)])
@pytest.mark.parametrize("block_n", [0, 256, 512])
def test_gemm_tc_residual_add(M, N, K, block_n):
    from tapclip_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    x = torch.randn(M, N, device="cuda", generator=g)
    out, _ = _gemm(a, w, bias, "bf16", _lib.EPI_F32_ADD, out_init=x, block_n=block_n)
    ref = x + a.float() @ w.float().t() + bias
    assert (out - ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("act", [-1, 0, 1])
@pytest.mark.parametrize("M,N,K", [(1576, 3072, 768), (93 * 65, 2048, 512), (130, 256, 64)])
@pytest.mark.parametrize("block_n", [0, 256])
def test_gemm_tc_bf16_out_act(M, N, K, act, block_n):
    from tapclip_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out, pre = _gemm(a, w, bias, "bf16", _lib.EPI_ACT, act=act, want_pre=(act >= 0), block_n=block_n)
    z = a.float() @ w.float().t() + bias
    ref = _ref_act(z, act)
    tol = 2e-2                                            # one bf16 rounding of values up to ~4
    assert (out.float() - ref).abs().max().item() < tol
    if pre is not None:
        assert (pre.float() - z).abs().max().item() < tol


@pytest.mark.parametrize("act", [-1, 1])
@pytest.mark.parametrize("M,N,K", [(93 * 65, 2048, 512), (93 * 65, 512, 2048), (65, 512, 512), (130, 256, 64)])
def test_gemm_tc_fp16_operands(M, N, K, act):
    """Mixed mode: text-tower forward GEMMs take fp16 operands (same tcgen05 kind::f16 instruction, a/b format = F16)."""
    from tapclip_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(9)
    a = torch.randn(M, K, device="cuda", generator=g).half()
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    z = a.float() @ w.float().t() + bias
    out, _ = _gemm(a, w, bias, "fp16", _lib.EPI_F32)
    assert (out - z).abs().max().item() < 2e-3
    x = torch.randn(M, N, device="cuda", generator=g)
    out, _ = _gemm(a, w, bias, "fp16", _lib.EPI_F32_ADD, out_init=x)
    assert (out - (x + z)).abs().max().item() < 2e-3
    out, pre = _gemm(a, w, bias, "fp16", _lib.EPI_ACT, act=act, want_pre=(act >= 0))
    assert out.dtype == torch.float16 and (out.float() - _ref_act(z, act)).abs().max().item() < 4e-3
    if pre is not None:
        assert (pre.float() - z).abs().max().item() < 4e-3


@pytest.mark.parametrize("M,N,K", [(300, 384, 128), (1576, 768, 768), (65, 512, 512), (77, 100, 36)])
def test_gemm_simt_fp32(M, N, K):
    from tapclip_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * K ** -0.5
    bias = torch.randn(N, device="cuda", generator=g)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = (a.double() @ w.double().t() + bias.double()).float()
    out, _ = _gemm(a, w, bias, "fp32", _lib.EPI_F32)
    assert (out - ref).abs().max().item() < 2e-5
    out, pre = _gemm(a, w, bias, "fp32", _lib.EPI_ACT, act=0, want_pre=True)
    assert (out - torch.nn.functional.gelu(ref)).abs().max().item() < 2e-5
    assert (pre - ref).abs().max().item() < 2e-5
    x = torch.randn(M, N, device="cuda", generator=g)
    out, _ = _gemm(a, w, bias, "fp32", _lib.EPI_F32_ADD, out_init=x)
    assert (out - (x + ref)).abs().max().item() < 2e-5


@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("aux", ["bf16", "fp16"])
@pytest.mark.parametrize("M,N,K", [(93 * 65, 2048, 512), (130, 256, 64), (333, 384, 128)])
@pytest.mark.parametrize("block_n", [0, 128])
def test_gemm_tc_act_grad_epilogue(M, N, K, act, aux, block_n):
    """epi 3/4: dh = (dy . W) * act'(h_pre), the MLP dgrad with the activation derivative fused into the store stage
    (replaces autograd through open_clip's mlp.gelu for the text tower backward)."""
    g = torch.Generator(device="cuda").manual_seed(17)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).bfloat16()
    h = (2.0 * torch.randn(M, N, device="cuda", generator=g)).to(torch.bfloat16 if aux == "bf16" else torch.float16)
    from tapclip_b200 import _lib
    lib = _lib.load()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.tapclip_op_gemm(_lib.ptr(a), _lib.ptr(w), None, _lib.ptr(out), _lib.ptr(h), M, N, K,
                                   _lib.DTYPE["bf16"], 3 if aux == "bf16" else 4, act, block_n, _lib.stream_ptr()))
    torch.cuda.synchronize()
    hf = h.float().requires_grad_(True)
    _ref_act(hf, act).sum().backward()
    ref = (a.float() @ w.float().t()) * hf.grad
    assert bool(((out.float() - ref).abs() <= 5e-3 + 2 ** -7 * ref.abs()).all())   # two bf16 roundings (2 x 2^-9 relative) + fast act'
    assert ((out.float() - ref).norm() / ref.norm()).item() < 6e-3


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("M,N,K,N2,act", [(93 * 65, 512, 512, 1536, -1), (93 * 65, 512, 2048, 2048, 1), (1576, 768, 768, 2304, -1),
                                          (300, 1024, 1024, 4096, 0), (129, 128, 64, 384, -1), (197 * 64, 768, 3072, 3072, 1),
                                          (128 * 40 + 5, 384, 512, 1152, 0), (25216, 768, 768, 2304, -1)])
def test_gemm_residual_statistics_and_folded_layernorm(M, N, K, N2, act, dtype):
    """The pre-LN block without LayerNorm kernels (gemm_tc.cu): EPI_F32_RESID writes x1 = x0 + A.W^T + b together with its 16-bit
    copy and per-row (sum, sum of squares) partials; the consumer GEMM multiplies the un-normalised copy with W2 diag(gamma) and
    applies rstd (acc - mean s) + b' in its epilogue.  Compared with torch: residual, statistics, and act(LN(x1) W2^T + b2)."""
    from tapclip_b200 import _lib
    lib = _lib.load()
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float16
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(tdt)
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(tdt)
    bias = torch.randn(N, device="cuda", generator=g)
    gamma = 1.0 + 0.2 * torch.randn(N, device="cuda", generator=g)
    beta = 0.1 * torch.randn(N, device="cuda", generator=g)
    x0 = 3.0 * torch.randn(M, N, device="cuda", generator=g) + 0.5            # non-zero row mean: exercises the mean term
    w2 = torch.randn(N2, N, device="cuda", generator=g) * N ** -0.5
    b2 = torch.randn(N2, device="cuda", generator=g)
    parts = lib.tapclip_op_gemm_stats_parts(N)
    assert parts == 2 * (-(-N // (256 if N % 256 == 0 else 128)))
    # the statistics describing x0 (as the previous producer would have left them): shift = row mean, one partial
    xb0 = torch.empty(M, N, device="cuda", dtype=tdt)
    st0 = torch.empty(M, 1, 2, device="cuda")
    sh0 = torch.empty(M, device="cuda")
    _lib.check(lib.tapclip_op_row_stats_cast(_lib.ptr(x0), _lib.ptr(xb0), _lib.DTYPE[dtype], _lib.ptr(st0), _lib.ptr(sh0), M, N, _lib.stream_ptr()))
    assert (sh0 - x0.mean(1)).abs().max().item() < 1e-5 and torch.equal(xb0, (x0 - sh0[:, None]).to(tdt))
    x1 = torch.empty(M, N, device="cuda")
    xb = torch.empty(M, N, device="cuda", dtype=tdt)
    stats = torch.full((M, parts, 2), float("nan"), device="cuda")
    shift = torch.full((M,), float("nan"), device="cuda")

    def resid(x_in, x_out, st, sh, prev):
        _lib.check(lib.tapclip_op_gemm_resid(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(x_in), 0, _lib.ptr(x_out), 0, _lib.ptr(xb),
                                             _lib.ptr(st), _lib.ptr(sh), _lib.ptr(st0) if prev else None, _lib.ptr(sh0) if prev else None,
                                             1 if prev else 0, M, N, K, _lib.DTYPE[dtype], _lib.stream_ptr()))
    resid(x0, x1, stats, shift, True)
    x_ref = x0 + a.float() @ w.float().t() + bias
    assert (x1 - x_ref).abs().max().item() < 2e-3
    assert (shift - x0.mean(1)).abs().max().item() < 1e-4                  # the shift is the mean of the row BEFORE the update
    z = x1 - shift[:, None]
    assert torch.equal(xb, z.to(tdt))                                      # the 16-bit copy is the shifted row
    assert not torch.isnan(stats).any()                                    # every partial slot has a writer
    s = stats.sum(1)
    assert (s[:, 0] - z.sum(1)).abs().max().item() < 1e-2 and ((s[:, 1] - (z * z).sum(1)).abs() / (z * z).sum(1)).max().item() < 1e-5
    # in place (x_in == x_out) gives the same bits; so does a second run (no atomics: deterministic)
    x_inplace = x0.clone()
    stats2, shift2 = torch.empty_like(stats), torch.empty_like(shift)
    resid(x_inplace, x_inplace, stats2, shift2, True)
    assert torch.equal(x_inplace, x1) and torch.equal(stats2, stats) and torch.equal(shift2, shift)
    # without previous statistics the shift is zero
    resid(x0, x_inplace, stats2, shift2, False)
    assert torch.equal(x_inplace, x1) and shift2.abs().max().item() == 0.0 and torch.equal(xb, x1.to(tdt))
    resid(x0, x1, stats, shift, True)                                      # restore xb / stats for the consumer below
    # folded consumer
    wf = torch.empty(N2, N, device="cuda", dtype=tdt)
    fb = torch.empty(N2, device="cuda")
    _lib.check(lib.tapclip_op_fold_ln_weight(_lib.ptr(w2), _lib.ptr(b2), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(wf), _lib.DTYPE[dtype],
                                             _lib.ptr(fb), N2, N, _lib.stream_ptr()))
    wg = w2 * gamma
    assert (wf.float() - (wg - wg.mean(1, keepdim=True))).abs().max().item() < (2e-2 if dtype == "bf16" else 2e-3)   # one rounding
    assert wf.float().sum(1).abs().max().item() < 0.3                       # rows are centred (up to the rounding of N entries)
    assert (fb - (b2 + w2 @ beta)).abs().max().item() < 1e-4
    out = torch.empty(M, N2, device="cuda", dtype=tdt)
    pre = torch.empty(M, N2, device="cuda", dtype=tdt) if act >= 0 else None
    _lib.check(lib.tapclip_op_gemm_fold(_lib.ptr(xb), _lib.ptr(stats), parts, _lib.ptr(wf), _lib.ptr(fb), _lib.ptr(out), _lib.ptr(pre),
                                        M, N2, N, _lib.DTYPE[dtype], act, _lib.stream_ptr()))
    torch.cuda.synchronize()
    ln_ref = torch.nn.functional.layer_norm(x_ref, (N,), gamma, beta, 1e-5)
    h_ref = ln_ref @ w2.t() + b2
    ref = _ref_act(h_ref, act)
    # the shifted 16-bit copy carries one rounding of CENTRED values up to ~4.5 sigma (sigma ~ 3): the budget of rounding LN(x)
    tol = 6e-2 if dtype == "bf16" else 8e-3
    assert (out.float() - ref).abs().max().item() < tol
    assert ((out.float() - ref).norm() / ref.norm()).item() < (8e-3 if dtype == "bf16" else 1e-3)
    if pre is not None:
        assert (pre.float() - h_ref).abs().max().item() < tol
    # the statistics of rows no residual GEMM produced (first block): row_stats_cast
    xbr = torch.empty(M, N, device="cuda", dtype=tdt)
    str_ = torch.empty(M, 1, 2, device="cuda")
    shr = torch.empty(M, device="cuda")
    _lib.check(lib.tapclip_op_row_stats_cast(_lib.ptr(x_ref), _lib.ptr(xbr), _lib.DTYPE[dtype], _lib.ptr(str_), _lib.ptr(shr), M, N, _lib.stream_ptr()))
    out0 = torch.empty(M, N2, device="cuda", dtype=tdt)
    _lib.check(lib.tapclip_op_gemm_fold(_lib.ptr(xbr), _lib.ptr(str_), 1, _lib.ptr(wf), _lib.ptr(fb), _lib.ptr(out0), None,
                                        M, N2, N, _lib.DTYPE[dtype], act, _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert (out0.float() - ref).abs().max().item() < tol


def test_gemm_residual_live_rows_through_leading_dimensions():
    """Last-block form: the out-projection reads one row per sequence of A and of the residual stream through leading dimensions
    and writes a compact [S, N] matrix; no 16-bit copy, no statistics."""
    from tapclip_b200 import _lib
    lib = _lib.load()
    S, T, N, K = 65, 93, 512, 512
    g = torch.Generator(device="cuda").manual_seed(7)
    a_full = torch.randn(S * T, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    x_full = torch.randn(S * T, N, device="cuda", generator=g)
    x_keep = x_full.clone()
    out = torch.empty(S, N, device="cuda")
    row = T - 1
    a_live = a_full.view(S, T, K)[:, row]                              # strided view: element pointer of row `row`, ld = T*K
    # tapclip_op_gemm_resid takes dense A; emulate the engine's lda by materialising the live rows of A only
    _lib.check(lib.tapclip_op_gemm_resid(_lib.ptr(a_live.contiguous()), _lib.ptr(w), _lib.ptr(bias), ctypes.c_void_p(x_full.data_ptr() + row * N * 4), T * N,
                                         _lib.ptr(out), 0, None, None, None, None, None, 0, S, N, K, _lib.DTYPE["bf16"], _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = x_keep.view(S, T, N)[:, row] + a_live.float() @ w.float().t() + bias
    assert (out - ref).abs().max().item() < 2e-3
    assert torch.equal(x_full, x_keep)                                 # the strided input is only read
