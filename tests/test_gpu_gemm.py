"""K1 parity: the tcgen05/TMEM/TMA GEMM (and the fp32 SIMT GEMM) against torch matmul, through the C ABI."""
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
@pytest.mark.parametrize("M,N,K", [(93 * 65, 512, 512), (93 * 65, 512, 2048), (1576, 768, 768), (300, 1024, 1024), (129, 512, 64),
                                   (197 * 64, 768, 3072), (128 * 40 + 5, 512, 512)])
def test_gemm_residual_layernorm_epilogue(M, N, K, dtype):
    """K1-LN (gemm_ln.cu): x += A.W^T + bias, ln_out = LayerNorm(x) in one launch; the CTAs owning the column slices of a
    128-row block exchange their row statistics through distributed shared memory (cluster of N/256 CTAs)."""
    from tapclip_b200 import _lib
    lib = _lib.load()
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float16
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(tdt)
    w = (torch.randn(N, K, device="cuda", generator=g) * K ** -0.5).to(tdt)
    bias = torch.randn(N, device="cuda", generator=g)
    gamma = 1.0 + 0.2 * torch.randn(N, device="cuda", generator=g)
    beta = 0.1 * torch.randn(N, device="cuda", generator=g)
    x0 = 3.0 * torch.randn(M, N, device="cuda", generator=g) + 0.5
    x = x0.clone()
    ln = torch.empty(M, N, device="cuda", dtype=tdt)
    xc = torch.empty(M, N, device="cuda")
    _lib.check(lib.tapclip_op_gemm_resid_ln(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(x),
                                            _lib.ptr(ln), _lib.ptr(xc), M, N, K, _lib.DTYPE[dtype], _lib.stream_ptr()))
    torch.cuda.synchronize()
    x_ref = x0 + a.float() @ w.float().t() + bias
    ln_ref = torch.nn.functional.layer_norm(x_ref, (N,), gamma, beta, 1e-5)
    assert (x - x_ref).abs().max().item() < 2e-3
    assert torch.equal(xc, x)
    tol = 3e-2 if dtype == "bf16" else 4e-3          # one 16-bit rounding of values up to ~5
    assert (ln.float() - ln_ref).abs().max().item() < tol
