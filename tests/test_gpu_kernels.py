"""Per-kernel parity against plain torch fp32 (LayerNorm fwd/bwd, attention fwd/bwd + probe epilogue, K3)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from tapclip_b200 import _lib as L
    return L, L.load()


@pytest.mark.parametrize("rows,d", [(1576, 768), (6045, 512), (33, 128), (4, 1024)])
@pytest.mark.parametrize("dtype", ["fp32", "bf16", "fp16"])
def test_layernorm_fwd_bwd(rows, d, dtype):
    L, lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(rows + d)
    x = torch.randn(rows, d, device="cuda", generator=g) * 2 + 0.5
    gamma = 1 + 0.1 * torch.randn(d, device="cuda", generator=g)
    beta = 0.1 * torch.randn(d, device="cuda", generator=g)
    tdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[dtype]
    out = torch.empty(rows, d, device="cuda", dtype=tdt)
    xc = torch.empty_like(x)
    L.check(lib.tapclip_op_layernorm(L.ptr(x), d, L.ptr(gamma), L.ptr(beta), L.ptr(out), L.DTYPE[dtype], L.ptr(xc), rows, d, L.stream_ptr()))
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (d,), gamma, beta, 1e-5)
    tol = {"bf16": 3e-2, "fp16": 4e-3, "fp32": 1e-5}[dtype]
    assert (out.float() - ref).abs().max().item() < tol
    assert torch.equal(xc, x)
    if dtype == "fp16":
        return                                              # gradients are never fp16
    dy = torch.randn(rows, d, device="cuda", generator=g)
    ref.backward(dy)
    acc0 = torch.randn(rows, d, device="cuda", generator=g)
    acc = acc0.clone()
    cast = torch.empty(rows, d, device="cuda", dtype=tdt)
    L.check(lib.tapclip_op_layernorm_bwd(L.ptr(dy), L.ptr(x), L.ptr(gamma), L.ptr(acc), L.ptr(cast), L.DTYPE[dtype], rows, d, L.stream_ptr()))
    torch.cuda.synchronize()
    assert (acc - (acc0 + xr.grad)).abs().max().item() < 2e-5
    assert (cast.float() - acc).abs().max().item() < (5e-2 if dtype == "bf16" else 1e-7)


def _ref_attention(qkv, S, N, H):
    d = H * 64
    q, k, v = qkv.float().view(S, N, 3, H, 64).permute(2, 0, 3, 1, 4)       # [S,H,N,64]
    p = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(S * N, d)
    return o, p


@pytest.mark.parametrize("S,N,H", [(3, 197, 12), (5, 93, 8), (2, 17, 4), (2, 577, 2), (4, 82, 8), (1, 50, 12)])
@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
def test_attention_fwd_and_probes(S, N, H, dtype):
    L, lib = _lib()
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(S * N + H)
    tdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[dtype]
    qkv = (torch.randn(S * N, 3 * d, device="cuda", generator=g) * 1.5).to(tdt)
    ref_o, ref_p = _ref_attention(qkv, S, N, H)
    tol_o, tol_p = {"bf16": (2e-2, 2e-3), "fp16": (3e-3, 3e-4), "fp32": (2e-5, 2e-6)}[dtype]
    # CLS-row probe
    out = torch.empty(S * N, d, device="cuda", dtype=tdt)
    rows = torch.zeros(S, H, N, device="cuda")
    L.check(lib.tapclip_op_attention(L.ptr(qkv), L.ptr(out), L.DTYPE[dtype], S, N, H, L.PROBE_CLS_ROW, L.ptr(rows), 0, H * N, L.stream_ptr()))
    torch.cuda.synchronize()
    assert (out.float() - ref_o).abs().max().item() < tol_o
    assert (rows - ref_p[:, :, 0, :]).abs().max().item() < tol_p
    assert (rows.sum(-1) - 1).abs().max().item() < 1e-3
    # text-column probe
    P = min(16, N - 1)
    col = torch.zeros(S, H, P, device="cuda")
    out2 = torch.empty_like(out)
    L.check(lib.tapclip_op_attention(L.ptr(qkv), L.ptr(out2), L.DTYPE[dtype], S, N, H, L.PROBE_TEXT_COL, L.ptr(col), P, 0, L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(out2, out)
    assert (col - ref_p[:, :, :P, N - 1]).abs().max().item() < tol_p
    # K3: head-mean + softmax over P (clip_wrapper.py:36 + attribution_monitor.py:29-32)
    raw = torch.empty(S, P, device="cuda")
    attr = torch.empty(S, P, device="cuda")
    L.check(lib.tapclip_op_attribution(L.ptr(col), L.ptr(raw), L.ptr(attr), S, H, P, L.stream_ptr()))
    torch.cuda.synchronize()
    ref_raw = ref_p[:, :, :P, N - 1].mean(1)
    assert (raw - ref_raw).abs().max().item() < tol_p
    assert (attr - torch.softmax(ref_raw, -1)).abs().max().item() < tol_p


@pytest.mark.parametrize("S,N,H", [(86, 197, 6), (128, 93, 8), (100, 50, 12), (70, 208, 8), (90, 129, 6), (140, 64, 8),
                                   (20, 577, 12), (40, 300, 10), (9, 1000, 16), (44, 257, 8)])   # N > 256: the flash-style KV-loop kernel
@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
def test_attention_fwd_tcgen05_persistent_kernel(S, N, H, dtype):
    """Shapes with >= 1024 (sequence, head, q-tile) items take the persistent tcgen05 kernel (attention_tc.cu) in all three
    chunk-count instances (N <= 64 / 128 / 208), with both probes; the small shapes above exercise the mma.sync kernel."""
    assert S * H * ((N + 127) // 128) >= 1024
    L, lib = _lib()
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(S + N + H)
    tdt = {"bf16": torch.bfloat16, "fp16": torch.float16}[dtype]
    qkv = (torch.randn(S * N, 3 * d, device="cuda", generator=g) * 1.5).to(tdt)
    ref_o, ref_p = _ref_attention(qkv, S, N, H)
    tol_o, tol_p = {"bf16": (2e-2, 2e-3), "fp16": (3e-3, 3e-4)}[dtype]
    out = torch.empty(S * N, d, device="cuda", dtype=tdt)
    rows = torch.zeros(S, H, N, device="cuda")
    L.check(lib.tapclip_op_attention(L.ptr(qkv), L.ptr(out), L.DTYPE[dtype], S, N, H, L.PROBE_CLS_ROW, L.ptr(rows), 0, H * N, L.stream_ptr()))
    torch.cuda.synchronize()
    assert (out.float() - ref_o).abs().max().item() < tol_o
    assert (rows - ref_p[:, :, 0, :]).abs().max().item() < tol_p
    P = min(16, N - 1)
    col = torch.zeros(S, H, P, device="cuda")
    out2 = torch.empty_like(out)
    L.check(lib.tapclip_op_attention(L.ptr(qkv), L.ptr(out2), L.DTYPE[dtype], S, N, H, L.PROBE_TEXT_COL, L.ptr(col), P, 0, L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(out2, out)
    assert (col - ref_p[:, :, :P, N - 1]).abs().max().item() < tol_p
    out3 = torch.empty_like(out)
    L.check(lib.tapclip_op_attention(L.ptr(qkv), L.ptr(out3), L.DTYPE[dtype], S, N, H, L.PROBE_NONE, None, 0, 0, L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(out3, out)                                     # a probe never changes O


@pytest.mark.parametrize("S,N,H,probe", [(128, 197, 12, "cls"), (65, 93, 8, "text"), (70, 141, 8, "none"), (150, 50, 12, "none")])
def test_attention_fwd_tcgen05_is_reproducible(S, N, H, probe):
    """The persistent kernel hands work between its warps through mbarriers only (operand slots released when P.V completes,
    S issued in two column ranges, O leaving through per-warp TMA stores from reused staging): 60 back-to-back launches on
    the same input, each into a poisoned output, must agree bit for bit -- a missing hand-off shows up as a sporadic difference."""
    L, lib = _lib()
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(7)
    tdt = torch.bfloat16 if probe == "cls" else torch.float16
    qkv = (torch.randn(S * N, 3 * d, device="cuda", generator=g) * 1.5).to(tdt)
    mode = {"cls": L.PROBE_CLS_ROW, "text": L.PROBE_TEXT_COL, "none": L.PROBE_NONE}[probe]
    P = 16 if probe == "text" else 0
    first = None
    for it in range(60):
        out = torch.full((S * N, d), float("nan"), device="cuda", dtype=tdt)
        pb = torch.full((S * H * max(N, 16),), float("nan"), device="cuda")
        L.check(lib.tapclip_op_attention(L.ptr(qkv), L.ptr(out), L.DTYPE["bf16" if tdt == torch.bfloat16 else "fp16"], S, N, H, mode,
                                         L.ptr(pb) if probe != "none" else None, P, H * N if probe == "cls" else 0, L.stream_ptr()))
        if first is None:
            torch.cuda.synchronize()
            assert torch.isfinite(out.float()).all()
            first = (out.clone(), pb.clone())
        else:
            assert torch.equal(out, first[0]), f"launch {it} differs"
            if probe != "none":
                assert torch.equal(torch.nan_to_num(pb, nan=-1.0), torch.nan_to_num(first[1], nan=-1.0)), f"probe of launch {it} differs"


@pytest.mark.parametrize("S,N,H", [(5, 93, 8), (2, 82, 4), (3, 17, 2), (1, 128, 8)])
@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
def test_attention_bwd(S, N, H, dtype):
    """dtype names the saved qkv; gradients are bf16 for both 16-bit cases (mixed mode saves fp16 activations)."""
    L, lib = _lib()
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(N)
    tdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[dtype]
    gdt = torch.float32 if dtype == "fp32" else torch.bfloat16
    qkv = torch.randn(S * N, 3 * d, device="cuda", generator=g).to(tdt)
    do = torch.randn(S * N, d, device="cuda", generator=g).to(gdt)
    dqkv = torch.empty(S * N, 3 * d, device="cuda", dtype=gdt)
    L.check(lib.tapclip_op_attention_bwd(L.ptr(qkv), L.ptr(do), L.ptr(dqkv), L.DTYPE[dtype], S, N, H, L.stream_ptr()))
    torch.cuda.synchronize()
    x = qkv.float().clone().requires_grad_(True)
    o, _ = _ref_attention(x, S, N, H)
    o.backward(do.float())
    tol = 2e-5 if dtype == "fp32" else 3e-2
    assert (dqkv.float() - x.grad).abs().max().item() < tol * max(1.0, x.grad.abs().max().item())


@pytest.mark.parametrize("S,N,H", [(3, 197, 12), (2, 577, 16), (5, 50, 12), (2, 17, 4), (3, 130, 2),      # mma.sync / SIMT statistics
                                   (90, 197, 12), (40, 577, 16), (44, 257, 8),                           # tcgen05 forward kernels
                                   (50, 257, 8), (30, 700, 4), (100, 197, 12), (120, 130, 6)])           # + tcgen05 rollout step (N > 128, >= 96 CTAs)
@pytest.mark.parametrize("dtype", ["bf16", "fp16", "fp32"])
def test_rollout_step_and_softmax_statistics(S, N, H, dtype):
    """Rollout extension (rollout.cu): the softmax statistics from the stand-alone kernel and from the attention forward
    (emitted by the persistent / KV-loop tcgen05 kernels, completed by the statistics kernel elsewhere), and one layer of
    the CLS-row propagation r_out = 0.5 r + 0.5 mean_h r^T P_h against torch fp32 on the same (rounded) qkv."""
    if dtype == "fp32" and N > 608:
        pytest.skip("fp32 parity-mode attention kernels serve N <= 608")
    L, lib = _lib()
    d = H * 64
    g = torch.Generator(device="cuda").manual_seed(7 * S + N + H)
    tdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[dtype]
    qkv = (torch.randn(S * N, 3 * d, device="cuda", generator=g) * 1.5).to(tdt)
    q, k, _ = qkv.float().view(S, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    scores = q @ k.transpose(-1, -2) / 8.0                                    # [S,H,N,N]
    lse_ref = torch.logsumexp(scores, dim=-1) / math.log(2.0)
    p = torch.softmax(scores, dim=-1)
    lse = torch.full((S, H, N), float("nan"), device="cuda")
    L.check(lib.tapclip_op_attention_lse(L.ptr(qkv), None, L.ptr(lse), L.DTYPE[dtype], S, N, H, L.stream_ptr()))
    out = torch.empty(S * N, d, device="cuda", dtype=tdt)
    lse_f = torch.full((S, H, N), float("nan"), device="cuda")
    L.check(lib.tapclip_op_attention_lse(L.ptr(qkv), L.ptr(out), L.ptr(lse_f), L.DTYPE[dtype], S, N, H, L.stream_ptr()))
    torch.cuda.synchronize()
    tol_l = 2e-5 if dtype == "fp32" else 2e-4                                  # absolute, log2 units (16-bit: fp32 accumulation order)
    assert (lse - lse_ref).abs().max().item() < tol_l
    assert (lse_f - lse_ref).abs().max().item() < tol_l
    ref_o, _ = _ref_attention(qkv, S, N, H)
    assert (out.float() - ref_o).abs().max().item() < {"bf16": 2e-2, "fp16": 3e-3, "fp32": 2e-5}[dtype]
    # first step (r = e_0), a middle step (dense positive r with exact zeros in it), last step (CLS column dropped)
    r_mid = torch.rand(S, N, device="cuda", generator=g) * 0.01
    r_mid[:, 3] = 0.0
    for r_in, last in [(None, False), (r_mid, False), (r_mid, True)]:
        r = r_mid if r_in is not None else torch.zeros(S, N, device="cuda").index_fill_(1, torch.tensor([0], device="cuda"), 1.0)
        ref = 0.5 * r + 0.5 * torch.einsum("si,shij->sj", r, p) / H
        if last:
            ref = ref[:, 1:]
        r_out = torch.full_like(ref, float("nan"))
        L.check(lib.tapclip_op_rollout_step(L.ptr(qkv), L.ptr(lse_f), L.ptr(r_in), L.ptr(r_out), L.DTYPE[dtype], S, N, H, int(last), L.stream_ptr()))
        torch.cuda.synchronize()
        rel = ((r_out - ref).abs() / ref.abs().clamp_min(1e-12)).max().item()
        assert rel < (2e-5 if dtype == "fp32" else 5e-4), (rel, last)
