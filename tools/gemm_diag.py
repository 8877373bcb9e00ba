"""Diagnostics for the tcgen05 GEMM on a real B200: structured inputs that localise descriptor / swizzle /
epilogue-layout bugs.  Prints a compact report; exits 0 even on mismatch (it is a diagnostic, not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tapclip_b200 import _lib

lib = _lib.load()


def run(a, w, bias, epi, block_n, out_dtype=torch.float32, init=None):
    M, K = a.shape; N = w.shape[0]
    out = init.clone() if init is not None else torch.full((M, N), float("nan"), device="cuda", dtype=out_dtype)
    rc = lib.tapclip_op_gemm(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), None, M, N, K, 1, epi, -1, block_n, _lib.stream_ptr())
    if rc != 0:
        print("  launch error:", _lib.last_error()); return None
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print("  sync error:", e); return None
    return out


def report(tag, out, ref):
    if out is None:
        print(f"{tag}: FAILED TO RUN"); return
    err = (out.float() - ref).abs()
    bad = err > 1e-2
    print(f"{tag}: max_err={err.max().item():.4g} bad={int(bad.sum())}/{bad.numel()} nan={int(torch.isnan(out.float()).sum())}")
    if bad.any():
        idx = bad.nonzero()
        print("   first bad (m,n):", idx[:6].tolist())
        rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
        print("   bad rows mod 128 (first 16):", sorted(set((rows % 128).tolist()))[:16], " bad cols mod 64 (first 16):", sorted(set((cols % 64).tolist()))[:16])
        m, n = idx[0].tolist()
        print(f"   out[{m},{n}]={out[m, n].item():.5g} ref={ref[m, n].item():.5g}")


for block_n in (128, 256, 512):
    for (M, N, K) in [(128, min(block_n, 256), 64), (256, min(block_n, 256), 128), (128, min(block_n, 256), 512), (256, 2 * min(block_n, 256), 64), (1000, 768, 768)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        w = torch.randn(N, K, device="cuda", generator=g).bfloat16()
        ref = a.float() @ w.float().t()
        report(f"bn={block_n} rand  M={M} N={N} K={K} f32", run(a, w, None, 1, block_n), ref)
    # one-hot A: out[m, n] = w[n, m % K]  -> exposes K-offset / swizzle mistakes exactly
    M, N, K = 256, min(block_n, 256), 64
    a = torch.zeros(M, K, device="cuda"); a[torch.arange(M), torch.arange(M) % K] = 1
    w = torch.arange(N * K, device="cuda", dtype=torch.float32).reshape(N, K) % 251
    report(f"bn={block_n} onehot f32", run(a.bfloat16(), w.bfloat16(), None, 1, block_n), a @ w.t())
    x0 = torch.ones(M, N, device="cuda")
    report(f"bn={block_n} onehot f32-add", run(a.bfloat16(), w.bfloat16(), None, 2, block_n, init=x0), x0 + a @ w.t())
    report(f"bn={block_n} onehot bf16", run(a.bfloat16(), w.bfloat16(), None, 0, block_n, out_dtype=torch.bfloat16), (a @ w.t()))
print("gemm_diag done")
