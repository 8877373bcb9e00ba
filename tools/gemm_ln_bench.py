"""Residual GEMM + LayerNorm: two launches (K1 with the red.add epilogue, then the LayerNorm kernel) vs the fused cluster kernel
(gemm_ln.cu).  CUDA-graph replay, no host time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tapclip_b200 import _lib
lib = _lib.load()
reps = 20
for tag, M, N, K, dt in [("txt out", 6045, 512, 512, "fp16"), ("txt proj", 6045, 512, 2048, "fp16"), ("img out", 25216, 768, 768, "bf16"),
                         ("img proj", 25216, 768, 3072, "bf16"), ("L/14 out", 128 * 577, 1024, 1024, "bf16")]:
    tdt = torch.bfloat16 if dt == "bf16" else torch.float16
    a = torch.randn(M, K, device="cuda").to(tdt); w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(tdt)
    bias = torch.randn(N, device="cuda"); gamma = torch.ones(N, device="cuda"); beta = torch.zeros(N, device="cuda")
    x = torch.randn(M, N, device="cuda"); ln = torch.empty(M, N, device="cuda", dtype=tdt)
    D = _lib.DTYPE[dt]

    def two():
        _lib.check(lib.tapclip_op_gemm(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(x), None, M, N, K, D, 2, -1, 0, _lib.stream_ptr()))
        _lib.check(lib.tapclip_op_layernorm(_lib.ptr(x), N, _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(ln), D, None, M, N, _lib.stream_ptr()))

    def fused():
        _lib.check(lib.tapclip_op_gemm_resid_ln(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(x),
                                                _lib.ptr(ln), None, M, N, K, D, _lib.stream_ptr()))

    out = []
    for fn in (two, fused):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / reps * 1e3)
    print(f"{tag:9s} M={M:6d} N={N:5d} K={K:5d}: GEMM+LN {out[0]:7.1f} us   fused {out[1]:7.1f} us")
