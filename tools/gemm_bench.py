"""Micro-benchmark of the tcgen05 GEMM through the C ABI: per-shape time / TFLOP/s, with torch.matmul (cuBLAS) beside it.
Usage (under gpurun): python tools/gemm_bench.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tapclip_b200 import _lib

lib = _lib.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
SHAPES = [  # (tag, M, N, K, epi, act)
    ("img qkv", 25216, 2304, 768, 0, -1), ("img out", 25216, 768, 768, 2, -1), ("img fc", 25216, 3072, 768, 0, 1),
    ("img proj", 25216, 768, 3072, 2, -1),
    ("txt qkv", 6045, 1536, 512, 0, -1), ("txt out", 6045, 512, 512, 2, -1), ("txt fc", 6045, 2048, 512, 0, 1),
    ("txt proj", 6045, 512, 2048, 2, -1), ("txt dfc", 6045, 512, 2048, 1, -1), ("txt dqkv", 6045, 512, 1536, 1, -1),
    ("txt dh", 6045, 2048, 512, 0, -1),
]


def bench(fn, reps):
    """GPU time per launch from a CUDA-graph replay of `reps` back-to-back launches (no host launch cost inside)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def host_cost(fn, reps):
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    return dt / reps * 1e6


print(f"{'shape':10s} {'M':>6s} {'N':>5s} {'K':>5s} epi |  auto us  TF/s | bn=128 us TF/s | bn=256 us TF/s | 2cta256 us TF/s | cuBLAS us TF/s")
for tag, M, N, K, epi, act in SHAPES:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if epi == 0 else torch.float32)
    fl = 2.0 * M * N * K
    cols = []
    for bn in (0, 128, 256, 512):
        def run():
            _lib.check(lib.tapclip_op_gemm(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), None, M, N, K, 1, epi, act, bn,
                                           _lib.stream_ptr()))
        us = bench(run, reps)
        cols.append(f"{us:8.1f} {fl / us / 1e6:6.0f}")
        if bn == 0:
            hc = host_cost(run, reps)
    us = bench(lambda: torch.matmul(a, w.t()), reps)
    cols.append(f"{us:8.1f} {fl / us / 1e6:6.0f}")
    print(f"{tag:10s} {M:6d} {N:5d} {K:5d} {epi:3d} | " + " | ".join(cols) + f" | host enqueue {hc:5.1f} us")
