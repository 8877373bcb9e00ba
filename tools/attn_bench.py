"""Micro-benchmark of the attention forward kernels (CUDA-graph replay): tcgen05 vs mma.sync, vision and text shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tapclip_b200 import _lib
lib = _lib.load()
reps = 20
for tag, S, N, H in [("vision B=128", 128, 197, 12), ("text C=65", 65, 93, 8), ("vision B=256", 256, 197, 12), ("ViT-L/14@336 B=64", 64, 577, 16)]:
    d = H * 64
    qkv = torch.randn(S * N, 3 * d, device="cuda").bfloat16()
    out = torch.empty(S * N, d, device="cuda", dtype=torch.bfloat16)
    def run():
        _lib.check(lib.tapclip_op_attention(_lib.ptr(qkv), _lib.ptr(out), 1, S, N, H, 0, None, 0, 0, _lib.stream_ptr()))
    for _ in range(3): run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    fl = 4.0 * S * H * N * N * 64
    print(f"{tag:14s} impl={os.environ.get('TAPCLIP_ATTN_IMPL','auto'):4s} {us:8.1f} us  {fl/us/1e6:7.1f} TFLOP/s")
