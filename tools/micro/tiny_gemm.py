import os, sys
sys.path.insert(0, "/root/repo")
import torch
from tapclip_b200 import _lib
lib = _lib.load()
reps = 50
for (M, N, K, epi) in [(128, 256, 64, 1), (128, 256, 512, 1), (6045, 512, 512, 1), (6045, 512, 512, 2), (6045, 512, 512, 0), (18944, 256, 512, 1), (148*128, 256, 64, 1)]:
    a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if epi == 0 else torch.float32)
    def run():
        _lib.check(lib.tapclip_op_gemm(_lib.ptr(a), _lib.ptr(w), None, _lib.ptr(out), None, M, N, K, 1, epi, -1, 0, _lib.stream_ptr()))
    for _ in range(3): run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"M={M:6d} N={N:4d} K={K:4d} epi={epi}: {e0.elapsed_time(e1)/reps*1e3:7.2f} us per launch (graph replay, PDL={os.environ.get('TAPCLIP_PDL','1')})")
