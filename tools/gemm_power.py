"""Sustained GEMM loop (3 s) with nvidia-smi clock/power sampling: is the epilogue cost a power/clock effect?"""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tapclip_b200 import _lib
lib = _lib.load()
M, N, K = 25216, 2304, 768
a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
bias = torch.randn(N, device="cuda"); out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
def run():
    _lib.check(lib.tapclip_op_gemm(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), None, M, N, K, 1, 0, -1, 256, _lib.stream_ptr()))
for _ in range(3): run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(50): run()
samples = []
stop = False
def sampler():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    while not stop:
        ln = p.stdout.readline()
        if ln: samples.append(ln.strip())
    p.terminate()
t = threading.Thread(target=sampler); t.start()
time.sleep(0.5)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 0
t0 = time.time(); e0.record()
while time.time() - t0 < 3.0:
    g.replay(); reps += 1
    if reps % 8 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
stop = True; t.join()
us = e0.elapsed_time(e1) * 1e3 / (reps * 50)
print(f"dbg={os.environ.get('TAPCLIP_GEMM_DEBUG','0')}: {us:.1f} us/GEMM  {2.0*M*N*K/us/1e6:.0f} TFLOP/s sustained; clocks/power samples (last 12): {samples[-12:]}")
