"""ncu target: the residual (EPI_F32_RESID) and folded-LayerNorm GEMMs next to their plain forms at the image-tower shapes of
the C2 step (out-projection 25216x768x768, c_fc 25216x3072x768), 3 launches each.  Usage (under gpurun):
    python tools/one_gemm.py && ncu --set full --import-source on -k regex:gemm_tc_kernel -c 12 -o gpurun_out/gemm_epi python tools/one_gemm.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tapclip_b200 import _lib

lib = _lib.load()
P, S = _lib.ptr, _lib.stream_ptr
M, d = 25216, 768
DT = _lib.DTYPE["bf16"]
dev = "cuda"
x = torch.randn(M, d, device=dev)
xb = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
parts = lib.tapclip_op_gemm_stats_parts(d)
stats = torch.zeros(M, parts, 2, device=dev)
stats_b, shift_a, shift_b = torch.zeros_like(stats), torch.zeros(M, device=dev), torch.zeros(M, device=dev)
attn = torch.randn(M, d, device=dev).bfloat16()
ln = torch.randn(M, d, device=dev).bfloat16()
h = torch.empty(M, 4 * d, device=dev, dtype=torch.bfloat16)
w_o = (torch.randn(d, d, device=dev) * d ** -0.5).bfloat16()
w_fc = torch.randn(4 * d, d, device=dev) * d ** -0.5
b_o, b_fc = torch.randn(d, device=dev) * 0.1, torch.randn(4 * d, device=dev) * 0.1
g1, b1 = torch.ones(d, device=dev), torch.zeros(d, device=dev)
wf = torch.empty(4 * d, d, device=dev, dtype=torch.bfloat16)
fb = torch.empty(4 * d, device=dev)
_lib.check(lib.tapclip_op_fold_ln_weight(P(w_fc), P(b_fc), P(g1), P(b1), P(wf), DT, P(fb), 4 * d, d, S()))
w_fc16 = w_fc.bfloat16()
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # L2 flush between launches
for _ in range(3):
    big.zero_()
    _lib.check(lib.tapclip_op_gemm_resid(P(attn), P(w_o), P(b_o), P(x), 0, P(x), 0, P(xb), P(stats), P(shift_a), P(stats_b), P(shift_b), parts,
                                         M, d, d, DT, S()))
for _ in range(3):
    big.zero_()
    _lib.check(lib.tapclip_op_gemm(P(attn), P(w_o), P(b_o), P(x), None, M, d, d, DT, 2, -1, 0, S()))
for _ in range(3):
    big.zero_()
    _lib.check(lib.tapclip_op_gemm_fold(P(xb), P(stats), parts, P(wf), P(fb), P(h), None, M, 4 * d, d, DT, 1, S()))
for _ in range(3):
    big.zero_()
    _lib.check(lib.tapclip_op_gemm(P(ln), P(w_fc16), P(b_fc), P(h), None, M, 4 * d, d, DT, 0, 1, 0, S()))
torch.cuda.synchronize()
print("ok")
