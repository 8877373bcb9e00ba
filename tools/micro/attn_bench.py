"""Developer tool: one attention layer through the op-level C-ABI (tapclip_op_attention), timed with CUDA events over inputs
rotated through a ring larger than L2, and checked against torch SDPA.  Shapes: ViT-B/16 image tower (S=128, N=197, H=12, bf16)
and the text tower of the C2 step (S=130, N=93, H=8, fp16).
  TAPCLIP_ATTN_IMPL=1|2 python tools/micro/attn_bench.py     # 1 = mma.sync kernel, 2 = tcgen05 kernel, unset = dispatch
"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = C.CDLL(os.environ.get("TAPCLIP_LIB", os.path.join(ROOT, "tapclip_b200", "lib", "libtapclip.so")))
lib.tapclip_op_attention.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]
lib.tapclip_op_attention.restype = C.c_int32
DT = {torch.bfloat16: 1, torch.float16: 2}


def run(S, N, H, dtype, iters=40, probe=0):
    d = H * 64
    ring = max(2, int(400e6 // (S * N * 3 * d * 2)) + 1)
    torch.manual_seed(0)
    qkv = [torch.randn(S * N, 3 * d, device="cuda").to(dtype) for _ in range(ring)]
    out = torch.empty(S * N, d, device="cuda", dtype=dtype)
    st = torch.cuda.current_stream().cuda_stream
    pb = torch.zeros(S * H * N, device="cuda") if probe else None      # 1: text-column probe (P = 16), 2: CLS row
    call = lambda t: lib.tapclip_op_attention(t.data_ptr(), out.data_ptr(), DT[dtype], S, N, H, probe, pb.data_ptr() if probe else None,
                                              16 if probe == 1 else 0, H * N if probe == 2 else 0, st)
    assert call(qkv[0]) == 0
    q, k, v = (t.view(S, N, H, 64).transpose(1, 2).float() for t in qkv[0].view(S, N, 3, d).unbind(2))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(S * N, d)
    err = (out.float() - ref).abs().max().item()
    for i in range(5): call(qkv[i % ring])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(iters): call(qkv[i % ring])
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    print(f"S={S} N={N} H={H} {str(dtype)[6:]} probe={probe}: {us:7.1f} us per layer   max|out - sdpa_fp32| = {err:.2e}   "
          f"({4 * S * H * N * N * 64 / us / 1e6:.0f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    print("TAPCLIP_ATTN_IMPL =", os.environ.get("TAPCLIP_ATTN_IMPL", "(auto)"))
    run(128, 197, 12, torch.bfloat16)
    run(128, 197, 12, torch.bfloat16, probe=2)
    run(130, 93, 8, torch.float16)
    run(65, 93, 8, torch.float16)
    run(65, 93, 8, torch.float16, probe=1)
    run(65, 141, 8, torch.float16)
    run(512, 197, 12, torch.bfloat16)
    run(128, 50, 12, torch.bfloat16)
