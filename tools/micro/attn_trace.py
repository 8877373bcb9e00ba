"""Developer tool: phase-by-phase cycle trace of the persistent tcgen05 attention kernel (CTA 0, both softmax groups).
Builds a private copy of the library with -DTAPCLIP_ATTN_TRACE into build/ and runs one vision-layer launch.

  python tools/micro/attn_trace.py build     # here (no GPU): compile build/libtapclip_trace.so
  python tools/micro/attn_trace.py           # on the GPU box: run + print
"""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "build", "libtapclip_trace.so")
SRC = os.path.join(ROOT, "tapclip_b200", "csrc")
if len(sys.argv) > 1 and sys.argv[1] == "build":
    srcs = "capi.cu engine.cu gemm_tc.cu gemm_simt.cu norm.cu attention.cu elementwise.cu preprocess.cu".split()
    objs = [os.path.join(ROOT, "build", "obj", s.replace(".cu", ".o")) for s in srcs]
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler",
                           "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr", "-DTAPCLIP_ATTN_TRACE", "-I", os.path.join(ROOT, "include"),
                           "-shared", "-o", LIB, os.path.join(SRC, "attention_tc.cu")] + objs + ["-lcudart"])
    sys.exit(0)
import torch
lib = C.CDLL(LIB)
S, N, H = (int(os.environ.get("S", 128)), int(os.environ.get("N", 197)), 12)
d = H * 64
qkv = torch.randn(S * N, 3 * d, device="cuda").bfloat16()
out = torch.empty(S * N, d, device="cuda", dtype=torch.bfloat16)
lib.tapclip_op_attention.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]
os.environ["TAPCLIP_ATTN_IMPL"] = "2"
for _ in range(3):
    assert lib.tapclip_op_attention(qkv.data_ptr(), out.data_ptr(), 1, S, N, H, 0, None, 0, 0, None) == 0
torch.cuda.synchronize()
buf = (C.c_longlong * 512)()
assert lib.tapclip_debug_attn_trace(buf) == 0
for g in range(2):
    print(f"group {g} (warp q=0 of CTA 0), cycles")
    base = buf[(g * 16) * 16]
    for it in range(12):
        t = [buf[(g * 16 + it) * 16 + k] for k in range(16)]
        if t[0] == 0: break
        print(f"  item {it:2d} start {t[0]-base:7d} | S wait {t[3]-t[0]:5d} | ld {t[4]-t[3]:4d} | max {t[6]-t[4]:4d} | exp+st {t[5]-t[6]:5d} | arrive {t[7]-t[5]:4d} | O wait {t[8]-t[7]:5d} | drain {t[9]-t[8]:5d}")
