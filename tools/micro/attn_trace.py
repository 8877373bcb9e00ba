"""Developer tool: phase-by-phase cycle trace of the persistent tcgen05 attention kernel (CTA 0, both softmax groups).
Builds a private copy of the library with -DTAPCLIP_ATTN_TRACE into build/ and runs one vision-layer launch.

  python tools/micro/attn_trace.py build     # here (no GPU): compile build/libtapclip_trace.so
  python tools/micro/attn_trace.py           # on the GPU box: run + print
"""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "build", os.environ.get("TRACE_LIB", "libtapclip_trace.so"))
SRC = os.path.join(ROOT, "tapclip_b200", "csrc")
if len(sys.argv) > 1 and sys.argv[1] == "build":
    srcs = [f for f in sorted(os.listdir(SRC)) if f.endswith(".cu") and f != "attention_tc.cu"]
    objs = [os.path.join(ROOT, "build", "obj", s.replace(".cu", ".o")) for s in srcs]
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler",
                           "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr", "-DTAPCLIP_ATTN_TRACE", *os.environ.get("TRACE_DEFS", "").split(), "-I", os.path.join(ROOT, "include"),
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
buf = (C.c_longlong * 2048)()
assert lib.tapclip_debug_attn_trace(buf) == 0
base = min(buf[((g * 4 + q) * 16) * 16] for g in range(2) for q in range(4) if buf[((g * 4 + q) * 16) * 16])
for g in range(2):
    for q in range(4):
        print(f"group {g} warp q={q} of CTA 0, cycles")
        for it in range(1, 16):
            t = [buf[((g * 4 + q) * 16 + it) * 16 + k] for k in range(16)]
            if t[0] == 0: break
            print(f"  item {it:2d} start {t[0]-base:7d} | S wait {t[3]-t[0]:5d} | ld {t[4]-t[3]:4d} | max {t[6]-t[4]:4d} | exp+st {t[5]-t[6]:5d} | arrive {t[7]-t[5]:4d} | O wait {t[8]-t[7]:5d} | drain {t[9]-t[8]:5d} | P at {t[5]-base:7d} O at {t[8]-base:7d} | O ld {t[10]-t[8]:4d} sts {t[11]-t[10]:4d} lds {t[12]-t[11]:4d} stg {t[13]-t[12]:4d}")
# one-line summary per group (warp q=0): mean phase lengths over items 2..8 and the item period
for g in range(2):
    rows = []
    for it in range(2, 9):
        t = [buf[((g * 4) * 16 + it) * 16 + k] for k in range(16)]
        if t[0] == 0: break
        rows.append(t)
    if len(rows) > 1:
        m = lambda f: sum(f(t) for t in rows) / len(rows)
        print(f"SUMMARY g{g}: period {(rows[-1][0] - rows[0][0]) / (len(rows) - 1):6.0f} | S wait {m(lambda t: t[3]-t[0]):5.0f} | ld {m(lambda t: t[4]-t[3]):5.0f} | max(+turn wait) {m(lambda t: t[6]-t[4]):5.0f} | "
              f"exp+st {m(lambda t: t[5]-t[6]):5.0f} | arrive {m(lambda t: t[7]-t[5]):5.0f} | O wait {m(lambda t: t[8]-t[7]):5.0f} | drain {m(lambda t: t[9]-t[8]):5.0f} "
              f"(O ld {m(lambda t: t[10]-t[8]):4.0f}, scale+sts {m(lambda t: t[11]-t[10]):4.0f}, lds {m(lambda t: t[12]-t[11]):4.0f}, stg {m(lambda t: t[13]-t[12]):4.0f}, fence {m(lambda t: t[9]-t[13]):4.0f})")
