"""Micro-benchmark of the LayerNorm-free block (EPI_F32_RESID + folded-LayerNorm GEMMs) against the separate-kernel form,
per kernel and as the GEMM chain of one residual block (attention left out), at the image- and text-tower shapes of the C2 step.
CUDA-graph replay, no host time.  Usage (under gpurun): python tools/fused_ln_bench.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tapclip_b200 import _lib

lib = _lib.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
S = _lib.stream_ptr
P = _lib.ptr


def bench(fn, reps):
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


for tag, M, d, dt_name in (("image tower (B=128)", 25216, 768, "bf16"), ("text tower (C=65)", 6045, 512, "fp16"), ("text tower (C=345)", 32085, 512, "fp16")):
    tdt = torch.bfloat16 if dt_name == "bf16" else torch.float16
    DT = _lib.DTYPE[dt_name]
    dev = "cuda"
    x = torch.randn(M, d, device=dev)
    xb = torch.empty(M, d, device=dev, dtype=tdt)
    ln = torch.empty(M, d, device=dev, dtype=tdt)
    parts = lib.tapclip_op_gemm_stats_parts(d)
    stats = [torch.zeros(M, parts, 2, device=dev) for _ in range(2)]
    shift = [torch.zeros(M, device=dev) for _ in range(2)]
    cur = [0]
    qkv = torch.empty(M, 3 * d, device=dev, dtype=tdt)
    attn = torch.randn(M, d, device=dev).to(tdt)
    h = torch.empty(M, 4 * d, device=dev, dtype=tdt)
    g1, b1 = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    mk = lambda n, k: (torch.randn(n, k, device=dev) * k ** -0.5)
    w_qkv, w_o, w_fc, w_pr = mk(3 * d, d), mk(d, d), mk(4 * d, d), mk(d, 4 * d)
    c = lambda t: t.to(tdt)
    wq16, wo16, wf16, wp16 = c(w_qkv), c(w_o), c(w_fc), c(w_pr)
    bq, bo, bf_, bp = (torch.randn(n, device=dev) * 0.1 for n in (3 * d, d, 4 * d, d))

    def fold(w, b):
        wf = torch.empty_like(w, dtype=tdt); fb = torch.empty(w.shape[0], device=dev)
        _lib.check(lib.tapclip_op_fold_ln_weight(P(w), P(b), P(g1), P(b1), P(wf), DT, P(fb), w.shape[0], w.shape[1], S()))
        return wf, fb
    fq, ff = fold(w_qkv, bq), fold(w_fc, bf_)

    def gemm(a, w, b, out, Mm, N, K, epi, act):
        _lib.check(lib.tapclip_op_gemm(P(a), P(w), P(b), P(out), None, Mm, N, K, DT, epi, act, 0, S()))

    def lnk(out):
        _lib.check(lib.tapclip_op_layernorm(P(x), d, P(g1), P(b1), P(out), DT, None, M, d, S()))

    def resid(a, w, b, K):
        c0, c1 = cur[0], cur[0] ^ 1          # read the set describing x, write the other one (as the engine does)
        _lib.check(lib.tapclip_op_gemm_resid(P(a), P(w), P(b), P(x), 0, P(x), 0, P(xb), P(stats[c1]), P(shift[c1]), P(stats[c0]), P(shift[c0]),
                                             parts, M, d, K, DT, S()))
        cur[0] = c1

    def foldg(f, out, N, act):
        _lib.check(lib.tapclip_op_gemm_fold(P(xb), P(stats[cur[0]]), parts, P(f[0]), P(f[1]), P(out), None, M, N, d, DT, act, S()))

    resid(attn, wo16, bo, d)          # valid statistics for the folded GEMMs
    rows = [
        ("layernorm", lambda: lnk(ln)),
        ("qkv  plain", lambda: gemm(ln, wq16, bq, qkv, M, 3 * d, d, 0, -1)),
        ("qkv  folded", lambda: foldg(fq, qkv, 3 * d, -1)),
        ("out  += (red)", lambda: gemm(attn, wo16, bo, x, M, d, d, 2, -1)),
        ("out  resid+stats", lambda: resid(attn, wo16, bo, d)),
        ("fc   plain", lambda: gemm(ln, wf16, bf_, h, M, 4 * d, d, 0, 1)),
        ("fc   folded", lambda: foldg(ff, h, 4 * d, 1)),
        ("proj += (red)", lambda: gemm(h, wp16, bp, x, M, d, 4 * d, 2, -1)),
        ("proj resid+stats", lambda: resid(h, wp16, bp, 4 * d)),
    ]
    print(f"== {tag}: M={M} d={d} {dt_name}")
    t = {}
    for name, fn in rows:
        t[name] = bench(fn, reps)
        print(f"   {name:18s} {t[name]:8.1f} us")

    def chain_plain():
        lnk(ln); gemm(ln, wq16, bq, qkv, M, 3 * d, d, 0, -1); gemm(attn, wo16, bo, x, M, d, d, 2, -1)
        lnk(ln); gemm(ln, wf16, bf_, h, M, 4 * d, d, 0, 1); gemm(h, wp16, bp, x, M, d, 4 * d, 2, -1)

    def chain_fused():
        foldg(fq, qkv, 3 * d, -1); resid(attn, wo16, bo, d); foldg(ff, h, 4 * d, 1); resid(h, wp16, bp, 4 * d)

    cp, cf = bench(chain_plain, max(reps // 4, 3)), bench(chain_fused, max(reps // 4, 3))
    print(f"   block chain (no attention): separate LayerNorm {cp:8.1f} us   folded {cf:8.1f} us   ({cp - cf:+.1f} us per block)")
    x.zero_()
