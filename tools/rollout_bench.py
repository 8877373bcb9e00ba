"""Per-launch time of one rollout step (rollout.cu) and of the statistics kernel at the bench shapes.
Usage: python tools/rollout_bench.py [S N H]   (default: ViT-L/14@336 B=512 and ViT-B/16 B=128)"""
import sys

import torch

sys.path.insert(0, ".")
from tapclip_b200 import _lib  # noqa: E402


def run(S, N, H, reps=10):
    lib = _lib.load()
    d = H * 64
    qkv = (torch.randn(S * N, 3 * d, device="cuda") * 1.5).to(torch.bfloat16)
    out = torch.empty(S * N, d, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(S, H, N, device="cuda")
    r_in = torch.rand(S, N, device="cuda")
    r_out = torch.empty(S, N, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2]

    t_fwd = timed(lambda: _lib.check(lib.tapclip_op_attention(_lib.ptr(qkv), _lib.ptr(out), 1, S, N, H, 0, None, 0, 0, _lib.stream_ptr())))
    rows = torch.empty(S, H, N, device="cuda")
    t_fwd_cls = timed(lambda: _lib.check(lib.tapclip_op_attention(_lib.ptr(qkv), _lib.ptr(out), 1, S, N, H, 2, _lib.ptr(rows), 0, H * N, _lib.stream_ptr())))
    t_fwd_lse = timed(lambda: _lib.check(lib.tapclip_op_attention_lse(_lib.ptr(qkv), _lib.ptr(out), _lib.ptr(lse), 1, S, N, H, _lib.stream_ptr())))
    t_lse = timed(lambda: _lib.check(lib.tapclip_op_attention_lse(_lib.ptr(qkv), None, _lib.ptr(lse), 1, S, N, H, _lib.stream_ptr())))
    t_step = timed(lambda: _lib.check(lib.tapclip_op_rollout_step(_lib.ptr(qkv), _lib.ptr(lse), _lib.ptr(r_in), _lib.ptr(r_out), 1, S, N, H, 0, _lib.stream_ptr())))
    n_exp = S * H * N * N
    print(f"S={S} N={N} H={H}: attention fwd {t_fwd:.3f} ms, fwd+CLS probe {t_fwd_cls:.3f} ms, fwd+lse {t_fwd_lse:.3f} ms, statistics alone {t_lse:.3f} ms, "
          f"rollout step {t_step:.3f} ms ({n_exp / t_step * 1e-9:.2f} T exp2/s, {2 * n_exp * 64 / t_step * 1e-9:.0f} TFLOP/s QK^T)")


if __name__ == "__main__":
    if len(sys.argv) == 4:
        run(*map(int, sys.argv[1:]))
    else:
        run(512, 577, 16)
        run(128, 197, 12)
