"""Experiment: does running two halves of the class set on two streams (two engines) shorten the text tower's latency chain?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import tapclip_b200 as tb

C, P = 65, 16
clip = tb.CLIPWrapper("ViT-B-16-quickgelu", None, "cuda", seed=0, attribution="intended", dtype="mixed")
clip2 = tb.CLIPWrapper("ViT-B-16-quickgelu", None, "cuda", seed=0, attribution="intended", dtype="mixed")
torch.manual_seed(4)
model = tb.FullModel([f"class_{i:03d}" for i in range(C)], clip, prompt_len=P, cache_text_features=False)
ctx, tok = model.prompt_learner.flat_ctx(), model.prompt_learner.flat_tok()
e1, e2 = clip.engine, clip2.engine
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h = (C + 1) // 2
dfeat = torch.randn(C, 512, device="cuda") * 1e-3


def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def fwd_full():
    e1.text_forward(ctx, tok, "intended", True)


def fwd_split():
    main = torch.cuda.current_stream()
    s1.wait_stream(main); s2.wait_stream(main)
    with torch.cuda.stream(s1):
        e1.text_forward(ctx[:h], tok[:h], "intended", True)
    with torch.cuda.stream(s2):
        e2.text_forward(ctx[h:], tok[h:], "intended", True)
    main.wait_stream(s1); main.wait_stream(s2)


def bwd_full():
    e1.text_backward(dfeat, C, P)


def bwd_split():
    main = torch.cuda.current_stream()
    s1.wait_stream(main); s2.wait_stream(main)
    with torch.cuda.stream(s1):
        e1.text_backward(dfeat[:h].contiguous(), h, P)
    with torch.cuda.stream(s2):
        e2.text_backward(dfeat[h:].contiguous(), C - h, P)
    main.wait_stream(s1); main.wait_stream(s2)


print(f"text forward  full  : {timeit(fwd_full):.3f} ms"); t = timeit(bwd_full); print(f"text backward full  : {t:.3f} ms")
print(f"text forward  2-way : {timeit(fwd_split):.3f} ms"); t = timeit(bwd_split); print(f"text backward 2-way : {t:.3f} ms")

import time


def host_cost(fn, reps=10):
    """Host time of ONE call with an empty launch queue (synchronise before each call)."""
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    return 1e3 * sorted(ts)[len(ts) // 2]


images = torch.randn(128, 3, 224, 224, device="cuda")
print(f"host enqueue cost per call (median): text forward {host_cost(fwd_full):.3f} ms (170 launches), "
      f"text backward {host_cost(bwd_full):.3f} ms (88), image tower {host_cost(lambda: e1.encode_image(images)):.3f} ms (89)")
