"""Time the three phases of the C2 train step in isolation (CUDA events, 20 reps): image tower, text forward (both
passes, activations saved), text backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tapclip_b200 as tb

B, C, P = 128, 65, 16
clip = tb.CLIPWrapper("ViT-B-16-quickgelu", None, "cuda", seed=0, attribution="intended", dtype="mixed")
torch.manual_seed(4)
model = tb.FullModel([f"class_{i:03d}" for i in range(C)], clip, prompt_len=P, cache_text_features=False)
eng = clip.engine
images = torch.randn(B, 3, 224, 224, device="cuda")
ctx, tok = model.prompt_learner.flat_ctx(), model.prompt_learner.flat_tok()
dfeat = torch.randn(C, 512, device="cuda") * 1e-3


def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


n0 = eng.launch_count; eng.encode_image(images); n_img = eng.launch_count - n0
n0 = eng.launch_count; eng.text_forward(ctx, tok, "intended", True); n_txt = eng.launch_count - n0
n0 = eng.launch_count; eng.text_backward(dfeat, C, P); n_bwd = eng.launch_count - n0
print(f"image tower   B={B}: {timeit(lambda: eng.encode_image(images)):7.3f} ms  ({n_img} launches)")
print(f"text forward  C={C}: {timeit(lambda: eng.text_forward(ctx, tok, 'intended', True)):7.3f} ms  ({n_txt} launches, 2 passes)")
print(f"text forward literal: {timeit(lambda: eng.text_forward(ctx, tok, 'literal', True)):7.3f} ms")
print(f"text backward C={C}: {timeit(lambda: eng.text_backward(dfeat, C, P)):7.3f} ms  ({n_bwd} launches)")
