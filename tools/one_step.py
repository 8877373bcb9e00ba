"""A few C2 train steps and nothing else (for `ncu --metrics gpu__time_duration.sum` launch lists: no e2e / profiling passes).
usage: python tools/one_step.py [n_steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tapclip_b200 as tb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B, C, P = 128, 65, 16
clip = tb.CLIPWrapper("ViT-B-16-quickgelu", None, "cuda", seed=0, attribution="intended", dtype="mixed")
torch.manual_seed(4)
model = tb.FullModel([f"class_{i:03d}" for i in range(C)], clip, prompt_len=P, cache_text_features=False)
if os.environ.get("TAPCLIP_NO_OVERLAP") == "1":      # image tower first, then the text side, on one stream (ncu -s/-c by launch order)
    model.overlap_towers = False
opt = tb.FusedAdamW(model, lr=2e-3, weight_decay=0.01)
model.train()
g = torch.Generator().manual_seed(1)
images = [torch.randn(B, 3, 224, 224, generator=g).cuda() for _ in range(2)]
labels = [torch.randint(0, C, (B,), generator=g).cuda() for _ in range(2)]
for i in range(n):
    out = model(images[i % 2], labels[i % 2])
    opt.zero_grad()
    out["loss"].backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(out["loss"]))
