"""Two configs[3] forwards (ViT-L/14@336, B=512, CLS probes + rollout) and nothing else, for ncu launch lists.
usage: python tools/one_step_c4.py [n_steps] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tapclip_b200 as tb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
C, P = 65, 16
clip = tb.CLIPWrapper("ViT-L-14-336-quickgelu", None, "cuda", seed=0, attribution="intended", dtype="mixed")
torch.manual_seed(4)
model = tb.FullModel([f"class_{i:03d}" for i in range(C)], clip, prompt_len=P, cache_text_features=False, image_attribution="rollout")
model.eval()
g = torch.Generator().manual_seed(1)
images = torch.randn(B, 3, 336, 336, generator=g).cuda()
with torch.no_grad():
    for i in range(n):
        out = model(images)
torch.cuda.synchronize()
print("logits", float(out["logits"].abs().mean()))
