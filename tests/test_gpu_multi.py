"""N>1 on real GPUs (NCCL): data-parallel images + class-sharded text must reproduce the single-GPU logits, loss and
ctx gradients.  Needs >= 2 visible GPUs (skipped otherwise); run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import subprocess
import sys
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["TAPCLIP_ROOT"])
import tapclip_b200 as tb
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
C, P, B = 7, 4, 4                                     # ragged class shards (7 over 2 ranks), 4 images per rank
clip = tb.CLIPWrapper("mini-16", None, "cuda", seed=0, attribution="intended", dtype="fp32")
torch.manual_seed(4)
model = tb.FullModel([f"class_{i:03d}" for i in range(C)], clip, prompt_len=P)
g = torch.Generator().manual_seed(1)
images = torch.randn(B * world, 3, 64, 64, generator=g).cuda()
labels = torch.randint(0, C, (B * world,), generator=g).cuda()
model.train()
opt = tb.FusedAdamW(model, lr=2e-3, weight_decay=0.01)
mine = slice(rank * B, (rank + 1) * B)
for step in range(3):                                 # three steps: both parities of the fused gather's slots, updated ctx
    out = model(images[mine].contiguous(), labels[mine].contiguous())
    opt.zero_grad()
    out["loss"].backward()
    if step < 2:
        opt.step()
grad = torch.stack([p.grad for p in model.prompt_learner.context_bank.values()])
model.eval()
model.cache_text_features = False
with torch.no_grad():
    ev = model(images[mine].contiguous())["logits"]
# a second model on the same CLIPWrapper (its own gather buffers) must not disturb the first one
torch.manual_seed(9)
other = tb.FullModel([f"class_{i:03d}" for i in range(C + 2)], clip, prompt_len=P, cache_text_features=False).eval()
with torch.no_grad():
    other(images[mine].contiguous())
    ev2 = model(images[mine].contiguous())["logits"]
assert torch.equal(ev, ev2)
torch.save({"logits": out["logits"].detach().cpu(), "loss": out["loss"].detach().cpu(), "grad": grad.cpu(),
            "sgrad": model.logit_scale.grad.cpu(), "eval_logits": ev.cpu(), "fused_gather": model._tg is not None,
            "epoch": model._tg.epoch if model._tg is not None else 0},
           os.path.join(os.environ["TAPCLIP_OUT"], f"r{rank}.pt"))
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("fused_gather", ["1", "0"])
def test_two_gpu_nccl_matches_single_gpu(fused_gather):
    """fused_gather=1: the text features cross NVLink as peer stores from the head kernel into every rank's symmetric buffer and the
    logits kernel waits on epoch flags (parallel.TextGather, no NCCL call on that path); 0: one NCCL all-gather."""
    import tapclip_b200 as tb
    world, C, P, B = 2, 7, 4, 4
    with tempfile.TemporaryDirectory() as d:
        script = os.path.join(d, "worker.py")
        open(script, "w").write(WORKER)
        env = dict(os.environ, TAPCLIP_ROOT=ROOT, TAPCLIP_OUT=d, TAPCLIP_FUSED_GATHER=fused_gather)
        subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", script], check=True, env=env, timeout=600)
        parts = [torch.load(os.path.join(d, f"r{r}.pt")) for r in range(world)]
    clip = tb.CLIPWrapper("mini-16", None, "cuda", seed=0, attribution="intended", dtype="fp32")
    torch.manual_seed(4)
    model = tb.FullModel([f"class_{i:03d}" for i in range(C)], clip, prompt_len=P, distributed=False)
    g = torch.Generator().manual_seed(1)
    images = torch.randn(B * world, 3, 64, 64, generator=g).cuda()
    labels = torch.randint(0, C, (B * world,), generator=g).cuda()
    model.train()
    opt = tb.FusedAdamW(model, lr=2e-3, weight_decay=0.01)
    for step in range(3):
        out = model(images, labels)
        opt.zero_grad()
        out["loss"].backward()
        if step < 2:
            opt.step()
    grad = torch.stack([p.grad for p in model.prompt_learner.context_bank.values()]).cpu()
    model.eval()
    with torch.no_grad():
        ev = model(images)["logits"].cpu()
    assert all(p["fused_gather"] == (fused_gather == "1") for p in parts)
    if fused_gather == "1":
        assert all(p["epoch"] == 5 for p in parts)                 # three train steps + two uncached eval forwards
    assert (torch.cat([p["eval_logits"] for p in parts], 0) - ev).abs().max().item() < 1e-5
    logits = torch.cat([p["logits"] for p in parts], 0)
    assert (logits - out["logits"].detach().cpu()).abs().max().item() < 1e-5
    for p in parts:
        assert abs(p["loss"].item() - out["loss"].item()) < 1e-5
        assert ((p["grad"] - grad).norm() / grad.norm()).item() < 1e-4
        assert abs(p["sgrad"].item() - model.logit_scale.grad.item()) < 1e-5
    assert torch.equal(parts[0]["grad"], parts[1]["grad"])
