"""The N>1 path on CPU: world_size-2 gloo processes, engine replaced by the oracle-backed test double.

Checks: class-sharded text + all-gather == unsharded; data-parallel image split == single process;
all-reduced text-feature gradient + all-gathered ctx gradients == single-process gradients; ragged shards (C odd).
"""
import os
import socket
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, C, B_local, mode):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tapclip_b200 as tb
    from fake_engine import FakeWrapper
    from helpers import build_oracle
    from oracle.tapclip_oracle import class_names, synthetic_images, synthetic_labels
    from tapclip_b200.parallel import ClassSharding, all_gather_rows
    # ragged all-gather of rows
    shard = ClassSharding.current()
    lo, hi = shard.bounds(C)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 3)
    full = all_gather_rows(local, shard, C)
    assert torch.equal(full[:, 0], torch.arange(C, dtype=torch.float32))
    # data-parallel + class-sharded train step
    ow, _ = build_oracle("mini-16", C, 4, mode)
    torch.manual_seed(4)
    model = tb.FullModel(class_names(C), FakeWrapper(ow, mode), prompt_len=4)
    images, labels = synthetic_images(B_local * world, 64), synthetic_labels(B_local * world, C)
    sl = slice(rank * B_local, (rank + 1) * B_local)
    model.train()
    # count the collectives of one train step: text-feature all-gather, ONE packed all-reduce (d T^, d logit_scale, loss),
    # ctx-gradient all-gather
    calls = {"all_reduce": 0, "all_gather": 0}
    real_ar, real_ag = dist.all_reduce, dist.all_gather_into_tensor
    def count_ar(*a, **k): calls["all_reduce"] += 1; return real_ar(*a, **k)
    def count_ag(*a, **k): calls["all_gather"] += 1; return real_ag(*a, **k)
    dist.all_reduce, dist.all_gather_into_tensor = count_ar, count_ag
    out = model(images[sl], labels[sl])
    out["loss"].backward()
    dist.all_reduce, dist.all_gather_into_tensor = real_ar, real_ag
    g = torch.stack([model.prompt_learner.context_bank[n].grad for n in class_names(C)])
    sgrad = model.logit_scale.grad.detach().clone()
    # a loss that also uses the logits directly takes the general backward path (gradients not pre-reduced in forward)
    model.zero_grad()
    out2 = model(images[sl], labels[sl])
    (2.0 * out2["loss"] + 0.1 * out2["logits"].square().sum() / (B_local * world)).backward()
    g2 = torch.stack([model.prompt_learner.context_bank[n].grad for n in class_names(C)])
    torch.save({"logits": out["logits"].detach(), "loss": out["loss"].detach(), "grad": g, "sgrad": sgrad, "calls": calls,
                "grad2": g2, "sgrad2": model.logit_scale.grad.detach()}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,C", [("intended", 5), ("literal", 4)])
def test_two_rank_gloo_matches_single_process(mode, C):
    sys.path.insert(0, HERE)
    from helpers import build_oracle
    from oracle.tapclip_oracle import class_names, synthetic_images, synthetic_labels
    world, B_local = 2, 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d, C, B_local, mode), nprocs=world, join=True)
        parts = [torch.load(os.path.join(d, f"r{r}.pt")) for r in range(world)]
    _, om = build_oracle("mini-16", C, 4, mode)
    om.train()
    images, labels = synthetic_images(B_local * world, 64), synthetic_labels(B_local * world, C)
    ref = om.forward_dedup(images, labels)
    ref["loss"].backward()
    g_ref = torch.stack([om.prompt_learner.context_bank[n].grad for n in class_names(C)])
    logits = torch.cat([p["logits"] for p in parts], 0)
    assert (logits - ref["logits"]).abs().max().item() < 5e-5
    for p in parts:                                   # every replica holds the global loss and the full identical gradient
        assert abs(p["loss"].item() - ref["loss"].item()) < 1e-5
        assert ((p["grad"] - g_ref).norm() / g_ref.norm()).item() < 1e-4
        assert abs(p["sgrad"].item() - om.logit_scale.grad.item()) < 1e-5
    assert torch.equal(parts[0]["grad"], parts[1]["grad"])
    assert parts[0]["calls"] == {"all_reduce": 1, "all_gather": 2}        # three collectives per train step (five before the packing)
    # general path: d/dctx of 2*loss + 0.1*mean_b sum_c logits^2
    om.zero_grad()
    ref2 = om.forward_dedup(images, labels)
    (2.0 * ref2["loss"] + 0.1 * ref2["logits"].square().sum() / (B_local * world)).backward()
    g_ref2 = torch.stack([om.prompt_learner.context_bank[n].grad for n in class_names(C)])
    for p in parts:
        assert ((p["grad2"] - g_ref2).norm() / g_ref2.norm()).item() < 1e-4
        assert abs(p["sgrad2"].item() - om.logit_scale.grad.item()) < 1e-4
