#!/usr/bin/env python
"""bench.py — the TAP-CLIP hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train_c2|fwd_c1|fwd_b128|eval_c3|fwd_c4|train_c5]

One "step" = one pass of the hot path over one batch of synthetic input.  Default workload = BASELINE.json
configs[1]: ViT-B/16 prompt-tuning train step (attribution-instrumented forward + backward to the ctx vectors +
AdamW on the ctx bank), batch 128 per GPU, 65 classes, 16 ctx tokens, bf16 tensor-core operands.

N > 1 is launched by the driver as `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...`
(one rank per GPU, NCCL); images are data-parallel (weak scaling: 128 per GPU), class prompts are sharded.

Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, per-GPU batch, classes, ctx tokens, train?, description)
    "train_c2": ("ViT-B-16-quickgelu", 128, 65, 16, True,
                 "BASELINE configs[1]: ViT-B/16 prompt-tuning train step, batch 128/GPU, 65 classes, 16 ctx tokens, "
                 "fwd + attention-attribution + bwd to ctx + AdamW"),
    "fwd_c1": ("ViT-B-16-quickgelu", 8, 65, 16, False,
               "BASELINE configs[0]: ViT-B/16 attribution-instrumented forward, batch 8, 65 classes, 16 ctx tokens"),
    "fwd_b128": ("ViT-B-16-quickgelu", 128, 65, 16, False,
                 "north-star target shape: ViT-B/16 attribution-instrumented FORWARD (text attribution + per-layer CLS-row image "
                 "probes), batch 128, 65 classes, 16 ctx tokens"),
    "eval_c3": ("ViT-B-16-quickgelu", 256, 345, 16, False,
                "BASELINE configs[2]: ViT-B/16 cross-domain eval, 345 classes, batch 256/GPU, class-sharded text encoder"),
    "fwd_c4": ("ViT-L-14-336-quickgelu", 512, 65, 16, False,
               "BASELINE configs[3]: ViT-L/14@336 attribution-instrumented forward, batch 512/GPU, 65 classes, CLS-row probes of "
               "all 24 layers + attention rollout"),
    "train_c5": ("ViT-B-16-quickgelu", 128, 345, 16, True,
                 "BASELINE configs[4]: ViT-B/16 few-shot prompt-tuning train step, 345 classes, batch 128/GPU, class-sharded text "
                 "tower, ctx-gradient all-gather"),
}
IMAGE_ATTRIBUTION = {"fwd_c4": "rollout", "fwd_b128": "cls"}          # workloads whose step also emits the image-side attribution


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_sustained": d.get("bf16_tflops_sustained"), "bf16_burst": d.get("bf16_tflops"), "hbm": d.get("hbm_gbs"),
                "source": "MEASURED_PEAKS.json (of measured)"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "B200_PROFILING.md fallback (of fallback)"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own schedule (oracle.forward_as_written is a line-for-line,
# bit-exact restatement of models/model_wrapper.py:28-100; /root/reference itself cannot travel to the GPU box)
# ----------------------------------------------------------------------------------------------------------
_CPU_MODELS = {}


def cpu_reference_sample(model_name, B, C, P, train, sample_b, sample_c, reps=1):
    """Times the reference schedule on the host cores on a (sample_b, sample_c) sub-grid and extrapolates to (B, C).

    The two text loops are exactly linear in B*C and the image tower in B (model_wrapper.py:48,55), so
        T(B, C) = t_image(sample_b) * B/sample_b + (T(sample) - t_image(sample)) * (B*C)/(sample_b*sample_c).
    """
    import torch
    from oracle.clip_standin import StandInCLIPWrapper, get_config
    from oracle.tapclip_oracle import OracleFullModel, class_names, synthetic_images, synthetic_labels
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = get_config(model_name)
    key = (model_name, P, sample_b, sample_c)
    if key not in _CPU_MODELS:                                               # built once per process (the reference arm times it K+W times)
        wrapper = StandInCLIPWrapper(model_name, device="cpu", seed=0, attribution="intended")
        torch.manual_seed(4)
        model = OracleFullModel(class_names(sample_c), wrapper, prompt_len=P)
        model.train()
        opt = torch.optim.AdamW(model.prompt_learner.parameters(), lr=2e-3, weight_decay=0.01)
        images, labels = synthetic_images(sample_b, cfg.image_size), synthetic_labels(sample_b, sample_c)
        with torch.no_grad():
            wrapper.encode_image(images)                                     # warm-up (thread pool, allocator)
        _CPU_MODELS[key] = (wrapper, model, opt, images, labels)
    wrapper, model, opt, images, labels = _CPU_MODELS[key]
    t_img, t_all = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        with torch.no_grad():
            wrapper.encode_image(images)
        t_img.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        if train:
            out = model.forward_as_written(images, labels)
            opt.zero_grad(); out["loss"].backward(); opt.step()
        else:
            with torch.no_grad():
                model.forward_as_written(images)
        t_all.append(time.perf_counter() - t0)
    ti, ta = min(t_img), min(t_all)
    t_full = ti * (B / sample_b) + max(ta - ti, 0.0) * (B * C) / (sample_b * sample_c)
    return {"seconds_sample": ta, "seconds_image_part": ti, "seconds_full_extrapolated": t_full, "images_per_s": B / t_full,
            "cores": cores,
            "sample": f"reference schedule (oracle.forward_as_written{' + backward + AdamW' if train else ''}) timed on B'={sample_b}, "
                      f"C'={sample_c}, P={P}; extrapolated linearly to B={B}, C={C} (text loops ~ B*C, image tower ~ B)"}


def run_reference_arm(args, wl):
    model_name, B, C, P, train, desc = WORKLOADS[wl]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_sample(model_name, B, C, P, train, sample_b=4, sample_c=8)      # ~2 s of host work per step
        if i >= args.warmup:
            vals.append(last["images_per_s"])
    v = statistics.median(vals)
    line = {
        "impl": "reference", "metric": "images_per_sec", "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * B / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "steps_per_s": v / B,
        "config": {"workload": desc, "model": model_name, "batch_per_gpu": B, "n_cls": C, "prompt_len": P, "attribution": "intended"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": last["cores"], "kind": "port", "sample": last["sample"]},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    import tapclip_b200 as tb
    from tapclip_b200.configs import flops_per_image, flops_per_text_sequence, get_model_config

    model_name, B, C, P, train, desc = WORKLOADS[wl]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device=torch.device("cuda", local_rank))
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    dev = torch.device("cuda", local_rank)
    cfg = get_model_config(model_name)

    clip = tb.CLIPWrapper(model_name, None, "cuda", seed=0, attribution="intended", dtype=args.dtype)
    torch.manual_seed(4)
    model = tb.FullModel([f"class_{i:03d}" for i in range(C)], clip, prompt_len=P, cache_text_features=False,
                         image_attribution=IMAGE_ATTRIBUTION.get(wl))
    if os.environ.get("TAPCLIP_NO_OVERLAP") == "1":          # measurement switch: both towers on the caller's stream
        model.overlap_towers = False
    opt = tb.FusedAdamW(model, lr=2e-3, weight_decay=0.01) if train else None
    model.train(train)

    # synthetic inputs: a ring of distinct batches larger than L2 (126 MB) so no step finds its input cached
    n_ring = 4
    g = torch.Generator().manual_seed(1 + rank)
    host_images = [torch.randn(B, 3, cfg.image_size, cfg.image_size, generator=g).pin_memory() for _ in range(n_ring)]
    host_labels = [torch.randint(0, C, (B,), generator=g).pin_memory() for _ in range(n_ring)]
    dev_images = [t.to(dev) for t in host_images]
    dev_labels = [t.to(dev) for t in host_labels]

    def step(images, labels):
        if train:
            out = model(images, labels)
            opt.zero_grad()
            out["loss"].backward()
            opt.step()
            return out["loss"]
        with torch.no_grad():
            return model(images)["logits"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, fn):
        """n calls of fn(i) bracketed by barrier+sync, CUDA events on the current stream; returns max-over-ranks ms."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def resident(i):
        step(dev_images[i % n_ring], dev_labels[i % n_ring])

    # End-to-end: every step moves its own inputs host->device (pinned memory) and its result device->host.
    # As in a DataLoader(pin_memory=True) + non_blocking loop, the copy of batch i+1 is issued on a copy stream while
    # step i computes; the step's result is copied to pinned memory asynchronously and consumed one step later.
    copy_stream = torch.cuda.Stream(device=dev)
    slots = [{"im": torch.empty_like(dev_images[0]), "lb": torch.empty_like(dev_labels[0]), "ready": torch.cuda.Event(),
              "free": torch.cuda.Event()} for _ in range(2)]
    res_host = [(torch.zeros((), dtype=torch.float32) if train else torch.zeros(B, C, dtype=torch.float32)).pin_memory() for _ in range(2)]
    res_done = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"primed": -1, "checksum": 0.0}

    def prefetch(i):
        s = slots[i % 2]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(s["free"])
            s["im"].copy_(host_images[i % n_ring], non_blocking=True)
            s["lb"].copy_(host_labels[i % n_ring], non_blocking=True)
            s["ready"].record(copy_stream)
        e2e_state["primed"] = i

    def end_to_end(i):
        if e2e_state["primed"] < i:
            prefetch(i)
        prefetch(i + 1)
        s = slots[i % 2]
        main = torch.cuda.current_stream()
        main.wait_event(s["ready"])
        res = step(s["im"], s["lb"])
        s["free"].record(main)
        res_host[i % 2].copy_(res.detach(), non_blocking=True)
        res_done[i % 2].record(main)
        if i > 0:
            res_done[(i - 1) % 2].synchronize()                       # consume the previous step's result on the host
            e2e_state["checksum"] += float(res_host[(i - 1) % 2].sum())

    def e2e_flush(n):
        res_done[(n - 1) % 2].synchronize()
        e2e_state["checksum"] += float(res_host[(n - 1) % 2].sum())
        e2e_state["primed"] = -1
        for s in slots:
            s["free"] = torch.cuda.Event()

    for i in range(args.warmup):
        resident(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = clip.engine.launch_count
    ms_resident = timed(args.steps, resident)
    launches = clip.engine.launch_count - n0
    for i in range(3):
        end_to_end(i)
    e2e_flush(3)

    def e2e_loop(i):
        end_to_end(i)
        if i == args.steps - 1:
            e2e_flush(args.steps)
    ms_e2e = timed(args.steps, e2e_loop)
    clocks = sampler.stop() if rank == 0 else None

    # per-launch CUDA-event timing of the tensor-core kernels over the same K steps (roofline numbers); the towers run
    # back to back on one stream here so that each kernel's events time that kernel alone
    model.overlap_towers = False
    clip.engine.profile(True)
    ms_serial = timed(args.steps, resident)
    clip.engine.profile(False)
    prof = clip.engine.profile_report()
    model.overlap_towers = True

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    imgs_per_step = B * world
    value = imgs_per_step * args.steps / (ms_resident / 1e3)
    e2e_value = imgs_per_step * args.steps / (ms_e2e / 1e3)
    # algorithmic FLOPs of one step on ONE GPU (de-duplicated schedule, intended attribution => 2 text passes)
    c_local = -(-C // world)
    f_img, f_txt = flops_per_image(cfg), flops_per_text_sequence(cfg, P + cfg.context_length)
    flops_step = B * f_img + 2 * c_local * f_txt + 2 * B * C * cfg.embed_dim + (c_local * f_txt if train else 0)
    # executed FLOPs: the last vision block runs its out-projection and MLP on the CLS row only (dead-row elimination)
    # (and the text feature pass / its backward on position T-1 only)
    t_len = P + cfg.context_length
    dead = 0.0 if os.environ.get("TAPCLIP_DEAD_ROWS") == "0" else (
        B * (cfg.vision_tokens - 1) * 18.0 * cfg.vision_width ** 2
        + c_local * (t_len - 1) * 18.0 * cfg.text_width ** 2 * (2 if train else 1))
    gemm = prof.get("gemm", {"launches": 0, "ms": 0.0, "flops": 0.0})
    gemm_tflops = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else None
    top = sorted(((k, v) for k, v in prof.get("shapes", {}).items()), key=lambda kv: -kv[1]["ms"])[:8]
    line = {
        "metric": "images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_resident / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.dtype == "bf16" else "bf16 (image tower + all gradients) / fp16 (text-tower forward operands); fp32 accumulate + residual",
        "data": "synthetic", "steps_per_s": args.steps / (ms_resident / 1e3),
        "config": {"workload": desc, "model": model_name, "batch_per_gpu": B, "global_batch": imgs_per_step, "n_cls": C,
                   "prompt_len": P, "attribution": "intended", "optimizer": "FusedAdamW(lr=2e-3, wd=0.01)" if train else None,
                   "parallelism": f"dp{world} images + class-sharded text" if world > 1 else "single GPU",
                   "l2_policy": f"ring of {n_ring} distinct input batches ({n_ring * B * 3 * cfg.image_size ** 2 * 4 / 1e6:.0f} MB) > 126 MB L2; "
                                "per-step activations (>1 GB) exceed L2"},
        "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": B * 3 * cfg.image_size ** 2 * 4 + B * 8, "d2h_bytes_per_step": 4 if train else B * C * 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05/TMEM/TMA bf16 GEMM, all shapes of the step)",
            "achieved": gemm_tflops, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
            "frac": (gemm_tflops / peaks["bf16_sustained"]) if gemm_tflops else None,
            # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full` capture of the
            # four image-tower shapes, launch-weighted; their algorithmic bytes (A + W + output, residual read+write) average 215 MB, so the operands are
            # L2-resident and nothing is re-read from HBM.  Only reported for the workloads that capture describes.
            "traffic": 162.9e6 if wl in ("train_c2", "fwd_b128") else None,
            "traffic_note": "launch-weighted mean over the 45 image-tower GEMM launches of a step, profiles/r01n_ncu_full_gemm_tc_vision_layer0.csv",

            "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
            "gemm_launches_per_step": gemm["launches"] / args.steps, "gemm_ms_per_step": gemm["ms"] / args.steps,
            "gemm_share_of_step": gemm["ms"] / ms_serial if ms_serial else None,
            "ms_per_step_single_stream_profiled": ms_serial / args.steps,
            "attention_fwd_ms_per_step": prof.get("attention_fwd", {}).get("ms", 0.0) / args.steps,
            "attention_bwd_ms_per_step": prof.get("attention_bwd", {}).get("ms", 0.0) / args.steps,
            "step_algorithmic_tflop_per_gpu": flops_step / 1e12,
            "step_executed_tflop_per_gpu": (flops_step - dead) / 1e12,
            "step_tflops_achieved_per_gpu": flops_step / (ms_resident / args.steps * 1e-3) / 1e12,
            "step_frac_of_peak": flops_step / (ms_resident / args.steps * 1e-3) / 1e12 / peaks["bf16_sustained"],
            # per-launch DRAM traffic of the four image-tower GEMM shapes from the committed `ncu --set full` capture
            # (profiles/r01n_ncu_full_gemm_tc_vision_layer0.csv: dram__bytes_read.sum + dram__bytes_write.sum, MB)
            "ncu_dram_mb_per_launch": {"qkv M=25216 N=2304 K=768": 100.5, "out M=25216 N=768 K=768": 137.6,
                                       "fc M=25216 N=3072 K=768": 139.7, "proj M=25216 N=768 K=3072": 279.3},
            "top_shapes": [{"shape": k, "launches": v["launches"], "ms": round(v["ms"], 4),
                            "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["ms"] > 0 else None} for k, v in top],
        },
    }
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_sample(model_name, B, C, P, train, sample_b=8, sample_c=32)     # ~10 s of host work on 16 cores
        line["cpu_baseline"] = {"value": cb["images_per_s"], "unit": "images/s", "cores": cb["cores"], "kind": "port",
                                "sample": cb["sample"], "seconds_sample": cb["seconds_sample"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train_c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dtype", default="mixed", choices=["mixed", "bf16"],
                    help="'mixed' (default, meets the 1e-2 logit bar) or 'bf16' (bf16 operands everywhere)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args, args.workload)
    else:
        run_ours(args, args.workload)


if __name__ == "__main__":
    main()
