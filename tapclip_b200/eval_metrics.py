"""Drop-in for ``utils/eval_metrics.py`` (evaluate_accuracy :7-41, evaluate_per_class_accuracy :44-73) with the
consumer fused on the device.

The reference pulls every prediction to the host (``.item()`` per sample, eval_metrics.py:22-29,58-63).  Here the argmax
and the overall / per-class correct and total counters stay on the GPU (``tapclip_argmax_count``: int32 ``[C]`` arrays updated
with atomics) and ONE device->host read per epoch returns them.  Same signatures, same return values, same printed report.
"""
from __future__ import annotations

import torch


class AccuracyCounters:
    """Device-side accumulators of one evaluation epoch."""

    def __init__(self, n_cls: int, device):
        self.correct = torch.zeros(1, dtype=torch.int32, device=device)
        self.class_correct = torch.zeros(n_cls, dtype=torch.int32, device=device)
        self.class_total = torch.zeros(n_cls, dtype=torch.int32, device=device)

    def update(self, engine, logits, labels):
        engine.argmax_count(logits, labels, counters=(self.correct, self.class_correct, self.class_total))

    def read(self):
        """The epoch's single device->host transfer: (correct, per-class correct list, per-class total list)."""
        packed = torch.cat([self.correct, self.class_correct, self.class_total]).cpu()
        n = self.class_correct.numel()
        return int(packed[0]), packed[1:1 + n].tolist(), packed[1 + n:].tolist()


def _engine_of(model):
    eng = getattr(getattr(model, "clip", None), "engine", None)
    if eng is None:
        raise TypeError("evaluate_accuracy needs a tapclip_b200.FullModel (the counters live in its engine)")
    return eng


def _run_epoch(model, dataloader, device):
    model.eval()
    counters, total = None, 0
    for images, labels in dataloader:
        images = images.to(device, non_blocking=True)
        labels = labels.to(device, non_blocking=True)
        logits = model(images)["logits"]
        if counters is None:
            counters = AccuracyCounters(logits.shape[1], logits.device)
        counters.update(_engine_of(model), logits, labels.long())
        total += labels.numel()                          # eval_metrics.py:23 (host-side count: no device read)
    if counters is None:
        return 0, [], [], 0
    return (*counters.read(), total)


@torch.no_grad()
def evaluate_accuracy(model, dataloader, device):
    """utils/eval_metrics.py:7-41: overall accuracy in percent (returned) and the per-class report (printed)."""
    correct, cls_ok, cls_n, total = _run_epoch(model, dataloader, device)
    acc = 100.0 * correct / total if total > 0 else 0.0
    print(f"🎯 Overall Accuracy: {acc:.2f}%")
    print("📊 Per-Class Accuracy:")
    for cls, (ok, n) in enumerate(zip(cls_ok, cls_n)):
        if n > 0:                                        # the reference iterates over the classes that occurred
            print(f" - Class {cls:2d}: {100.0 * ok / n:.2f}% ({ok}/{n})")
    return acc


@torch.no_grad()
def evaluate_per_class_accuracy(model, dataloader, device, class_names=None):
    """utils/eval_metrics.py:44-73: {class name or index string: accuracy in percent} for the classes that occurred."""
    _, cls_ok, cls_n, _ = _run_epoch(model, dataloader, device)
    return {(class_names[cls] if class_names else str(cls)): 100.0 * ok / n
            for cls, (ok, n) in enumerate(zip(cls_ok, cls_n)) if n > 0}
