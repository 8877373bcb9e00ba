"""Multi-GPU plumbing of the hot path: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch on B200; gloo in the CPU tests).

* images are data-parallel (each rank encodes its own batch; frozen weights are replicated);
* class prompts are sharded: rank r owns classes ``bounds(n_cls)``; text features ``[C_r, E]`` are all-gathered
  (C*E*4 bytes: 133 KB at C=65, 707 KB at C=345 — latency-bound);
* backward: the text-feature gradient ``[C, E]`` is all-reduced, each rank back-propagates its own class shard and
  the ctx gradients ``[C_r, P, D]`` are all-gathered so every replica's optimizer sees the full identical gradient.

Shards may be ragged (C not divisible by the world size): rows are padded to the largest shard for the
collective and stripped afterwards.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class ClassSharding:
    rank: int = 0
    world: int = 1
    group: object = None

    @staticmethod
    def current(distributed=None, group=None) -> "ClassSharding":
        on = dist.is_available() and dist.is_initialized() if distributed is None else bool(distributed)
        if not on:
            return ClassSharding()
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("distributed=True needs torch.distributed.init_process_group() first")
        return ClassSharding(dist.get_rank(group), dist.get_world_size(group), group)

    def bounds(self, n_cls: int, rank: int | None = None):
        """Contiguous balanced partition: the first (n_cls % world) ranks own one extra class."""
        r = self.rank if rank is None else rank
        base, extra = divmod(n_cls, self.world)
        lo = r * base + min(r, extra)
        return lo, lo + base + (1 if r < extra else 0)

    def max_shard(self, n_cls: int) -> int:
        return -(-n_cls // self.world)

    def global_batch(self, local_batch: int) -> int:
        """Per-rank batches are assumed equal (weak scaling); the loss is the mean over world*local samples."""
        return local_batch * self.world


_UNPAD_INDEX = {}


def _unpad_index(shard: ClassSharding, n_rows_total: int, device) -> torch.Tensor:
    """Row indices that pick every rank's real rows out of the gathered [world * max_shard, W] block (cached)."""
    key = (shard.world, n_rows_total, str(device))
    idx = _UNPAD_INDEX.get(key)
    if idx is None:
        m = shard.max_shard(n_rows_total)
        rows = []
        for r in range(shard.world):
            lo, hi = shard.bounds(n_rows_total, r)
            rows.extend(range(r * m, r * m + (hi - lo)))
        idx = _UNPAD_INDEX[key] = torch.tensor(rows, dtype=torch.int64, device=device)
    return idx


def all_gather_rows(local: torch.Tensor, shard: ClassSharding, n_rows_total: int) -> torch.Tensor:
    """Concatenate every rank's ``[rows_r, W]`` block (rows_r = shard.bounds) into ``[n_rows_total, W]``: ONE collective into a
    pre-sized buffer.  Even shards land at their final offsets (no copy before or after); ragged shards are gathered at a
    fixed pitch of max_shard rows and compacted with one cached index_select."""
    if shard.world == 1:
        return local
    width = local.shape[1]
    m = shard.max_shard(n_rows_total)
    even = n_rows_total % shard.world == 0
    src = local.contiguous()
    if not even and local.shape[0] < m:                      # only the short ranks of a ragged partition pad their block
        src = local.new_zeros(m, width)
        src[: local.shape[0]] = local
    out = local.new_empty(shard.world * m, width)
    dist.all_gather_into_tensor(out, src, group=shard.group)
    return out if even else out.index_select(0, _unpad_index(shard, n_rows_total, local.device))


class TextGather:
    """Fused text-feature all-gather: every rank owns a symmetric buffer (``torch.distributed._symmetric_memory``:
    one allocation mapped into every peer's address space over NVLink / NVSwitch).  The engine's head kernel stores the
    normalised text features of this rank's classes straight into EVERY rank's buffer (peer stores from the L2-norm epilogue)
    and publishes an epoch flag; the logits kernel waits on the flags.  No NCCL call, no extra launch, no barrier.

    Buffer layout (bytes): two feature slots ``[n_cls, E]`` fp32 (used by epoch parity: a rank one step ahead never overwrites rows a
    slower peer still reads), each padded to 128 bytes, then 8 int32 flags.
    """

    def __init__(self, engine, shard: ClassSharding, n_cls: int, embed_dim: int, device):
        import torch.distributed._symmetric_memory as symm
        self.n_cls, self.embed_dim = n_cls, embed_dim
        self.slot_floats = ((n_cls * embed_dim * 4 + 127) // 128 * 128) // 4
        self.buf = symm.empty(2 * self.slot_floats + 8, dtype=torch.float32, device=device)
        self.buf.zero_()
        group = shard.group if shard.group is not None else dist.group.WORLD
        self.hdl = symm.rendezvous(self.buf, group)
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=shard.group)                   # every rank's zeroed buffer is mapped before the first peer store
        self.engine, self.rank = engine, shard.rank
        self.epoch = 0

    def next_epoch(self) -> int:
        """Called once per forward on every rank (ranks run in lockstep: same sequence everywhere).  Also points the engine at
        THIS gather's buffers: several FullModels may share one CLIPWrapper / engine (train + eval model), each with its own."""
        if getattr(self.engine, "_active_text_gather", None) is not self:
            self.engine.set_text_gather(list(self.hdl.buffer_ptrs), self.rank, self.n_cls)
            self.engine._active_text_gather = self
        self.epoch += 1
        return self.epoch

    def slot(self, epoch: int) -> torch.Tensor:
        """This rank's ``[n_cls, E]`` view of the slot that holds (or will hold) the features of ``epoch``."""
        o = (epoch & 1) * self.slot_floats
        return self.buf[o: o + self.n_cls * self.embed_dim].view(self.n_cls, self.embed_dim)

    @staticmethod
    def available(device) -> bool:
        if os.environ.get("TAPCLIP_FUSED_GATHER", "1") == "0":
            return False
        if torch.device(device).type != "cuda" or not (dist.is_available() and dist.is_initialized()):
            return False
        if dist.get_backend() != "nccl" or dist.get_world_size() > 8:
            return False
        try:
            import importlib
            importlib.import_module("torch.distributed._symmetric_memory")
        except Exception:
            return False
        return True


def all_reduce_sum_(t: torch.Tensor, shard: ClassSharding) -> torch.Tensor:
    if shard.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=shard.group)
    return t
